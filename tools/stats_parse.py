import sys,re
ctas = []
def flush_ctas():
    global ctas
    if not ctas: return
    ctas.sort(key=lambda r: -r[1])
    tot = sorted(r[1] for r in ctas)
    print('  per-CTA total cycles: median %.3g, p90 %.3g, p99 %.3g, max %.3g' % (tot[len(tot)//2], tot[int(len(tot)*0.9)], tot[int(len(tot)*0.99)], tot[-1]))
    print('  slowest CTAs: id total | coupled steps, coupled-solve cycles, detection cycles, mem-solver steps, steps with >=3 contacts, cable cycles')
    for r in ctas[:12]: print('   ', r)
    slow = [r for r in ctas if r[0] >= 2048]
    if slow:
        print('  one-warp-per-environment launch (ids from 2048): %d CTAs, slowest %.3g, median %.3g' % (len(slow), max(r[1] for r in slow), sorted(r[1] for r in slow)[len(slow)//2]))
        usual = sorted((r for r in ctas if r[0] < 2048), key=lambda r: -r[1])
        print('  usual layout: slowest %.3g, p99 %.3g, median %.3g' % (usual[0][1], usual[len(usual)//100][1], usual[len(usual)//2][1]))
    ctas = []
for line in open(sys.argv[1]):
    if line.startswith('TGB_CTA'):
        ctas.append([int(x) for x in line.split()[1:]]); continue
    if line.startswith('episodes'): flush_ctas()
    if line.startswith('TGB_STATS'):
        v=[int(x) for x in line.split()[1:]]
        ws=v[50]
        print('warp-steps %d body %.0f cable %.0f gather %.0f cycles/warp-step; total/warp %.0f'%(ws, v[100]/ws, v[101]/ws, v[102]/ws, v[51]/max(1,v[54])/1000))
        print('  per-warp cycles/step: mean %.0f  slowest %.0f  fastest %.0f' % (v[51]/max(1,v[54])/1000, v[52]/1000, v[53]/1000))
        print('  pair detection %.0f cycles/warp-step, %.2f pair tests/warp-step, %.3f emissions/warp-step' % (v[103]/ws, v[58]/ws, v[59]/ws))
        cs = max(1, v[55])
        print('  coupled warp-steps %.1f%% (%.1f%% of them with the shared-memory solver), %.2f coupled envs per coupled warp-step, %.0f cycles per coupled solve'
              % (100.0*v[55]/ws, 100.0*v[56]/cs, v[57]/cs, v[104]/cs))
        print('  coupled solve sections (cycles per coupled solve): own rows %.0f, pair set-up %.0f, W %.0f, split impulse %.0f, iterations %.0f; body phase before it %.0f'
              % (v[105]/cs, v[106]/cs, v[107]/cs, v[108]/cs, v[109]/cs, v[110]/cs))
        print('  contact-count histogram (max nc of a lane per warp-step):', [round(100.0*x/ws,1) for x in v[0:9]])
    elif line.startswith('episodes'): print(line.strip())

flush_ctas()
