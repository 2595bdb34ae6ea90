import sys,re
for line in open(sys.argv[1]):
    if line.startswith('TGB_STATS'):
        v=[int(x) for x in line.split()[1:]]
        ws=v[50]
        print('warp-steps %d body %.0f cable %.0f gather %.0f cycles/warp-step; total/warp %.0f'%(ws, v[100]/ws, v[101]/ws, v[102]/ws, v[51]/max(1,v[54])/1000))
    elif line.startswith('episodes'): print(line.strip())
