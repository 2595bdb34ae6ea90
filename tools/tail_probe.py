"""What a persistently coupled environment (rods leaning on each other) costs per launch, by layout: the bench workload is run for
`pre` launches, the environments that touch during launch pre+1 are selected, and small worlds of only those environments (every
warp resident, so a launch lasts as long as its slowest warp) are timed over the same launch with 8, 16 and 32 lanes per environment.
Trajectories do not depend on the layout (tested bit for bit), so the same environments touch at the same steps in every world.
usage: python tools/tail_probe.py [pre_launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ntrtsim_b200 as nt
import bench

pre = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_env, H = 4096, 1000
model, rods, cab, ctrl, params, b_alg = bench.build_workload("superball", n_env, bench.SEED)


def run(sel, lanes, wide=None):
    w = nt.BatchedWorld(model, len(sel), precision="f32", lanes_per_env=lanes, wide=wide)
    w.set_rod_state(np.ascontiguousarray(rods[sel])); w.set_cable_state(np.ascontiguousarray(cab[sel]))
    w.set_controller(ctrl, np.ascontiguousarray(params[sel]))
    for _ in range(pre):
        w.step(model.dt, H)
    w.min_clearance(reset=True)
    ms = w.step_timed(model.dt, H, 1)[0]
    clr = w.min_clearance()
    geo = w.launch_geometry()
    w.close()
    return ms, clr, geo


ms, clr, geo = run(np.arange(n_env), 0)
touch = np.where(clr <= 0.02)[0]
print("all %d envs: launch %d takes %.3f ms, geometry %s; %d envs touch during it" % (n_env, pre + 1, ms, geo, len(touch)))
free = np.where(clr > 0.5)[0][:len(touch)]
for name, sel in (("touching", touch), ("never closer than 0.5", free)):
    if len(sel) == 0:
        continue
    for lanes, wide in ((8, False), (8, True), (16, False), (32, False)):
        ms, c2, geo = run(sel, lanes, wide)
        print("%-24s %4d envs lanes %2d wide %-5s geometry %-18s launch %.3f ms (%.0f cycles/step at 1.965 GHz); touching again: %d"
              % (name, len(sel), lanes, wide, geo, ms, ms * 1.965e6 / H, int((c2 <= 0.02).sum())))
