import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ntrtsim_b200 as nt
import bench
n_env = 640
model, rods, cab, ctrl, params, _ = bench.build_workload("superball", n_env, bench.SEED)
def run(split, lanes, cuts):
    w = nt.BatchedWorld(model, n_env, precision="f32", tail_split=split, wide=False, lanes_per_env=lanes)
    w.set_controller(ctrl, params); w.set_rod_state(rods); w.set_cable_state(cab)
    for c in cuts: w.step(model.dt, c)
    r = w.get_rod_state(); clr = w.min_clearance(); n = w.launch_count(); w.close()
    return r, clr, n
ref, clr, _ = run(0, 8, [250, 250, 250])
print("touching envs", (clr <= 0.02).sum())
for name, args in (("cuts of 100", (0, 8, [100] * 7 + [50])), ("lanes 32", (0, 32, [250] * 3)), ("lanes 32 cuts 100", (0, 32, [100] * 7 + [50])),
                   ("forced split", (1, 8, [250] * 3)), ("auto split", (-1, 8, [250] * 3))):
    r, c2, n = run(*args)
    bad = np.where((r != ref).any(axis=(1, 2)))[0]
    print(name, "launches", n, "envs that differ", len(bad), bad[:10], "of which touching", int((clr[bad] <= 0.02).sum()),
          "max diff", np.abs(r - ref).max())
