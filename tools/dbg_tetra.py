import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/oracle']
import numpy as np
import ntrtsim_b200 as nt
import bench
n_env, steps = 512, 6000
model, rods, cab, ctrl, params, _ = bench.build_workload("tetraspine", n_env, bench.SEED)
for prec in ("f32", "f64"):
    w = nt.BatchedWorld(model, n_env, precision=prec)
    w.set_rod_state(rods); w.set_cable_state(cab); w.set_cpg_controller(model.cpg, params)
    hist = []
    for k in range(12):
        w.step(model.dt, steps // 12)
        hist.append(w.min_clearance(reset=True))
    w.close()
    hist = np.array(hist)
    worst = np.argsort(hist.min(axis=0))[:4]
    print(prec, "worst envs", worst, "their clearance per 500-step window:")
    for e in worst: print("   env", e, np.round(hist[:, e], 3))
