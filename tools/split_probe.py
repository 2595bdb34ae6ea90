"""Launch durations of the bench workload, launch by launch, with the tail split's counts (TGB_SPLIT_DEBUG=1) and without the split."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ntrtsim_b200 as nt
import bench

n_env, H = 4096, 1000
model, rods, cab, ctrl, params, b_alg = bench.build_workload("superball", n_env, bench.SEED)
for split in (0, -1):
    w = nt.BatchedWorld(model, n_env, precision="f32", tail_split=split)
    w.set_rod_state(rods); w.set_cable_state(cab); w.set_controller(ctrl, params)
    ms = [w.step_timed(model.dt, H, 1)[0] for _ in range(10)]
    print("split", split, "ms per launch", ["%.2f" % m for m in ms], file=sys.stderr)
    w.close()
