"""Where does the end-to-end time of tgb_rollout_host go?  Times, for the bench workload: the step kernel alone from the
episode's initial state, the host copies alone, and the whole call for 1/2/4 chunks."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ntrtsim_b200 as nt
from ntrtsim_b200 import _capi as capi
import bench

n_env, H = 4096, 1000
model, rods, cab, ctrl, params, b_alg = bench.build_workload("superball", n_env, bench.SEED)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); v = t.numpy(); v[...] = a; return t, v
rt, rods_p = pinned(rods); ct, cab_p = pinned(cab)
w = nt.BatchedWorld(model, n_env, precision="f32")
ot, rod_out = pinned(np.zeros((n_env, w.R, capi.ROD_STATE_DOUBLES))); bt, cab_out = pinned(np.zeros((n_env, w.C, capi.CABLE_OBS_DOUBLES)))
w.set_controller(ctrl, params)
for rep in range(3):
    w.set_rod_state(rods); w.set_cable_state(cab)
    ms = [w.step_timed(model.dt, H, 1)[0] for _ in range(4)]
    print("kernel ms from the initial state, 4 consecutive episodes of H:", ["%.3f" % m for m in ms])
for chunks in (1, 2, 4, 1, 4):
    os.environ["TGB_ROLLOUT_CHUNKS"] = str(chunks)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); w.rollout_host(rods_p, cab_p, model.dt, H, rod_out, cab_out); ts.append(time.perf_counter() - t0)
    print("chunks", chunks, "rollout_host ms", ["%.3f" % (1e3 * t) for t in ts])
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); w.rollout_host(rods_p, cab_p, model.dt, 0, rod_out, cab_out); ts.append(time.perf_counter() - t0)
    print("chunks", chunks, "copies only (0 steps) ms", ["%.3f" % (1e3 * t) for t in ts])
w.close()
