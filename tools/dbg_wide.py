import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/oracle']
import numpy as np
import ntrtsim_b200 as nt
from ntrtsim_b200 import _capi as capi
from test_gpu_parity import device_rollout, escape_inputs
model = nt.build_example("Escape_c3")
n_env, steps = 70, 120
rods, cab, params = escape_inputs(model, n_env, seed=5)
for extra in (dict(), dict(warm_starting=0.0), dict(pair_contact_cap=0), dict(warm_starting=0.0, pair_contact_cap=0)):
    for precision in ("f32", "f64"):
        ref = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params, wide=False, **extra)
        ref2 = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params, wide=False, **extra)
        out = []
        for kw in (dict(lanes_per_env=16, wide=False), dict(lanes_per_env=32, block_threads=64), dict(block_threads=128, wide=False), dict(chunks=7, wide=False), dict(wide=True), dict(wide=True, chunks=7)):
            got = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params, **kw, **extra)
            d = np.abs(got[0] - ref[0]).max()
            bad = np.where(np.abs(got[0] - ref[0]).max(axis=(1, 2)) > 0)[0]
            out.append((kw, float(d), bad[:6].tolist(), len(bad)))
        print(extra, precision, "repeat identical:", np.array_equal(ref[0], ref2[0]))
        for o in out: print("    ", o)
