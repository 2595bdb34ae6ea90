#!/usr/bin/env python
"""Census of body-body proximity on the bench workload with the CPU oracle: per environment the smallest surface distance between
two rods over a horizon, with the pair contacts on (cap 6) and off (cap 0).  VERDICT r1 #1: 10/128 envs interpenetrated within H=1000."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import bench  # noqa: E402
import oracle_py  # noqa: E402
from ntrtsim_b200 import _capi as capi  # noqa: E402


def census(n_env, horizon, cap, workload="superball"):
    model, rods, cab, ctrl, params, _ = bench.build_workload(workload, n_env, bench.SEED)
    mins, rows = [], []
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_pair_contact_cap(cap)
        o.set_rod_state(rods[e])
        o.set_cable_state(cab[e])
        if ctrl:
            o.set_controller(ctrl, params[e])
        steps_with_rows = 0
        for _ in range(horizon // 50):
            o.step(model.dt, 50)
            steps_with_rows += o.num_pair_contacts() > 0
        mins.append(o.min_clearance())
        rows.append(steps_with_rows)
    return np.array(mins), np.array(rows)


if __name__ == "__main__":
    n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    horizon = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    wl = sys.argv[3] if len(sys.argv) > 3 else "superball"
    for cap in (0, 6):
        mins, rows = census(n_env, horizon, cap, wl)
        print("cap %d: %d envs x %d steps: clearance < 0 in %d envs, < -0.1 in %d, < -0.4 (one radius) in %d; min %.3f; envs with pair rows at a sampled step: %d"
              % (cap, n_env, horizon, (mins < 0).sum(), (mins < -0.1).sum(), (mins < -0.4).sum(), mins.min(), (rows > 0).sum()))
