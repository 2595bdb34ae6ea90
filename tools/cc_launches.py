"""A few steps of the contact-cable workload (for an ncu launch list: which of the two kernels of a step takes the time)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ntrtsim_b200 as nt  # noqa: E402

n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model, rods, cab, ctrl, params, _ = bench.build_workload("contact_cable", n_env, bench.SEED)
w = nt.BatchedWorld(model, n_env, precision="f32")
w.set_rod_state(rods)
w.step(model.dt, int(sys.argv[2]) if len(sys.argv) > 2 else 6)
w.sync()
print(w.get_contact_cable_obs()[:2])
