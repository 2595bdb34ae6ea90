"""Contact / warp-time statistics of the step kernel (needs a -DTGB_STATS build: tools/build_variant.sh stats -DTGB_STATS;
TGB_LIB_DIR=variants/stats TGB_STATS=1 [TGB_STATS_RESET=1] python tools/stats_probe.py [launches ...]).  The library prints the raw
counters on stderr when a world is destroyed (with TGB_STATS_RESET: those of its last launch only); this script runs the bench
workload for N launches per world so that they can be told apart."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntrtsim_b200 as nt
import bench

n_env, H = 4096, 1000
counts = [int(x) for x in sys.argv[1:]] or [1, 5]
model, rods, cab, ctrl, params, b_alg = bench.build_workload("superball", n_env, bench.SEED)
for episodes in counts:
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_rod_state(rods); w.set_cable_state(cab); w.set_controller(ctrl, params)
    ms = [w.step_timed(model.dt, H, 1)[0] for _ in range(episodes)]
    print("episodes", episodes, "ms", ["%.3f" % m for m in ms], file=sys.stderr)
    w.close()
