"""The reference suite's regression number test_integration/SpineTests/WorldConf_Spines_test.cpp:156 (EXPECT_EQ(dist, 215.278) after
15000 steps of the 12-vertebra kinematic spine, learning off) cannot be a constant of the reference itself:
  * with learning off NeuroAdapter::initialize loads logs/bestParameters-default_{node,edge}-i.nnw into the controllers
    (NeuroAdapter.cpp:57-67);
  * every shipped file is "a,b" without a trailing comma, and NeuroEvoMember::loadFromFile drops the value after the last comma
    (NeuroEvoMember.cpp:163-171: getline hits end of file), so only parameter 0 of each controller is set;
  * parameter 1 keeps what the constructor drew: rand() * 1.0 / RAND_MAX (NeuroEvoMember.cpp:47-48) after
    srand(rdtsc()) (NeuroEvolution.cpp:87) -- the time stamp counter.
So in the reference run the node controller's second action (the CPG amplitude, scaled to [-30, 30]) and the second action of all
100 edge controllers (the coupling phase offsets, scaled to [-pi, pi]) are fresh uniform random numbers on every execution.
This script draws them the same way for a population of environments -- parameter 0 from the shipped files, parameter 1 ~ U(0, 1)
per controller -- runs the 15000 steps of the test on the device and prints the distribution of the score the test reads
(XZ distance moved by segment 6, or -1 for a bogus episode).
usage: python tools/spine215_sweep.py [n_env] [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ntrtsim_b200 as nt
from ntrtsim_b200 import neuroevo

n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
precision = sys.argv[2] if len(sys.argv) > 2 else "f64"
steps = 15000
example = "FlemonsSpine"
g = np.load(os.path.join(ROOT, "tests", "golden", "flemonsspine_kinematic_cpg.npz"))
rng = np.random.default_rng(215278)
node = np.tile(g["node_actions"], (n_env, 1))
edge = np.tile(g["edge_actions"], (n_env, 1, 1))
# environment 0: the files' own second values (what a loader without the quirk would use); the others: the reference's behaviour
node[1:, 1] = rng.uniform(size=n_env - 1)
edge[1:, :, 1] = rng.uniform(size=(n_env - 1, edge.shape[1]))
model = nt.build_example(example)
model.cpg = nt.flemons_cpg(example)
rows = nt.flemons_param_rows(node, edge, example)
scores, status = neuroevo.evaluate_population(model, model.cpg, rows, steps, precision=precision, steps_per_launch=1000)
d = scores[:, 0]
ok = d >= 0
print("WorldConf_Spines (12 vertebrae, %d steps, %s): %d parameter draws" % (steps, precision, n_env))
print("  files' own second values (no loader quirk): score %.3f" % d[0])
print("  bogus episodes (score -1): %d of %d" % (int((~ok).sum()), n_env))
q = np.percentile(d[ok], [0, 1, 5, 25, 50, 75, 95, 99, 100])
print("  distance moved, valid episodes: min %.2f  p1 %.2f  p5 %.2f  p25 %.2f  median %.2f  p75 %.2f  p95 %.2f  p99 %.2f  max %.2f" % tuple(q))
print("  draws with a score of at least 215.278: %d (%.2f %%); within 1 %% of it: %d" % (int((d >= 215.278).sum()), 100.0 * (d >= 215.278).mean(),
                                                                                     int((np.abs(d - 215.278) < 2.15).sum())))
edges = [0, 5, 10, 20, 40, 80, 120, 160, 200, 215.278, 240, 280, 320, 400, 1e9]
h, _ = np.histogram(d[ok], bins=edges)
for lo, hi, c in zip(edges[:-1], edges[1:], h):
    print("  [%7.1f, %7s)  %5d  %s" % (lo, ("%.1f" % hi) if hi < 1e8 else "inf", c, "#" * int(60.0 * c / max(1, h.max()))))
best = int(np.argmax(d))
print("  best draw: env %d, node amplitude action %.4f, score %.3f" % (best, node[best, 1], d[best]))
amp = node[:, 1]
for lo, hi in ((0.0, 0.25), (0.25, 0.45), (0.45, 0.55), (0.55, 0.75), (0.75, 1.0)):
    m = ok & (amp >= lo) & (amp < hi)
    if m.any():
        print("  node amplitude action in [%.2f, %.2f): %4d draws, median score %.2f, max %.2f" % (lo, hi, int(m.sum()), np.median(d[m]), d[m].max()))
