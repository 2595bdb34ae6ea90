#!/usr/bin/env python
"""How sharp is the reference suite's ICRA 2015 hardware comparison (test_integration/ICRA2015Tests/ICRA2015_test.cpp:77-150) as a pin
for this simulator?  Runs the test's 60000 steps on the CPU oracle for an ensemble of initial states moved by 1e-9, with and without
warm starting, and prints the distribution of the six errors next to the reference's bounds.  (VERDICT r1, next #2a.)"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import ntrtsim_b200 as nt  # noqa: E402
from test_golden import ICRA_BOUND, icra_errors  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    model = nt.build_example("TetraSpineStatic_hf")
    row = nt.serialized_spine_params(model)
    names = ["outer front right", "outer front left", "outer front top", "inner front right", "inner front left", "inner front top"]
    for warm in (0.0, 0.85):
        with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
            e = np.array(list(ex.map(lambda s: icra_errors(model, row, warm, s), range(n))))
        print("warm starting %.2f, %d runs (seed 0 = the unperturbed run): members meeting all six reference bounds: %d" % (warm, n, int(np.all(e <= ICRA_BOUND, axis=1).sum())))
        for k, nm in enumerate(names):
            print("  %-18s bound %.4f | unperturbed %.4f | min %.4f median %.4f max %.4f | members within the bound %2d / %d"
                  % (nm, ICRA_BOUND[k], e[0, k], e[:, k].min(), np.median(e[:, k]), e[:, k].max(), int((e[:, k] <= ICRA_BOUND[k]).sum()), n))
