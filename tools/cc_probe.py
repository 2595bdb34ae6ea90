"""Per-checkpoint agreement of the contact-cable path (device, f64) with the oracle on the ContactCableDemo ensemble, and the oracle's own
sensitivity to a 1e-13 perturbation of its input (the yardstick: anchor insertions and removals amplify round-off)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import test_contact_cables as t  # noqa: E402
from helpers import rod_errors  # noqa: E402

m = t.demo_model()
states = t.demo_states(m, 32)
dev = t._device_run(m, states, "f64", 7000, 500)
ora = t._oracle_run(m, states, 7000, 500)
pert = states.copy()
pert[:, :, 0:3] += 1e-13 * np.random.default_rng(0).normal(size=pert[:, :, 0:3].shape)
orp = t._oracle_run(m, pert, 7000, 500)
for k, ((rd, od, ad), (ro, oo, ao), (rp, op, ap)) in enumerate(zip(dev, ora, orp)):
    print("step %5d  dev-vs-oracle rods %.2e  len %.2e  anchors equal %d/32 | oracle self-sensitivity rods %.2e anchors equal %d/32 | touching %d" % (
        (k + 1) * 500, max(rod_errors(rd, ro, 10.0)), np.abs(od[..., 1] - oo[..., 1]).max(), int((od[:, 0, 4] == oo[:, 0, 4]).sum()),
        max(rod_errors(rp, ro, 10.0)), int((op[:, 0, 4] == oo[:, 0, 4]).sum()), int((oo[:, 0, 4] > 0).sum())))
