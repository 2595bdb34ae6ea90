#!/usr/bin/env python
"""Contact cables: the REFERENCE's own src/core/tgBulletContactSpringCable.cpp + tgBulletSpringCableAnchor.cpp (compiled in place,
oracle/ref/build_ref.py) beside the oracle's restatement, on the ContactCableDemo ensemble.

Hybrid run, per step: the oracle's world step moves the bodies; the reference cable's anchor positions go to the oracle's geometric
manifold (tgo_contact_cable_points); the points fill the shim's btPersistentManifolds; the reference's own
tgBulletContactSpringCable::step runs (updateManifolds, updateAnchorList with its findNearestPastAnchor, pruneAnchors,
calculateAndApplyForce) and applies its impulses to the shim bodies; the velocities go back to the oracle's bodies.
Pure run: tgo_step.  What is compared: rod states and the anchor lists.

TEST INFRASTRUCTURE.  `python oracle/ref/cc_hybrid.py [n_env] [steps]` prints the comparison; make_cc_golden() writes
tests/golden/contact_cable_demo.npz (the hybrid run's checkpoints) for the boxes that have no /root/reference."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import ntrtsim_b200 as nt  # noqa: E402
from ntrtsim_b200 import _capi as capi  # noqa: E402
import oracle_py  # noqa: E402
import ref_py  # noqa: E402

DT = 0.001


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _quat_rotate(q, v):
    x, y, z, w = q
    u = np.array([x, y, z])
    return v + 2.0 * np.cross(u, np.cross(u, v) + w * v)


class HybridWorld(object):
    """bodies in the oracle (world half), the contact cable in the reference's own code (models half)"""

    def __init__(self, model, state):
        self.lib = ref_py.load()
        lib = self.lib
        vp, d, i = C.c_void_p, C.c_double, C.c_int
        lib.ref_add_contact_cable.argtypes = [vp, i, i, vp, vp, vp]
        lib.ref_cc_chain.argtypes = [vp, i, vp, i]
        lib.ref_cc_set_points.argtypes = [vp, i, i, vp]
        lib.ref_cc_step.argtypes = [vp, i, d]
        lib.ref_cc_get.argtypes = [vp, i, vp]
        lib.ref_cc_anchors.argtypes = [vp, i, vp, i]
        lib.ref_cc_set_control_input_dt.argtypes = [vp, i, d, d]
        self.model = model
        self.ora = oracle_py.OracleWorld(model, capi.WorldCfg)
        self.ora.set_rod_state(state)
        self.w = C.c_void_p(lib.ref_create())
        for b, s in zip(model.bodies, state):
            lib.ref_add_body(self.w, _p(np.ascontiguousarray(s)), b.mass, _p(np.array(list(b.inertia))))
        self.cc = []
        for c in model.contact_cables:
            pa = state[c.body_a, 0:3] + _quat_rotate(state[c.body_a, 3:7], np.array(list(c.local_a)))
            pb = state[c.body_b, 0:3] + _quat_rotate(state[c.body_b, 3:7], np.array(list(c.local_b)))
            pretension = (np.linalg.norm(pb - pa) - c.rest_length) * c.stiffness
            cfg = np.array([c.stiffness, c.damping, pretension, c.thickness, c.resolution, c.target_velocity])
            idx = lib.ref_add_contact_cable(self.w, c.body_a, c.body_b, _p(np.ascontiguousarray(pa)), _p(np.ascontiguousarray(pb)), _p(cfg))
            if idx < 0:
                raise RuntimeError(lib.ref_last_error().decode())
            self.cc.append(idx)

    def close(self):
        if self.w:
            self.lib.ref_destroy(self.w)
            self.w = None

    def step(self, n=1, control=None):
        """control: preferred length per contact cable, handed to the reference's tgBasicActuator::setControlInput(input, dt) before
        every step (a model-level controller's onStep runs before the children step, tgModel.cpp / tgSubject)"""
        lib, R = self.lib, len(self.model.bodies)
        for _ in range(n):
            self.ora.world_step(DT)
            if control is not None:
                st = self.ora.get_rod_state()
                for b in range(R):
                    lib.ref_set_body(self.w, b, _p(np.ascontiguousarray(st[b])))
                for k, idx in enumerate(self.cc):
                    if lib.ref_cc_set_control_input_dt(self.w, idx, float(control[k]), DT) != 0:
                        raise RuntimeError(lib.ref_last_error().decode())
            st = self.ora.get_rod_state()
            for b in range(R):
                lib.ref_set_body(self.w, b, _p(np.ascontiguousarray(st[b])))
            for k, idx in enumerate(self.cc):
                chain = np.zeros((16, 3))
                na = lib.ref_cc_chain(self.w, idx, _p(chain), 16)
                pts = np.ascontiguousarray(self.ora.contact_cable_points(k, chain[:na]))
                lib.ref_cc_set_points(self.w, idx, pts.shape[0], _p(pts))
                rc = lib.ref_cc_step(self.w, idx, DT)
                if rc != 0:
                    raise RuntimeError(lib.ref_last_error().decode())
            out = np.zeros(13)
            for b in range(R):
                lib.ref_get_body(self.w, b, _p(out))
                st[b] = out
            self.ora.set_rod_state(st)

    def rod_state(self):
        return self.ora.get_rod_state()

    def cable(self, k=0):
        out = np.zeros(4)
        self.lib.ref_cc_get(self.w, self.cc[k], _p(out))
        return out

    def anchors(self, k=0):
        out = np.zeros((16, 7))
        n = self.lib.ref_cc_anchors(self.w, self.cc[k], _p(out), 16)
        return out[:n]


def compare(n_env=8, steps=7000, chunk=250, seed=77, verbose=True):
    import test_contact_cables as t
    m = t.demo_model()
    states = t.demo_states(m, n_env, seed=seed)
    rows = []
    for e in range(n_env):
        hyb = HybridWorld(m, states[e])
        ora = oracle_py.OracleWorld(m, capi.WorldCfg)
        ora.set_rod_state(states[e])
        worst, same, checks, nmax = 0.0, 0, 0, 0
        for k in range(steps // chunk):
            hyb.step(chunk)
            ora.step(DT, chunk)
            rh, ro = hyb.rod_state(), ora.get_rod_state()
            err = np.abs(rh - ro).max()
            ah, ao = hyb.anchors(), ora.get_contact_cable_anchors(0)
            no = int(ora.get_contact_cable_obs()[0, 4])
            checks += 1
            same += int(ah.shape[0] == no and (no == 0 or (np.array_equal(ah[:, 0], ao[:no, 0]) and np.abs(ah[:, 1:] - ao[:no, 1:]).max() < 1e-5)))
            worst = max(worst, err)
            nmax = max(nmax, no)
        rows.append((e, worst, same, checks, nmax))
        if verbose:
            print("env %2d: max |hybrid - oracle| over rod states %.3e, anchor lists equal at %d / %d checkpoints, most anchors %d" % rows[-1])
        hyb.close()
    return rows


def make_cc_golden(path=None, n_env=4, steps=7000, chunk=500, seed=77):
    """checkpoints of the hybrid run (reference cable code + oracle world step) for tests/test_golden-style use without /root/reference"""
    import test_contact_cables as t
    m = t.demo_model()
    states = t.demo_states(m, n_env, seed=seed)
    rods = np.zeros((n_env, steps // chunk, len(m.bodies), 13))
    cab = np.zeros((n_env, steps // chunk, 4))
    anch = np.full((n_env, steps // chunk, 12, 7), 0.0)
    anch[..., 0] = -1.0
    for e in range(n_env):
        hyb = HybridWorld(m, states[e])
        for k in range(steps // chunk):
            hyb.step(chunk)
            rods[e, k] = hyb.rod_state()
            cab[e, k] = hyb.cable()
            a = hyb.anchors()[:12]
            anch[e, k, :a.shape[0]] = a
        hyb.close()
    path = path or os.path.join(ROOT, "tests", "golden", "contact_cable_demo.npz")
    np.savez_compressed(path, states=states, rods=rods, cable=cab, anchors=anch, chunk=chunk, seed=seed)
    return path


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    s = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
    compare(n, s)
