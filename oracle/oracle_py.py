"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        raise OSError("%s missing: run `make oracle/liboracle.so`" % path)
    lib = C.CDLL(path)
    vp, dbl, i = C.c_void_p, C.c_double, C.c_int
    lib.tgo_create.restype = vp
    lib.tgo_create.argtypes = [vp, vp]
    lib.tgo_destroy.argtypes = [vp]
    lib.tgo_add_model_flat.argtypes = [vp, vp, i, vp, i, vp, i]
    lib.tgo_num_bodies.argtypes = [vp]
    lib.tgo_num_cables.argtypes = [vp]
    lib.tgo_set_rod_state.argtypes = [vp, vp]
    lib.tgo_get_rod_state.argtypes = [vp, vp]
    lib.tgo_set_cable_state.argtypes = [vp, vp]
    lib.tgo_get_cable_obs.argtypes = [vp, vp]
    lib.tgo_set_controller.argtypes = [vp, i, vp, i]
    lib.tgo_set_control_input.argtypes = [vp, vp, dbl]
    lib.tgo_step.argtypes = [vp, dbl, i]
    lib.tgo_world_step.argtypes = [vp, dbl]
    lib.tgo_models_step.argtypes = [vp, dbl]
    lib.tgo_get_time.restype = dbl
    lib.tgo_get_time.argtypes = [vp]
    lib.tgo_mark_episode_start.argtypes = [vp]
    lib.tgo_get_scores.argtypes = [vp, vp]
    lib.tgo_num_contacts.argtypes = [vp]
    lib.tgo_num_pair_contacts.argtypes = [vp]
    lib.tgo_set_pair_contact_cap.argtypes = [vp, i]
    lib.tgo_set_default_pair_contact_cap.argtypes = [i]
    lib.tgo_set_warm_start.argtypes = [vp, dbl]
    lib.tgo_get_min_clearance.restype = dbl
    lib.tgo_get_min_clearance.argtypes = [vp]
    lib.tgo_reset_min_clearance.argtypes = [vp]
    lib.tgo_rollout_batch.argtypes = [vp, vp, vp, i, vp, i, vp, i, i, vp, i, vp, dbl, i, vp, vp, dbl, i, i, vp, vp, vp, vp, vp]
    lib.tgo_set_cpg_controller.argtypes = [vp, vp, vp, i]
    lib.tgo_get_status.argtypes = [vp]
    lib.tgo_get_cpg_state.argtypes = [vp, vp]
    lib.tgo_add_contact_cables.argtypes = [vp, vp, i]
    lib.tgo_set_batch_contact_cables.argtypes = [vp, i]
    lib.tgo_contact_cable_points.argtypes = [vp, i, vp, i, vp]
    lib.tgo_set_contact_cable_control_input.argtypes = [vp, vp, dbl]
    lib.tgo_num_contact_cables.argtypes = [vp]
    lib.tgo_get_contact_cable_obs.argtypes = [vp, vp]
    lib.tgo_get_contact_cable_anchors.argtypes = [vp, i, vp]
    lib.tgo_track_max_tension.argtypes = [vp, i]
    lib.tgo_get_max_tension.argtypes = [vp, vp]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleWorld(object):
    """One environment stepped by the CPU oracle."""

    def __init__(self, model, cfg_cls, solver_iterations=0):
        self._lib = load()
        self.model = model
        cfg = cfg_cls(model.gravity, 1, 1, 0, solver_iterations)
        self._h = C.c_void_p(self._lib.tgo_create(C.byref(cfg), C.byref(model.ground)))
        self._lib.tgo_add_model_flat(self._h, model.bodies, len(model.bodies), model.shapes, len(model.shapes), model.cables,
                                     len(model.cables))
        self.R, self.C = len(model.bodies), len(model.cables)
        self.n_cc = 0
        if getattr(model, "contact_cables", None) is not None:
            rc = self._lib.tgo_add_contact_cables(self._h, model.contact_cables, len(model.contact_cables))
            if rc != 0:
                raise ValueError("oracle: bad contact cable descriptor (%d)" % rc)
            self.n_cc = len(model.contact_cables)

    def get_contact_cable_obs(self):
        out = np.zeros((self.n_cc, 5))
        self._lib.tgo_get_contact_cable_obs(self._h, _p(out))
        return out

    def contact_cable_points(self, cable, chain):
        """the geometric manifold of a cable whose anchors stand at chain [n, 3]: [k, 8] = body, position, normal into the body, distance"""
        chain = np.ascontiguousarray(chain, dtype=np.float64)
        out = np.zeros((64, 8))
        k = self._lib.tgo_contact_cable_points(self._h, int(cable), _p(chain), chain.shape[0], _p(out))
        return out[:k]

    def world_step(self, dt):
        self._lib.tgo_world_step(self._h, dt)

    def models_step(self, dt):
        self._lib.tgo_models_step(self._h, dt)

    def set_contact_cable_control_input(self, inputs, dt):
        inputs = np.ascontiguousarray(inputs, dtype=np.float64)
        return self._lib.tgo_set_contact_cable_control_input(self._h, _p(inputs), dt)

    def get_contact_cable_anchors(self, cable):
        out = np.zeros((12, 7))
        self._lib.tgo_get_contact_cable_anchors(self._h, int(cable), _p(out))
        return out

    def close(self):
        if self._h:
            self._lib.tgo_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_rod_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        self._lib.tgo_set_rod_state(self._h, _p(s))

    def get_rod_state(self):
        out = np.empty((self.R, 13))
        self._lib.tgo_get_rod_state(self._h, _p(out))
        return out

    def set_cable_state(self, rp):
        rp = np.ascontiguousarray(rp, dtype=np.float64)
        self._lib.tgo_set_cable_state(self._h, _p(rp))

    def get_cable_obs(self):
        out = np.empty((self.C, 4))
        self._lib.tgo_get_cable_obs(self._h, _p(out))
        return out

    def set_controller(self, kind, params):
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._lib.tgo_set_controller(self._h, kind, _p(params), params.size)

    def track_max_tension(self, enable=True):
        self._lib.tgo_track_max_tension(self._h, 1 if enable else 0)

    def max_tension(self):
        out = np.zeros(self.C)
        self._lib.tgo_get_max_tension(self._h, _p(out))
        return out

    def set_control_input(self, inputs, dt):
        inputs = np.ascontiguousarray(inputs, dtype=np.float64)
        return self._lib.tgo_set_control_input(self._h, _p(inputs), dt)

    def step(self, dt, n=1):
        return self._lib.tgo_step(self._h, dt, n)

    def num_contacts(self):
        return self._lib.tgo_num_contacts(self._h)

    def num_pair_contacts(self):
        """rows between two bodies in the last world step"""
        return self._lib.tgo_num_pair_contacts(self._h)

    def set_pair_contact_cap(self, cap):
        self._lib.tgo_set_pair_contact_cap(self._h, int(cap))

    def set_warm_start(self, factor):
        """warm starting factor of the contact solver (Bullet: 0.85); 0 = no persistent manifold"""
        self._lib.tgo_set_warm_start(self._h, float(factor))

    def min_clearance(self, reset=False):
        """smallest surface distance between two bodies' shapes seen by any world step so far"""
        v = self._lib.tgo_get_min_clearance(self._h)
        if reset:
            self._lib.tgo_reset_min_clearance(self._h)
        return v

    def scores(self):
        out = np.empty(2)
        self._lib.tgo_get_scores(self._h, _p(out))
        return out

    def mark_episode_start(self):
        self._lib.tgo_mark_episode_start(self._h)

    def time(self):
        return self._lib.tgo_get_time(self._h)

    def set_cpg_controller(self, desc, params):
        params = np.ascontiguousarray(params, dtype=np.float64)
        return self._lib.tgo_set_cpg_controller(self._h, C.byref(desc), _p(params), params.size)

    def status(self):
        return self._lib.tgo_get_status(self._h)

    def cpg_state(self):
        out = np.empty((self.C, 3))
        self._lib.tgo_get_cpg_state(self._h, _p(out))
        return out


def set_default_pair_contact_cap(cap):
    """pair-contact row cap of oracle worlds created from now on (rollout_batch creates its own worlds); 0 = bodies do not collide"""
    load().tgo_set_default_pair_contact_cap(int(cap))


def rollout_batch(model, cfg_cls, rod_in, cable_in, dt, n_steps, ctrl_kind=0, ctrl_params=None, setup_input=None, setup_dt=0.0,
                  n_threads=1, solver_iterations=0, cpg_desc=None, want_scores=False):
    """n_env independent oracle environments from their own states; returns (rod_state, cable_obs) or, with
    want_scores, (rod_state, cable_obs, scores, status). cpg_desc: a ctypes tgb_cpg_desc for TGB_CTRL_CPG_IMPEDANCE."""
    lib = load()
    rod_in = None if rod_in is None else np.ascontiguousarray(rod_in, dtype=np.float64)
    cable_in = None if cable_in is None else np.ascontiguousarray(cable_in, dtype=np.float64)
    R, Cn = len(model.bodies), len(model.cables)
    n_env = rod_in.shape[0] if rod_in is not None else (cable_in.shape[0] if cable_in is not None else 1)
    per_env = 0
    if ctrl_params is not None:
        ctrl_params = np.ascontiguousarray(ctrl_params, dtype=np.float64).reshape(n_env, -1)
        per_env = ctrl_params.shape[1]
    if setup_input is not None:
        setup_input = np.ascontiguousarray(setup_input, dtype=np.float64)
    cfg = cfg_cls(model.gravity, 1, 1, 0, solver_iterations)
    rod_out = np.empty((n_env, R, 13))
    cab_out = np.empty((n_env, Cn, 4))
    scores = np.empty((n_env, 2))
    status = np.zeros(n_env, dtype=np.int32)
    cc = getattr(model, "contact_cables", None)
    lib.tgo_set_batch_contact_cables(cc, 0 if cc is None else len(cc))
    rc = lib.tgo_rollout_batch(C.byref(cfg), C.byref(model.ground), model.bodies, R, model.shapes, len(model.shapes), model.cables, Cn,
                               ctrl_kind, _p(ctrl_params), per_env, _p(setup_input), setup_dt, n_env, _p(rod_in), _p(cable_in), dt,
                               n_steps, n_threads, _p(rod_out), _p(cab_out), None if cpg_desc is None else C.byref(cpg_desc),
                               _p(scores), _p(status))
    if rc != 0:
        raise ValueError("oracle rollout failed: %d" % rc)
    if want_scores:
        return rod_out, cab_out, scores, status
    return rod_out, cab_out
