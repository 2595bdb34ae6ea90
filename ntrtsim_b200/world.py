"""BatchedWorld: numpy-facing wrapper of one tgb_world (N environments of one flattened NTRTsim model).
Every method is a single C-ABI call; all arithmetic happens in the CUDA library."""
import ctypes as C

import numpy as np

from . import _capi as capi


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _host_buffer(a, shape, what, writable=False):
    """A C-contiguous float64 array of exactly `shape`: the C ABI reads / writes raw doubles, so anything else is an error here
    rather than garbage there.  Inputs are converted; output buffers must already be right (a converted copy would be lost)."""
    if writable:
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable and a.size == int(np.prod(shape))):
            raise ValueError("%s must be a writable C-contiguous float64 array of %d elements" % (what, int(np.prod(shape))))
        return a
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.size != int(np.prod(shape)):
        raise ValueError("%s has %d elements, expected %s" % (what, a.size, "x".join(str(int(x)) for x in shape)))
    return a


def _rows(params, n, broadcast, what):
    """env_count of a parameter array: one row when broadcast, else exactly one row per environment of the range"""
    params = np.ascontiguousarray(params, dtype=np.float64)
    rows = 1 if params.ndim == 1 else int(np.prod(params.shape[:-1]))
    if broadcast:
        if rows != 1:
            raise ValueError("%s: a broadcast takes one parameter row, got %d" % (what, rows))
    elif rows != n:
        raise ValueError("%s: %d parameter rows for %d environments" % (what, rows, n))
    return params


class FlatModel(object):
    """The flattened descriptors of one model: what tgStructureInfo::buildInto hands to the world
    (reference: src/tgcreator/tgStructureInfo.cpp:277-289) plus the world settings of its app."""

    def __init__(self, bodies, shapes, cables, ground, gravity, dt, name="", contact_cables=None):
        self.bodies, self.shapes, self.cables = bodies, shapes, cables
        self.ground, self.gravity, self.dt, self.name = ground, gravity, dt, name
        # tgBulletContactSpringCable cables (tgBasicContactCableInfo): a ctypes array of tgb_contact_cable_desc, or None
        self.contact_cables = contact_cables if contact_cables is not None and len(contact_cables) else None

    @property
    def n_contact_cables(self):
        return 0 if self.contact_cables is None else len(self.contact_cables)

    @property
    def n_bodies(self):
        return len(self.bodies)

    @property
    def n_cables(self):
        return len(self.cables)

    def initial_rod_state(self):
        s = np.zeros((self.n_bodies, capi.ROD_STATE_DOUBLES))
        for i, b in enumerate(self.bodies):
            s[i, 0:3] = list(b.pos)
            s[i, 3:7] = list(b.quat)
        return s

    def initial_cable_state(self):
        return np.array([[c.rest_length, c.prev_length] for c in self.cables], dtype=np.float64).reshape(self.n_cables, 2)


class BatchedWorld(object):
    def __init__(self, model, n_env, precision="f32", device=0, solver_iterations=0, lanes_per_env=0, block_threads=0, wide=None,
                 pair_contact_cap=None, warm_starting=None, tail_split=None):
        self._lib = capi.load()
        self.model = model
        self.n_env = int(n_env)
        self.precision = precision
        cfg = capi.WorldCfg(model.gravity, self.n_env, capi.TGB_F64 if precision == "f64" else capi.TGB_F32, device, solver_iterations)
        h = C.c_void_p()
        capi.check(self._lib.tgb_create(C.byref(cfg), C.byref(model.ground), C.byref(h)))
        self._h = h
        try:
            capi.check(self._lib.tgb_add_model_flat(h, model.bodies, len(model.bodies), model.shapes, len(model.shapes),
                                                    model.cables, len(model.cables)))
            if getattr(model, "contact_cables", None) is not None:
                capi.check(self._lib.tgb_add_contact_cables(h, model.contact_cables, len(model.contact_cables)))
            if lanes_per_env:
                capi.check(self._lib.tgb_set_lanes_per_env(h, lanes_per_env))
            if block_threads:
                capi.check(self._lib.tgb_set_block_threads(h, block_threads))
            if wide is not None:
                capi.check(self._lib.tgb_set_wide_layout(h, 1 if wide else 0))
            if pair_contact_cap is not None:
                capi.check(self._lib.tgb_set_pair_contact_cap(h, int(pair_contact_cap)))
            if warm_starting is not None:
                capi.check(self._lib.tgb_set_warm_starting(h, float(warm_starting)))
            if tail_split is not None:
                capi.check(self._lib.tgb_set_tail_split(h, int(tail_split)))
            capi.check(self._lib.tgb_finalize(h))
        except Exception:
            self.close()
            raise
        self.R, self.C = model.n_bodies, model.n_cables

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tgb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state -------------------------------------------------------------------------------------------
    def set_rod_state(self, state, env_begin=0):
        state = np.ascontiguousarray(state, dtype=np.float64).reshape(-1, self.R, capi.ROD_STATE_DOUBLES)
        capi.check(self._lib.tgb_set_rod_state(self._h, env_begin, state.shape[0], _ptr(state)))

    def get_rod_state(self, env_begin=0, env_count=None):
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.empty((n, self.R, capi.ROD_STATE_DOUBLES))
        capi.check(self._lib.tgb_get_rod_state(self._h, env_begin, n, _ptr(out)))
        return out

    def get_contact_cable_obs(self, env_begin=0, env_count=None):
        """[n, n_contact_cables, 5]: rest length, actual length over all anchors, tension, velocity, number of sliding anchors"""
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.zeros((n, self.model.n_contact_cables, capi.CCABLE_OBS_DOUBLES))
        capi.check(self._lib.tgb_get_contact_cable_obs(self._h, env_begin, n, _ptr(out)))
        return out

    def set_contact_cable_control_input(self, inputs, dt, env_begin=0, env_count=None):
        """tgBasicActuator::setControlInput(input, dt) on the contact cables: [n_contact_cables] for every environment of the range, or
        one row per environment"""
        n = self.n_env - env_begin if env_count is None else env_count
        inputs = np.ascontiguousarray(inputs, dtype=np.float64)
        broadcast = inputs.ndim == 1
        want = self.model.n_contact_cables * (1 if broadcast else n)
        if inputs.size != want:
            raise ValueError("contact cable control input has %d elements, expected %d" % (inputs.size, want))
        capi.check(self._lib.tgb_set_contact_cable_control_input(self._h, env_begin, n, _ptr(inputs), 1 if broadcast else 0, float(dt)))

    def reset_contact_cables(self, rest_prev=None, env_begin=0, env_count=None):
        """a new episode for the contact cables: rest / previous lengths ([n, n_contact_cables, 2], default the model's), no anchors"""
        n = self.n_env - env_begin if env_count is None else env_count
        if rest_prev is not None:
            rest_prev = _host_buffer(rest_prev, (n, self.model.n_contact_cables, 2), "rest_prev")
        capi.check(self._lib.tgb_reset_contact_cables(self._h, env_begin, n, _ptr(rest_prev) if rest_prev is not None else None))

    def get_contact_cable_anchors(self, cable, env_begin=0, env_count=None):
        """[n, 12, 7]: body (-1 = unused slot), world position, contact normal of the cable's sliding anchors in order"""
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.zeros((n, capi.CCABLE_MAX_SLIDING, capi.CCABLE_ANCHOR_DOUBLES))
        capi.check(self._lib.tgb_get_contact_cable_anchors(self._h, int(cable), env_begin, n, _ptr(out)))
        return out

    def set_cable_state(self, rest_prev, env_begin=0):
        if self.C == 0:
            return  # a structure without cables has no cable state
        rest_prev = np.ascontiguousarray(rest_prev, dtype=np.float64).reshape(-1, self.C, 2)
        capi.check(self._lib.tgb_set_cable_state(self._h, env_begin, rest_prev.shape[0], _ptr(rest_prev)))

    def get_cable_obs(self, env_begin=0, env_count=None):
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.empty((n, self.C, capi.CABLE_OBS_DOUBLES))
        capi.check(self._lib.tgb_get_cable_obs(self._h, env_begin, n, _ptr(out)))
        return out

    # ---- control -----------------------------------------------------------------------------------------
    def set_controller(self, kind, params=None, env_begin=0, env_count=None, broadcast=False):
        n = self.n_env - env_begin if env_count is None else env_count
        if params is None:
            capi.check(self._lib.tgb_set_controller(self._h, kind, None, 0, env_begin, n, 0))
            return
        params = _rows(params, n, broadcast, "set_controller")
        per_env = params.shape[-1]
        capi.check(self._lib.tgb_set_controller(self._h, kind, _ptr(params), per_env, env_begin, n, 1 if broadcast else 0))

    def set_cpg_controller(self, desc, params, env_begin=0, env_count=None, broadcast=False):
        """desc: a CpgController (ntrtsim_b200.examples.tetraspine_cpg); params: [n, 2 + 2 * n_pairs] rows"""
        n = self.n_env - env_begin if env_count is None else env_count
        params = _rows(params, n, broadcast, "set_cpg_controller")
        capi.check(self._lib.tgb_set_cpg_controller(self._h, C.byref(desc.desc), _ptr(params), params.shape[-1], env_begin, n,
                                                    1 if broadcast else 0))

    def track_max_tension(self, enable=True):
        """hist = true for every cable: keep the running maximum of the tension history (BaseSpineModelLearning::getStringMaxTensions)"""
        capi.check(self._lib.tgb_track_max_tension(self._h, 1 if enable else 0))

    def max_tension(self, env_begin=0, env_count=None):
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.zeros((n, self.C))
        capi.check(self._lib.tgb_get_max_tension(self._h, env_begin, n, _ptr(out)))
        return out

    def get_cpg_state(self, env_begin=0, env_count=None):
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.empty((n, self.C, 3))
        capi.check(self._lib.tgb_get_cpg_state(self._h, env_begin, n, _ptr(out)))
        return out

    def set_control_input(self, inputs, dt, env_begin=0, env_count=None):
        n = self.n_env - env_begin if env_count is None else env_count
        inputs = np.ascontiguousarray(inputs, dtype=np.float64)
        broadcast = 1 if inputs.ndim == 1 else 0
        if inputs.size != (1 if broadcast else n) * self.C:
            raise ValueError("set_control_input: %d values for %d environments x %d cables" % (inputs.size, 1 if broadcast else n, self.C))
        capi.check(self._lib.tgb_set_control_input(self._h, env_begin, n, _ptr(inputs), broadcast, dt))

    # ---- stepping ----------------------------------------------------------------------------------------
    def step(self, dt, n_steps=1):
        capi.check(self._lib.tgb_step(self._h, dt, n_steps))

    def step_world(self, dt):
        capi.check(self._lib.tgb_step_world(self._h, dt))

    def step_models(self, dt):
        capi.check(self._lib.tgb_step_models(self._h, dt))

    def sync(self):
        capi.check(self._lib.tgb_sync(self._h))

    def step_timed(self, dt, n_steps, reps):
        ms = (C.c_float * reps)()
        capi.check(self._lib.tgb_step_timed(self._h, dt, n_steps, reps, ms))
        return [float(x) for x in ms]

    def rollout_host(self, rod_in, cable_in, dt, n_steps, rod_out=None, cable_out=None):
        """Host buffers in, host buffers out (the reference-facing end-to-end call)."""
        if rod_out is None:
            rod_out = np.empty((self.n_env, self.R, capi.ROD_STATE_DOUBLES))
        if cable_out is None:
            cable_out = np.empty((self.n_env, self.C, capi.CABLE_OBS_DOUBLES))
        if rod_in is not None:
            rod_in = _host_buffer(rod_in, (self.n_env, self.R, capi.ROD_STATE_DOUBLES), "rod_in")
        if cable_in is not None:
            cable_in = _host_buffer(cable_in, (self.n_env, self.C, 2), "cable_in")
        rod_out = _host_buffer(rod_out, (self.n_env, self.R, capi.ROD_STATE_DOUBLES), "rod_out", writable=True)
        cable_out = _host_buffer(cable_out, (self.n_env, self.C, capi.CABLE_OBS_DOUBLES), "cable_out", writable=True)
        capi.check(self._lib.tgb_rollout_host(self._h, _ptr(rod_in) if rod_in is not None else None,
                                              _ptr(cable_in) if cable_in is not None else None, dt, n_steps,
                                              _ptr(rod_out), _ptr(cable_out)))
        return rod_out, cable_out

    # ---- bookkeeping ---------------------------------------------------------------------------------------
    def time(self):
        t = C.c_double()
        capi.check(self._lib.tgb_get_time(self._h, C.byref(t)))
        return t.value

    def status(self, env_begin=0, env_count=None):
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.empty(n, dtype=np.int32)
        capi.check(self._lib.tgb_get_status(self._h, env_begin, n, _ptr(out)))
        return out

    def min_clearance(self, env_begin=0, env_count=None, reset=False):
        """smallest surface distance between two bodies' collision shapes seen by any step so far, per environment"""
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.empty(n)
        capi.check(self._lib.tgb_get_min_clearance(self._h, env_begin, n, _ptr(out), 1 if reset else 0))
        return out

    def set_height_window(self, lo, hi):
        capi.check(self._lib.tgb_set_height_window(self._h, lo, hi))

    def mark_episode_start(self):
        capi.check(self._lib.tgb_mark_episode_start(self._h))

    def scores(self, env_begin=0, env_count=None):
        n = self.n_env - env_begin if env_count is None else env_count
        out = np.empty((n, 2))
        capi.check(self._lib.tgb_get_scores(self._h, env_begin, n, _ptr(out)))
        return out

    def scores_device_ptr(self):
        p, nbytes = C.c_void_p(), C.c_size_t()
        capi.check(self._lib.tgb_scores_device_ptr(self._h, C.byref(p), C.byref(nbytes)))
        return p.value, nbytes.value

    def scores_refresh_async(self):
        """scores_device_ptr without the synchronisation: the refresh is queued on the world's stream"""
        p, nbytes = C.c_void_p(), C.c_size_t()
        capi.check(self._lib.tgb_scores_refresh_async(self._h, C.byref(p), C.byref(nbytes)))
        return p.value, nbytes.value

    def stream_handle(self):
        """the CUDA stream (cudaStream_t as an integer) all work of this world is issued on"""
        p = C.c_void_p()
        capi.check(self._lib.tgb_stream_handle(self._h, C.byref(p)))
        return p.value or 0

    def launch_count(self):
        n = C.c_int64()
        capi.check(self._lib.tgb_launch_count(self._h, C.byref(n)))
        return n.value

    def launch_geometry(self):
        a, b, g = C.c_int32(), C.c_int32(), C.c_int32()
        capi.check(self._lib.tgb_launch_geometry(self._h, C.byref(a), C.byref(b), C.byref(g)))
        return a.value, b.value, g.value

    def state_bytes_per_env(self):
        n = C.c_size_t()
        capi.check(self._lib.tgb_state_bytes_per_env(self._h, C.byref(n)))
        return n.value
