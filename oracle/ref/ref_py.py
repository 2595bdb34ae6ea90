"""ctypes binding of oracle/_ref/libntrt_ref.so: the reference's own cable / actuator / controller / CPG sources compiled
in place (oracle/ref/build_ref.py). TEST INFRASTRUCTURE; exists only in the container that has /root/reference."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("NTRT_REF_LIB", os.path.join(os.path.dirname(HERE), "_ref", "libntrt_ref.so"))
_lib = None


def available():
    return os.path.exists(LIB)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB)
        vp, d, i = C.c_void_p, C.c_double, C.c_int
        lib.ref_create.restype = vp
        lib.ref_destroy.argtypes = [vp]
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_add_body.argtypes = [vp, vp, d, vp]
        lib.ref_set_body.argtypes = [vp, i, vp]
        lib.ref_get_body.argtypes = [vp, i, vp]
        lib.ref_add_cable.argtypes = [vp, i, i, vp, vp, i, vp]
        lib.ref_get_cable.argtypes = [vp, i, vp]
        lib.ref_set_control_input_dt.argtypes = [vp, i, d, d]
        lib.ref_set_control_input.argtypes = [vp, i, d]
        lib.ref_move_motors.argtypes = [vp, i, d]
        lib.ref_tension_control.argtypes = [vp, i, d, d]
        lib.ref_impedance_control.argtypes = [vp, i, d, d, d, d, d, d]
        lib.ref_impedance_control.restype = d
        lib.ref_step_actuators.argtypes = [vp, d]
        lib.ref_pid_create.argtypes = [vp, i, d, d, d, i]
        lib.ref_impedance_control_pid.argtypes = [vp, i, d, d, d, d, d, d]
        lib.ref_impedance_control_pid.restype = d
        lib.ref_pid_control.argtypes = [vp, i, d, d]
        lib.ref_cpg_create.restype = vp
        lib.ref_cpg_create.argtypes = [i]
        lib.ref_cpg_destroy.argtypes = [vp]
        lib.ref_cpg_add_node.argtypes = [vp, vp]
        lib.ref_cpg_define_connections.argtypes = [vp, i, i, vp, vp, vp]
        lib.ref_cpg_update.argtypes = [vp, vp, i, d]
        lib.ref_cpg_value.restype = d
        lib.ref_cpg_value.argtypes = [vp, i]
        lib.ref_cpg_state.argtypes = [vp, vp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(x):
    return np.ascontiguousarray(np.array(list(x), dtype=np.float64))


class RefModels(object):
    """The models half of tgSimulation::step (controllers + every actuator's step) run by reference code on shim bodies.
    Built from the same flat descriptors as the device world; `pretension` is the Config value of the example (the
    reference constructor recomputes the rest length from it: tgSpringCable.cpp:49-59)."""

    def __init__(self, model, pretension, rod_state=None):
        from ntrtsim_b200 import _capi as capi
        self.lib = load()
        self.model = model
        self.h = C.c_void_p(self.lib.ref_create())
        st = model.initial_rod_state() if rod_state is None else np.ascontiguousarray(rod_state, dtype=np.float64)
        for b, body in enumerate(model.bodies):
            self.lib.ref_add_body(self.h, _p(np.ascontiguousarray(st[b])), body.mass, _p(_arr(body.inertia)))
        # anchors are constructed from WORLD points: body transform * local point, like tgBasicActuatorInfo hands them over
        from helpers_quat import quat_rotate
        for c in model.cables:
            wa = quat_rotate(st[c.body_a, 3:7], _arr(c.local_a)) + st[c.body_a, 0:3]
            wb = quat_rotate(st[c.body_b, 3:7], _arr(c.local_b)) + st[c.body_b, 0:3]
            cfg = _arr([c.stiffness, c.damping, pretension, c.max_tension, c.target_velocity, c.min_actual_length, c.min_rest_length,
                        c.motor_radius, c.motor_friction, c.motor_inertia, float(c.backdrivable)])
            kind = 2 if c.actuator == capi.TGB_ACT_KINEMATIC else 1
            idx = self.lib.ref_add_cable(self.h, c.body_a, c.body_b, _p(_arr(wa)), _p(_arr(wb)), kind, _p(cfg))
            if idx < 0:
                raise ValueError(self.lib.ref_last_error().decode())
        self.R, self.C = len(model.bodies), len(model.cables)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def set_bodies(self, st):
        st = np.ascontiguousarray(st, dtype=np.float64)
        for b in range(self.R):
            self.lib.ref_set_body(self.h, b, _p(st[b]))

    def get_bodies(self):
        out = np.empty((self.R, 13))
        for b in range(self.R):
            self.lib.ref_get_body(self.h, b, _p(out[b]))
        return out

    def cables(self):
        """[C, 8]: actuator rest, cable rest, length, tension, cable velocity, actuator velocity, start length, damping"""
        out = np.empty((self.C, 8))
        for c in range(self.C):
            self.lib.ref_get_cable(self.h, c, _p(out[c]))
        return out

    def set_control_input_dt(self, c, value, dt):
        return self.lib.ref_set_control_input_dt(self.h, c, value, dt)

    def set_control_input(self, c, value):
        return self.lib.ref_set_control_input(self.h, c, value)

    def move_motors(self, c, dt):
        self.lib.ref_move_motors(self.h, c, dt)

    def tension_control(self, c, dt, set_point):
        return self.lib.ref_tension_control(self.h, c, dt, set_point)

    def impedance_control(self, c, dt, T0, Kp, Kv, new_position, offset_vel):
        return self.lib.ref_impedance_control(self.h, c, dt, T0, Kp, Kv, new_position, offset_vel)

    def pid_create(self, c, p, i, d, tension_control=True):
        return self.lib.ref_pid_create(self.h, c, p, i, d, 1 if tension_control else 0)

    def impedance_control_pid(self, c, dt, T0, Kp, Kv, new_position, offset_vel):
        return self.lib.ref_impedance_control_pid(self.h, c, dt, T0, Kp, Kv, new_position, offset_vel)

    def pid_control(self, c, dt, set_point):
        self.lib.ref_pid_control(self.h, c, dt, set_point)

    def step_actuators(self, dt):
        return self.lib.ref_step_actuators(self.h, dt)


class RefCPG(object):
    def __init__(self, max_steps=200):
        self.lib = load()
        self.h = C.c_void_p(self.lib.ref_cpg_create(max_steps))
        self.n = 0

    def add_node(self, params7):
        self.n += 1
        return self.lib.ref_cpg_add_node(self.h, _p(_arr(params7)))

    def define_connections(self, node, conn, weights, phases):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.lib.ref_cpg_define_connections(self.h, node, len(conn), _p(conn), _p(_arr(weights)), _p(_arr(phases)))

    def update(self, desc_com, dt):
        return self.lib.ref_cpg_update(self.h, _p(_arr(desc_com)), self.n, dt)

    def values(self):
        return np.array([self.lib.ref_cpg_value(self.h, i) for i in range(self.n)])

    def state(self):
        out = np.empty(3 * self.n)
        self.lib.ref_cpg_state(self.h, _p(out))
        return out

    def close(self):
        if self.h:
            self.lib.ref_cpg_destroy(self.h)
            self.h = None
