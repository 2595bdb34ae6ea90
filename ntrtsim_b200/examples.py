"""The shipped example robots of the hot-path configs, built on the host through the C++ modelling API of this
repo (tgStructure / tgBuildSpec / tgStructureInfo in ntrtsim_b200/host, mirroring src/tgcreator of the reference)
and returned as flat descriptors ready for BatchedWorld."""
import ctypes as C

import numpy as np

from . import _capi as capi
from .world import FlatModel

EXAMPLES = ("3_prism", "3_prism_noground", "SUPERball", "SUPERball_flat", "Escape", "Escape_c3", "tsTestRig_basic",
            "tsTestRig_kinematic", "TetraSpine", "FlemonsSpine", "FlemonsSpine3", "TetraSpineStatic_hf")


def build_example(name):
    h = capi.load_host()
    m = h.tgh_build_example(name.encode())
    if not m:
        raise ValueError("cannot build example %r: %s" % (name, h.tgh_last_error().decode()))
    try:
        nb, ns, nc = C.c_int(), C.c_int(), C.c_int()
        h.tgh_counts(m, C.byref(nb), C.byref(ns), C.byref(nc))
        bodies = (capi.BodyDesc * nb.value)()
        shapes = (capi.ShapeDesc * max(1, ns.value))()
        cables = (capi.CableDesc * max(1, nc.value))()
        ground = capi.GroundDesc()
        g, dt = C.c_double(), C.c_double()
        h.tgh_copy(m, bodies, shapes, cables, C.byref(ground), C.byref(g), C.byref(dt))
        shapes = (capi.ShapeDesc * ns.value).from_buffer_copy(bytes(shapes)[: C.sizeof(capi.ShapeDesc) * ns.value])
        cables = (capi.CableDesc * nc.value).from_buffer_copy(bytes(cables)[: C.sizeof(capi.CableDesc) * nc.value])
        return FlatModel(bodies, shapes, cables, ground, g.value, dt.value, name)
    finally:
        h.tgh_free(m)


def tree_counts(name):
    """(number of tgRod, number of tgSpringCableActuator) descendants of the built tgModel tree."""
    h = capi.load_host()
    m = h.tgh_build_example(name.encode())
    if not m:
        raise ValueError(h.tgh_last_error().decode())
    try:
        r, a = C.c_int(), C.c_int()
        h.tgh_tree_counts(m, C.byref(r), C.byref(a))
        return r.value, a.value
    finally:
        h.tgh_free(m)


def escape_transform_actions(actions01, initial_length=9.0):
    """EscapeController::transformActions + applyActions (src/examples/craterEscape/EscapeController.cpp:142-194):
    [..., 32] numbers in [0,1] -> [..., 33] rows for TGB_CTRL_ESCAPE_SINE."""
    h = capi.load_host()
    a = np.ascontiguousarray(actions01, dtype=np.float64).reshape(-1, 32)
    out = np.empty((a.shape[0], capi.TGB_ESCAPE_PARAMS))
    for i in range(a.shape[0]):
        h.tgh_escape_transform_actions(a[i].ctypes.data_as(C.c_void_p), initial_length, out[i].ctypes.data_as(C.c_void_p))
    return out.reshape(tuple(np.shape(actions01)[:-1]) + (capi.TGB_ESCAPE_PARAMS,))


class CpgController(object):
    """Static part of a CPG + impedance controller (a ctypes tgb_cpg_desc plus the arrays it points to)."""

    def __init__(self, begin, target, pair, n_pairs, impedance, r_const, d_min, d_max, descending_command, control_time,
                 cable_control_step, max_steps, score_body, height_lo, height_hi, cable_control=0, pid=(0.0, 0.0, 0.0)):
        self.begin = np.ascontiguousarray(begin, dtype=np.int32)
        self.target = np.ascontiguousarray(target, dtype=np.int32)
        self.pair = np.ascontiguousarray(pair, dtype=np.int32)
        self.impedance = np.ascontiguousarray(impedance, dtype=np.float64)
        i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        self.n_nodes, self.n_pairs = len(self.begin) - 1, int(n_pairs)
        self.desc = capi.CpgDesc(self.n_nodes, len(self.target), self.begin.ctypes.data_as(i32p), self.target.ctypes.data_as(i32p),
                                 self.pair.ctypes.data_as(i32p), self.n_pairs, int(max_steps), self.impedance.ctypes.data_as(dp), r_const, d_min,
                                 d_max, descending_command, control_time, cable_control_step, int(score_body), int(cable_control), height_lo, height_hi,
                                 float(pid[0]), float(pid[1]), float(pid[2]))

    @property
    def params_per_env(self):
        return 2 + 2 * self.n_pairs


def tetraspine_cpg():
    """The controller of AppTetraSpineLearning (BaseSpineCPGControl::Config(3, 6, 6, 2, segment 1, controlTime .001) with
    TetraSpineCPGControl's impedance gains), flattened by the C++ host code for tgb_set_cpg_controller."""
    h = capi.load_host()
    m = h.tgh_build_example(b"TetraSpine")
    if not m:
        raise ValueError(h.tgh_last_error().decode())
    try:
        nn, nc, npairs = C.c_int(), C.c_int(), C.c_int()
        h.tgh_tetraspine_cpg(m, C.byref(nn), C.byref(nc), C.byref(npairs), None, None, None, None, None, None)
        begin = np.zeros(nn.value + 1, dtype=np.int32)
        target = np.zeros(nc.value, dtype=np.int32)
        pair = np.zeros(nc.value, dtype=np.int32)
        imp = np.zeros((nn.value, 4))
        sc = np.zeros(8)
        it = np.zeros(2, dtype=np.int32)
        vp = C.c_void_p
        rc = h.tgh_tetraspine_cpg(m, C.byref(nn), C.byref(nc), C.byref(npairs), begin.ctypes.data_as(vp), target.ctypes.data_as(vp),
                                  pair.ctypes.data_as(vp), imp.ctypes.data_as(vp), sc.ctypes.data_as(vp), it.ctypes.data_as(vp))
        if rc != 0:
            raise ValueError(h.tgh_last_error().decode())
        return CpgController(begin, target, pair, npairs.value, imp, sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], int(it[0]), int(it[1]), sc[6], sc[7])
    finally:
        h.tgh_free(m)


def flemons_cpg(name="FlemonsSpine"):
    """The controller of the kinematic tetrahedral-complex spine (KinematicSpineCPGControl with the Config of
    test_integration/SpineTests/WorldConf_Spines_test.cpp:110-147; tgCPGCableControl + tgPIDController(5000, 0, 10, tension control) on
    every tgKinematicActuator), flattened by the C++ host code for tgb_set_cpg_controller (cable_control = 1)."""
    h = capi.load_host()
    m = h.tgh_build_example(name.encode())
    if not m:
        raise ValueError(h.tgh_last_error().decode())
    try:
        nn, nc, npairs = C.c_int(), C.c_int(), C.c_int()
        if h.tgh_flemons_cpg(m, C.byref(nn), C.byref(nc), C.byref(npairs), None, None, None, None, None, None) != 0:
            raise ValueError(h.tgh_last_error().decode())
        begin = np.zeros(nn.value + 1, dtype=np.int32)
        target = np.zeros(nc.value, dtype=np.int32)
        pair = np.zeros(nc.value, dtype=np.int32)
        imp = np.zeros((nn.value, 4))
        sc = np.zeros(11)
        it = np.zeros(2, dtype=np.int32)
        vp = C.c_void_p
        rc = h.tgh_flemons_cpg(m, C.byref(nn), C.byref(nc), C.byref(npairs), begin.ctypes.data_as(vp), target.ctypes.data_as(vp),
                               pair.ctypes.data_as(vp), imp.ctypes.data_as(vp), sc.ctypes.data_as(vp), it.ctypes.data_as(vp))
        if rc != 0:
            raise ValueError(h.tgh_last_error().decode())
        return CpgController(begin, target, pair, npairs.value, imp, sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], int(it[0]), int(it[1]), sc[6], sc[7],
                             cable_control=1, pid=(sc[8], sc[9], sc[10]))
    finally:
        h.tgh_free(m)


def flemons_param_rows(node_actions, edge_actions, name="FlemonsSpine"):
    """scaleNodeActions + scaleEdgeActions for the kinematic tetrahedral-complex spine: node_actions [P, 2], edge_actions [P, 100, 2] in
    [0, 1] (resources/src/learningSpines/TetrahedralComplex/logs/bestParameters-default_*.nnw) -> rows [P, 2 + 2 * 192]."""
    h = capi.load_host()
    m = h.tgh_build_example(name.encode())
    if not m:
        raise ValueError(h.tgh_last_error().decode())
    try:
        na = np.ascontiguousarray(node_actions, dtype=np.float64).reshape(-1, 2)
        ea = np.ascontiguousarray(edge_actions, dtype=np.float64).reshape(na.shape[0], -1, 2)
        rows = np.zeros((na.shape[0], 2 + 2 * 192))
        vp = C.c_void_p
        for i in range(na.shape[0]):
            used = h.tgh_flemons_param_row(m, na[i].ctypes.data_as(vp), ea[i].ctypes.data_as(vp), ea.shape[1], rows[i].ctypes.data_as(vp))
            if used != ea.shape[1]:
                raise ValueError("scaleEdgeActions consumed %d of %d edge actions" % (used, ea.shape[1]))
        return rows
    finally:
        h.tgh_free(m)


def tetraspine_param_rows(node_actions, edge_actions):
    """scaleNodeActions + scaleEdgeActions (BaseSpineCPGControl.cpp:363-468): node_actions [P, 2], edge_actions [P, 57, 2], both in
    [0, 1] (the contents of the bestParameters-*_node/edge-*.nnw files) -> controller rows [P, 2 + 2 * 108]."""
    h = capi.load_host()
    na = np.ascontiguousarray(node_actions, dtype=np.float64).reshape(-1, 2)
    ea = np.ascontiguousarray(edge_actions, dtype=np.float64).reshape(na.shape[0], -1, 2)
    rows = np.zeros((na.shape[0], 2 + 2 * 108))
    vp = C.c_void_p
    for i in range(na.shape[0]):
        used = h.tgh_tetraspine_param_row(na[i].ctypes.data_as(vp), ea[i].ctypes.data_as(vp), ea.shape[1], rows[i].ctypes.data_as(vp))
        if used != ea.shape[1]:
            raise ValueError("scaleEdgeActions consumed %d of %d edge actions" % (used, ea.shape[1]))
    return rows
