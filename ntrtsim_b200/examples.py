"""The shipped example robots of the hot-path configs, built on the host through the C++ modelling API of this
repo (tgStructure / tgBuildSpec / tgStructureInfo in ntrtsim_b200/host, mirroring src/tgcreator of the reference)
and returned as flat descriptors ready for BatchedWorld."""
import ctypes as C

import numpy as np

from . import _capi as capi
from .world import FlatModel

EXAMPLES = ("3_prism", "3_prism_noground", "SUPERball", "SUPERball_flat", "Escape", "Escape_c3", "tsTestRig_basic",
            "tsTestRig_kinematic", "TetraSpine", "FlemonsSpine", "FlemonsSpine3", "TetraSpineStatic_hf", "Escape_crater", "Escape_crater_shallow", "SUPERball_stairs", "SUPERball_blocks", "SUPERball_wall", "ContactCableDemo")


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
        ncc = h.tgh_contact_cables(m, None, 0)
        cc = None
        if ncc > 0:
            cc = (capi.ContactCableDesc * ncc)()
            h.tgh_contact_cables(m, cc, ncc)
        return FlatModel(bodies, shapes, cables, ground, g.value, dt.value, name, contact_cables=cc)
    finally:
        h.tgh_free(m)


def build_yaml(path, gravity=98.1, dt=0.001):
    """A structure file of the YAML front-end (src/yamlbuilder/TensegrityModel.cpp) built by the C++ host API in the world of the
    reference's yamlbuilder app (BuildTensegrityModel.cpp:44-79: gravity 98.1, default box ground, dt 1e-3) -> FlatModel."""
    return build_example("yaml:%s@%r@%r" % (path, float(gravity), float(dt)))


def yaml_to_json(text):
    """The host YAML reader's parse of `text` as a JSON string (scalars as strings); ValueError on malformed input."""
    h = capi.load_host()
    n = h.tgh_yaml_to_json(text.encode(), None, 0)
    if n < 0:
        raise ValueError(h.tgh_last_error().decode())
    buf = C.create_string_buffer(n + 1)
    h.tgh_yaml_to_json(text.encode(), buf, n + 1)
    return buf.value.decode()


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


# resources/src/ICRA2015/static/controlVars.json of the reference, the file SerializedSpineControl::Config reads
# (SerializedSpineControl.cpp:51-163); values only, keyed like the file
ICRA2015_STATIC_CONTROL_VARS = {
    "inside_imp_ten": 0.0, "inside_imp_pos": 100.0, "inside_imp_vel": 100.0,
    "outside_imp_ten": 1000.0, "outside_imp_pos": 100.0, "outside_imp_vel": 100.0,
    "top_imp_ten": 1200.0, "top_imp_pos": 500.0, "top_imp_vel": 150.0,
    "in_top_tens": (2800.0, 1000.0), "in_left_tens": (1500.0, 1000.0), "in_right_tens": (1800.0, 800.0),
    "out_top_tens": (1500.0, 600.0), "out_left_tens": (1100.0, 800.0), "out_right_tens": (1000.0, 400.0),
    "in_top_length": (19.0, 18.0), "in_left_length": (21.5, 22.5), "in_right_length": (21.5, 22.5),
    "out_top_length": (18.5, 19.8), "out_left_length": (19.2, 18.8), "out_right_length": (19.3, 19.2),
    "offset_speed": 0.0, "cpg_amplitude": 0.0, "cpg_frequency": 2.51, "bodyWaves": 0.5,
    "top_offset": 0.0, "left_offset": 0.0, "right_offset": 0.0, "updateFrequency": 20.0,
}


def serialized_spine_params(model, cfg=None):
    """The TGB_CTRL_SERIALIZED_SPINE parameter row of SerializedSpineControl on TetraSpineStaticModel_hf
    (src/examples/IROS_2015/TetraSpineStatic/SerializedSpineControl.cpp:174-205,262-371): per-cable impedance gains / offset tension /
    control length by muscle group (cables are created per segment pair in the order outer right, outer left, outer top, inner right,
    inner left, inner top: TetraSpineStaticModel_hf.cpp:123-142), and the 2.3 kg load pulled on the front tip between simTime 30 and
    45 with 0.45 / 0.55 of it held back at the other two tips (markers 1, 2 and 5: TetraSpineStaticModel_hf.cpp:160-205)."""
    import math
    c = dict(ICRA2015_STATIC_CONTROL_VARS if cfg is None else cfg)
    C, R = model.n_cables, model.n_bodies
    segments = R
    row = np.zeros(capi.TGB_SPINE_HEADER + 5 * C)
    row[0:4] = [c["offset_speed"], c["cpg_amplitude"], c["cpg_frequency"], c["updateFrequency"]]
    row[4:8] = [30.0, 35.0, 40.0, 45.0]
    row[8:11] = [0.0, 0.0, 2.3 * 981.0]
    row[11] = 3
    # the tip node of segment s: (0, edge / 2 * tan(pi / 6), 25) moved by (0, 2, 100) + (s + 1) * (0, 0, -23)
    edge = 38.0
    for k, (body, factor) in enumerate(((0, 1.0), (1, -1.0 / 2.0 * 0.9), (2, -1.0 / 2.0 * 1.1))):
        tip = np.array([0.0, edge / 2.0 * math.tan(math.pi / 6.0) + 2.0, 25.0 + 100.0 - 23.0 * (body + 1)])
        local = tip - np.array(model.bodies[body].pos)          # compound bodies start with identity rotation
        row[12 + 5 * k: 12 + 5 * k + 5] = [body, local[0], local[1], local[2], factor]
    groups = ["out_right", "out_left", "out_top", "in_right", "in_left", "in_top"]
    gains = {"out_right": ("outside", 2), "out_left": ("outside", 1), "out_top": ("top", 0),
             "in_right": ("inside", 2), "in_left": ("inside", 1), "in_top": ("top", 0)}
    offsets = [c["top_offset"], c["left_offset"], c["right_offset"]]
    for cable in range(C):
        pair, g = divmod(cable, 6)          # pair = index of the string inside its group (stringList[i])
        name = groups[g]
        ctl, phase = gains[name]
        row[capi.TGB_SPINE_HEADER + 5 * cable: capi.TGB_SPINE_HEADER + 5 * cable + 5] = [
            c[name + "_tens"][pair], c[ctl + "_imp_pos"], c[ctl + "_imp_vel"], c[name + "_length"][pair],
            2 * c["bodyWaves"] * math.pi * pair / segments + offsets[phase]]
    return row
