"""The shipped example robots of the hot-path configs, built on the host through the C++ modelling API of this
repo (tgStructure / tgBuildSpec / tgStructureInfo in ntrtsim_b200/host, mirroring src/tgcreator of the reference)
and returned as flat descriptors ready for BatchedWorld."""
import ctypes as C

import numpy as np

from . import _capi as capi
from .world import FlatModel

EXAMPLES = ("3_prism", "3_prism_noground", "SUPERball", "SUPERball_flat", "Escape", "Escape_c3", "tsTestRig_basic",
            "tsTestRig_kinematic")


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
