"""ctypes binding of the C ABI declared in include/ntrtsim_b200.h (libntrtsim_b200.so) and of the host-side
model builder (libtgb_host.so).  The libraries are built in-tree by `make` / ntrtsim_b200.build.build_all();
loading fails loudly when they are missing -- there is no Python or CPU fallback for the step."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("TGB_LIB_DIR") or os.path.join(_HERE, "lib")  # TGB_LIB_DIR: kernel-variant A/B runs (tools/)

TGB_OK, TGB_ERR_INVALID_ARGUMENT, TGB_ERR_STATE, TGB_ERR_CUDA, TGB_ERR_UNSUPPORTED, TGB_ERR_RUNTIME = range(6)
TGB_F32, TGB_F64 = 0, 1
TGB_GROUND_EMPTY, TGB_GROUND_BOX, TGB_GROUND_PLANE = 0, 1, 2
TGB_ACT_PASSIVE, TGB_ACT_BASIC, TGB_ACT_KINEMATIC = 0, 1, 2
TGB_CTRL_NONE, TGB_CTRL_SETPOINT, TGB_CTRL_TENSION, TGB_CTRL_ESCAPE_SINE, TGB_CTRL_TORQUE, TGB_CTRL_CPG_IMPEDANCE, TGB_CTRL_SERIALIZED_SPINE = range(7)
TGB_SPINE_HEADER = 27
TGB_ESCAPE_PARAMS = 33
ROD_STATE_DOUBLES = 13
CABLE_OBS_DOUBLES = 4


class WorldCfg(C.Structure):
    _fields_ = [("gravity", C.c_double), ("n_env", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
                ("solver_iterations", C.c_int32)]


class GroundDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("normal", C.c_double * 3), ("offset", C.c_double), ("friction", C.c_double),
                ("rolling_friction", C.c_double), ("restitution", C.c_double), ("margin", C.c_double)]


class BodyDesc(C.Structure):
    _fields_ = [("mass", C.c_double), ("inertia", C.c_double * 3), ("pos", C.c_double * 3), ("quat", C.c_double * 4),
                ("friction", C.c_double), ("rolling_friction", C.c_double), ("restitution", C.c_double),
                ("shape_begin", C.c_int32), ("shape_count", C.c_int32), ("gyroscopic", C.c_int32), ("reserved", C.c_int32)]


class ShapeDesc(C.Structure):
    _fields_ = [("radius", C.c_double), ("half_length", C.c_double), ("margin", C.c_double), ("pos", C.c_double * 3),
                ("quat", C.c_double * 4), ("kind", C.c_int32), ("reserved", C.c_int32), ("half_extent_z", C.c_double)]


class CableDesc(C.Structure):
    _fields_ = [("body_a", C.c_int32), ("body_b", C.c_int32), ("local_a", C.c_double * 3), ("local_b", C.c_double * 3),
                ("stiffness", C.c_double), ("damping", C.c_double), ("rest_length", C.c_double), ("prev_length", C.c_double),
                ("actuator", C.c_int32), ("backdrivable", C.c_int32), ("max_tension", C.c_double),
                ("target_velocity", C.c_double), ("min_actual_length", C.c_double), ("min_rest_length", C.c_double),
                ("motor_radius", C.c_double), ("motor_friction", C.c_double), ("motor_inertia", C.c_double)]


class ContactCableDesc(C.Structure):
    _fields_ = [("body_a", C.c_int32), ("body_b", C.c_int32), ("local_a", C.c_double * 3), ("local_b", C.c_double * 3),
                ("stiffness", C.c_double), ("damping", C.c_double), ("rest_length", C.c_double), ("prev_length", C.c_double),
                ("thickness", C.c_double), ("resolution", C.c_double), ("max_tension", C.c_double), ("target_velocity", C.c_double),
                ("min_actual_length", C.c_double), ("min_rest_length", C.c_double)]


CCABLE_MAX_SLIDING = 12
CCABLE_OBS_DOUBLES = 5
CCABLE_ANCHOR_DOUBLES = 7


class CpgDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_couplings", C.c_int32), ("coupling_begin", C.POINTER(C.c_int32)),
                ("coupling_target", C.POINTER(C.c_int32)), ("coupling_pair", C.POINTER(C.c_int32)), ("n_pairs", C.c_int32),
                ("max_steps", C.c_int32), ("impedance", C.POINTER(C.c_double)), ("r_const", C.c_double), ("d_min", C.c_double),
                ("d_max", C.c_double), ("descending_command", C.c_double), ("control_time", C.c_double),
                ("cable_control_step", C.c_double), ("score_body", C.c_int32), ("cable_control", C.c_int32), ("height_lo", C.c_double),
                ("height_hi", C.c_double), ("pid_kp", C.c_double), ("pid_ki", C.c_double), ("pid_kd", C.c_double)]


class TgbError(RuntimeError):
    def __init__(self, status, message):
        RuntimeError.__init__(self, "tgb status %d: %s" % (status, message))
        self.status = status


# every symbol include/ntrtsim_b200.h declares; tests check that the library exports all of them
ABI_SYMBOLS = [
    "tgb_abi_version", "tgb_last_error", "tgb_device_count", "tgb_create", "tgb_destroy", "tgb_add_model_flat",
    "tgb_finalize", "tgb_set_cpg_controller", "tgb_get_cpg_state", "tgb_num_bodies", "tgb_num_cables", "tgb_num_envs", "tgb_set_rod_state", "tgb_get_rod_state",
    "tgb_set_cable_state", "tgb_get_cable_obs", "tgb_set_controller", "tgb_set_control_input", "tgb_step", "tgb_step_world", "tgb_step_models", "tgb_sync",
    "tgb_rollout_host", "tgb_track_max_tension", "tgb_get_max_tension", "tgb_get_time", "tgb_get_status", "tgb_mark_episode_start", "tgb_get_scores",
    "tgb_scores_device_ptr", "tgb_scores_refresh_async", "tgb_step_timed", "tgb_launch_count", "tgb_launch_geometry", "tgb_set_lanes_per_env",
    "tgb_set_block_threads", "tgb_set_wide_layout", "tgb_state_bytes_per_env", "tgb_set_height_window", "tgb_stream_handle",
    "tgb_set_pair_contact_cap", "tgb_get_min_clearance", "tgb_set_warm_starting", "tgb_set_tail_split",
    "tgb_add_contact_cables", "tgb_num_contact_cables", "tgb_get_contact_cable_obs", "tgb_get_contact_cable_anchors",
    "tgb_set_contact_cable_control_input", "tgb_reset_contact_cables",
]

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i32 = C.c_int32

_lib = None
_host = None


def lib_path():
    return os.path.join(LIB_DIR, "libntrtsim_b200.so")


def host_lib_path():
    return os.path.join(LIB_DIR, "libtgb_host.so")


def load():
    """Load libntrtsim_b200.so (the CUDA library). Raises OSError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise OSError("%s is missing: build it with `make` or ntrtsim_b200.build.build_all(); there is no CPU path" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.tgb_abi_version.restype = C.c_int
    lib.tgb_last_error.restype = C.c_char_p
    lib.tgb_device_count.restype = C.c_int
    lib.tgb_create.argtypes = [C.POINTER(WorldCfg), C.POINTER(GroundDesc), C.POINTER(_vp)]
    lib.tgb_destroy.argtypes = [_vp]
    lib.tgb_add_model_flat.argtypes = [_vp, C.POINTER(BodyDesc), _i32, C.POINTER(ShapeDesc), _i32, C.POINTER(CableDesc), _i32]
    lib.tgb_finalize.argtypes = [_vp]
    for name in ("tgb_num_bodies", "tgb_num_cables", "tgb_num_envs"):
        getattr(lib, name).argtypes = [_vp, C.POINTER(_i32)]
    lib.tgb_set_rod_state.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_get_rod_state.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_set_cable_state.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_get_cable_obs.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_set_controller.argtypes = [_vp, _i32, _vp, _i32, _i32, _i32, _i32]
    lib.tgb_set_control_input.argtypes = [_vp, _i32, _i32, _vp, _i32, C.c_double]
    lib.tgb_get_cpg_state.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_set_cpg_controller.argtypes = [_vp, C.POINTER(CpgDesc), _vp, _i32, _i32, _i32, _i32]
    lib.tgb_step.argtypes = [_vp, C.c_double, _i32]
    lib.tgb_sync.argtypes = [_vp]
    lib.tgb_step_world.argtypes = [_vp, C.c_double]
    lib.tgb_step_models.argtypes = [_vp, C.c_double]
    lib.tgb_track_max_tension.argtypes = [_vp, _i32]
    lib.tgb_get_max_tension.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_rollout_host.argtypes = [_vp, _vp, _vp, C.c_double, _i32, _vp, _vp]
    lib.tgb_get_time.argtypes = [_vp, _dp]
    lib.tgb_get_status.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_mark_episode_start.argtypes = [_vp]
    lib.tgb_get_scores.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_scores_device_ptr.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]
    lib.tgb_scores_refresh_async.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]
    lib.tgb_step_timed.argtypes = [_vp, C.c_double, _i32, _i32, C.POINTER(C.c_float)]
    lib.tgb_launch_count.argtypes = [_vp, C.POINTER(C.c_int64)]
    lib.tgb_launch_geometry.argtypes = [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]
    lib.tgb_set_lanes_per_env.argtypes = [_vp, _i32]
    lib.tgb_set_block_threads.argtypes = [_vp, _i32]
    lib.tgb_set_wide_layout.argtypes = [_vp, _i32]
    lib.tgb_state_bytes_per_env.argtypes = [_vp, C.POINTER(C.c_size_t)]
    lib.tgb_set_height_window.argtypes = [_vp, C.c_double, C.c_double]
    lib.tgb_stream_handle.argtypes = [_vp, C.POINTER(_vp)]
    lib.tgb_set_pair_contact_cap.argtypes = [_vp, _i32]
    lib.tgb_get_min_clearance.argtypes = [_vp, _i32, _i32, _vp, _i32]
    lib.tgb_set_warm_starting.argtypes = [_vp, C.c_double]
    lib.tgb_set_tail_split.argtypes = [_vp, _i32]
    lib.tgb_add_contact_cables.argtypes = [_vp, C.POINTER(ContactCableDesc), _i32]
    lib.tgb_num_contact_cables.argtypes = [_vp, C.POINTER(_i32)]
    lib.tgb_get_contact_cable_obs.argtypes = [_vp, _i32, _i32, _vp]
    lib.tgb_get_contact_cable_anchors.argtypes = [_vp, _i32, _i32, _i32, _vp]
    lib.tgb_set_contact_cable_control_input.argtypes = [_vp, _i32, _i32, _vp, _i32, C.c_double]
    lib.tgb_reset_contact_cables.argtypes = [_vp, _i32, _i32, _vp]
    for name in ABI_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("tgb_abi_version", "tgb_last_error", "tgb_device_count"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def load_host():
    """Load libtgb_host.so: the C++ host API (tgStructure / tgBuildSpec / tgModel ...) with its C helpers."""
    global _host
    if _host is not None:
        return _host
    load()
    path = host_lib_path()
    if not os.path.exists(path):
        raise OSError("%s is missing: build it with `make`" % path)
    h = C.CDLL(path, mode=C.RTLD_GLOBAL)
    h.tgh_last_error.restype = C.c_char_p
    h.tgh_build_example.restype = _vp
    h.tgh_build_example.argtypes = [C.c_char_p]
    h.tgh_free.argtypes = [_vp]
    h.tgh_counts.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    h.tgh_copy.argtypes = [_vp, C.POINTER(BodyDesc), C.POINTER(ShapeDesc), C.POINTER(CableDesc), C.POINTER(GroundDesc), _dp, _dp]
    h.tgh_contact_cables.argtypes = [_vp, C.POINTER(ContactCableDesc), C.c_int]
    h.tgh_contact_cables.restype = C.c_int
    h.tgh_tree_counts.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    h.tgh_escape_transform_actions.argtypes = [_vp, C.c_double, _vp]
    h.tgh_escape_transform_actions.restype = None
    h.tgh_run_example.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _vp, _vp]
    h.tgh_run_example.restype = C.c_int
    h.tgh_run_example_stepwise.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _vp, _vp]
    h.tgh_run_example_stepwise.restype = C.c_int
    h.tgh_run_example_logged.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_double, C.c_int, C.c_char_p, C.c_int]
    h.tgh_run_example_logged.restype = C.c_int
    h.tgh_tetraspine_cpg.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _vp, _vp, _vp, _vp, _vp, _vp]
    h.tgh_tetraspine_param_row.argtypes = [_vp, _vp, C.c_int, _vp]
    h.tgh_flemons_cpg.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _vp, _vp, _vp, _vp, _vp, _vp]
    h.tgh_flemons_param_row.argtypes = [_vp, _vp, _vp, C.c_int, _vp]
    h.tgh_quaternion_between.argtypes = [_vp, _vp, _vp, _vp, _vp]
    h.tgh_quaternion_between.restype = None
    h.tgh_yaml_to_json.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    h.tgh_yaml_to_json.restype = C.c_int
    _host = h
    return h


def check(status):
    if status != TGB_OK:
        raise TgbError(status, load().tgb_last_error().decode("utf-8", "replace"))
