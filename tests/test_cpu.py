"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, the host-side model builder
reproduces the reference's construction arithmetic, and the oracle satisfies the invariants and reference-test pins
that exist for this path."""
import ctypes as C
import math
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ntrtsim_b200 as nt  # noqa: E402
from ntrtsim_b200 import _capi as capi  # noqa: E402
import oracle_py  # noqa: E402
from helpers import random_poses, quat_rotate  # noqa: E402


# ------------------------------------------------------------------------------------------------ the boundary
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ntrtsim_b200.h")).read()
    declared = set(re.findall(r"\b(tgb_[a-z0-9_]+)\s*\(", header))
    declared -= {"tgb_world"}
    assert declared == set(capi.ABI_SYMBOLS), declared ^ set(capi.ABI_SYMBOLS)
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    # 2: tgb_cpg_desc gained cable_control and the PID gains; 3: box shapes (tgb_shape_desc.half_extent_z); 4: contact between bodies
    assert lib.tgb_abi_version() == 6


def test_create_fails_loudly_without_a_gpu():
    lib = capi.load()
    if lib.tgb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    model = nt.build_example("3_prism")
    with pytest.raises(capi.TgbError) as ei:
        nt.BatchedWorld(model, 4)
    assert ei.value.status == capi.TGB_ERR_CUDA
    assert "no CPU path" in str(ei.value)


def test_product_never_touches_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "ntrtsim_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "oracle_py" not in text and "tg_oracle" not in text and "liboracle" not in text, os.path.join(base, f)


def test_create_argument_validation():
    lib = capi.load()
    h = C.c_void_p()
    model = nt.build_example("3_prism")
    cfg = capi.WorldCfg(981.0, 0, 0, 0, 0)
    assert lib.tgb_create(C.byref(cfg), C.byref(model.ground), C.byref(h)) == capi.TGB_ERR_INVALID_ARGUMENT
    assert lib.tgb_create(None, C.byref(model.ground), C.byref(h)) == capi.TGB_ERR_INVALID_ARGUMENT
    cfg = capi.WorldCfg(981.0, 1, 7, 0, 0)
    assert lib.tgb_create(C.byref(cfg), C.byref(model.ground), C.byref(h)) == capi.TGB_ERR_INVALID_ARGUMENT


# ------------------------------------------------------------------------------------------------ model builder
def test_prism_builder_matches_reference_construction():
    m = nt.build_example("3_prism")
    assert (m.n_bodies, m.n_cables) == (3, 9) and nt.tree_counts("3_prism") == (3, 9)
    # mass = (float)(pi r^2 L rho): tgRodInfo.cpp:116-123 through tgBulletUtil.cpp:39-42
    lengths = [math.sqrt(10 ** 2 + 20 ** 2), math.sqrt(5 ** 2 + 20 ** 2 + 10 ** 2), math.sqrt(5 ** 2 + 20 ** 2 + 10 ** 2)]
    for b, L in zip(m.bodies, lengths):
        mass = float(np.float32(math.pi * 0.31 ** 2 * L * 0.2))
        assert b.mass == mass
        # btCylinderShape(r, L/2, r)::calculateLocalInertia, axis Y
        assert abs(b.inertia[1] - 0.5 * mass * 0.31 ** 2) < 1e-12
        assert abs(b.inertia[0] - (mass / 12 * L ** 2 + mass / 4 * 0.31 ** 2)) < 1e-9
        assert b.inertia[0] == b.inertia[2] and b.gyroscopic == 1
    assert abs(m.bodies[0].mass - 1.3501694) < 1e-6 and abs(m.bodies[1].mass - 1.3835119) < 1e-6
    # rest = |p2 - p1| - pretension / k with the anchors moved to the rod surface; prev = rest (tgSpringCable.cpp:49-59)
    st = m.initial_rod_state()
    for c in m.cables:
        pa = quat_rotate(st[c.body_a, 3:7], np.array(list(c.local_a))) + st[c.body_a, 0:3]
        pb = quat_rotate(st[c.body_b, 3:7], np.array(list(c.local_b))) + st[c.body_b, 0:3]
        assert abs(np.linalg.norm(pb - pa) - 500.0 / 1000.0 - c.rest_length) < 1e-9
        assert c.prev_length == c.rest_length and c.actuator == capi.TGB_ACT_BASIC
        # anchor sits on the cylinder surface: distance from the rod axis == radius
        assert abs(math.hypot(c.local_a[0], c.local_a[2]) - 0.31) < 1e-9
    assert m.ground.kind == capi.TGB_GROUND_BOX and m.ground.offset == 1.5 and m.ground.friction == 0.5


def test_superball_builder_quirks():
    t6 = nt.build_example("SUPERball")
    esc = nt.build_example("Escape_c3")
    assert t6.n_cables == 25 and esc.n_cables == 24  # duplicated (2,4)/(4,2) pair: T6Model.cpp:143,153
    assert abs(t6.bodies[0].mass - 3.4978709) < 1e-6
    assert abs(t6.bodies[0].inertia[0] - 82.746) < 1e-2 and abs(t6.bodies[0].inertia[1] - 0.16807) < 1e-4
    # "pitch" of AppSUPERball.cpp:53-58 goes through btQuaternion::setEuler(yaw, pitch, roll): a tilt about X
    n = list(t6.ground.normal)
    assert abs(n[0]) < 1e-12 and abs(n[1] - math.cos(math.pi / 15)) < 1e-12 and abs(abs(n[2]) - math.sin(math.pi / 15)) < 1e-12
    # EscapeModel.cpp:184-185 positional slip: pretension 0 -> rest length == start length
    st = esc.initial_rod_state()
    c = esc.cables[0]
    pa = quat_rotate(st[c.body_a, 3:7], np.array(list(c.local_a))) + st[c.body_a, 0:3]
    pb = quat_rotate(st[c.body_b, 3:7], np.array(list(c.local_b))) + st[c.body_b, 0:3]
    assert abs(np.linalg.norm(pb - pa) - c.rest_length) < 1e-9
    assert c.max_tension == 100000 and c.target_velocity == 10000 and c.stiffness == 3152.36 and c.damping == 50.0


def test_motor_rig_builder():
    m = nt.build_example("tsTestRig_kinematic")
    assert m.n_bodies == 2 and m.n_cables == 1
    masses = sorted(b.mass for b in m.bodies)
    assert masses[0] == 0.0 and masses[1] > 0.0  # density 0 -> static body (tsTestRig.cpp:113,135)
    assert m.cables[0].actuator == capi.TGB_ACT_KINEMATIC and m.cables[0].motor_radius == 1.0 and m.cables[0].backdrivable == 0


def test_escape_transform_actions():
    a = np.full(32, 0.5)
    p = nt.escape_transform_actions(a, 9.0)
    # amplitude in L0*(0.9 -+ 0.5), angular frequency in [0.3, 20], phase change in [-pi, pi], dcOffset forced to 1
    assert np.allclose(p[0:4], [9.0 * 0.9, 10.15, 0.0, 1.0]) and p[32] == 9.0


# ------------------------------------------------------------------------------------------------ oracle pins
def _oracle(model, rods=None, cab=None):
    o = oracle_py.OracleWorld(model, capi.WorldCfg)
    if rods is not None:
        o.set_rod_state(rods)
    if cab is not None:
        o.set_cable_state(cab)
    return o


def test_tgutil_quaternion_between_opposite_vectors():
    """test/tgcreator/tgUtil_test.cpp:72-97 (tgUtilTest.testQuaternion) on the host API's tgUtil: rotating by the quaternion between two
    opposite vectors maps one onto the other, for the up / down pair btTransforms are defined against and for an arbitrary pair;
    the reference's pass mark is btVector3::fuzzyZero, |d|^2 < SIMD_EPSILON (double build: DBL_EPSILON)."""
    h = capi.load_host()

    def between(a, b, v):
        a, b, v = (np.array(x, dtype=np.float64) for x in (a, b, v))
        q, r = np.zeros(4), np.zeros(3)
        vp = C.c_void_p
        h.tgh_quaternion_between(a.ctypes.data_as(vp), b.ctypes.data_as(vp), v.ctypes.data_as(vp), q.ctypes.data_as(vp), r.ctypes.data_as(vp))
        return q, r
    up, down = (0.0, 1.0, 0.0), (0.0, -1.0, 0.0)
    q, result = between(down, up, down)
    assert np.sum((np.array(up) - result) ** 2) < np.finfo(np.float64).eps
    assert abs(np.linalg.norm(q) - 1.0) < 1e-15
    start, end = (1.0, -1.0, 2.0), (-1.0, 1.0, -2.0)
    q, result = between(start, end, end)
    assert np.sum((np.array(start) - result) ** 2) < np.finfo(np.float64).eps


def test_oracle_step_rejects_nonpositive_dt():
    o = _oracle(nt.build_example("3_prism"))
    assert o.step(0.0, 1) == capi.TGB_ERR_INVALID_ARGUMENT and o.step(-1.0, 1) == capi.TGB_ERR_INVALID_ARGUMENT


def test_oracle_first_step_is_gravity_only_then_double_force():
    """World step runs before the cables (tgSimulation.cpp:205-211): the first world step sees gravity only; the first
    cable force is 2 k stretch (prevLength = restLength, damping clamped to |k stretch|)."""
    m = nt.build_example("3_prism_noground")
    o = _oracle(m)
    s0 = o.get_rod_state()
    o.step(1e-3, 1)
    s1 = o.get_rod_state()
    g, dt = 981.0, 1e-3
    # positions moved by exactly the gravity-only velocity g dt over dt
    assert np.allclose(s1[:, 1] - s0[:, 1], -g * dt * dt, atol=1e-12)
    assert np.allclose(s1[:, 0] - s0[:, 0], 0.0, atol=1e-15)
    # velocity after step 1 = gravity + impulses of 2 * 500 on each incident cable end
    st = s1
    J = np.zeros((3, 3))
    for c in m.cables:
        pa = quat_rotate(st[c.body_a, 3:7], np.array(list(c.local_a))) + st[c.body_a, 0:3]
        pb = quat_rotate(st[c.body_b, 3:7], np.array(list(c.local_b))) + st[c.body_b, 0:3]
        d = pb - pa
        L = np.linalg.norm(d)
        stretch = L - c.rest_length
        f = d / L * (2.0 * c.stiffness * stretch)
        J[c.body_a] += f * dt
        J[c.body_b] -= f * dt
    for b in range(3):
        v = np.array([0, -g * dt, 0]) + J[b] / m.bodies[b].mass
        assert np.allclose(s1[b, 7:10], v, atol=1e-9)


def test_oracle_momentum_conserved_without_gravity():
    import copy
    m = copy.copy(nt.build_example("3_prism_noground"))
    m.gravity = 0.0
    rods = random_poses(m, 1, seed=3)[0]
    o = _oracle(m, rods, m.initial_cable_state())
    mass = np.array([b.mass for b in m.bodies])
    o.step(1e-3, 2000)
    s = o.get_rod_state()
    p = (s[:, 7:10] * mass[:, None]).sum(axis=0)
    assert np.abs(p).max() < 1e-10  # MuscleNP_test.cpp:108-112 bound
    # and the structure has rung down: cable tensions settle near equilibrium, all positive
    assert (o.get_cable_obs()[:, 2] > 0).all()


def test_oracle_free_fall_and_gyroscopic_spin():
    """A single free rod (cables detached by zero stiffness is not possible, so use the rig's dynamic rod with the cable
    slack): free fall x = -g/2 t(t+dt) for semi-implicit Euler; spin about the symmetry axis is steady under the
    gyroscopic term."""
    import copy
    m = copy.copy(nt.build_example("tsTestRig_basic"))
    m.ground = capi.GroundDesc(capi.TGB_GROUND_EMPTY, (0.0, 1.0, 0.0), 0.0, 0.5, 0.0, 0.0, 0.04)
    dyn = [i for i, b in enumerate(m.bodies) if b.mass > 0][0]
    rods = m.initial_rod_state()
    rods[dyn, 10:13] = quat_rotate(rods[dyn, 3:7], np.array([0.0, 7.0, 0.0]))  # spin about the rod axis
    cab = m.initial_cable_state() * 0 + 1000.0  # rest length 1000: slack for the whole fall, no force
    o = _oracle(m, rods, cab)
    n, dt = 500, 1e-3
    o.step(dt, n)
    s = o.get_rod_state()
    t = n * dt
    assert abs((s[dyn, 1] - rods[dyn, 1]) - (-0.5 * 981.0 * t * (t + dt))) < 1e-9
    assert abs(s[dyn, 8] + 981.0 * t) < 1e-9
    assert np.allclose(s[dyn, 10:13], rods[dyn, 10:13], atol=1e-9)


def test_oracle_rod_rests_on_the_ground():
    """3-prism dropped on the box ground (config 1): after the impact transient the lowest rim point of every rod ends
    within the contact margin band of the surface and the structure is at rest."""
    m = nt.build_example("3_prism")
    o = _oracle(m)
    o.step(1e-3, 10000)
    s = o.get_rod_state()
    assert np.abs(s[:, 7:10]).max() < 1.0, s[:, 7:10]  # settled to a slow creep (impact speed was ~130)
    assert o.num_contacts() >= 3
    lows = []
    for b, body in enumerate(m.bodies):
        sh = m.shapes[body.shape_begin]
        axis = quat_rotate(s[b, 3:7], np.array([0.0, 1.0, 0.0]))
        ends = [s[b, 0:3] + axis * sh.half_length, s[b, 0:3] - axis * sh.half_length]
        lows.append(min(e[1] for e in ends) - sh.radius * math.sqrt(max(0.0, 1 - axis[1] ** 2)))
    assert min(lows) > 1.5 - 0.1 and min(lows) < 1.5 + 0.1, lows


def test_oracle_box_comes_to_rest_on_a_face(tmp_path):
    """A tgBox dropped slightly tilted: it lands on corners, tips onto its lowest face and stays there with
    the face's four corners as the manifold, its centre one half extent (the smallest: height 0.5) above the surface."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_yaml import write
    from ntrtsim_b200 import examples
    path = write(tmp_path, "box.yaml", """
nodes:
  p0: [0, 2.4, -1.5]
  p1: [0.2, 2.5, 1.5]
pair_groups:
  box:
    - [p0, p1]
builders:
  box:
    class: tgBoxInfo
    parameters:
      width: 0.8
      height: 0.5
      density: 0.3
      friction: 0.7
      restitution: 0.0
""")
    m = examples.build_yaml(path)
    o = _oracle(m)
    o.step(1e-3, 3000)
    s = o.get_rod_state()
    assert np.abs(s[0, 7:13]).max() < 0.05, s[0, 7:13]
    assert o.num_contacts() == 4
    # the local axes: Y along the pair, X = width, Z = height; at rest one of them points up
    up = [abs(quat_rotate(s[0, 3:7], np.eye(3)[k])[1]) for k in range(3)]
    k = int(np.argmax(up))
    assert up[k] > 0.999
    half = [0.8, 1.5, 0.5][k]
    assert abs(s[0, 1] - (1.5 + half)) < 0.05, (s[0, 1], half)


def test_crater_obstacle_models():
    """tgCraterDeep / tgCraterShallow (src/models/obstacles): four static boxes a quarter turn apart, sunk by 5; as the second model of
    the AppEscape simulation they follow EscapeModel's six rods in the flattened world.  Deep: walls 20 from the origin.  Shallow
    (tgCraterShallow.cpp:124-168): boxes along the diagonal between (-25, -7, -25) and (25, 7, 25) turned an eighth -- low ramps; the
    exact descriptors are held against the reference's own source in tests/test_reference_sources.py."""
    m = nt.build_example("Escape_crater")
    assert (m.n_bodies, m.n_cables) == (10, 24)
    walls = [i for i, b in enumerate(m.bodies) if b.mass == 0.0]
    assert walls == [6, 7, 8, 9]
    pos = m.initial_rod_state()[walls, 0:3]
    assert np.allclose(np.sort(np.abs(pos[:, [0, 2]]).max(axis=1)), 20.0) and np.allclose(pos[:, 1], 2.5)
    assert np.allclose(np.abs(pos[:, [0, 2]]).min(axis=1), 0.0, atol=1e-12)
    for i in walls:
        s = m.shapes[m.bodies[i].shape_begin]
        assert (s.kind, s.radius, s.half_length, s.half_extent_z) == (2, 15.0, 17.5, 15.0)
        assert (m.bodies[i].friction, m.bodies[i].rolling_friction, m.bodies[i].restitution) == (1.0, 0.01, 0.2)
    m = nt.build_example("Escape_crater_shallow")
    assert (m.n_bodies, m.n_cables) == (10, 24)
    ramps = [i for i, b in enumerate(m.bodies) if b.mass == 0.0]
    assert ramps == [6, 7, 8, 9]
    pos = m.initial_rod_state()[ramps, 0:3]
    half_diagonal = 0.5 * math.sqrt(50.0 ** 2 + 14.0 ** 2 + 50.0 ** 2)
    assert np.allclose(pos[:, 1], -5.0)                                                       # centred 5 below the origin's height
    assert np.allclose(pos.sum(axis=0)[[0, 2]], 0.0, atol=1e-9)                               # a quarter turn apart about the vertical
    assert np.allclose(np.hypot(pos[:, 0], pos[:, 2]), np.hypot(pos[0, 0], pos[0, 2]))
    for i in ramps:
        s = m.shapes[m.bodies[i].shape_begin]
        assert (s.kind, s.radius, s.half_extent_z) == (2, 10.0, 10.0) and abs(s.half_length - half_diagonal) < 1e-9
        assert (m.bodies[i].friction, m.bodies[i].rolling_friction, m.bodies[i].restitution) == (1.0, 0.01, 0.2)


def test_oracle_bodies_rest_on_a_static_obstacle(tmp_path):
    """Static tgBox obstacles (tgCraterDeep / tgBlockField are built from them): a rod and a ball dropped on a
    block come to rest on its top face (y = 3.5) instead of falling through to the ground, the block itself never moves, and a
    rod dropped beside the block slides off its side face and ends on the ground without entering the block."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_yaml import OBSTACLE_SCENE, write
    from ntrtsim_b200 import examples
    m = examples.build_yaml(write(tmp_path, "scene.yaml", OBSTACLE_SCENE))
    static = [i for i, b in enumerate(m.bodies) if b.mass == 0.0]
    assert len(static) == 1 and m.shapes[m.bodies[static[0]].shape_begin].kind == 2
    moving = [i for i in range(m.n_bodies) if i not in static]
    rods = [i for i in moving if m.shapes[m.bodies[i].shape_begin].kind == 0]
    balls = [i for i in moving if m.shapes[m.bodies[i].shape_begin].kind == 1]
    assert (len(rods), len(balls)) == (2, 1)
    o = _oracle(m)
    start = o.get_rod_state().copy()
    o.step(1e-3, 3000)
    s = o.get_rod_state()
    assert np.array_equal(s[static[0]], start[static[0]])
    top_rod = max(rods, key=lambda i: start[i, 1])
    side_rod = min(rods, key=lambda i: start[i, 1])
    assert abs(s[top_rod, 1] - (3.5 + 0.2)) < 0.03, s[top_rod, 0:3]          # lying on the top face
    assert abs(s[balls[0], 1] - (3.5 + 0.5)) < 0.03, s[balls[0], 0:3]        # resting on the top face
    assert np.abs(s[[top_rod, balls[0]], 7:13]).max() < 0.05
    assert s[side_rod, 1] < 2.5 and abs(s[side_rod, 2]) > 1.0 + 0.2 - 0.05, s[side_rod, 0:3]  # beside the block, on the ground
    assert np.isfinite(s).all()


@pytest.mark.parametrize("kind,command", [("basic", 5.0), ("kinematic", -580.0)])
def test_oracle_motor_timestep_independence(kind, command):
    """MotorTimestep_test.cpp:81-215: 1000 steps @1e-3 vs 500 @2e-3 vs 5000 @2e-4: rest length equal within 0.03."""
    m = nt.build_example("tsTestRig_" + kind)
    ctrl = capi.TGB_CTRL_SETPOINT if kind == "basic" else capi.TGB_CTRL_TORQUE
    rest = {}
    for dt, n in ((1e-3, 1000), (2e-3, 500), (2e-4, 5000)):
        o = _oracle(m)
        o.set_controller(ctrl, np.array([command]))
        o.step(dt, n)
        assert abs(o.time() - 1.0) < 1e-9
        rest[dt] = o.get_cable_obs()[0, 0]
    assert abs(rest[2e-3] - rest[1e-3]) <= 0.03 and abs(rest[2e-4] - rest[1e-3]) <= 0.03, rest
    if kind == "basic":
        assert abs(rest[1e-3] - 5.0) < 1e-9  # commanded rest length reached (tsTestRig.cpp:187)


def test_oracle_sine_controller_moves_motors_twice():
    """EscapeController.cpp:100-117: setControlInput(L, dt) then moveMotors(dt) again -> two motor steps per sim step."""
    m = nt.build_example("Escape_c3")
    o = _oracle(m)
    params = np.zeros(33)
    params[3::4][:8] = 1.0  # dcOffset
    params[32] = 9.0
    for c in range(8):
        params[c * 4 + 0] = 0.0   # amplitude 0 -> target = dc = 1 -> clamped to L0/2 = 4.5
        params[c * 4 + 1] = 1.0
    o.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params)
    rest0 = o.get_cable_obs()[:, 0].copy()
    o.step(1e-3, 1)
    rest1 = o.get_cable_obs()[:, 0]
    # targetVelocity 10000 * dt 1e-3 = 10 per motor move: the rest length (~10.4) reaches 4.5 within the first move
    assert np.allclose(rest1, 4.5, atol=1e-12), rest1
    # tighter motors: tVel such that one move covers less than the gap -> exactly two steps taken
    for c in m.cables:
        c.target_velocity = 1000.0
    o = _oracle(m)
    o.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params)
    o.step(1e-3, 1)
    assert np.allclose(o.get_cable_obs()[:, 0], rest0 - 2.0, atol=1e-12)


def test_oracle_batch_rollout_matches_single():
    m = nt.build_example("Escape_c3")
    rods = random_poses(m, 5, seed=9, translate=1.0, lift=(0, 3), yaw_only=True)
    cab = np.tile(m.initial_cable_state(), (5, 1, 1))
    params = nt.escape_transform_actions(np.random.default_rng(1).uniform(size=(5, 32)))
    ro, co = oracle_py.rollout_batch(m, capi.WorldCfg, rods, cab, 1e-3, 100, capi.TGB_CTRL_ESCAPE_SINE, params, n_threads=3)
    for e in range(5):
        o = _oracle(m, rods[e], cab[e])
        o.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params[e])
        o.step(1e-3, 100)
        assert np.array_equal(o.get_rod_state(), ro[e]) and np.array_equal(o.get_cable_obs(), co[e])


def test_data_logger_schema_and_errors(tmp_path):
    """tgDataLogger2 + tgRodSensorInfo + tgSpringCableActuatorSensorInfo of the host API write the reference's CSV schema
    (src/sensors/tgDataLogger2.cpp:107-170, tgRodSensor.cpp:73-119, tgSpringCableActuatorSensor.cpp:60-85): one line of free text, then
    "time,<sensor index>_<type>(<tags>).<field>,..." -- checked without a device (zero steps), with the reference's exceptions."""
    import ctypes as C
    h = capi.load_host()
    buf = C.create_string_buffer(512)
    rc = h.tgh_run_example_logged(b"3_prism", 1, 0, 0, str(tmp_path / "prism").encode(), 0.0, 0, buf, 512)
    assert rc == 0, h.tgh_last_error()
    name = buf.value.decode()
    assert re.match(r".*/prism_\d{8}_\d{6}\.txt$", name)
    lines = open(name).read().split("\n")
    assert re.match(r"tgDataLogger2 started logging at time \d{8}_\d{6}, with 12 sensors on 1 senseable objects\.$", lines[0])
    cols = lines[1].split(",")
    assert cols[0] == "time" and cols[-1] == ""               # every field ends with a comma
    assert cols[1:8] == ["0_rod(rod).X", "0_rod(rod).Y", "0_rod(rod).Z", "0_rod(rod).Euler1", "0_rod(rod).Euler2", "0_rod(rod).Euler3",
                         "0_rod(rod).mass"]
    assert cols[22:25] == ["3_SCA(muscle).RestLen", "3_SCA(muscle).CurrLen", "3_SCA(muscle).Tension"]
    assert len(cols) == 1 + 3 * 7 + 9 * 3 + 1
    assert lines[2:] == [""]                                   # no samples before the first step
    assert h.tgh_run_example_logged(b"3_prism", 1, 0, 0, b"", 0.0, 0, buf, 512) == capi.TGB_ERR_INVALID_ARGUMENT
    assert b"File name cannot be the empty string" in h.tgh_last_error()
    assert h.tgh_run_example_logged(b"3_prism", 1, 0, 0, str(tmp_path / "x").encode(), -1.0, 0, buf, 512) == capi.TGB_ERR_INVALID_ARGUMENT
    assert b"Time interval must be nonnegative" in h.tgh_last_error()
    assert h.tgh_run_example_logged(b"3_prism", 1, 0, 0, b"/nonexistent_dir/x", 0.0, 0, buf, 512) == capi.TGB_ERR_RUNTIME
    assert b"Log file could not be opened" in h.tgh_last_error()


def test_data_logger_compound_rigid_sensors(tmp_path):
    """tgCompoundRigidSensorInfo (src/sensors/tgCompoundRigidSensor{,Info}.cpp): the members of an auto-compounded body share a tag
    "compound_XXXXXX" (tgRigidAutoCompound.cpp:146-157) and the logger gets one compound sensor per tag -- TetraSpine: 3 segments of 6
    rods -- with the headings compound(tag).X ... .mass, next to the 18 rod and 12 cable sensors."""
    import ctypes as C
    h = capi.load_host()
    buf = C.create_string_buffer(512)
    rc = h.tgh_run_example_logged(b"TetraSpine", 1, 0, 0, str(tmp_path / "spine").encode(), 0.0, 0, buf, 512)
    assert rc == 0, h.tgh_last_error()
    lines = open(buf.value.decode()).read().split("\n")
    assert re.match(r".* with 33 sensors on 1 senseable objects\.$", lines[0]), lines[0]
    cols = lines[1].split(",")[1:-1]
    compound = [c for c in cols if "_compound(" in c]
    tags = sorted(set(re.findall(r"compound\((compound_\w{6})\)", " ".join(compound))))
    assert len(tags) == 3 and len(compound) == 3 * 7
    first = [c.split(".", 1)[1] for c in compound[:7]]
    assert first == ["X", "Y", "Z", "YawDegrees", "PitchDegrees", "RollDegrees", "mass"]
    # every rod of a segment carries its segment's compound tag
    rods = [c for c in cols if "_rod(" in c and c.endswith(".X")]
    assert len(rods) == 18
    for t in tags:
        assert sum(1 for c in rods if t in c) == 6


# ------------------------------------------------------------------------------------------------ contact: analytic known answers
# Bullet is not in the reference tree, so the contact half of the step cannot be pinned against it; these tests pin
# it against mechanics instead: Coulomb friction, rolling without slipping, and the sliding threshold on an incline.
BALL = """
nodes:
  ball: [0, 2.0, 0]
builders:
  ball:
    class: tgSphereInfo
    parameters:
      radius: 0.5
      density: 1.0
      friction: 1.0
      roll_friction: 0.0
      restitution: 0.0
"""

BRICK = """
nodes:
  p0: [0, 0, -1]
  p1: [0, 0, 1]
pair_groups:
  box:
    - [p0, p1]
builders:
  box:
    class: tgBoxInfo
    parameters:
      width: 0.6
      height: 0.3
      density: 1.0
      friction: 0.9
      roll_friction: 0.0
      restitution: 0.0
"""


def _build_text(tmp_path, name, text):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_yaml import write
    from ntrtsim_b200 import examples
    return examples.build_yaml(write(tmp_path, name, text))


def sliding_ball_expectations(model, v0=10.0):
    mu = model.bodies[0].friction * model.ground.friction  # 1.0 x 0.5 (btBoxGround never applies its Config friction: Bullet default)
    g = model.gravity
    t_roll = 2.0 * v0 / (7.0 * mu * g)  # solid sphere (I = 2/5 m r^2): slipping ends when v = 5/7 v0
    return mu, g, t_roll, 5.0 / 7.0 * v0


def test_oracle_sliding_ball_ends_rolling_at_five_sevenths(tmp_path):
    """A ball set sliding at v0 on the flat ground decelerates at mu g while it slips, spins up, and rolls without slipping at
    5/7 v0 from t = 2 v0 / (7 mu g) on (rolling friction 0)."""
    m = _build_text(tmp_path, "ball.yaml", BALL)
    assert abs(m.bodies[0].inertia[0] - 0.4 * m.bodies[0].mass * 0.25) < 1e-12  # btSphereShape::calculateLocalInertia
    mu, g, t_roll, v_roll = sliding_ball_expectations(m)
    rods = m.initial_rod_state()
    rods[0, 7] = 10.0
    o = _oracle(m, rods)
    n_mid = int(0.5 * t_roll / m.dt)
    o.step(m.dt, n_mid)
    s = o.get_rod_state()
    assert abs(s[0, 7] - (10.0 - mu * g * n_mid * m.dt)) < 0.02 * 10.0          # Coulomb deceleration while slipping
    assert abs(s[0, 12] - (-2.5 * mu * g * n_mid * m.dt / 0.5)) < 0.02 * 20.0   # spin-up: dw/dt = -(5/2) mu g / r about z
    o.step(m.dt, int(2.0 * t_roll / m.dt))
    s = o.get_rod_state()
    assert abs(s[0, 7] - v_roll) < 0.01 * v_roll, s[0, 7:13]
    assert abs(s[0, 12] + s[0, 7] / 0.5) < 0.01 * v_roll / 0.5                  # v = -w_z r: no slip
    v_then = s[0, 7]
    o.step(m.dt, 200)
    assert abs(o.get_rod_state()[0, 7] - v_then) < 1e-3 * v_then                # and it keeps rolling
    assert abs(o.get_rod_state()[0, 1] - 2.0) < 0.03


def incline(model, theta):
    """The model's ground turned into a slope of angle theta rising along +x, through the point (0, 1.5, 0)."""
    model.ground.normal[0], model.ground.normal[1], model.ground.normal[2] = -math.sin(theta), math.cos(theta), 0.0
    model.ground.offset = 1.5 * math.cos(theta)
    return model


def brick_on_incline_state(model, theta):
    """The brick (half extents 0.6 across, 0.3 thick after the builder's frame: local Z = thickness) lying on the slope at rest."""
    rods = model.initial_rod_state()
    n = np.array([-math.sin(theta), math.cos(theta), 0.0])
    # builder frame: local Y along the pair (world z), local X = width -> world x, local Z = height -> world -y; turn about z by theta
    q_tilt = np.array([0.0, 0.0, math.sin(theta / 2), math.cos(theta / 2)])
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import quat_mul
    rods[0, 3:7] = quat_mul(q_tilt[None], rods[0, 3:7][None])[0]
    rods[0, 0:3] = np.array([0.0, 1.5, 0.0]) + n * (0.3 + 0.01)
    return rods


@pytest.mark.parametrize("theta,slides", [(0.30, False), (0.60, True)])
def test_oracle_brick_on_incline_sticks_below_the_friction_angle(tmp_path, theta, slides):
    """mu = 0.9 x 0.5 = 0.45 (friction angle 0.42 rad): a brick on a 0.30 rad slope stays put, on a 0.60 rad slope it slides down with
    a = g (sin theta - mu cos theta)."""
    m = incline(_build_text(tmp_path, "brick.yaml", BRICK), theta)
    mu = m.bodies[0].friction * m.ground.friction
    o = _oracle(m, brick_on_incline_state(m, theta))
    o.step(m.dt, 100)  # settle onto the slope
    s0 = o.get_rod_state().copy()
    T = 0.2
    o.step(m.dt, int(T / m.dt))
    s1 = o.get_rod_state()
    down = np.array([-math.cos(theta), -math.sin(theta), 0.0])
    travelled = float(np.dot(s1[0, 0:3] - s0[0, 0:3], down))
    v_down = float(np.dot(s1[0, 7:10], down))
    if not slides:
        # a velocity-level friction row without position correction creeps a little (3 mm here); frictionless it would travel 0.58
        assert abs(travelled) < 0.01 and abs(v_down) < 0.05, (travelled, v_down)
    else:
        a = m.gravity * (math.sin(theta) - mu * math.cos(theta))
        v_start = float(np.dot(s0[0, 7:10], down))
        assert abs((v_down - v_start) - a * T) < 0.03 * a * T, (v_down - v_start, a * T)


def test_oracle_hanging_rod_settles_at_weight_over_stiffness():
    """The motor rig left alone (src/examples/motorModel/tsTestRig.cpp: a rod hanging from a static rod by one cable, k = 1000,
    c = 10): at rest the cable carries the rod's weight, stretch = m g / k -- whatever the damping law does on the way."""
    m = nt.build_example("tsTestRig_basic")
    weight = [b.mass for b in m.bodies if b.mass != 0.0][0] * m.gravity
    o = _oracle(m)
    o.step(m.dt, 6000)
    rest, length, tension, _ = o.get_cable_obs()[0]
    assert abs(tension - weight) < 2e-3 * weight, (tension, weight)
    assert abs((length - rest) - weight / m.cables[0].stiffness) < 2e-3 * weight / m.cables[0].stiffness


# ------------------------------------------------------------------------------------------------ contact between bodies
# tgWorldBulletPhysicsImpl.cpp:82-107 collides every pair of bodies of the world (rod x rod, vertebra x vertebra).  Known answers:
# crossed rods rest on each other one diameter apart, a rod-rod impact without gravity conserves momentum, and a 3-prism under its
# own pretension folds until its rods touch (they used to pass through each other).
CROSSED_RODS = """
nodes:
  a0: [-3, 0.5, 0]
  a1: [3, 0.5, 0]
  b0: [0.3, 2.0, -3]
  b1: [0.3, 2.0, 3]
pair_groups:
  rod:
    - [a0, a1]
    - [b0, b1]
builders:
  rod:
    class: tgRodInfo
    parameters:
      radius: 0.5
      density: 1.0
      friction: 1.0
      roll_friction: 0.0
      restitution: 0.0
"""


def segment_distance(p1, q1, p2, q2):
    """distance between the segments p1-q1 and p2-q2 (dense sampling; test-side check, independent of the oracle's closed form)"""
    s = np.linspace(0.0, 1.0, 801)
    a = p1[None] + (q1 - p1)[None] * s[:, None]
    b = p2[None] + (q2 - p2)[None] * s[:, None]
    return np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)).min()


def rod_axes(model, rods):
    from helpers import quat_rotate
    out = []
    for i, b in enumerate(model.bodies):
        sh = model.shapes[b.shape_begin]
        ax = quat_rotate(rods[i, 3:7][None], np.array([[0.0, 1.0, 0.0]]))[0]
        out.append((rods[i, 0:3] - ax * sh.half_length, rods[i, 0:3] + ax * sh.half_length, sh.radius))
    return out


RAILS = """
nodes:
  a0: [-3, 2.0, -1.5]
  a1: [3, 2.0, -1.5]
  c0: [-3, 2.0, 1.5]
  c1: [3, 2.0, 1.5]
  b0: [0.3, 3.5, -3]
  b1: [0.3, 3.5, 3]
pair_groups:
  rail:
    - [a0, a1]
    - [c0, c1]
  rod:
    - [b0, b1]
builders:
  rail:
    class: tgRodInfo
    parameters:
      radius: 0.5
      density: 0.0
      friction: 1.0
      restitution: 0.0
  rod:
    class: tgRodInfo
    parameters:
      radius: 0.5
      density: 1.0
      friction: 1.0
      roll_friction: 0.0
      restitution: 0.0
"""


def test_oracle_crossed_rods_rest_one_diameter_apart(tmp_path):
    """a rod dropped across two static rods (rails) comes to rest on them, axes one diameter apart"""
    m = _build_text(tmp_path, "rails.yaml", RAILS)
    assert len(m.bodies) == 3 and [b.mass == 0.0 for b in m.bodies] == [True, True, False]
    o = _oracle(m)
    o.step(m.dt, 1500)
    s = o.get_rod_state()
    assert np.abs(s[2, 7:13]).max() < 0.05, s[2, 7:13]           # at rest
    axes = rod_axes(m, s)
    for rail in (0, 1):
        gap = segment_distance(axes[rail][0], axes[rail][1], axes[2][0], axes[2][1]) - axes[rail][2] - axes[2][2]
        assert -0.045 < gap < 0.021, gap                          # within the contact band of the solver (slop 0.04, breaking threshold 0.02)
    assert abs(s[2, 1] - 3.0) < 0.05                              # on top of the rails (the ground's top face is at y = 1.5)
    assert o.num_pair_contacts() == 2 and o.min_clearance() > -0.1
    # with the pair contacts switched off it falls through the static rods onto the ground
    o2 = _oracle(m)
    o2.set_pair_contact_cap(0)
    o2.step(m.dt, 1500)
    assert abs(o2.get_rod_state()[2, 1] - 2.0) < 0.05


def test_oracle_rod_rod_impact_conserves_momentum(tmp_path):
    m = _build_text(tmp_path, "crossed.yaml", CROSSED_RODS)
    m.ground.kind = capi.TGB_GROUND_EMPTY
    m.gravity = 0.0
    rods = m.initial_rod_state()
    rods[1, 8] = -8.0                                              # the upper rod flies down onto the lower one, off centre
    o = _oracle(m, rods)
    mass = np.array([b.mass for b in m.bodies])

    def momenta(s):
        lin = (mass[:, None] * s[:, 7:10]).sum(0)
        from helpers import quat_rotate
        ang = np.zeros(3)
        for i, b in enumerate(m.bodies):
            wl = quat_rotate(s[i, 3:7][None] * np.array([-1, -1, -1, 1.0]), s[i, 10:13][None])[0]
            Lw = quat_rotate(s[i, 3:7][None], (np.array(list(b.inertia)) * wl)[None])[0]
            ang += Lw + mass[i] * np.cross(s[i, 0:3], s[i, 7:10])
        return lin, ang
    lin0, ang0 = momenta(o.get_rod_state())
    touched = False
    for _ in range(400):
        o.step(m.dt, 1)
        touched = touched or o.num_pair_contacts() > 0
    s = o.get_rod_state()
    lin1, ang1 = momenta(s)
    assert touched and s[0, 8] < -1.0 and s[1, 8] > -7.0, s[:, 8]  # the struck rod moves off, the striker slows down
    assert np.abs(lin1 - lin0).max() < 1e-9 * np.abs(lin0).max()
    # the gyroscopic term is integrated explicitly (btRigidBody::computeGyroscopicForce): angular momentum drifts at its order only
    assert np.abs(ang1 - ang0).max() < 2e-3 * max(1.0, np.abs(ang0).max()), (ang0, ang1)
    (p1, q1, r1), (p2, q2, r2) = rod_axes(m, s)
    assert segment_distance(p1, q1, p2, q2) - r1 - r2 > -0.05      # and they did not pass through each other


def test_oracle_prism_rods_touch_instead_of_passing_through():
    """3_prism without ground: the pretensioned strings fold the prism until its rods meet (step ~180 of the golden drop)."""
    m = nt.build_example("3_prism_noground")
    g = np.load(os.path.join(ROOT, "tests", "golden", "prism_drop.npz"))
    o = _oracle(m, g["rods0"])
    o.step(m.dt, 600)
    assert o.min_clearance() > -0.08, o.min_clearance()
    o2 = _oracle(m, g["rods0"])
    o2.set_pair_contact_cap(0)
    o2.step(m.dt, 600)
    assert o2.min_clearance() < -0.3, o2.min_clearance()          # without pair contacts the rods interpenetrate deeply


# ------------------------------------------------------------------------------------------------ learning outer loop
def test_neuroevolution_population_bookkeeping(tmp_path):
    """Ordering, the lenient maximum, the int-truncated average, mutation of the worst by copies of the best, and the .nnw format with
    the loader's dropped last value (src/learning/NeuroEvolution/*.cpp)."""
    from ntrtsim_b200.neuroevo import NeuroEvoPopulation, load_nnw
    rng = np.random.default_rng(0)
    pop = NeuroEvoPopulation(6, 10, rng, compare_average_scores=False, leniency_coef=0.5)
    assert pop.params.shape == (6, 10) and (pop.params >= 0).all() and (pop.params <= 1).all()
    for m, s in enumerate([3.0, 9.0, 1.0, 7.0, 5.0, 2.0]):
        pop.update_scores(m, (s, 0.0))
    pop.update_scores(1, (4.0, 0.0))           # a worse repeat: max decays to 0.5 * 9 + 0.5 * 4
    assert pop.max_score[1] == 6.5 and pop.max_score1[1] == 9.0
    before = pop.params.copy()
    order = pop.order()
    assert list(order) == [3, 1, 4, 0, 5, 2]   # by maximum: 7, 6.5, 5, 3, 2, 1
    assert np.array_equal(pop.params, before[order])
    pop2 = NeuroEvoPopulation(2, 3, rng, compare_average_scores=True)
    for s in (0.9, 0.9, 0.9):
        pop2.update_scores(0, (s, 0.0))
    pop2.update_scores(1, (1.2, 0.0))
    pop2.order()
    assert pop2.average_score[0] == 1.0 and pop2.average_score[1] == 0.0   # int accumulator: 0.9 + 0.9 + 0.9 -> 0; 1.2 -> 1
    best = pop.params[:2].copy()
    pop.mutate(2)
    changed = pop.params[[5, 4]] != best
    assert (np.abs(pop.params[[5, 4]] - best) <= 0.2).all() and (pop.params >= 0).all() and (pop.params <= 1).all()
    assert changed.mean() < 0.75               # each parameter moves with probability 1/4 overall
    with pytest.raises(ValueError):
        pop.mutate(4)
    path = str(tmp_path / "best.nnw")
    pop.save_best(path)
    assert open(path).read().count(",") == 9
    loaded = load_nnw(path, 10, defaults=np.full(10, -1.0))
    assert np.allclose(loaded[:9], pop.params[0, :9], rtol=1e-5) and loaded[9] == -1.0   # the last value is never read


def test_neuroevolution_generation_loop_improves_a_known_fitness():
    """The batched generation loop on a fitness with a known optimum: first generation tests everybody, later ones only the renewed
    members; the leader's score never decreases when every member keeps a single score."""
    from ntrtsim_b200.neuroevo import NeuroEvolution
    target = np.linspace(0.1, 0.9, 8)
    calls = []

    def evaluate(rows):
        calls.append(len(rows))
        return np.stack([10.0 - np.abs(rows - target).sum(axis=1), np.zeros(len(rows))], axis=1)
    evo = NeuroEvolution(n_params=8, population_size=40, number_of_elements_to_mutate=20, seed=3, leniency_coef=0.0)
    leaders = []
    for _ in range(60):
        evo.step(evaluate)
        leaders.append(evo.population.max_score[0])
    assert calls[0] == 40 and set(calls[1:]) == {20}
    assert leaders[-1] > leaders[0] + 0.3 and leaders[-1] > 9.0
    assert evo.log[-1][0] == 60 * 20 and len(evo.log) == 60


def test_learning_configuration_files(tmp_path):
    """configuration::readFile semantics and a NeuroEvolution set up from such a file; the shipped TetraSpine node configuration when
    the reference tree is here."""
    from ntrtsim_b200.neuroevo import read_configuration, neuroevolution_from_configuration
    p = tmp_path / "Config.ini"
    p.write_text("#config file\n; also a comment\n\nlearning=1\n  populationSize = 10 \nnumberOfElementsToMutate=4\nleniencyCoef=0.25\n"
                 "compareAverageScores=1\n=nokey\npopulationSize=12\n")
    c = read_configuration(str(p))
    assert c == {"learning": "1", "populationSize": "12", "numberOfElementsToMutate": "4", "leniencyCoef": "0.25", "compareAverageScores": "1"}
    evo = neuroevolution_from_configuration(str(p), n_params=5)
    assert len(evo.population) == 12 and evo.num_mutate == 4 and evo.population.leniency == 0.25 and evo.population.compare_average_scores
    assert read_configuration(str(tmp_path / "missing.ini")) == {}
    shipped = "/root/reference/resources/src/learningSpines/TetraSpine/nodeConfig.ini"
    if os.path.exists(shipped):
        evo = neuroevolution_from_configuration(shipped, n_params=2)
        assert len(evo.population) == 2 and evo.num_mutate == 1 and evo.population.leniency == 1.0
