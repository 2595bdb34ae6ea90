"""Contact cables: tgBulletContactSpringCable / tgBasicContactCableInfo.

CPU part: the oracle's restatement against what the reference suite pins for this path --
test_integration/MuscleNP/MuscleNP_test.cpp:77-114 (ContactCableDemo, no gravity, no ground, 10 000 steps: linear momentum conserved,
energy not increased) -- and against the plain cable it must equal while it touches nothing.
GPU part: the CUDA path (step kernel + contact-cable kernel through the C ABI) against the oracle on the same seeded inputs.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ntrtsim_b200 as nt  # noqa: E402
from ntrtsim_b200 import _capi as capi  # noqa: E402
from ntrtsim_b200.world import FlatModel  # noqa: E402
import oracle_py  # noqa: E402
from helpers import rod_errors  # noqa: E402

DT = 0.001


def demo_model():
    return nt.build_example("ContactCableDemo")


def demo_states(model, n_env, seed=77, spread=0.15):
    """ContactCableDemo.cpp:104-110: the central impulse (0.4, 0, 0.4) on the first two bodies, per environment scaled and turned a
    little so that the cable meets the third rod at different times and angles."""
    rng = np.random.default_rng(seed)
    s = np.tile(model.initial_rod_state()[None], (n_env, 1, 1))
    for e in range(n_env):
        imp = np.array([0.4, 0.0, 0.4]) * (1.0 if e == 0 else 1.0 + spread * rng.uniform(-1, 1))
        if e > 0:
            imp = imp + spread * 0.4 * rng.uniform(-1, 1, size=3)
        for b in (0, 1):
            s[e, b, 7:10] = imp / model.bodies[b].mass
    return s


def masses(model):
    return np.array([b.mass for b in model.bodies])


def momentum(model, st):
    return (st[..., 7:10] * masses(model)[:, None]).sum(axis=-2)


def demo_energy(model, st):
    """ContactCableDemo::getEnergy (ContactCableDemo.cpp:184-198): sum of |v|^2 m over the rigids"""
    return ((st[..., 7:10] ** 2).sum(axis=-1) * masses(model)).sum(axis=-1)


# ------------------------------------------------------------------------------------------------ CPU: builder and oracle
def test_contact_cable_builder_matches_the_reference_construction():
    m = demo_model()
    assert (m.n_bodies, m.n_cables, m.n_contact_cables) == (3, 0, 1)
    c = m.contact_cables[0]
    assert (c.body_a, c.body_b) == (0, 1)
    # tgBulletContactSpringCable.h:89-90 defaults; rest = |p2 - p1| - pretension / k with pretension 0; prev = rest
    assert (c.thickness, c.resolution) == (0.001, 0.1)
    assert (c.stiffness, c.damping) == (1000.0, 0.0)
    # both ends on the surface of their rigid (tgBasicContactCableInfo.cpp:90-91 getConnectionPoint): radius 0.25 off the axis
    assert abs(np.hypot(c.local_a[0], c.local_a[2]) - 0.25) < 1e-12 and abs(np.hypot(c.local_b[0], c.local_b[2]) - 0.25) < 1e-12
    assert abs(c.rest_length - 8.0) < 1e-9 and c.prev_length == c.rest_length
    # the model tree holds one tgBasicActuator (tgBasicContactCableInfo::createModel) beside the three rigids
    assert nt.examples.tree_counts("ContactCableDemo")[1] == 1


def test_muscle_np_momentum_and_energy_like_the_reference_suite():
    """MuscleNP_test.cpp:77-114"""
    m = demo_model()
    w = oracle_py.OracleWorld(m, capi.WorldCfg)
    w.set_rod_state(demo_states(m, 1)[0])
    st0 = w.get_rod_state()
    p0, e0 = momentum(m, st0), demo_energy(m, st0)
    max_anchors, energies = 0, [e0]
    for _ in range(100):
        assert w.step(DT, 100) == 0
        max_anchors = max(max_anchors, int(w.get_contact_cable_obs()[0, 4]))
        energies.append(demo_energy(m, w.get_rod_state()))
    st = w.get_rod_state()
    p1 = momentum(m, st)
    # EXPECT_FLOAT_EQ on x and z (4 float ulps), EXPECT_NEAR 1e-10 on y
    assert abs(p1[0] - p0[0]) <= 4 * np.spacing(np.float32(p0[0])) and abs(p1[2] - p0[2]) <= 4 * np.spacing(np.float32(p0[2]))
    assert abs(p1[1] - p0[1]) < 1e-10
    # far tighter in fact: the anchor forces of a step sum to zero (tgBulletContactSpringCable.cpp:183-187)
    assert np.abs(p1 - p0).max() < 1e-12
    assert energies[-1] <= e0                                    # EXPECT_LE(energyEnd, energyStart)
    assert max(energies) <= e0 * (1 + 1e-3)                      # (the demo's measure is the translational part only: it may swing back up on the way, never beyond the start)
    assert max_anchors >= 2                                      # the cable did meet the third rod and bent around it
    assert st[2, 7:10].dot(st[2, 7:10]) > 1e-4                   # and set it moving


def test_muscle_np_energy_bound_holds_for_an_ensemble_of_launches():
    """EXPECT_LE(energyEnd, energyStart) (MuscleNP_test.cpp:113, the demo's own measure sum m |v|^2) for 24 perturbed launches"""
    m = demo_model()
    states = demo_states(m, 24, seed=5)
    for s0 in states:
        w = oracle_py.OracleWorld(m, capi.WorldCfg)
        w.set_rod_state(s0)
        w.step(DT, 10000)
        assert demo_energy(m, w.get_rod_state()) <= demo_energy(m, s0)
        assert np.abs(momentum(m, w.get_rod_state()) - momentum(m, s0)).max() < 1e-12


def test_sliding_anchors_sit_on_the_body_and_the_cable_pulls_not_pushes():
    m = demo_model()
    w = oracle_py.OracleWorld(m, capi.WorldCfg)
    w.set_rod_state(demo_states(m, 1)[0])
    seen = 0
    for _ in range(800):
        w.step(DT, 10)
        a = w.get_contact_cable_anchors(0)
        st = w.get_rod_state()
        n = int(w.get_contact_cable_obs()[0, 4])
        assert np.all(a[:n, 0] == 2) and np.all(a[n:, 0] == -1)   # only the free rod is ever touched; unused slots are marked
        if n == 0:
            continue
        seen += 1
        # distance of every anchor from the rod's axis segment = its radius (the capsule the restatement collides with), within the
        # sideways slides of a step
        from helpers import quat_rotate
        axis = quat_rotate(st[2, 3:7][None], np.array([[0.0, 1.0, 0.0]]))[0]
        for k in range(n):
            rel = a[k, 1:4] - st[2, 0:3]
            t = np.clip(rel.dot(axis), -0.75, 0.75)
            assert abs(np.linalg.norm(rel - t * axis) - 0.25) < 0.04   # the sideways slides move in the tangent plane: a few hundredths off the round cap
            assert abs(np.linalg.norm(a[k, 4:7]) - 1.0) < 1e-9
    assert seen > 50


def test_a_contact_cable_that_touches_nothing_is_a_plain_cable():
    """While no body is near, tgBulletContactSpringCable::calculateAndApplyForce (two anchors) is tgBulletSpringCable's:
    same tension, same clamped damping, same impulses.  (The slack branch differs in form only: force 0 either way.)"""
    m = demo_model()
    cc = m.contact_cables[0]
    cc.damping = 5.0
    cc.rest_length = cc.prev_length = 7.5   # pretension so that the cable works from the first step
    plain = (capi.CableDesc * 1)()
    p = plain[0]
    p.body_a, p.body_b = cc.body_a, cc.body_b
    for k in range(3):
        p.local_a[k], p.local_b[k] = cc.local_a[k], cc.local_b[k]
    p.stiffness, p.damping, p.rest_length, p.prev_length = cc.stiffness, cc.damping, cc.rest_length, cc.prev_length
    p.actuator = capi.TGB_ACT_PASSIVE
    p.max_tension, p.target_velocity, p.min_actual_length, p.min_rest_length = 1e9, 100.0, 0.1, 0.1
    p.motor_radius, p.motor_inertia = 1.0, 1.0
    m_plain = FlatModel(m.bodies, m.shapes, plain, m.ground, m.gravity, m.dt, "plain")
    wa, wb = oracle_py.OracleWorld(m, capi.WorldCfg), oracle_py.OracleWorld(m_plain, capi.WorldCfg)
    s = demo_states(m, 1)[0]
    s[:, 7:10] *= 0.1     # slowly: the cable stays clear of the third rod for the whole run
    wa.set_rod_state(s)
    wb.set_rod_state(s)
    for _ in range(30):   # 300 steps: the pretension pulls the two ends together; they are still 3 units apart then
        wa.step(DT, 10)
        wb.step(DT, 10)
        assert int(wa.get_contact_cable_obs()[0, 4]) == 0
        assert np.array_equal(wa.get_rod_state(), wb.get_rod_state())
        assert np.allclose(wa.get_contact_cable_obs()[0, :3], wb.get_cable_obs()[0, :3], rtol=0, atol=1e-12)


def total_mechanical_energy(model, st, obs, k):
    """translation + rotation of the rigids + the cable's spring energy 1/2 k stretch^2 over its whole path"""
    from helpers import quat_rotate
    inertia = np.array([list(b.inertia) for b in model.bodies])
    ke = 0.5 * (masses(model) * (st[:, 7:10] ** 2).sum(axis=1)).sum()
    w_body = quat_rotate(st[:, 3:7] * np.array([-1.0, -1.0, -1.0, 1.0]), st[:, 10:13])
    kr = 0.5 * (inertia * w_body ** 2).sum()
    stretch = max(0.0, obs[0, 1] - obs[0, 0])
    return ke + kr + 0.5 * k * stretch * stretch


def test_a_cable_sliding_around_a_body_does_no_work():
    """Sliding anchors are frictionless and the cable has no damping here: the anchor forces T (unit(forward) + unit(back)) push
    along the bisector of the bend, and what the cable takes from the bodies' motion it holds as stretch.  Through the whole
    encounter -- anchors created, slid, removed -- the mechanical energy of bodies + cable stays within a few percent (the residue
    is the path-length jump of an insertion, capped at 2 x resolution), for every member of an ensemble of launches."""
    m = demo_model()
    states = demo_states(m, 16, seed=5)
    ratios = []
    for e in range(16):
        w = oracle_py.OracleWorld(m, capi.WorldCfg)
        w.set_rod_state(states[e])
        e0 = total_mechanical_energy(m, w.get_rod_state(), w.get_contact_cable_obs(), 1000.0)
        n_max = 0
        for _ in range(100):
            w.step(DT, 100)
            n_max = max(n_max, int(w.get_contact_cable_obs()[0, 4]))
        assert n_max >= 2
        ratios.append(total_mechanical_energy(m, w.get_rod_state(), w.get_contact_cable_obs(), 1000.0) / e0)
    ratios = np.array(ratios)
    assert np.abs(ratios - 1.0).max() < 0.05, ratios      # measured: worst +3.3 %
    assert abs(np.median(ratios) - 1.0) < 0.005           # measured: median +0.15 %


def test_control_input_on_a_contact_cable_is_the_basic_actuators():
    """tgBasicContactCableInfo::createModel wraps the contact cable in a tgBasicActuator: setControlInput(input, dt) -> moveMotors uses the
    cable's actual length (tgBasicActuator.cpp:149-225).  While the cable touches nothing that is a plain tgBasicActuator cable."""
    m = demo_model()
    cc = m.contact_cables[0]
    assert (cc.max_tension, cc.target_velocity, cc.min_actual_length, cc.min_rest_length) == (600000000.0, 100.0, 0.1, 0.1)
    cc.target_velocity = 2.0
    plain = (capi.CableDesc * 1)()
    p = plain[0]
    p.body_a, p.body_b = cc.body_a, cc.body_b
    for k in range(3):
        p.local_a[k], p.local_b[k] = cc.local_a[k], cc.local_b[k]
    p.stiffness, p.damping, p.rest_length, p.prev_length = cc.stiffness, cc.damping, cc.rest_length, cc.prev_length
    p.actuator = capi.TGB_ACT_BASIC
    p.max_tension, p.target_velocity, p.min_actual_length, p.min_rest_length = cc.max_tension, cc.target_velocity, 0.1, 0.1
    p.motor_radius, p.motor_inertia = 1.0, 1.0
    m_plain = FlatModel(m.bodies, m.shapes, plain, m.ground, m.gravity, m.dt, "plain")
    wa, wb = oracle_py.OracleWorld(m, capi.WorldCfg), oracle_py.OracleWorld(m_plain, capi.WorldCfg)
    s = demo_states(m, 1)[0]
    s[:, 7:10] *= 0.1
    wa.set_rod_state(s)
    wb.set_rod_state(s)
    for i in range(200):
        target = 7.8 if i < 150 else 8.1      # reel in at 2 units / s (reached after 100 steps), then pay out for 50
        assert wa.set_contact_cable_control_input(np.array([target]), DT) == 0
        assert wb.set_control_input(np.array([target]), DT) == 0
        wa.step(DT, 1)
        wb.step(DT, 1)
    assert int(wa.get_contact_cable_obs()[0, 4]) == 0
    assert np.array_equal(wa.get_rod_state(), wb.get_rod_state())
    assert np.allclose(wa.get_contact_cable_obs()[0, :3], wb.get_cable_obs()[0, :3], rtol=0, atol=1e-12)
    assert abs(wa.get_contact_cable_obs()[0, 0] - 8.0) > 0.05   # the rest length did move
    assert wa.set_contact_cable_control_input(np.array([-1.0]), DT) != 0   # "Rest length is negative."


def test_contact_cable_descriptor_checks():
    m = demo_model()
    lib = oracle_py.load()
    for field, bad in (("stiffness", 0.0), ("damping", -1.0), ("resolution", 0.0), ("thickness", -1.0), ("body_b", 7)):  # the oracle's checks
        mm = demo_model()
        setattr(mm.contact_cables[0], field, bad)
        with pytest.raises(ValueError):
            oracle_py.OracleWorld(mm, capi.WorldCfg)
    assert lib.tgo_num_contact_cables is not None and m.n_contact_cables == 1


def test_yaml_front_end_builds_contact_cables(tmp_path):
    """TensegrityModel.cpp:609-611,731-734: tgBasicContactCableInfo shares tgBasicActuatorInfo's parameters"""
    text = """
nodes:
  a: [0, 2, 0]
  b: [0, 4, 0]
  c: [0, 12, 0]
  d: [0, 14, 0]
pair_groups:
  rod:
    - [a, b]
    - [c, d]
  wire:
    - [b, c]
builders:
  rod:
    class: tgRodInfo
    parameters:
      density: 1
      radius: 0.25
  wire:
    class: tgBasicContactCableInfo
    parameters:
      stiffness: 500
      damping: 2
      pretension: 100
"""
    path = os.path.join(str(tmp_path), "cc.yaml")
    open(path, "w").write(text)
    m = nt.examples.build_yaml(path)
    assert (m.n_bodies, m.n_cables, m.n_contact_cables) == (2, 0, 1)
    c = m.contact_cables[0]
    assert (c.stiffness, c.damping) == (500.0, 2.0)
    assert abs(c.rest_length - (8.0 - 100.0 / 500.0)) < 1e-9


PRISM_WITH_CONTACT_CABLES = """
nodes:
  bottom1: [-5, 0, 0]
  bottom2: [5, 0, 0]
  bottom3: [0, 0, 8.66]
  top1: [-5, 5, 0]
  top2: [5, 5, 0]
  top3: [0, 5, 8.66]
pair_groups:
  prism_rod:
    - [bottom1, top2]
    - [bottom2, top3]
    - [bottom3, top1]
  horizontal_string:
    - [bottom1, bottom2]
    - [bottom2, bottom3]
    - [bottom1, bottom3]
    - [top1, top2]
    - [top2, top3]
    - [top1, top3]
  vertical_string:
    - [bottom1, top1]
    - [bottom2, top2]
    - [bottom3, top3]
builders:
  prism_rod:
    class: tgRodInfo
    parameters:
      density: 0.688
      radius: 0.31
  horizontal_string:
    class: tgBasicActuatorInfo
    parameters:
      stiffness: 1000
      damping: 10
      pretension: 3000
  vertical_string:
    class: tgBasicContactCableInfo
    parameters:
      stiffness: 1000
      damping: 10
      pretension: 1000
"""


def prism_with_contact_cables(tmp_path):
    """resources/YamlStructures/BaseStructures/3Prism.yaml with its three vertical strings built as contact cables: two-anchor cables,
    contact cables, gravity, the ground and contact between the rods in one world."""
    path = os.path.join(str(tmp_path), "3PrismContact.yaml")
    open(path, "w").write(PRISM_WITH_CONTACT_CABLES)
    return nt.examples.build_yaml(path)


def test_contact_cables_beside_plain_cables_in_one_model(tmp_path):
    m = prism_with_contact_cables(tmp_path)
    assert (m.n_bodies, m.n_cables, m.n_contact_cables) == (3, 6, 3)
    w = oracle_py.OracleWorld(m, capi.WorldCfg)
    s = m.initial_rod_state()
    s[:, 1] += 3.0
    w.set_rod_state(s)
    seen = 0
    for _ in range(60):
        assert w.step(DT, 10) == 0
        o = w.get_contact_cable_obs()
        assert np.all(np.isfinite(w.get_rod_state())) and np.all(o[:, 2] >= 0.0)
        seen = max(seen, int(o[:, 4].max()))
    # it fell onto the ground and stays above it, in one piece
    y = w.get_rod_state()[:, 1]
    assert y.max() < 5.4 and y.min() > 0.3 and np.abs(w.get_rod_state()[:, 7:13]).max() < 50.0
    print("most sliding anchors on a cable:", seen)


# ------------------------------------------------------------------------------------------------ the reference's own cable code
def test_oracle_reproduces_the_reference_cable_code_golden():
    """tests/golden/contact_cable_demo.npz = checkpoints of hybrid runs in which the REFERENCE's own src/core/tgBulletContactSpringCable.cpp
    and tgBulletSpringCableAnchor.cpp (compiled in place, oracle/ref/build_ref.py) do the cable's step -- updateManifolds,
    updateAnchorList with its findNearestPastAnchor, pruneAnchors, calculateAndApplyForce -- on the contact points of the oracle's
    geometric manifold, the oracle's world step between them (oracle/ref/cc_hybrid.py).  The pure oracle has to reproduce them: rod
    states, cable length and tension, and the sliding anchors themselves (body, place, normal)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "contact_cable_demo.npz"))
    m = demo_model()
    chunk = int(g["chunk"])
    touched = 0
    for e in range(g["states"].shape[0]):
        w = oracle_py.OracleWorld(m, capi.WorldCfg)
        w.set_rod_state(g["states"][e])
        for k in range(g["rods"].shape[1]):
            w.step(DT, chunk)
            assert np.abs(w.get_rod_state() - g["rods"][e, k]).max() < 1e-7
            o = w.get_contact_cable_obs()[0]
            assert abs(o[1] - g["cable"][e, k, 1]) < 1e-7 and abs(o[2] - g["cable"][e, k, 2]) < 1e-4
            assert int(o[4]) + 2 == int(g["cable"][e, k, 3])
            a = w.get_contact_cable_anchors(0)
            assert np.array_equal(a[:, 0], g["anchors"][e, k, :, 0])
            assert np.abs(a[:, 1:] - g["anchors"][e, k, :, 1:]).max() < 1e-6
            touched = max(touched, int(o[4]))
    assert touched >= 3


def test_reference_cable_code_live_against_the_oracle():
    """the same comparison live, where oracle/_ref/libntrt_ref.so exists (this container: built from /root/reference; the GPU box: it
    travels prebuilt)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref"))
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libntrt_ref.so not built (no /root/reference here)")
    import cc_hybrid
    if not hasattr(ref_py.load(), "ref_add_contact_cable"):
        pytest.skip("oracle/_ref/libntrt_ref.so predates the contact-cable harness")
    for e, worst, same, checks, nmax in cc_hybrid.compare(n_env=3, steps=6000, chunk=500, seed=123, verbose=False):
        assert worst < 1e-6 and same == checks, (e, worst, same, checks)
    # with a controller: tgBasicActuator::setControlInput(input, dt) before every step reels the cable in while it bends around the rod
    model = demo_model()
    rods = demo_states(model, 2, seed=11)
    rods[:, :2, 7:10] *= 4.0   # reach the rod within a thousand steps
    model.contact_cables[0].target_velocity = 2.0
    for e in range(2):
        hyb = cc_hybrid.HybridWorld(model, rods[e])
        ora = oracle_py.OracleWorld(model, capi.WorldCfg)
        ora.set_rod_state(rods[e])
        most = 0
        for k in range(1800):
            target = np.array([8.0 - 0.0008 * k])
            hyb.step(1, control=target)
            ora.world_step(DT)
            assert ora.set_contact_cable_control_input(target, DT) == 0
            ora.models_step(DT)
            if k % 100 == 99:
                assert np.abs(hyb.rod_state() - ora.get_rod_state()).max() < 1e-7
                o = ora.get_contact_cable_obs()[0]
                assert abs(hyb.cable()[0] - o[0]) < 1e-12 and hyb.anchors().shape[0] == int(o[4])
                most = max(most, int(o[4]))
        assert most >= 2
        hyb.close()


# ------------------------------------------------------------------------------------------------ GPU
def _device_run(model, states, precision, steps, chunk):
    w = nt.BatchedWorld(model, states.shape[0], precision=precision)
    try:
        w.set_rod_state(states)
        out = []
        for _ in range(steps // chunk):
            w.step(DT, chunk)
            w.sync()
            out.append((w.get_rod_state(), w.get_contact_cable_obs(), w.get_contact_cable_anchors(0)))
        return out
    finally:
        w.close()


def _oracle_run(model, states, steps, chunk):
    worlds = [oracle_py.OracleWorld(model, capi.WorldCfg) for _ in range(states.shape[0])]
    for w, s in zip(worlds, states):
        w.set_rod_state(s)
    out = []
    for _ in range(steps // chunk):
        for w in worlds:
            w.step(DT, chunk)
        out.append((np.stack([w.get_rod_state() for w in worlds]), np.stack([w.get_contact_cable_obs() for w in worlds]),
                    np.stack([w.get_contact_cable_anchors(0) for w in worlds])))
    return out


@pytest.mark.gpu
def test_contact_cable_demo_f64_matches_the_oracle():
    """Double precision, 32 environments, 7000 steps (the cable meets the rod from step ~3500 on and bends around it): rod states, cable
    length / tension and the anchor lists themselves, checkpoint by checkpoint.  Anchor insertions and removals amplify round-off, so the
    yardstick after first contact is the oracle's own sensitivity to a 1e-13 perturbation of its input (tools/cc_probe.py; measured on
    B200: device-vs-oracle 2e-11 ... 2.6e-6 beside a self-sensitivity of 2e-11 ... 4.4e-6, identical anchor lists throughout)."""
    m = demo_model()
    states = demo_states(m, 32)
    dev = _device_run(m, states, "f64", 7000, 500)
    ora = _oracle_run(m, states, 7000, 500)
    pert = states.copy()
    pert[:, :, 0:3] += 1e-13 * np.random.default_rng(0).normal(size=pert[:, :, 0:3].shape)
    orp = _oracle_run(m, pert, 7000, 500)
    touched = 0
    for (rd, od, ad), (ro, oo, ao), (rp, _, _) in zip(dev, ora, orp):
        sens = max(rod_errors(rp, ro, 10.0))
        tol = max(1e-9, 20.0 * sens)
        assert tol < 2e-4
        assert max(rod_errors(rd, ro, 10.0)) < tol
        assert np.array_equal(od[..., 4], oo[..., 4])                   # the same number of sliding anchors everywhere
        assert np.abs(od[..., :2] - oo[..., :2]).max() < tol            # rest, length
        assert np.abs(od[..., 2] - oo[..., 2]).max() < 1000.0 * tol     # tension = 1000 x stretch
        assert np.array_equal(ad[..., 0], ao[..., 0])                   # on the same bodies
        assert np.abs(ad[..., 1:] - ao[..., 1:]).max() < 100.0 * tol    # at the same places with the same normals
        touched = max(touched, int((oo[..., 4] > 0).sum()))
    assert touched >= 16


@pytest.mark.gpu
def test_contact_cable_demo_f32_conserves_momentum_and_follows_the_oracle():
    m = demo_model()
    states = demo_states(m, 256, seed=5)
    dev = _device_run(m, states, "f32", 6000, 1000)
    p0 = momentum(m, states)
    e0 = demo_energy(m, states)
    rd, od, ad = dev[-1]
    # MuscleNP_test.cpp:108-112 on the device: single precision keeps the momentum to float round-off over 6000 steps (measured 1.8e-6)
    assert np.abs(momentum(m, rd) - p0).max() < 2e-5
    # EXPECT_LE(energyEnd, energyStart), MuscleNP_test.cpp:113, for every member of the ensemble (single precision: within 1e-3)
    assert np.all(demo_energy(m, rd) <= e0 * (1 + 1e-3))
    assert max((o[1][..., 4] > 0).sum() for o in dev) >= 64
    # Trajectories: this demo's cable starts exactly at its rest length, so its tension (k x 1e-6 of stretch on a length of 8) sits at
    # the resolution of a float position and single precision follows the oracle only statistically.  With the cable pretensioned
    # (rest 7.5: 500 units of tension) the comparison is the stated FP32 one: 1e-4 of the structure size / of the tension.
    mp = demo_model()
    mp.contact_cables[0].rest_length = mp.contact_cables[0].prev_length = 7.5
    sp = demo_states(mp, 32, seed=6)
    d = _device_run(mp, sp, "f32", 300, 100)
    o = _oracle_run(mp, sp, 300, 100)
    for (rd, od, _), (ro, oo, _) in zip(d, o):
        e = rod_errors(rd, ro, 10.0)
        assert e[0] < 1e-4 and e[1] < 1e-4 and e[2] < 1e-3 and e[3] < 1e-3   # velocities reach ~10 and ~1 here
        assert (np.abs(od[..., 2] - oo[..., 2]) / np.maximum(oo[..., 2], 100.0)).max() < 1e-3
        assert np.array_equal(od[..., 4], oo[..., 4])


@pytest.mark.gpu
def test_contact_cable_control_input_on_the_device_matches_the_oracle():
    """A host-side controller reeling the contact cable in while it bends around the rod: tgb_set_contact_cable_control_input every
    step for 1500 steps (f64), against the oracle."""
    m = demo_model()
    states = demo_states(m, 8, seed=11)
    states[:, :2, 7:10] *= 4.0
    w = nt.BatchedWorld(m, 8, precision="f64")
    worlds = [oracle_py.OracleWorld(m, capi.WorldCfg) for _ in range(8)]
    try:
        w.set_rod_state(states)
        for o, s in zip(worlds, states):
            o.set_rod_state(s)
        for i in range(1500):
            target = np.array([8.0 - 0.0005 * i])
            w.set_contact_cable_control_input(target, DT)
            w.step(DT, 1)
            for o in worlds:
                assert o.set_contact_cable_control_input(target, DT) == 0
                o.step(DT, 1)
        w.sync()
        rd, od = w.get_rod_state(), w.get_contact_cable_obs()
        ro = np.stack([o.get_rod_state() for o in worlds])
        oo = np.stack([o.get_contact_cable_obs() for o in worlds])
        assert (oo[..., 4] > 0).any()
        assert np.array_equal(od[..., 4], oo[..., 4])
        assert np.abs(od[..., 0] - oo[..., 0]).max() < 1e-9 and abs(oo[0, 0, 0] - 8.0) > 0.5
        assert max(rod_errors(rd, ro, 10.0)) < 1e-6
        with pytest.raises(capi.TgbError):
            w.set_contact_cable_control_input(np.array([-1.0]), DT)
    finally:
        w.close()


@pytest.mark.gpu
def test_prism_with_contact_cables_f64_matches_the_oracle(tmp_path):
    """plain cables + contact cables + gravity + ground + rod-rod contact in one world, 16 random poses, 400 steps"""
    from helpers import random_poses
    m = prism_with_contact_cables(tmp_path)
    states = random_poses(m, 16, seed=21, translate=1.0, lift=(3.0, 6.0))
    w = nt.BatchedWorld(m, 16, precision="f64")
    worlds = [oracle_py.OracleWorld(m, capi.WorldCfg) for _ in range(16)]
    pert = [oracle_py.OracleWorld(m, capi.WorldCfg) for _ in range(16)]
    try:
        w.set_rod_state(states)
        noise = 1e-13 * np.random.default_rng(1).normal(size=states[:, :, 0:3].shape)
        for o, p, s, dn in zip(worlds, pert, states, noise):
            o.set_rod_state(s)
            s2 = s.copy()
            s2[:, 0:3] += dn
            p.set_rod_state(s2)
        anchors, compared = 0, 0
        for _ in range(16):
            w.step(DT, 25)
            w.sync()
            for o, p in zip(worlds, pert):
                o.step(DT, 25)
                p.step(DT, 25)
            rd, cd, od = w.get_rod_state(), w.get_cable_obs(), w.get_contact_cable_obs()
            ro = np.stack([o.get_rod_state() for o in worlds])
            rp = np.stack([o.get_rod_state() for o in pert])
            co = np.stack([o.get_cable_obs() for o in worlds])
            oo = np.stack([o.get_contact_cable_obs() for o in worlds])
            tol = max(1e-9, 20.0 * max(rod_errors(rp, ro, 10.0)))
            if tol > 1e-5:
                break   # landing on the ground with rods and cables touching amplifies a 1e-13 perturbation beyond use from here on
            compared += 1
            assert max(rod_errors(rd, ro, 10.0)) < tol
            assert np.abs(cd[..., :2] - co[..., :2]).max() < tol and np.abs(od[..., :2] - oo[..., :2]).max() < tol
            assert (od[..., 4] != oo[..., 4]).sum() <= (1 if tol > 1e-7 else 0)
            anchors = max(anchors, int(oo[..., 4].max()))
        print("checkpoints compared:", compared, "most sliding anchors on a cable:", anchors)
        assert compared >= 4
    finally:
        w.close()


@pytest.mark.gpu
def test_contact_cable_world_steps_through_the_halves_and_rollout_host():
    """tgb_step(w, dt, n) == n x (tgb_step_world, tgb_step_models) == tgb_rollout_host for a world with a contact cable"""
    m = demo_model()
    states = demo_states(m, 8, seed=9)
    states[:, :2, 7:10] *= 4.0   # reach the rod within 1200 steps
    a = nt.BatchedWorld(m, 8, precision="f64")
    b = nt.BatchedWorld(m, 8, precision="f64")
    c = nt.BatchedWorld(m, 8, precision="f64")
    try:
        a.set_rod_state(states)
        b.set_rod_state(states)
        a.step(DT, 1200)
        for _ in range(1200):
            b.step_world(DT)
            b.step_models(DT)
        a.sync()
        b.sync()
        assert np.array_equal(a.get_rod_state(), b.get_rod_state())
        assert np.array_equal(a.get_contact_cable_obs(), b.get_contact_cable_obs())
        assert (a.get_contact_cable_obs()[..., 4] > 0).any()
        rods_out, _ = c.rollout_host(states, np.zeros((8, 0, 2)), DT, 1200)
        assert np.array_equal(rods_out, a.get_rod_state())
        # a new episode on the same world: bodies back, cables afresh (no anchors, previous length = rest) -> the same trajectory
        a.set_rod_state(states)
        a.reset_contact_cables()
        assert (a.get_contact_cable_obs()[..., 4] == 0).all()
        a.step(DT, 1200)
        a.sync()
        assert np.array_equal(a.get_rod_state(), rods_out)
    finally:
        a.close()
        b.close()
        c.close()


@pytest.mark.gpu
def test_host_api_drop_in_runs_a_contact_cable_model_on_the_device(tmp_path):
    """tgSimulation::run of ContactCableDemo through the host mirror (tgWorldB200Impl finalises with the contact cables; the
    tgBasicActuator of the cable reads its length and tension through the contact-cable observations), and the tgDataLogger2 line of
    the same run."""
    h = capi.load_host()
    m = demo_model()
    n_env, steps = 3, 60
    rod = np.empty((n_env, m.n_bodies, 13))
    cab = np.empty((n_env, max(1, m.n_cables), 4))
    for fn in (h.tgh_run_example, h.tgh_run_example_stepwise):
        rc = fn(b"ContactCableDemo", n_env, capi.TGB_F64, steps, rod.ctypes.data_as(C.c_void_p), cab.ctypes.data_as(C.c_void_p))
        assert rc == 0, h.tgh_last_error()
        w = nt.BatchedWorld(m, n_env, precision="f64")
        try:
            w.step(DT, steps)
            w.sync()
            assert np.array_equal(w.get_rod_state(), rod)
        finally:
            w.close()
    buf = C.create_string_buffer(512)
    rc = h.tgh_run_example_logged(b"ContactCableDemo", n_env, capi.TGB_F64, 20, str(tmp_path / "cc").encode(), 0.0, 1, buf, 512)
    assert rc == 0, h.tgh_last_error()
    lines = [ln for ln in open(buf.value.decode()).read().split("\n") if ln]
    assert len(lines) == 2 + 20
    header = lines[1].split(",")
    cols = [i for i, c in enumerate(header) if "RestLen" in c or "CurrLen" in c or "Tension" in c]
    assert len(cols) >= 3, header
    last = lines[-1].split(",")
    vals = {header[i]: float(last[i]) for i in cols}
    rest = [v for k, v in vals.items() if "RestLen" in k][0]
    curr = [v for k, v in vals.items() if "CurrLen" in k][0]
    assert abs(rest - 8.0) < 1e-9 and abs(curr - 8.0) < 1e-6
