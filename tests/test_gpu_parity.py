"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded inputs.

Tolerances.  The double-precision instantiation evaluates the oracle's expression tree without FMA contraction, so it
is held to round-off level (1e-9 on positions that reach a few hundred units).  The single-precision instantiation is
the throughput configuration; BASELINE.json's north star states "within a stated FP32 tolerance (e.g. 1e-4 relative)":
cable tensions are held to 1e-4 of max(|T|, typical pretension level) and positions to 1e-4 of the structure size.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ntrtsim_b200 as nt  # noqa: E402
from ntrtsim_b200 import _capi as capi  # noqa: E402
import oracle_py  # noqa: E402
from helpers import random_poses, rod_errors, tension_rel_error, align_quat_sign  # noqa: E402

pytestmark = pytest.mark.gpu

F32_REL = 1e-4


def oracle_rollout(model, rods, cab, steps, ctrl_kind=0, params=None, dt=None, setup_input=None, threads=8):
    return oracle_py.rollout_batch(model, capi.WorldCfg, rods, cab, dt or model.dt, steps, ctrl_kind, params,
                                   setup_input=setup_input, setup_dt=dt or model.dt, n_threads=threads)


def device_rollout(model, rods, cab, steps, precision, ctrl_kind=0, params=None, dt=None, setup_input=None, chunks=1, **kw):
    n_env = rods.shape[0]
    w = nt.BatchedWorld(model, n_env, precision=precision, **kw)
    try:
        w.set_rod_state(rods)
        w.set_cable_state(cab)
        if setup_input is not None:
            w.set_control_input(setup_input, dt or model.dt)
        if ctrl_kind:
            w.set_controller(ctrl_kind, params)
        per = steps // chunks
        for i in range(chunks):
            w.step(dt or model.dt, per if i < chunks - 1 else steps - per * (chunks - 1))
        w.sync()
        return w.get_rod_state(), w.get_cable_obs(), w.status()
    finally:
        w.close()


def oracle_self_sensitivity(model, rods, cab, steps, ctrl_kind=0, params=None, eps=1e-13, seed=0):
    """How far the ORACLE moves when its own input moves by eps: the yardstick for device-vs-oracle agreement on trajectories that
    amplify round-off (manifold points switching on and off).  Returns (unperturbed rods, unperturbed cables, max rod deviation)."""
    ro, co = oracle_rollout(model, rods, cab, steps, ctrl_kind, params)
    pert = rods.copy()
    pert[:, :, 0:3] += eps * np.random.default_rng(seed).normal(size=pert[:, :, 0:3].shape)
    rp, _ = oracle_rollout(model, pert, cab, steps, ctrl_kind, params)
    return ro, co, max(rod_errors(rp, ro, 1.0))


def tiled_cables(model, n_env):
    return np.tile(model.initial_cable_state(), (n_env, 1, 1))


# ------------------------------------------------------------------------------------------------ contact-free
@pytest.mark.parametrize("steps", [1, 2, 200])
def test_prism_drop_f64_roundoff(steps):
    """config 2 (3-prism, no ground, random poses), double precision: round-off level agreement."""
    model = nt.build_example("3_prism_noground")
    n_env = 67  # not a multiple of the envs per CTA: exercises the tail groups
    rods, cab = random_poses(model, n_env, seed=1234), tiled_cables(model, n_env)
    rd, cd, st = device_rollout(model, rods, cab, steps, "f64")
    ro, co = oracle_rollout(model, rods, cab, steps)
    e = rod_errors(rd, ro, 1.0)
    assert max(e) < 1e-9, e
    assert np.abs(cd - co).max() < 1e-7, np.abs(cd - co).max()
    assert not st.any()


def test_prism_first_step_damping_artefact():
    """prevLength starts at the rest length (tgSpringCable.cpp:59): the first applied force is 2 k stretch. After one
    step the observable is the velocity change of the rods, compared with the oracle in double."""
    model = nt.build_example("3_prism_noground")
    rods, cab = model.initial_rod_state()[None], tiled_cables(model, 1)
    rd, cd, _ = device_rollout(model, rods, cab, 1, "f64")
    ro, co = oracle_rollout(model, rods, cab, 1)
    assert np.abs(rd - align_quat_sign(ro, rd)).max() < 1e-11
    # tension reported without the damping term: k * pretension/k = 500 on every cable
    assert np.allclose(cd[0, :, 2], 500.0, atol=1e-6)


def test_prism_drop_f32_full_horizon():
    """config 2, single precision, full horizon H = 1000 steps (each env falls ~490 units): stated FP32 tolerance."""
    model = nt.build_example("3_prism_noground")
    n_env, steps = 4096, 1000   # BASELINE config 2 at its full size
    rods, cab = random_poses(model, n_env, seed=1234), tiled_cables(model, n_env)
    rd, cd, st = device_rollout(model, rods, cab, steps, "f32", chunks=4)
    ro, co = oracle_rollout(model, rods, cab, steps)
    # positions relative to the structure (its size is ~20 units): compare after removing the common fall
    terr = tension_rel_error(cd, co, floor=100.0)
    lerr = np.abs(cd[..., 1] - co[..., 1]).max() / 20.0
    rerr = np.abs(cd[..., 0] - co[..., 0]).max() / 20.0
    perr = rod_errors(rd, ro, 20.0)
    print("f32 full horizon: tension rel %.3e, length rel %.3e, rest %.3e, rods %s" % (terr, lerr, rerr, perr))
    assert terr < F32_REL and lerr < F32_REL and rerr < F32_REL
    assert perr[0] < F32_REL and perr[1] < F32_REL
    assert not st.any()


def test_prism_per_step_trace_f32():
    """Per-step cable tensions / rest lengths / rod states over a short horizon from identical initial conditions."""
    model = nt.build_example("3_prism_noground")
    n_env, steps = 32, 60
    rods, cab = random_poses(model, n_env, seed=99), tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    ow = [oracle_py.OracleWorld(model, capi.WorldCfg) for _ in range(n_env)]
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    for e in range(n_env):
        ow[e].set_rod_state(rods[e])
        ow[e].set_cable_state(cab[e])
    worst_t = worst_p = 0.0
    for s in range(steps):
        w.step(model.dt, 1)
        cd, rd = w.get_cable_obs(), w.get_rod_state()
        for e in range(n_env):
            ow[e].step(model.dt, 1)
        co = np.stack([o.get_cable_obs() for o in ow])
        ro = np.stack([o.get_rod_state() for o in ow])
        worst_t = max(worst_t, tension_rel_error(cd, co, floor=100.0))
        worst_p = max(worst_p, rod_errors(rd, ro, 20.0)[0])
    w.close()
    print("per-step trace: worst tension rel %.3e, worst position rel %.3e" % (worst_t, worst_p))
    assert worst_t < F32_REL and worst_p < F32_REL


# ------------------------------------------------------------------------------------------------ contact + control
def escape_inputs(model, n_env, seed):
    rods = random_poses(model, n_env, seed=seed, translate=1.0, lift=(0.0, 5.0), yaw_only=True)
    rng = np.random.default_rng(seed + 1)
    params = nt.escape_transform_actions(rng.uniform(size=(n_env, 32)), 9.0)
    return rods, tiled_cables(model, n_env), params


@pytest.mark.parametrize("steps", [1, 50, 400])
def test_superball_sine_contact_f64(steps):
    """config 3 (24-cable SUPERball, sine controller, box ground): the double path follows the oracle through contact."""
    model = nt.build_example("Escape_c3")
    n_env = 48
    rods, cab, params = escape_inputs(model, n_env, seed=2345)
    rd, cd, st = device_rollout(model, rods, cab, steps, "f64", capi.TGB_CTRL_ESCAPE_SINE, params)
    ro, co, own = oracle_self_sensitivity(model, rods, cab, steps, capi.TGB_CTRL_ESCAPE_SINE, params)
    e = rod_errors(rd, ro, 1.0)
    print("superball f64 %d steps: rods %s cables %.3e; oracle vs its 1e-13-perturbed self %.3e" % (steps, e, np.abs(cd - co).max(), own))
    # round-off level while the contacts are young; at 400 steps (rods also touch each other by then) no further from the oracle than
    # 20 x what the oracle does to itself under a 1e-13 perturbation
    tol = 1e-7 if steps <= 50 else max(1e-7, 20.0 * own)
    assert max(e) < tol, (e, own)
    assert tension_rel_error(cd, co, floor=10.0) < tol


def test_superball_sine_contact_f32_short_horizon():
    """config 3, single precision, short horizon (200 steps): stated FP32 tolerance.  Contact makes the
    trajectory sensitive (a contact point appearing one step earlier changes the impulses), so the bound is held on the
    bulk of the batch: at least 99% of the environments within tolerance, and every environment within 1e-2."""
    model = nt.build_example("Escape_c3")
    n_env, steps = 256, 200
    rods, cab, params = escape_inputs(model, n_env, seed=2345)
    rd, cd, st = device_rollout(model, rods, cab, steps, "f32", capi.TGB_CTRL_ESCAPE_SINE, params)
    ro, co = oracle_rollout(model, rods, cab, steps, capi.TGB_CTRL_ESCAPE_SINE, params)
    per_env_t = (np.abs(cd[..., 2] - co[..., 2]) / np.maximum(np.abs(co[..., 2]), 100.0)).max(axis=1)
    rdl = align_quat_sign(rd, ro)
    per_env_p = np.abs(rdl[..., 0:3] - ro[..., 0:3]).max(axis=(1, 2)) / 17.0
    frac_ok = np.mean((per_env_t < F32_REL) & (per_env_p < F32_REL))
    print("superball f32: frac within 1e-4 = %.4f, worst tension %.3e, worst pos %.3e, median tension %.3e"
          % (frac_ok, per_env_t.max(), per_env_p.max(), np.median(per_env_t)))
    assert frac_ok >= 0.99
    assert per_env_t.max() < 1e-2 and per_env_p.max() < 1e-2
    assert not st.any()


def test_t6_25_cables_tension_controller_f64():
    """Shipped SUPERball (25 cables incl. the duplicated pair) on the pitched ground with T6TensionController."""
    model = nt.build_example("SUPERball")
    n_env, steps = 16, 300
    rods = random_poses(model, n_env, seed=5, translate=0.5, lift=(0.0, 2.0), yaw_only=True)
    cab = tiled_cables(model, n_env)
    params = np.full((n_env, 1), 1000.0)  # AppSUPERball.cpp: T6TensionController(1000)
    rd, cd, _ = device_rollout(model, rods, cab, steps, "f64", capi.TGB_CTRL_TENSION, params)
    ro, co = oracle_rollout(model, rods, cab, steps, capi.TGB_CTRL_TENSION, params)
    assert max(rod_errors(rd, ro, 1.0)) < 1e-7
    assert np.abs(cd - co).max() < 1e-6


def test_prism_on_ground_f64():
    """config 1 world (3-prism over the box ground) replicated: drop, impact and rest."""
    model = nt.build_example("3_prism")
    n_env, steps = 8, 600
    rods = random_poses(model, n_env, seed=11, translate=0.2, rotate=False)
    cab = tiled_cables(model, n_env)
    # the rods of the prism meet at step ~181 (tgb_pairs.cuh); from then on rod-rod and rod-ground manifolds switch points and the
    # trajectory amplifies round-off: a 1e-13 perturbation of the ORACLE's own input grows to 1e-9 by step 200 and to 5e-5 by step 300
    # (measured, tools/pair_contact_census.py's scenario).  Tight agreement is therefore required up to step 220, and at the full
    # horizon the device may deviate from the oracle as far as the oracle deviates from itself under that perturbation (x20).
    rd, cd, _ = device_rollout(model, rods, cab, 220, "f64")
    ro, co = oracle_rollout(model, rods, cab, 220)
    assert max(rod_errors(rd, ro, 1.0)) < 1e-7 and tension_rel_error(cd, co, floor=10.0) < 1e-7
    rd, cd, _ = device_rollout(model, rods, cab, steps, "f64")
    ro, co = oracle_rollout(model, rods, cab, steps)
    pert = rods.copy()
    pert[:, :, 0:3] += 1e-13 * np.random.default_rng(0).normal(size=pert[:, :, 0:3].shape)
    rp, _ = oracle_rollout(model, pert, cab, steps)
    own = max(rod_errors(rp, ro, 1.0))
    dev = max(rod_errors(rd, ro, 1.0))
    print("prism on ground f64, %d steps: device vs oracle %.3e, oracle vs perturbed oracle %.3e" % (steps, dev, own))
    assert dev < max(1e-6, 20.0 * own)
    assert tension_rel_error(cd, co, floor=10.0) < max(1e-6, 20.0 * own)


# ------------------------------------------------------------------------------------------------ motors (config 5)
@pytest.mark.parametrize("kind", ["basic", "kinematic"])
@pytest.mark.parametrize("dt", [2e-3, 1e-3, 2e-4, 1e-4])
def test_motor_rig_matches_oracle(kind, dt):
    """tsTestRig (MotorTimestep_test.cpp:81-215): setControlInput(5.0, dt) / setControlInput(-580) every step, 1 s."""
    model = nt.build_example("tsTestRig_" + kind)
    n_env = 8
    steps = int(round(1.0 / dt))
    rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    cab = tiled_cables(model, n_env)
    ctrl = capi.TGB_CTRL_SETPOINT if kind == "basic" else capi.TGB_CTRL_TORQUE
    params = np.full((n_env, 1), 5.0 if kind == "basic" else -580.0)
    rd, cd, _ = device_rollout(model, rods, cab, steps, "f64", ctrl, params, dt=dt)
    ro, co = oracle_rollout(model, rods, cab, steps, ctrl, params, dt=dt)
    assert np.abs(cd - co).max() < 1e-8
    rd32, cd32, _ = device_rollout(model, rods, cab, steps, "f32", ctrl, params, dt=dt)
    assert np.abs(cd32[..., 0] - co[..., 0]).max() < 1e-3  # rest length; the reference test's own bound is 0.03


@pytest.mark.parametrize("kind", ["basic", "kinematic"])
def test_motor_rig_timestep_independence_on_device(kind):
    """The reference's own assertion, on the device: rest length after 1 s equal across dt within 0.03."""
    model = nt.build_example("tsTestRig_" + kind)
    ctrl = capi.TGB_CTRL_SETPOINT if kind == "basic" else capi.TGB_CTRL_TORQUE
    params = np.full((4, 1), 5.0 if kind == "basic" else -580.0)
    rest = {}
    for dt in (2e-3, 1e-3, 2e-4, 1e-4):
        rods, cab = np.tile(model.initial_rod_state(), (4, 1, 1)), tiled_cables(model, 4)
        _, cd, _ = device_rollout(model, rods, cab, int(round(1.0 / dt)), "f32", ctrl, params, dt=dt)
        rest[dt] = cd[0, 0, 0]
    for dt in rest:
        assert abs(rest[dt] - rest[1e-3]) <= 0.03, rest


# ------------------------------------------------------------------------------------------------ API behaviour
def test_control_input_and_errors():
    model = nt.build_example("3_prism_noground")
    n_env = 5
    rods, cab = random_poses(model, n_env, seed=3), tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    target = np.linspace(5.0, 9.0, model.n_cables)
    w.set_control_input(target, 1e-3)
    obs = w.get_cable_obs()
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_rod_state(rods[e])
        o.set_cable_state(cab[e])
        assert o.set_control_input(target, 1e-3) == 0
        assert np.abs(o.get_cable_obs() - obs[e]).max() < 1e-10
    with pytest.raises(capi.TgbError) as ei:
        w.set_control_input(-target, 1e-3)
    assert ei.value.status == capi.TGB_ERR_INVALID_ARGUMENT  # "Rest length is negative."
    with pytest.raises(capi.TgbError) as ei:
        w.step(0.0, 1)
    assert ei.value.status == capi.TGB_ERR_INVALID_ARGUMENT  # tgSimulation.cpp:197-200
    with pytest.raises(capi.TgbError):
        w.step(-1e-3, 1)
    w.step(1e-3, 0)  # zero steps: a no-op
    assert w.launch_count() == 0
    assert abs(w.time()) == 0.0
    w.step(1e-3, 3)
    assert abs(w.time() - 3e-3) < 1e-15
    w.close()


def test_rollout_host_equals_set_step_get():
    model = nt.build_example("Escape_c3")
    n_env, steps = 33, 40
    rods, cab, params = escape_inputs(model, n_env, seed=77)
    a = device_rollout(model, rods, cab, steps, "f32", capi.TGB_CTRL_ESCAPE_SINE, params)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params)
    rd, cd = w.rollout_host(rods, cab, model.dt, steps)
    w.close()
    assert np.array_equal(rd, a[0]) and np.array_equal(cd, a[1])


def test_pipelined_rollout_host_equals_set_step_get(monkeypatch):
    """For large batches (or with TGB_ROLLOUT_CHUNKS set) tgb_rollout_host cuts the batch into chunks on separate streams
    (copies of one chunk overlap the steps of another); a ragged size must give bit-identical results to one launch
    over the whole batch."""
    monkeypatch.setenv("TGB_ROLLOUT_CHUNKS", "4")
    model = nt.build_example("Escape_c3")
    n_env, steps = 1500 + 3, 30
    rods, cab, params = escape_inputs(model, n_env, seed=78)
    a = device_rollout(model, rods, cab, steps, "f32", capi.TGB_CTRL_ESCAPE_SINE, params)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params)
    rd, cd = w.rollout_host(rods, cab, model.dt, steps)
    assert w.launch_count() == 4
    assert abs(w.time() - steps * model.dt) < 1e-12
    # a second rollout from the same inputs (controller time restarted): no state leaks between chunks or calls
    w.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params)
    rd2, cd2 = w.rollout_host(rods, cab, model.dt, steps)
    w.close()
    assert np.array_equal(rd, a[0]) and np.array_equal(cd, a[1])
    assert np.array_equal(rd2, a[0]) and np.array_equal(cd2, a[1])


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_bit_identical_across_launch_geometry(precision):
    """The arithmetic of a body / cable does not depend on which lane runs it: results are bit-identical across
    lanes-per-env, CTA size and launch chunking."""
    model = nt.build_example("Escape_c3")
    n_env, steps = 70, 120
    rods, cab, params = escape_inputs(model, n_env, seed=5)
    ref = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params, wide=False)
    for kw in (dict(lanes_per_env=16, wide=False), dict(lanes_per_env=32, block_threads=64), dict(block_threads=128, wide=False), dict(chunks=7, wide=False)):
        got = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params, **kw)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), kw
    # The four-warp layout is a separate instantiation of the kernel.  Double precision is compiled without multiply-add contraction
    # and is bit-identical to the narrow layout.  In single precision the compiler is free to contract a * b + c * d either way and
    # does so differently in the two instantiations (seen in the SASS, round 2), so the layouts agree at round-off level, not bit for
    # bit: identical with itself across chunking, and within the FP32 tolerance of the narrow layout over a short horizon.
    wide = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params, wide=True)
    wide7 = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params, wide=True, chunks=7)
    assert np.array_equal(wide[0], wide7[0]) and np.array_equal(wide[1], wide7[1])
    if precision == "f64":
        assert np.array_equal(wide[0], ref[0]) and np.array_equal(wide[1], ref[1])
    else:
        short_n = device_rollout(model, rods, cab, 40, precision, capi.TGB_CTRL_ESCAPE_SINE, params, wide=False)
        short_w = device_rollout(model, rods, cab, 40, precision, capi.TGB_CTRL_ESCAPE_SINE, params, wide=True)
        assert rod_errors(short_w[0], short_n[0], 17.0)[0] < 1e-5 and tension_rel_error(short_w[1], short_n[1], floor=100.0) < F32_REL


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("example,ctrl", [("3_prism_noground", 0), ("3_prism", 0), ("SUPERball", capi.TGB_CTRL_TENSION),
                                          ("tsTestRig_kinematic", capi.TGB_CTRL_TORQUE)])
def test_wide_layout_is_bit_identical(precision, example, ctrl):
    """The four-warp layout (body warp + shared cable phase; tgb_set_wide_layout) against the narrow one: 4 and 8 lanes per
    environment, with and without ground, environment counts that leave tail groups, a launch cut in chunks."""
    model = nt.build_example(example)
    n_env, steps = 37, 150
    rods = random_poses(model, n_env, seed=9, translate=0.5, lift=(0.0, 1.0), yaw_only=True)
    cab = tiled_cables(model, n_env)
    params = None
    if ctrl == capi.TGB_CTRL_TENSION:
        params = np.full((n_env, 1), 1000.0)
    elif ctrl == capi.TGB_CTRL_TORQUE:
        params = np.random.default_rng(4).uniform(-2.0, 2.0, size=(n_env, model.n_cables))
    kw = dict(lanes_per_env=4) if model.n_bodies <= 2 else {}
    ref = device_rollout(model, rods, cab, steps, precision, ctrl, params, wide=False, **kw)
    got = [device_rollout(model, rods, cab, steps, precision, ctrl, params, wide=True, **kw, **extra) for extra in (dict(), dict(chunks=4))]
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1]) and np.array_equal(got[0][2], got[1][2])
    if precision == "f64":
        assert np.array_equal(got[0][0], ref[0]) and np.array_equal(got[0][1], ref[1]) and np.array_equal(got[0][2], ref[2])
    else:
        # single precision: round-off level agreement between the two instantiations (see test_bit_identical_across_launch_geometry)
        short_n = device_rollout(model, rods, cab, 30, precision, ctrl, params, wide=False, **kw)
        short_w = device_rollout(model, rods, cab, 30, precision, ctrl, params, wide=True, **kw)
        assert rod_errors(short_w[0], short_n[0], 17.0)[0] < 1e-5 and np.array_equal(short_w[2], short_n[2])
    assert np.isfinite(ref[0]).all()


def test_scores_and_status():
    model = nt.build_example("Escape_c3")
    n_env, steps = 12, 300
    rods, cab, params = escape_inputs(model, n_env, seed=21)
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params)
    w.set_height_window(1.0, 25.0)
    w.step(model.dt, steps)
    sc = w.scores()
    st = w.status()
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_rod_state(rods[e])
        o.set_cable_state(cab[e])
        o.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params[e])
        o.step(model.dt, steps)
        assert abs(o.scores()[0] - sc[e, 0]) < 1e-7
        assert sc[e, 1] == 0.0
    assert not (st & 1).any()
    w.close()


def test_host_api_drop_in_runs_on_device():
    """tgSimulation::run of the example app through the C++ host API == direct C-ABI rollout."""
    import ctypes as C
    h = capi.load_host()
    model = nt.build_example("3_prism")
    n_env, steps = 4, 100
    rod = np.empty((n_env, model.n_bodies, 13))
    cab = np.empty((n_env, model.n_cables, 4))
    rc = h.tgh_run_example(b"3_prism", n_env, capi.TGB_F64, steps, rod.ctypes.data_as(C.c_void_p), cab.ctypes.data_as(C.c_void_p))
    assert rc == 0, h.tgh_last_error()
    rods0 = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    rd, cd, _ = device_rollout(model, rods0, tiled_cables(model, n_env), steps, "f64")
    assert np.array_equal(rd, rod) and np.array_equal(cd, cab)


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_momentum_conservation_zero_gravity():
    """4096 envs (BASELINE.json batch size), zero gravity, no ground: cables are internal forces, so the total linear
    momentum of every environment stays zero and its mass centre does not move (MuscleNP_test.cpp:108-112 style)."""
    import copy
    model = copy.copy(nt.build_example("3_prism_noground"))
    model.gravity = 0.0
    n_env, steps = 4096, 500
    rods, cab = random_poses(model, n_env, seed=42), tiled_cables(model, n_env)
    mass = np.array([b.mass for b in model.bodies])
    for precision, tol in (("f64", 1e-9), ("f32", 2e-3)):
        rd, cd, st = device_rollout(model, rods, cab, steps, precision)
        p = (rd[:, :, 7:10] * mass[None, :, None]).sum(axis=1)
        com0 = (rods[:, :, 0:3] * mass[None, :, None]).sum(axis=1) / mass.sum()
        com1 = (rd[:, :, 0:3] * mass[None, :, None]).sum(axis=1) / mass.sum()
        print(precision, "max |p| %.3e  max com drift %.3e" % (np.abs(p).max(), np.abs(com1 - com0).max()))
        assert np.abs(p).max() < tol and np.abs(com1 - com0).max() < tol
        assert not st.any()


def test_full_size_superball_sample_against_oracle():
    """config 3 at full batch size: every 64th environment of a 4096-env run checked against the oracle (f64)."""
    model = nt.build_example("Escape_c3")
    n_env, steps = 4096, 100
    rods, cab, params = escape_inputs(model, n_env, seed=2345)
    rd, cd, st = device_rollout(model, rods, cab, steps, "f64", capi.TGB_CTRL_ESCAPE_SINE, params)
    idx = np.arange(0, n_env, 64)
    ro, co = oracle_rollout(model, rods[idx], cab[idx], steps, capi.TGB_CTRL_ESCAPE_SINE, params[idx])
    assert max(rod_errors(rd[idx], ro, 1.0)) < 1e-7
    assert not st.any()


# ------------------------------------------------------------------------------------------------ config 4: TetraSpine + CPG
def spine_inputs(n_env, seed=3456):
    """Shipped TetraSpine parameters (tests/golden/tetraspine_cpg.npz carries the .nnw values), every member perturbed by
    N(0, 0.03) clipped to [0, 1] like NeuroEvoMember::mutate."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "tetraspine_cpg.npz"))
    rng = np.random.default_rng(seed)
    node = np.clip(g["node_actions"][None] + rng.normal(0, 0.03, size=(n_env, 2)), 0, 1)
    edge = np.clip(g["edge_actions"][None] + rng.normal(0, 0.03, size=(n_env, 57, 2)), 0, 1)
    node[0], edge[0] = g["node_actions"], g["edge_actions"]
    return nt.tetraspine_param_rows(node, edge)


@pytest.mark.parametrize("steps", [1, 10, 60])
def test_tetraspine_cpg_f64(steps):
    model, ctl = nt.build_example("TetraSpine"), nt.tetraspine_cpg()
    n_env = 21
    rows = spine_inputs(n_env)
    rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    cab = tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_cpg_controller(ctl, rows)
    w.step(model.dt, steps)
    rd, cd, x = w.get_rod_state(), w.get_cable_obs(), w.get_cpg_state()
    sc, st = w.scores(), w.status()
    w.close()
    ro, co, so, sto = oracle_py.rollout_batch(model, capi.WorldCfg, rods, cab, model.dt, steps, capi.TGB_CTRL_CPG_IMPEDANCE, rows,
                                              cpg_desc=ctl.desc, want_scores=True, n_threads=8)
    xo = []
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_cpg_controller(ctl.desc, rows[e])
        o.step(model.dt, steps)
        xo.append(o.cpg_state())
    print("tetraspine f64 %d steps: rods %s, cables %.2e, cpg %.2e" % (steps, rod_errors(rd, ro, 1.0), np.abs(cd - co).max(),
                                                                         np.abs(x - np.array(xo)).max()))
    assert np.abs(x - np.array(xo)).max() < 1e-9           # CPG is decoupled from the bodies: tight at any horizon
    assert max(rod_errors(rd, ro, 1.0)) < 1e-6             # stiff contact amplifies round-off ~10x / 25 steps (tests/test_golden.py)
    assert tension_rel_error(cd, co, floor=10.0) < 1e-6
    assert np.abs(sc - so).max() < 1e-6 and np.array_equal(st & 6, sto & 6)


def test_tetraspine_cpg_f32_short_horizon():
    model, ctl = nt.build_example("TetraSpine"), nt.tetraspine_cpg()
    n_env, steps = 64, 30
    rows = spine_inputs(n_env)
    rods, cab = np.tile(model.initial_rod_state(), (n_env, 1, 1)), tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_cpg_controller(ctl, rows)
    w.step(model.dt, steps)
    rd, cd, x = w.get_rod_state(), w.get_cable_obs(), w.get_cpg_state()
    w.close()
    ro, co = oracle_py.rollout_batch(model, capi.WorldCfg, rods, cab, model.dt, steps, capi.TGB_CTRL_CPG_IMPEDANCE, rows, cpg_desc=ctl.desc,
                                     n_threads=8)
    perr = rod_errors(rd, ro, 38.1)
    lerr = np.abs(cd[..., 1] - co[..., 1]).max() / 38.1
    rerr = np.abs(cd[..., 0] - co[..., 0]).max() / 38.1
    print("tetraspine f32 %d steps: rods %s length %.2e rest %.2e" % (steps, perr, lerr, rerr))
    assert perr[0] < F32_REL and perr[1] < F32_REL and lerr < F32_REL and rerr < F32_REL


@pytest.mark.parametrize("precision,steps", [("f64", 100), ("f32", 40)])
def test_sphere_and_rod_compound_on_the_ground(precision, steps):
    """SURVEY.md 8(f) rank 2 (first part): compound bodies of rods and spheres.  The hardware-faithful tetrahedral spine of the ICRA 2015
    comparison (3 bodies of 8 cylinders + 6 point-mass spheres, pretensioned cables, no controller) drops 0.18 onto its corner
    spheres and rod ends (first contact around step 19): sphere-ground and cylinder-ground contacts in the same manifold set.  The
    resting contact of a many-point compound chatters (no persistent manifold, no warm start) and amplifies round-off: f64 agrees to
    1e-12 over 100 steps and to 1e-2 after 300, so the horizons are short."""
    model = nt.build_example("TetraSpineStatic_hf")
    assert sum(1 for s in model.shapes if s.kind == 1) == 18 and sum(1 for s in model.shapes if s.kind == 0) == 24
    n_env = 24
    rods = random_poses(model, n_env, seed=31, translate=0.5, lift=(0.0, 1.0), yaw_only=True)
    cab = tiled_cables(model, n_env)
    rd, cd, st = device_rollout(model, rods, cab, steps, precision)
    ro, co = oracle_rollout(model, rods, cab, steps)
    e = rod_errors(rd, ro, 38.0 if precision == "f32" else 1.0)
    terr = tension_rel_error(cd, co, floor=100.0)
    print("sphere+rod compound %s %d steps: rods %s tension %.2e, min height %.3f" % (precision, steps, e, terr, ro[:, :, 1].min()))
    assert not st.any()
    if precision == "f64":
        assert max(e) < 1e-7 and terr < 1e-7
    else:
        assert e[0] < F32_REL and e[1] < F32_REL and terr < 5 * F32_REL


def flemons_inputs(example, n_env, seed=13):
    """controller rows for a population of the kinematic tetrahedral-complex spine: the shipped parameters, mutated like
    NeuroEvoMember::mutate (member 0 = the shipped parameters)"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "%s_kinematic_cpg.npz" % example.lower()))
    rng = np.random.default_rng(seed)
    node = np.clip(g["node_actions"][None] + rng.normal(0, 0.03, size=(n_env, 2)), 0, 1)
    edge = np.clip(g["edge_actions"][None] + rng.normal(0, 0.03, size=(n_env, 100, 2)), 0, 1)
    node[0], edge[0] = g["node_actions"], g["edge_actions"]
    return nt.flemons_param_rows(node, edge, example)


@pytest.mark.parametrize("example,steps", [("FlemonsSpine3", 1), ("FlemonsSpine3", 60), ("FlemonsSpine", 60)])
def test_flemons_kinematic_cpg_pid_f64(example, steps):
    """tgKinematicActuator cables driven by CPG -> impedance -> tgPIDController (cable_control 1), 3 and 12
    vertebrae (12 compound bodies of 4 cylinders, 88 cables, 16 lanes per environment): double path vs the oracle."""
    model, ctl = nt.build_example(example), nt.flemons_cpg(example)
    n_env = 9
    rows = flemons_inputs(example, n_env)
    rods, cab = np.tile(model.initial_rod_state(), (n_env, 1, 1)), tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_cpg_controller(ctl, rows)
    w.step(model.dt, steps)
    rd, cd, x = w.get_rod_state(), w.get_cable_obs(), w.get_cpg_state()
    sc, st = w.scores(), w.status()
    w.close()
    ro, co, so, sto = oracle_py.rollout_batch(model, capi.WorldCfg, rods, cab, model.dt, steps, capi.TGB_CTRL_CPG_IMPEDANCE, rows,
                                              cpg_desc=ctl.desc, want_scores=True, n_threads=8)
    xo = []
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_cpg_controller(ctl.desc, rows[e])
        o.step(model.dt, steps)
        xo.append(o.cpg_state())
    print("%s f64 %d steps: rods %s, cables %.2e, cpg %.2e" % (example, steps, rod_errors(rd, ro, 1.0), np.abs(cd - co).max(),
                                                               np.abs(x - np.array(xo)).max()))
    assert np.abs(x - np.array(xo)).max() < 1e-9
    assert max(rod_errors(rd, ro, 1.0)) < 1e-6       # the PID loop amplifies round-off ~10x / 20 steps (tests/test_golden.py)
    assert tension_rel_error(cd, co, floor=10.0) < 1e-6
    assert np.abs(cd[..., 3] - co[..., 3]).max() < 1e-6   # motor velocity
    assert np.abs(sc - so).max() < 1e-6 and np.array_equal(st & 6, sto & 6)


def test_flemons_kinematic_cpg_pid_f32_short_horizon():
    """Single precision on the PID-controlled kinematic spine.  tgPIDController differentiates the tension error over dt = 1e-3 with
    kD = 10 (and kP = 5000): the tension (length - rest) * k carries the single-precision resolution of a length of ~11 on k = 1000,
    ~1e-3, which the derivative term turns into torque noise of the order of 10 and the non-backdrivable motor's switching logic into
    per-environment divergence as soon as the cables go taut (step ~21; measured: motor velocity off by 5 at step 25).  The stated
    FP32 tolerance is therefore only meaningful until then -- 20 steps here; the double-precision instantiation is the one to use
    for this controller (tests above).  Later steps are checked for sanity only: finite state, rest lengths within the motor limits."""
    example = "FlemonsSpine"
    model, ctl = nt.build_example(example), nt.flemons_cpg(example)
    n_env, steps = 32, 20
    rows = flemons_inputs(example, n_env)
    rods, cab = np.tile(model.initial_rod_state(), (n_env, 1, 1)), tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_cpg_controller(ctl, rows)
    w.step(model.dt, steps)
    rd, cd = w.get_rod_state(), w.get_cable_obs()
    ro, co = oracle_py.rollout_batch(model, capi.WorldCfg, rods, cab, model.dt, steps, capi.TGB_CTRL_CPG_IMPEDANCE, rows, cpg_desc=ctl.desc,
                                     n_threads=8)
    perr = rod_errors(rd, ro, 20.0)
    lerr = np.abs(cd[..., 1] - co[..., 1]).max() / 20.0
    rerr = np.abs(cd[..., 0] - co[..., 0]).max() / 20.0
    print("flemons f32 %d steps: rods %s length %.2e rest %.2e" % (steps, perr, lerr, rerr))
    assert perr[0] < F32_REL and perr[1] < F32_REL and lerr < F32_REL and rerr < F32_REL
    w.step(model.dt, 180)
    rd, cd, st = w.get_rod_state(), w.get_cable_obs(), w.status()
    w.close()
    assert np.isfinite(rd).all() and np.isfinite(cd).all() and not (st & 1).any()
    assert cd[..., 0].min() >= 0.1 - 1e-6 and cd[..., 0].max() < 40.0        # rest length: floor minRestLength, no runaway
    assert np.abs(cd[..., 3]).max() <= 12.0 * 1.5                           # motor velocity stays of the order of the target velocity


def test_population_evaluation_matches_oracle_scores():
    """NeuroEvo-style evaluate(params) -> scores on the device vs the oracle, incl. a member that leaves the height window."""
    from ntrtsim_b200.neuroevo import evaluate_population
    model, ctl = nt.build_example("TetraSpine"), nt.tetraspine_cpg()
    n_env, steps = 10, 50
    rows = spine_inputs(n_env)
    scores, status = evaluate_population(model, ctl, rows, steps, precision="f64")
    rods, cab = np.tile(model.initial_rod_state(), (n_env, 1, 1)), tiled_cables(model, n_env)
    _, _, so, sto = oracle_py.rollout_batch(model, capi.WorldCfg, rods, cab, model.dt, steps, capi.TGB_CTRL_CPG_IMPEDANCE, rows,
                                            cpg_desc=ctl.desc, want_scores=True)
    assert scores.shape == (n_env, 2) and np.abs(scores - so).max() < 1e-6
    assert np.array_equal(status & 6, sto & 6)
    # bogus episode: a height window the spine is outside of from the first step -> score -1 (BaseSpineCPGControl.cpp:296-299)
    import copy
    ctl2 = nt.tetraspine_cpg()
    ctl2.desc.height_lo, ctl2.desc.height_hi = 100.0, 200.0
    s2, st2 = evaluate_population(model, ctl2, rows[:3], 5, precision="f32")
    assert np.all(s2[:, 0] == -1.0) and np.all(st2 & 2)


@pytest.mark.parametrize("kind", ["basic", "kinematic"])
def test_host_step_code_runs_between_the_halves(kind):
    """tsTestRig::step commands its motor on the host every step (tsTestRig.cpp:181-188). Stepping through the host API's
    tgSimulation::step (world half -> host model walk -> models half) must equal the device-side controller rollout and
    the oracle: the host code runs exactly where the reference runs it, not one step late."""
    import ctypes as C
    h = capi.load_host()
    name = "tsTestRig_" + kind
    model = nt.build_example(name)
    n_env, steps = 3, 400
    rod = np.empty((n_env, model.n_bodies, 13))
    cab = np.empty((n_env, model.n_cables, 4))
    rc = h.tgh_run_example_stepwise(name.encode(), n_env, capi.TGB_F64, steps, rod.ctypes.data_as(C.c_void_p), cab.ctypes.data_as(C.c_void_p))
    assert rc == 0, h.tgh_last_error()
    ctrl = capi.TGB_CTRL_SETPOINT if kind == "basic" else capi.TGB_CTRL_TORQUE
    params = np.full((n_env, 1), 5.0 if kind == "basic" else -580.0)
    rods0 = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    rd, cd, _ = device_rollout(model, rods0, tiled_cables(model, n_env), steps, "f64", ctrl, params)
    assert np.array_equal(rd, rod) and np.array_equal(cd, cab)
    ro, co = oracle_rollout(model, rods0, tiled_cables(model, n_env), steps, ctrl, params)
    assert np.abs(cab - co).max() < 1e-8


def test_world_and_models_halves_compose_to_a_step():
    model = nt.build_example("Escape_c3")
    n_env, steps = 9, 25
    rods, cab, params = escape_inputs(model, n_env, seed=31)
    ref = device_rollout(model, rods, cab, steps, "f32", capi.TGB_CTRL_ESCAPE_SINE, params)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params)
    for _ in range(steps):
        w.step_world(model.dt)
        w.step_models(model.dt)
    assert abs(w.time() - steps * model.dt) < 1e-12
    assert np.array_equal(w.get_rod_state(), ref[0]) and np.array_equal(w.get_cable_obs(), ref[1])
    w.close()


# ------------------------------------------------------------------------------------------------ ICRA 2015 hardware comparison
ICRA2015_HARDWARE_N = [14.1811, 15.5914, 24.6800, 22.0161, 18.0986, 34.3954]   # outer right/left/top, inner right/left/top (front)
ICRA2015_REFERENCE_BOUND = [0.2251, 0.3501, 0.19, 0.0601, 0.0715, 0.06]        # ICRA2015_test.cpp:133-150


def test_serialized_spine_control_matches_oracle_f64():
    """SerializedSpineControl (per-cable impedance control + the load ramp applied at markers) and the tension history maximum:
    device (double) vs oracle over the first 400 steps, and with the clock moved to the loaded phase."""
    model = nt.build_example("TetraSpineStatic_hf")
    row = nt.serialized_spine_params(model)
    for ramp in ((30.0, 35.0, 40.0, 45.0), (0.05, 0.1, 0.25, 0.3)):     # the second: the load acts inside the short horizon
        r = row.copy()
        r[4:8] = ramp
        n_env, steps = 6, 100
        rods = random_poses(model, n_env, seed=41, translate=0.3, lift=(0.0, 0.5), yaw_only=True)
        cab = tiled_cables(model, n_env)
        w = nt.BatchedWorld(model, n_env, precision="f64")
        w.set_rod_state(rods)
        w.set_cable_state(cab)
        w.set_controller(capi.TGB_CTRL_SERIALIZED_SPINE, np.tile(r, (n_env, 1)))
        w.track_max_tension(True)
        for _ in range(4):          # several launches: simTime / updateTime travel through the frame record
            w.step(model.dt, steps // 4)
        rd, cd, mt = w.get_rod_state(), w.get_cable_obs(), w.max_tension()
        w.close()
        for e in range(n_env):
            o = oracle_py.OracleWorld(model, capi.WorldCfg)
            o.set_rod_state(rods[e])
            o.set_cable_state(cab[e])
            o.set_controller(capi.TGB_CTRL_SERIALIZED_SPINE, r)
            o.track_max_tension(True)
            o.step(model.dt, steps)
            ro = align_quat_sign(o.get_rod_state(), rd[e])
            assert np.abs(rd[e] - ro).max() < 1e-7, (ramp, e, np.abs(rd[e] - ro).max())
            assert np.abs(cd[e] - o.get_cable_obs()).max() < 1e-6
            assert np.abs(mt[e] - o.max_tension()).max() < 1e-6


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_icra2015_hardware_comparison(precision):
    """The reference's own whole-path integration test, test_integration/ICRA2015Tests/ICRA2015_test.cpp:77-150: the hardware-faithful
    3-segment spine under SerializedSpineControl for 60000 steps; the maxima of the six front cables' tension histories are compared
    with the tensions measured on the robot.  The bounds in that test are the reference simulator's own errors plus a hair.  60 s of
    resting contact make the run chaotic at the percent level, so a single run meeting or missing the two bounds without room is a
    coin toss (tests/test_golden.py, profiles/icra2015_ensemble_r02.txt).  Here the batch does what it is good at: 16 copies of the
    robot whose initial positions differ by 1e-9 run the test at once, and the ensemble is asserted like the oracle's -- the three
    bounds with room for every member, every member within every bound widened by one percentage point, the ensemble median within
    0.006 of every bound -- plus agreement of the device's medians with the oracle's."""
    model = nt.build_example("TetraSpineStatic_hf")
    row = nt.serialized_spine_params(model)
    n = 16
    rods = np.tile(model.initial_rod_state(), (n, 1, 1))
    for seed in range(1, n):
        rods[seed, :, 0:3] += 1e-9 * np.random.default_rng(seed).normal(size=rods[seed, :, 0:3].shape)
    w = nt.BatchedWorld(model, n, precision=precision)
    w.set_rod_state(rods)
    w.set_controller(capi.TGB_CTRL_SERIALIZED_SPINE, np.tile(row, (n, 1)))
    w.track_max_tension(True)
    for _ in range(6):
        w.step(model.dt, 10000)
    mt = w.max_tension()
    assert not w.status().any()
    w.close()
    hardware, bound = np.array(ICRA2015_HARDWARE_N), np.array(ICRA2015_REFERENCE_BOUND)
    err = np.abs((hardware[None] - mt[:, :6] / 100.0) / hardware[None])
    print("ICRA2015 %s, ensemble of %d: unperturbed %s median %s min %s max %s; members meeting all six reference bounds %d"
          % (precision, n, np.round(err[0], 4), np.round(np.median(err, axis=0), 4), np.round(err.min(axis=0), 4), np.round(err.max(axis=0), 4),
             int(np.all(err <= bound, axis=1).sum())))
    assert np.all(err[:, [0, 2, 4]] <= bound[[0, 2, 4]]), err
    # (single precision spreads a little wider than double: measured 0.342 ... 0.3605 on the outer front left cable)
    assert np.all(err <= bound + (0.01 if precision == "f64" else 0.013)), err
    assert np.all(np.median(err, axis=0) <= bound + 0.0065), np.median(err, axis=0)
    from test_golden import icra_errors
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as ex:
        eo = np.array(list(ex.map(lambda sd: icra_errors(model, row, 0.0, sd), range(8))))
    print("oracle ensemble of 8: median", np.round(np.median(eo, axis=0), 4))
    assert np.abs(np.median(err, axis=0) - np.median(eo, axis=0)).max() < 0.006


def test_data_logger_samples_match_the_device_state(tmp_path):
    """tgDataLogger2 through tgSimulation::run on the device: one CSV line per step (time interval 0), the reference's default
    stream formatting (6 significant digits); the last line equals the final state read through the C ABI, for a chosen environment."""
    import ctypes as C
    h = capi.load_host()
    buf = C.create_string_buffer(512)
    steps, n_env = 40, 3
    rc = h.tgh_run_example_logged(b"3_prism", n_env, capi.TGB_F64, steps, str(tmp_path / "prism").encode(), 0.0, 2, buf, 512)
    assert rc == 0, h.tgh_last_error()
    lines = [ln for ln in open(buf.value.decode()).read().split("\n") if ln]
    assert len(lines) == 2 + steps
    n_cols = len(lines[1].split(","))
    rows = np.array([[float(x) for x in ln.split(",")[:-1]] for ln in lines[2:]])
    assert rows.shape == (steps, n_cols - 1)
    assert np.allclose(rows[:, 0], 1e-3 * np.arange(1, steps + 1), rtol=1e-5)
    model = nt.build_example("3_prism")
    rods = np.empty((n_env, model.n_bodies, 13))
    cab = np.empty((n_env, model.n_cables, 4))
    assert h.tgh_run_example_stepwise(b"3_prism", n_env, capi.TGB_F64, steps, rods.ctypes.data_as(C.c_void_p), cab.ctypes.data_as(C.c_void_p)) == 0
    last = rows[-1]
    for b in range(model.n_bodies):
        assert np.allclose(last[1 + 7 * b: 4 + 7 * b], rods[2, b, 0:3], rtol=2e-5, atol=1e-6)       # centre of mass
        assert abs(last[7 + 7 * b] - model.bodies[b].mass) < 1e-5 * model.bodies[b].mass           # mass
    for c in range(model.n_cables):
        o = 1 + 7 * model.n_bodies + 3 * c
        assert np.allclose(last[o: o + 3], cab[2, c, 0:3], rtol=2e-5, atol=1e-6)                    # rest length, length, tension


def test_data_logger_compound_samples_match_the_device_state(tmp_path):
    """tgCompoundRigidSensor on the device: the logged position of every TetraSpine segment is its body's centre of mass, its mass
    column the sum of its six members' mass() (each the whole body's, as in the reference), its yaw / pitch / roll the rotation since
    the logger was set up."""
    import ctypes as C
    import re
    h = capi.load_host()
    buf = C.create_string_buffer(512)
    steps = 25
    rc = h.tgh_run_example_logged(b"TetraSpine", 2, capi.TGB_F64, steps, str(tmp_path / "spine").encode(), 0.0, 1, buf, 512)
    assert rc == 0, h.tgh_last_error()
    lines = [ln for ln in open(buf.value.decode()).read().split("\n") if ln]
    cols = lines[1].split(",")[:-1]
    last = [float(x) for x in lines[-1].split(",")[:-1]]
    model = nt.build_example("TetraSpine")
    rods = np.empty((2, model.n_bodies, 13))
    assert h.tgh_run_example_stepwise(b"TetraSpine", 2, capi.TGB_F64, steps, rods.ctypes.data_as(C.c_void_p), None) == 0
    xs = [i for i, c in enumerate(cols) if re.search(r"_compound\(compound_\w{6}\)\.X$", c)]
    assert len(xs) == 3
    seen = set()
    for i in xs:
        com = np.array(last[i: i + 3])
        body = int(np.argmin(np.abs(rods[1, :, 0:3] - com).max(axis=1)))
        seen.add(body)
        assert np.allclose(com, rods[1, body, 0:3], rtol=2e-5, atol=1e-6)
        assert abs(last[i + 6] - 6 * model.bodies[body].mass) < 1e-5 * 6 * model.bodies[body].mass
        assert all(abs(a) < 30.0 for a in last[i + 3: i + 6])  # a few degrees of settling, not radians or garbage
    assert seen == {0, 1, 2}


# ------------------------------------------------------------------------------------------------ YAML front-end (8f rank 3)
@pytest.mark.parametrize("precision,steps", [("f64", 150), ("f32", 60)])
def test_yaml_built_tower_matches_oracle(tmp_path, precision, steps):
    """A structure file of the reference's YAML format (three prisms joined by node_edge bonds: 9 rods, 27 strings) built by
    TensegrityModel, replicated with random poses, dropped on the box ground of the yamlbuilder app's world (the lowest rod ends touch
    from the first steps on).  The chattering resting contact amplifies a 1e-13 perturbation to 1e-11 after 150 steps, 1e-8 after 200
    and 1e-3 after 300 (measured on the oracle), hence the horizon."""
    from test_yaml import BASE_PRISM, TOWER, write
    from ntrtsim_b200 import examples
    write(tmp_path, "base/Prism.yaml", BASE_PRISM)
    model = examples.build_yaml(write(tmp_path, "tower.yaml", TOWER))
    n_env = 32
    rods = random_poses(model, n_env, seed=77, translate=0.5, lift=(-1.2, 0.5), yaw_only=True)
    cab = tiled_cables(model, n_env)
    rd, cd, st = device_rollout(model, rods, cab, steps, precision)
    ro, co = oracle_rollout(model, rods, cab, steps)
    e = rod_errors(rd, ro, 24.0 if precision == "f32" else 1.0)
    terr = tension_rel_error(cd, co, floor=100.0)
    print("yaml tower %s %d steps: rods %s tension %.2e, min height %.3f" % (precision, steps, e, terr, ro[:, :, 1].min()))
    assert not st.any()
    if precision == "f64":
        assert max(e) < 1e-7 and terr < 1e-7
    else:
        assert e[0] < F32_REL and e[1] < F32_REL and terr < 5 * F32_REL


def box_cart_inputs(model, n_env, seed=88):
    """Random yaw, a small random tilt about a horizontal axis (no two corners of a box at the same height: a tie at the contact
    threshold would be decided by round-off) and a height that leaves the box 0.05 .. 0.3 above the ground."""
    from helpers import quat_mul, quat_rotate
    rods = random_poses(model, n_env, seed=seed, translate=0.5, lift=(-0.48, -0.3), yaw_only=True)
    rng = np.random.default_rng(seed + 1)
    ang = rng.uniform(0.01, 0.05, size=n_env) * rng.choice([-1.0, 1.0], size=n_env)
    az = rng.uniform(0, 2 * np.pi, size=n_env)
    tilt = np.stack([np.sin(ang / 2) * np.cos(az), np.zeros(n_env), np.sin(ang / 2) * np.sin(az), np.cos(ang / 2)], axis=1)
    centre = rods[:, :, 0:3].mean(axis=1, keepdims=True)
    tb = np.broadcast_to(tilt[:, None, :], (n_env, model.n_bodies, 4))
    rods[:, :, 0:3] = quat_rotate(tb, rods[:, :, 0:3] - centre) + centre
    rods[:, :, 3:7] = quat_mul(tb, rods[:, :, 3:7])
    return rods


@pytest.mark.parametrize("precision,steps", [("f64", 80), ("f32", 45)])
def test_box_shapes_match_oracle(tmp_path, precision, steps):
    """SURVEY.md 8(f) rank 2 (second part): tgBoxInfo bodies.  A box with two rods tied to it by four pretensioned strings, built
    from a structure file; the box lands on a corner within the first 30..60 steps and tips onto an edge and a face while the strings
    pull the rods in.  Corner contacts switching on and off make this the most sensitive scenario of the suite: a 1e-13 perturbation
    of the oracle's input grows to 1e-12 after 60 steps, 2e-9 after 90 and 7e-4 after 120, hence the horizons."""
    from test_yaml import BOX_CART, write
    from ntrtsim_b200 import examples
    model = examples.build_yaml(write(tmp_path, "cart.yaml", BOX_CART))
    assert sum(1 for s in model.shapes if s.kind == 2) == 1
    n_env = 48
    rods = box_cart_inputs(model, n_env)
    cab = tiled_cables(model, n_env)
    rd, cd, st = device_rollout(model, rods, cab, steps, precision)
    ro, co = oracle_rollout(model, rods, cab, steps)
    e = rod_errors(rd, ro, 6.0 if precision == "f32" else 1.0)
    terr = tension_rel_error(cd, co, floor=50.0)
    print("box cart %s %d steps: rods %s tension %.2e, min height %.3f" % (precision, steps, e, terr, ro[:, :, 1].min()))
    assert not st.any()
    if precision == "f64":
        assert max(e) < 1e-7 and terr < 1e-7
    else:
        assert e[0] < F32_REL and e[1] < F32_REL and terr < 5 * F32_REL


# ------------------------------------------------------------------------------------------------ sizes and edge cases
def tall_tower_yaml(n_prisms):
    names = "/".join("p%d" % (i + 1) for i in range(n_prisms))
    return """
substructures:
  %s:
    path: base/Prism.yaml
    offset: [0, 12, 0]
    translation: [0, 3, 0]
bond_groups:
  prism_string:
    %s/node_edge:
      - [top1, bottom1/bottom2]
      - [top2, bottom2/bottom3]
      - [top3, bottom3/bottom1]
""" % (names, names)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_largest_supported_structure(tmp_path, precision):
    """32 rigid bodies per environment is the limit of the step kernel (one lane per body): ten bonded prisms = 30 rods and 90
    strings through the YAML front-end, one warp per environment, environment counts that do not fill the last CTA."""
    from test_yaml import BASE_PRISM, write
    from ntrtsim_b200 import examples
    write(tmp_path, "base/Prism.yaml", BASE_PRISM)
    model = examples.build_yaml(write(tmp_path, "tower10.yaml", tall_tower_yaml(10)))
    assert (model.n_bodies, model.n_cables) == (30, 90)
    n_env, steps = 5, 60
    rods = random_poses(model, n_env, seed=3, translate=0.3, lift=(-1.2, 0.0), yaw_only=True)
    cab = tiled_cables(model, n_env)
    rd, cd, st = device_rollout(model, rods, cab, steps, precision)
    ro, co = oracle_rollout(model, rods, cab, steps)
    e = rod_errors(rd, ro, 90.0 if precision == "f32" else 1.0)
    terr = tension_rel_error(cd, co, floor=100.0)
    print("tower of 10 %s: rods %s tension %.2e" % (precision, e, terr))
    assert not st.any()
    if precision == "f64":
        assert max(e) < 1e-7 and terr < 1e-7
    else:
        assert e[0] < F32_REL and e[1] < F32_REL and terr < 5 * F32_REL
    # one body more than a warp has lanes is refused, loudly
    too_big = examples.build_yaml(write(tmp_path, "tower11.yaml", tall_tower_yaml(11)))
    with pytest.raises(capi.TgbError, match="more than 32 rigid bodies"):
        nt.BatchedWorld(too_big, 1)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_single_body_without_cables_and_single_environment(tmp_path, precision):
    """Degenerate sizes: one compound body (three rods welded at a node, no cables at all), one environment, and a zero-step call."""
    from test_yaml import write
    from ntrtsim_b200 import examples
    model = examples.build_yaml(write(tmp_path, "tripod.yaml", """
nodes:
  hub: [0, 4, 0]
  f1: [2, 2.2, 0]
  f2: [-1, 2.2, 1.7]
  f3: [-1, 2.3, -1.7]
pair_groups:
  rod:
    - [hub, f1]
    - [hub, f2]
    - [hub, f3]
builders:
  rod:
    class: tgRodInfo
    parameters:
      radius: 0.2
      density: 0.5
"""))
    assert (model.n_bodies, model.n_cables, len(model.shapes)) == (1, 0, 3)
    rods = model.initial_rod_state()[None].copy()
    cab = np.zeros((1, 0, 2))
    w = nt.BatchedWorld(model, 1, precision=precision)
    try:
        w.set_rod_state(rods)
        w.step(model.dt, 0)  # a zero-step call changes nothing
        w.sync()
        before = w.get_rod_state()
        with pytest.raises(capi.TgbError):
            w.step(model.dt, -1)
        w.step(model.dt, 150)  # impact at step ~100, at rest by 160; the resting chatter amplifies round-off from there on
        w.sync()
        rd = w.get_rod_state()
        assert w.get_cable_obs().shape == (1, 0, capi.CABLE_OBS_DOUBLES)
    finally:
        w.close()
    assert np.allclose(before, rods, atol=1e-6 if precision == "f32" else 0)
    ro, _ = oracle_rollout(model, rods, cab, 150)
    e = rod_errors(rd, ro, 4.0 if precision == "f32" else 1.0)
    print("tripod %s: %s, height %.3f" % (precision, e, ro[0, 0, 1]))
    if precision == "f64":
        assert max(e) < 1e-7
    else:
        assert e[0] < 5 * F32_REL and e[1] < 5 * F32_REL   # 50 steps of impact and a three-point stance in single precision


# ------------------------------------------------------------------------------------------------ static obstacles (8f rank 2)
@pytest.mark.parametrize("precision,steps", [("f64", 45), ("f32", 40)])
def test_static_obstacle_contacts_match_oracle(tmp_path, precision, steps):
    """A static block (tgBoxInfo, density 0) with a rod and a ball dropping on its top face and a rod hitting its side face: obstacle
    contacts carry their own normal, friction and restitution through the solver.  The static body stays bit-exactly in place.
    A rod landing flat on a face is a four-point contact that chatters: a 1e-13 perturbation of the oracle's input grows to 6e-10 after
    60 steps and 7e-8 after 80, hence the horizons (first contacts: side face at step ~17, top face from step ~30)."""
    from test_yaml import OBSTACLE_SCENE, write
    from ntrtsim_b200 import examples
    model = examples.build_yaml(write(tmp_path, "scene.yaml", OBSTACLE_SCENE))
    static = [i for i, b in enumerate(model.bodies) if b.mass == 0.0]
    n_env = 24
    base = model.initial_rod_state()
    rng = np.random.default_rng(123)
    rods = np.tile(base, (n_env, 1, 1))
    moving = [i for i in range(model.n_bodies) if i not in static]
    # every moving body gets its own small random offset and initial velocity: contacts switch on at different steps (top rod and
    # ball 0.05 .. 0.4 above the top face; the low rod 0.1 from the side face and moving towards it)
    rods[:, moving, 0:3] += rng.uniform(-0.05, 0.05, size=(n_env, len(moving), 3)) + np.array([0.0, -0.25, 0.0])
    rods[:, moving, 7:10] = rng.uniform(-0.5, 0.5, size=(n_env, len(moving), 3))
    side = min((i for i in moving if model.shapes[model.bodies[i].shape_begin].kind == 0), key=lambda i: base[i, 1])
    rods[:, side, 1] += 0.25
    rods[:, side, 2] -= 0.6
    rods[:, side, 9] = -6.0
    cab = np.zeros((n_env, 0, 2))
    rd, _, st = device_rollout(model, rods, cab, steps, precision)
    ro, _ = oracle_rollout(model, rods, cab, steps)
    e = rod_errors(rd, ro, 6.0 if precision == "f32" else 1.0)
    print("obstacle scene %s %d steps: %s" % (precision, steps, e))
    assert not st.any()
    assert np.array_equal(rd[:, static], rods[:, static]) if precision == "f64" else np.allclose(rd[:, static], rods[:, static], atol=1e-6)
    if precision == "f64":
        assert max(e) < 1e-7
    else:
        assert e[0] < F32_REL and e[1] < F32_REL


@pytest.mark.parametrize("precision,steps", [("f64", 300), ("f32", 150)])
def test_escape_crater_app_matches_oracle(precision, steps):
    """AppEscape complete (src/examples/craterEscape/AppEscape.cpp:59-97): EscapeModel with its sine controller and tgCraterDeep as a
    second model of the same simulation -- four static tgBox walls around a 10 x 10 pit, rim at y = 20.  The balls are dropped from
    just above the rim and land on the walls' top faces and inner edges (the 17-long rods do not fit into the pit)."""
    model = nt.build_example("Escape_crater")
    dyn = [i for i, b in enumerate(model.bodies) if b.mass != 0.0]
    assert (model.n_bodies, len(dyn), model.n_cables) == (10, 6, 24)
    assert all(model.shapes[model.bodies[i].shape_begin].kind == 2 for i in range(model.n_bodies) if i not in dyn)
    n_env = 16
    rng = np.random.default_rng(5)
    rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    rods[:, dyn, 0:3] += rng.uniform(-0.3, 0.3, size=(n_env, 1, 3)) + np.array([0.0, 14.0, 0.0])
    cab = tiled_cables(model, n_env)
    params = nt.escape_transform_actions(rng.uniform(size=(n_env, 32)), 9.0)
    rd, cd, st = device_rollout(model, rods, cab, steps, precision, capi.TGB_CTRL_ESCAPE_SINE, params)
    ro, co = oracle_rollout(model, rods, cab, steps, capi.TGB_CTRL_ESCAPE_SINE, params)
    # per environment: the sampled surface point of a rod jumps when the obstacle feature closest to a station switches between a
    # face and an edge, and an environment that sits on such a switch amplifies round-off at once (on the oracle alone a 1e-15
    # perturbation of this very input moves one of the 16 environments by 5e-7 within 100 steps, the others by < 1e-10)
    scale = 17.0 if precision == "f32" else 1.0
    rda = align_quat_sign(rd, ro)
    per_env = np.maximum(np.abs(rda[..., 0:3] - ro[..., 0:3]).reshape(n_env, -1).max(axis=1) / scale,
                         np.abs(rda[..., 3:7] - ro[..., 3:7]).reshape(n_env, -1).max(axis=1))
    tper = (np.abs(cd[..., 2] - co[..., 2]) / np.maximum(np.abs(co[..., 2]), 2452.0)).max(axis=1)
    print("Escape + crater %s %d steps: pose error per env %s, tension %s, lowest rod centre %.2f"
          % (precision, steps, np.sort(per_env)[[0, n_env // 2, -2, -1]], np.sort(tper)[[n_env // 2, -1]], ro[:, dyn, 1].min()))
    assert not st.any()
    assert ro[:, dyn, 1].min() > 20.0  # held up by the rim
    tol = 1e-8 if precision == "f64" else F32_REL
    assert np.sum(per_env < tol) >= n_env - 2 and per_env.max() < 1e-3
    assert np.sum(tper < (1e-7 if precision == "f64" else 5 * F32_REL)) >= n_env - 2 and tper.max() < 1e-2


@pytest.mark.parametrize("scene", ["SUPERball_stairs", "SUPERball_blocks", "SUPERball_wall"])
def test_obstacle_models_match_oracle(scene):
    """tgStairs, tgBlockField and Wall (src/models/obstacles) next to the SUPERball with its T6TensionController: up to 20 static
    boxes per environment (26 bodies: one warp per environment).  Same per-environment criterion as the crater test."""
    model = nt.build_example(scene)
    dyn = [i for i, b in enumerate(model.bodies) if b.mass != 0.0]
    n_env, steps = 12, 250
    rng = np.random.default_rng(17)
    rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    rods[:, dyn, 0:3] += rng.uniform(-0.5, 0.5, size=(n_env, 1, 3)) + np.array([0.0, 1.0, 0.0])
    cab = tiled_cables(model, n_env)
    params = np.full((n_env, 1), 1000.0)
    rd, cd, st = device_rollout(model, rods, cab, steps, "f64", capi.TGB_CTRL_TENSION, params)
    ro, co = oracle_rollout(model, rods, cab, steps, capi.TGB_CTRL_TENSION, params)
    rda = align_quat_sign(rd, ro)
    per_env = np.maximum(np.abs(rda[..., 0:3] - ro[..., 0:3]).reshape(n_env, -1).max(axis=1),
                         np.abs(rda[..., 3:7] - ro[..., 3:7]).reshape(n_env, -1).max(axis=1))
    print("%s: pose error per env %s" % (scene, np.sort(per_env)[[0, n_env // 2, -2, -1]]))
    assert not st.any()
    walls = [i for i in range(model.n_bodies) if i not in dyn]
    assert np.allclose(rd[:, walls], rods[:, walls], rtol=0, atol=1e-13)  # stored relative to the env frame (pinned to the first static body)
    # round-off level except for the environments that sit on a face / edge switch of the sampling scheme (the wall's top edge
    # under a rod: 3 of 12 environments at 1e-8 .. 4e-7 after 250 steps)
    assert np.sum(per_env < 1e-8) >= (2 * n_env) // 3 and per_env.max() < 1e-4


# ------------------------------------------------------------------------------------------------ contact: analytic known answers
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_sliding_ball_and_incline_known_answers_on_device(tmp_path, precision):
    """The analytic pins of the contact half (tests/test_cpu.py: Coulomb deceleration and 5/7 v0 rolling speed of a sliding ball; a brick
    sticks below the friction angle and slides with g (sin theta - mu cos theta) above it) asked of the CUDA path itself."""
    import math
    from test_cpu import BALL, BRICK, _build_text, sliding_ball_expectations, incline, brick_on_incline_state
    m = _build_text(tmp_path, "ball.yaml", BALL)
    mu, g, t_roll, v_roll = sliding_ball_expectations(m)
    n_env = 8
    v0 = np.linspace(6.0, 13.0, n_env)  # every environment its own launch speed
    rods = np.tile(m.initial_rod_state(), (n_env, 1, 1))
    rods[:, 0, 7] = v0
    steps = int(3.0 * (2.0 * v0.max() / (7.0 * mu * g)) / m.dt)
    rd, _, st = device_rollout(m, rods, np.zeros((n_env, 0, 2)), steps, precision)
    assert not st.any()
    assert np.allclose(rd[:, 0, 7], 5.0 / 7.0 * v0, rtol=0.01)
    assert np.allclose(rd[:, 0, 12], -rd[:, 0, 7] / 0.5, rtol=0.01)
    for theta, slides in ((0.30, False), (0.60, True)):
        b = incline(_build_text(tmp_path, "brick.yaml", BRICK), theta)
        mub = b.bodies[0].friction * b.ground.friction
        start = np.tile(brick_on_incline_state(b, theta), (4, 1, 1))
        w = nt.BatchedWorld(b, 4, precision=precision)
        try:
            w.set_rod_state(start)
            w.step(b.dt, 100)
            s0 = w.get_rod_state()
            w.step(b.dt, 200)
            s1 = w.get_rod_state()
        finally:
            w.close()
        down = np.array([-math.cos(theta), -math.sin(theta), 0.0])
        dv = (s1[:, 0, 7:10] - s0[:, 0, 7:10]) @ down
        if slides:
            a = b.gravity * (math.sin(theta) - mub * math.cos(theta))
            assert np.allclose(dv, a * 0.2, rtol=0.03), dv
        else:
            assert np.abs((s1[:, 0, 0:3] - s0[:, 0, 0:3]) @ down).max() < 0.01 and np.abs(s1[:, 0, 7:10] @ down).max() < 0.05


def test_neuroevolution_generations_on_device():
    """The learning outer loop end to end: a population of TetraSpine controllers (2 node + 57 x 2 edge actions in
    [0, 1]) evolved for three generations, every generation's episodes evaluated in one batched launch sequence."""
    from ntrtsim_b200.neuroevo import NeuroEvolution, evaluate_population
    model, ctl = nt.build_example("TetraSpine"), nt.tetraspine_cpg()
    sizes = []

    def evaluate(rows):
        sizes.append(len(rows))
        params = nt.tetraspine_param_rows(rows[:, :2], rows[:, 2:].reshape(len(rows), 57, 2))
        scores, status = evaluate_population(model, ctl, params, n_steps=400)
        assert np.isfinite(scores).all()
        return scores
    evo = NeuroEvolution(n_params=2 + 114, population_size=24, number_of_elements_to_mutate=8, seed=11)
    first = evo.step(evaluate)
    best0 = evo.population.max_score[0]
    assert best0 == first[:, 0].max()
    evo.step(evaluate)
    evo.step(evaluate)
    assert sizes == [24, 8, 8]
    assert evo.population.max_score[0] >= best0 or evo.population.max_score[0] >= 0.5 * (best0 + first[:, 0].min())
    assert len(evo.log) == 3 and evo.log[0][0] == 24


# ------------------------------------------------------------------------------------------------ contact between bodies
# tgWorldBulletPhysicsImpl.cpp:82-107: every pair of bodies of the world collides.  tgb_pairs.cuh against the oracle's
# generateBodyPairContacts, and the known answers of tests/test_cpu.py asked of the CUDA path.
def _text_model(tmp_path, name, text):
    from test_cpu import _build_text
    return _build_text(tmp_path, name, text)


@pytest.mark.parametrize("precision,steps", [("f64", 400), ("f32", 150)])
def test_rod_on_static_rods_matches_oracle(tmp_path, precision, steps):
    """a rod dropped across two static rods: dynamic x static pair rows next to nothing else (the rod never reaches the ground)"""
    from test_cpu import RAILS
    model = _text_model(tmp_path, "rails.yaml", RAILS)
    n_env = 8
    rng = np.random.default_rng(7)
    rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    rods[:, 2, 0] += rng.uniform(-0.5, 0.5, size=n_env)   # where it crosses the rails
    rods[:, 2, 1] += rng.uniform(0.0, 1.0, size=n_env)
    cab = np.zeros((n_env, 0, 2))
    w = nt.BatchedWorld(model, n_env, precision=precision)
    w.set_rod_state(rods)
    w.step(model.dt, steps)
    rd, clr = w.get_rod_state(), w.min_clearance()
    w.close()
    ro = np.empty_like(rd)
    clr_o = np.empty(n_env)
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_rod_state(rods[e])
        o.step(model.dt, steps)
        ro[e], clr_o[e] = o.get_rod_state(), o.min_clearance()
    e = rod_errors(rd, ro, 1.0)
    print("rod on rails %s %d steps: %s clearance dev %s oracle %s" % (precision, steps, e, clr, clr_o))
    tol = 1e-7 if precision == "f64" else 2e-3
    assert max(e[0], e[1]) < tol and max(e[2], e[3]) < tol * 100, e
    # the clearance is exact whenever a pair is within the contact band; further away the device keeps the distance of the last
    # time it had to test the pair (tgb_pairs.cuh: slack), which is never smaller than the true minimum
    band = clr_o < 0.02
    assert band.sum() >= 6
    assert np.abs(clr - clr_o)[band].max() < (1e-6 if precision == "f64" else 1e-3)
    assert (clr[~band] >= clr_o[~band] - 1e-3).all()
    assert (ro[:, 2, 1] > 2.9).all() and (rd[:, 2, 1] > 2.9).all()   # resting on the rails (ground top face at 1.5, rails at 2.0)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_rod_rod_impact_conserves_momentum_on_device(tmp_path, precision):
    from test_cpu import CROSSED_RODS
    model = _text_model(tmp_path, "crossed.yaml", CROSSED_RODS)
    model.ground.kind = capi.TGB_GROUND_EMPTY
    model.gravity = 0.0
    n_env = 5
    rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    rods[:, 1, 8] = -8.0
    rods[:, 1, 0] += np.linspace(-1.0, 1.0, n_env)         # hit at different places
    mass = np.array([b.mass for b in model.bodies])
    w = nt.BatchedWorld(model, n_env, precision=precision)
    w.set_rod_state(rods)
    w.step(model.dt, 400)
    rd, clr = w.get_rod_state(), w.min_clearance()
    w.close()
    p0 = (mass[None, :, None] * rods[:, :, 7:10]).sum(1)
    p1 = (mass[None, :, None] * rd[:, :, 7:10]).sum(1)
    print("rod-rod impact %s: momentum drift %.3e, clearance %s" % (precision, np.abs(p1 - p0).max(), clr))
    assert np.abs(p1 - p0).max() < (1e-9 if precision == "f64" else 2e-4) * np.abs(p0).max()
    assert (rd[:, 0, 8] < -1.0).all() and (rd[:, 1, 8] > -7.0).all()   # the struck rod moves off, the striker slows down
    assert (clr < 0.021).all() and (clr > -0.06).all()                # they touched and did not pass through each other
    if precision == "f64":
        ro = np.stack([_oracle_states(model, rods[e], 400) for e in range(n_env)])
        assert max(rod_errors(rd, ro, 1.0)) < 1e-7


def _oracle_states(model, rods, steps, cab=None, ctrl=0, params=None, cap=None):
    o = oracle_py.OracleWorld(model, capi.WorldCfg)
    if cap is not None:
        o.set_pair_contact_cap(cap)
    o.set_rod_state(rods)
    if cab is not None:
        o.set_cable_state(cab)
    if ctrl:
        o.set_controller(ctrl, params)
    o.step(model.dt, steps)
    return o.get_rod_state()


def test_prism_rods_meet_f64():
    """config 2 beyond the contact-free part: the pretensioned strings fold the 3-prism until its rods meet (step ~180 for the model
    pose); the double path follows the oracle through the rod-rod contact, and the clearance never goes below the solver's slop."""
    model = nt.build_example("3_prism_noground")
    n_env, steps = 24, 450
    rods, cab = random_poses(model, n_env, seed=1234), tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.step(model.dt, steps)
    rd, cd, clr = w.get_rod_state(), w.get_cable_obs(), w.min_clearance()
    w.close()
    ro, co = oracle_rollout(model, rods, cab, steps)
    e = rod_errors(rd, ro, 1.0)
    print("prism rods meet f64: %s cables %.3e clearance min %.4f max %.4f" % (e, np.abs(cd - co).max(), clr.min(), clr.max()))
    assert (clr < 0.02).all() and (clr > -0.08).all()
    assert max(e) < 1e-6, e
    # switched off (ABI <= 3 behaviour) the rods pass through each other
    w = nt.BatchedWorld(model, n_env, precision="f64", pair_contact_cap=0)
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.step(model.dt, steps)
    r0 = w.get_rod_state()
    w.close()
    ro0 = np.stack([_oracle_states(model, rods[e], steps, cab[e], cap=0) for e in range(n_env)])
    assert max(rod_errors(r0, ro0, 1.0)) < 1e-7


def test_bench_workload_rods_keep_their_clearance():
    """on the benchmark workload (seed 2345) rods used to pass through each other in ~8 % of the environments within
    one launch of 1000 steps.  With the pair contacts no environment goes deeper than the solver's slop, in single precision at the
    benchmark's own settings; the double path agrees with the oracle on the environments that had a rod-rod contact."""
    sys.path.insert(0, ROOT)
    import bench
    n_env, steps = 256, 1000
    model, rods, cab, ctrl, params, _ = bench.build_workload("superball", n_env, bench.SEED)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_controller(ctrl, params)
    w.step(model.dt, steps)
    clr32 = w.min_clearance()
    w.close()
    w = nt.BatchedWorld(model, n_env, precision="f32", pair_contact_cap=-1)   # the test runs, no row is generated
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_controller(ctrl, params)
    w.step(model.dt, steps)
    clr_off = w.min_clearance()
    w.close()
    print("bench workload, %d envs x %d steps: clearance < 0 in %d envs (pairs on) / %d (off); deepest %.3f / %.3f"
          % (n_env, steps, (clr32 < 0).sum(), (clr_off < 0).sum(), clr32.min(), clr_off.min()))
    assert clr32.min() > -0.3 and (clr32 < -0.1).sum() <= 2   # squeezed by the cables a contact gives a little (10 PGS iterations, erp 0.2)
    assert (clr_off < -0.4).sum() >= 5      # the hole this closes: rods more than a radius inside each other
    # double precision against the oracle on the environments whose rods met within the first 400 steps
    steps64 = 400
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_controller(ctrl, params)
    w.step(model.dt, steps64)
    rd, clr64 = w.get_rod_state(), w.min_clearance()
    w.close()
    met = np.where(clr64 < 0.02)[0]
    assert len(met) >= 3, len(met)
    ro = np.stack([_oracle_states(model, rods[e], steps64, cab[e], ctrl, params[e]) for e in met])
    e = rod_errors(rd[met], ro, 1.0)
    # rod-rod and rod-ground contacts together amplify round-off (see test_prism_on_ground_f64): the yardstick is the oracle's own
    # response to a 1e-13 perturbation of its input
    pert = rods[met].copy()
    pert[:, :, 0:3] += 1e-13 * np.random.default_rng(0).normal(size=pert[:, :, 0:3].shape)
    rp = np.stack([_oracle_states(model, pert[k], steps64, cab[e], ctrl, params[e]) for k, e in enumerate(met)])
    own = rod_errors(rp, ro, 1.0)
    print("bench workload f64, %d envs with a rod-rod contact within %d steps: device vs oracle %s, oracle vs perturbed oracle %s"
          % (len(met), steps64, e, own))
    assert max(e) < max(1e-6, 20.0 * max(own)), (e, own)
    # and tight agreement while the contacts are young: 150 steps
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_controller(ctrl, params)
    w.step(model.dt, 150)
    rd150, clr150 = w.get_rod_state(), w.min_clearance()
    w.close()
    met150 = np.where(clr150 < 0.02)[0]
    ro150 = np.stack([_oracle_states(model, rods[e], 150, cab[e], ctrl, params[e]) for e in met150])
    e150 = rod_errors(rd150[met150], ro150, 1.0)
    print("bench workload f64, %d envs with a rod-rod contact within 150 steps: %s" % (len(met150), e150))
    assert len(met150) >= 1 and max(e150) < 1e-7, e150


# ------------------------------------------------------------------------------------------------ persistent manifolds, warm starting
@pytest.mark.parametrize("example,ctrl,steps", [("Escape_c3", capi.TGB_CTRL_ESCAPE_SINE, 220), ("TetraSpineStatic_hf", capi.TGB_CTRL_SERIALIZED_SPINE, 150),
                                                ("3_prism", 0, 225)])
def test_warm_starting_matches_oracle_f64(example, ctrl, steps):
    """tgb_set_warm_starting(0.85) = Bullet's sequential-impulse default (btContactSolverInfo::m_warmstartingFactor over persistent
    manifold points) against the oracle with the same factor: single-contact rod ends, the many-point feet of the spine, and the
    3-prism whose rods meet (pair rows) -- double precision, round-off level; and it does change the trajectory."""
    model = nt.build_example(example)
    n_env = 12
    if example == "TetraSpineStatic_hf":
        rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
        params = np.tile(nt.serialized_spine_params(model), (n_env, 1))
    else:
        rods = random_poses(model, n_env, seed=21, translate=0.5, lift=(0.0, 2.0), yaw_only=True)
        params = nt.escape_transform_actions(np.random.default_rng(22).uniform(size=(n_env, 32)), 9.0) if ctrl else None
    cab = tiled_cables(model, n_env)
    out = {}
    for warm in (0.85, 0.0):
        w = nt.BatchedWorld(model, n_env, precision="f64", warm_starting=warm)
        w.set_rod_state(rods)
        w.set_cable_state(cab)
        if ctrl:
            w.set_controller(ctrl, params)
        w.step(model.dt, steps // 2)
        w.step(model.dt, steps - steps // 2)   # the manifolds travel between launches
        out[warm] = w.get_rod_state()
        w.close()
    ro = np.empty_like(out[0.85])
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_warm_start(0.85)
        o.set_rod_state(rods[e])
        o.set_cable_state(cab[e])
        if ctrl:
            o.set_controller(ctrl, params[e])
        o.step(model.dt, steps)
        ro[e] = o.get_rod_state()
    e = rod_errors(out[0.85], ro, 1.0)
    changed = max(rod_errors(out[0.85], out[0.0], 1.0))
    print("warm starting %s %d steps: device vs oracle %s; warm vs cold on the device %.3e" % (example, steps, e, changed))
    assert max(e) < 1e-7, e
    assert changed > 1e-6


def test_warm_starting_f32_and_layouts():
    """single precision with warm starting: FP32 tolerance against the oracle over a short contact horizon, and identical results
    whatever the lanes per environment and however the episode is cut into launches"""
    model = nt.build_example("Escape_c3")
    n_env, steps = 64, 150
    rods, cab, params = escape_inputs(model, n_env, seed=31)
    ref = device_rollout(model, rods, cab, steps, "f32", capi.TGB_CTRL_ESCAPE_SINE, params, wide=False, warm_starting=0.85)
    for kw in (dict(lanes_per_env=16, wide=False), dict(chunks=5, wide=False), dict(block_threads=128, wide=False)):
        got = device_rollout(model, rods, cab, steps, "f32", capi.TGB_CTRL_ESCAPE_SINE, params, warm_starting=0.85, **kw)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), kw
    ro = np.empty_like(ref[0])
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_warm_start(0.85)
        o.set_rod_state(rods[e])
        o.set_cable_state(cab[e])
        o.set_controller(capi.TGB_CTRL_ESCAPE_SINE, params[e])
        o.step(model.dt, steps)
        ro[e] = o.get_rod_state()
    per_env = np.abs(align_quat_sign(ref[0], ro)[..., 0:3] - ro[..., 0:3]).max(axis=(1, 2)) / 17.0
    print("warm starting f32: fraction within 1e-4 %.3f, worst %.3e" % (np.mean(per_env < F32_REL), per_env.max()))
    assert np.mean(per_env < F32_REL) >= 0.95 and per_env.max() < 1e-2   # 62 of 64 measured; the others are contacts switching a step apart


def test_superball_f32_contact_parity_statistics_full_size():
    """BASELINE config 3 at its full size (4096 environments, the benchmark's seed and controller draws), single precision against the
    oracle: what fraction of the environments stays within the FP32 tolerance (1e-4 of the structure size / of max(|T|, 100)) after
    200 steps and after the benchmark's 1000 steps per launch, and what the others have in common.  Contact makes trajectories
    sensitive (a manifold point appearing a step earlier changes the impulses), so this is a statistic, asserted as such."""
    sys.path.insert(0, ROOT)
    import bench
    n_env = 4096
    model, rods, cab, ctrl, params, _ = bench.build_workload("superball", n_env, bench.SEED)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_controller(ctrl, params)
    out = {}
    done = 0
    for steps in (200, 1000):
        w.step(model.dt, steps - done)
        done = steps
        out[steps] = (w.get_rod_state(), w.get_cable_obs(), w.min_clearance())
    assert not (w.status() & 1).any()
    w.close()
    for steps in (200, 1000):
        rd, cd, clr = out[steps]
        ro, co = oracle_rollout(model, rods, cab, steps, ctrl, params, threads=os.cpu_count() or 8)
        pos = np.abs(align_quat_sign(rd, ro)[..., 0:3] - ro[..., 0:3]).max(axis=(1, 2)) / 17.0
        ten = (np.abs(cd[..., 2] - co[..., 2]) / np.maximum(np.abs(co[..., 2]), 100.0)).max(axis=1)
        worst = np.maximum(pos, ten)
        touched = clr < 0.02
        frac = [float(np.mean(worst < t)) for t in (1e-4, 1e-3, 1e-2)]
        print("config 3, 4096 envs, f32 vs oracle after %d steps: within 1e-4 %.4f, 1e-3 %.4f, 1e-2 %.4f; median %.2e; envs with a rod-rod "
              "contact so far %d, of the %d envs beyond 1e-4 %d had one" % (steps, frac[0], frac[1], frac[2], np.median(worst), int(touched.sum()),
                                                                           int((worst >= 1e-4).sum()), int((touched & (worst >= 1e-4)).sum())))
        # the yardstick for the long horizon: the oracle against itself with its input moved by one single-precision ulp of the structure
        # size (1e-6) -- what any single-precision computation of these trajectories can be expected to hold
        pert = rods.copy()
        pert[:, :, 0:3] += 1e-6 * np.random.default_rng(1).normal(size=pert[:, :, 0:3].shape)
        rp, cp = oracle_rollout(model, pert, cab, steps, ctrl, params, threads=os.cpu_count() or 8)
        pos_o = np.abs(align_quat_sign(rp, ro)[..., 0:3] - ro[..., 0:3]).max(axis=(1, 2)) / 17.0
        ten_o = (np.abs(cp[..., 2] - co[..., 2]) / np.maximum(np.abs(co[..., 2]), 100.0)).max(axis=1)
        own = np.maximum(pos_o, ten_o)
        frac_own = [float(np.mean(own < t)) for t in (1e-4, 1e-3, 1e-2)]
        print("    oracle vs oracle with its input moved by 1e-6: within 1e-4 %.4f, 1e-3 %.4f, 1e-2 %.4f; median %.2e"
              % (frac_own[0], frac_own[1], frac_own[2], np.median(own)))
        if steps == 200:
            assert frac[0] >= 0.985 and frac[2] >= 0.999
        else:
            # The north star asks for the tolerance over a short horizon (above) and over the full horizon for contact-free rollouts
            # (test_prism_drop_f32_full_horizon).  With contact the full horizon is a statistic: measured 48 % / 79 % / 98 % within
            # 1e-4 / 1e-3 / 1e-2 (median 1.1e-4), against 79 % / 91 % / 99 % for the oracle after ONE perturbation of 1e-6 at the start --
            # single precision perturbs every step.  Held here: the bulk stays bounded.
            assert frac[1] >= 0.70 and frac[2] >= 0.95 and np.median(worst) < 5e-4


@pytest.mark.parametrize("ground", [True, False])
def test_compound_bodies_touch_matches_oracle_f64(ground):
    """Contact between COMPOUND bodies (TetraSpine: three bodies of six cylinders): the middle vertebra starts pushed 11.5 ... 12.5 into
    the first one's cage, so rods of the two meet within ~30 steps.  With the ground every body also stands on several points (the
    shared-memory coupled solver); without it the pair rows are alone (the register solvers).  Out-of-line emission over all pairs of
    shapes in both.  Double precision against the oracle over the collision."""
    n_env, steps = 8, 60
    model = nt.build_example("TetraSpine")
    if not ground:
        model.ground.kind = capi.TGB_GROUND_EMPTY
        model.gravity = 0.0
    rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
    rods[:, 1, 2] += np.linspace(11.5, 12.5, n_env)
    cab = tiled_cables(model, n_env)
    w = nt.BatchedWorld(model, n_env, precision="f64")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.step(model.dt, steps)
    rd, clr = w.get_rod_state(), w.min_clearance()
    w.close()
    ro = np.empty_like(rd)
    clr_o, rows = np.empty(n_env), np.zeros(n_env, dtype=int)
    for e in range(n_env):
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_rod_state(rods[e])
        o.set_cable_state(cab[e])
        for _ in range(steps):
            o.step(model.dt, 1)
            rows[e] = max(rows[e], o.num_pair_contacts())
        ro[e], clr_o[e] = o.get_rod_state(), o.min_clearance()
    e = rod_errors(rd, ro, 1.0)
    print("compound bodies, ground %s: most pair rows per env %s, clearance oracle %s device %s, rods %s" % (ground, rows, np.round(clr_o, 4), np.round(clr, 4), e))
    assert (rows >= 1).sum() >= 4 and rows.max() >= 2
    assert np.abs(clr - clr_o).max() < 1e-6
    assert max(e) < 1e-7, e


def test_population_members_vertebrae_contact_statistics():
    """BASELINE config 4's workload (mutated TetraSpine controllers) in single precision, 512 members x 6000 steps: in most members the
    vertebrae never meet; where they do they rest against each other (clearance ~0).  One member of this draw (250) saturates its
    cables at the 7000-unit tension limit and jams its vertebrae into each other -- more force than four rows per pair of bodies and ten
    iterations hold; the oracle does exactly the same (clearance -1.27, with 32 rows allowed still -0.8), so it is the model's
    behaviour at that operating point, not a device defect: asserted as agreement with the oracle on the worst members."""
    sys.path.insert(0, ROOT)
    import bench
    n_env, steps = 512, 6000
    model, rods, cab, ctrl, params, _ = bench.build_workload("tetraspine", n_env, bench.SEED)
    w = nt.BatchedWorld(model, n_env, precision="f32")
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    w.set_cpg_controller(model.cpg, params)
    w.step(model.dt, steps)
    clr, st = w.min_clearance(), w.status()
    w.close()
    worst = np.argsort(clr)[:3]
    clr_o = []
    for e in worst:
        o = oracle_py.OracleWorld(model, capi.WorldCfg)
        o.set_rod_state(rods[e])
        o.set_cable_state(cab[e])
        o.set_cpg_controller(model.cpg.desc, params[e])
        o.step(model.dt, steps)
        clr_o.append(o.min_clearance())
    print("tetraspine population, %d members x %d steps: vertebrae touched in %d, deeper than 0.1 in %d; worst members %s device %s oracle %s"
          % (n_env, steps, int((clr < 0.02).sum()), int((clr < -0.1).sum()), worst, np.round(clr[worst], 3), np.round(clr_o, 3)))
    assert int((clr < -0.1).sum()) <= 3
    assert np.abs(clr[worst] - np.array(clr_o)).max() < 0.15


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_tail_split_is_bit_identical(precision):
    """The tail split of long launches (tgb_set_tail_split: environments whose rods leaned on each other in the last launch get a warp
    of their own on a second stream, the others run from a list in the usual layout) changes durations only: rod states, cable
    observations, clearances and the world's status are bit for bit those of unsplit launches -- through tgb_step and through
    tgb_rollout_host, in the forced mode (every launch, threshold of one step) and in the automatic one."""
    sys.path.insert(0, ROOT)
    import bench
    n_env, steps = 640, 250
    model, rods, cab, ctrl, params, _ = bench.build_workload("superball", n_env, bench.SEED)

    def run(split, host):
        w = nt.BatchedWorld(model, n_env, precision=precision, tail_split=split, wide=False)   # (the split is for the narrow layouts)
        w.set_controller(ctrl, params)
        if host:
            r, c = rods.copy(), cab.copy()
            for _ in range(3):
                ro, co = w.rollout_host(r, c, model.dt, steps)
                r = ro
                c = np.ascontiguousarray(co[..., 0:2])
        else:
            w.set_rod_state(rods)
            w.set_cable_state(cab)
            for _ in range(3):
                w.step(model.dt, steps)
            ro, co = w.get_rod_state(), w.get_cable_obs()
        # (the clearance is measured on the steps a pair is tested on; beyond the contact band those depend on the warp's other pairs)
        out = (ro, co, np.minimum(w.min_clearance(), 0.02), w.status(), w.launch_count())
        w.close()
        return out

    for host in (False, True):
        ref = run(0, host)
        assert (ref[2] <= 0.02).sum() >= 5          # rods do meet in this sample
        for mode in (1, -1):
            got = run(mode, host)
            for a, b in zip(ref[:4], got[:4]):
                assert np.array_equal(a, b)
            if mode == 1:
                assert got[4] > ref[4]               # the split launches ran: lists + two kernels per launch
