"""The oracle against golden vectors produced by the REFERENCE's own sources (oracle/ref/make_golden.py: the cable /
anchor / actuator / tgTensionController code of /root/reference compiled unchanged; fixtures in tests/golden/).
This pins pieces (1) cables and (2) motors + controllers of the oracle to the reference's arithmetic."""
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
from helpers import align_quat_sign  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = [
    ("prism_drop.npz", "3_prism_noground", capi.TGB_CTRL_NONE, None),
    ("motor_basic_dt0.001.npz", "tsTestRig_basic", capi.TGB_CTRL_SETPOINT, "command"),
    ("motor_basic_dt0.0002.npz", "tsTestRig_basic", capi.TGB_CTRL_SETPOINT, "command"),
    ("motor_kinematic_dt0.001.npz", "tsTestRig_kinematic", capi.TGB_CTRL_TORQUE, "command"),
    ("motor_kinematic_dt0.0002.npz", "tsTestRig_kinematic", capi.TGB_CTRL_TORQUE, "command"),
    ("superball_tension.npz", "SUPERball", capi.TGB_CTRL_TENSION, "set_point"),
    ("escape_sine.npz", "Escape_c3", capi.TGB_CTRL_ESCAPE_SINE, "params"),
]


def golden_case(fixture, example, ctrl, param_key):
    g = np.load(os.path.join(GOLDEN, fixture))
    model = nt.build_example(example)
    params = None
    if param_key is not None:
        params = np.atleast_1d(g[param_key]).astype(np.float64)
        if ctrl in (capi.TGB_CTRL_SETPOINT, capi.TGB_CTRL_TORQUE):
            params = np.full(model.n_cables, params[0])
    return g, model, params


@pytest.mark.parametrize("fixture,example,ctrl,param_key", CASES)
def test_oracle_reproduces_reference_trace(fixture, example, ctrl, param_key):
    g, model, params = golden_case(fixture, example, ctrl, param_key)
    o = oracle_py.OracleWorld(model, capi.WorldCfg)
    o.set_rod_state(g["rods0"])
    if ctrl:
        o.set_controller(ctrl, params)
    dt = float(g["dt"])
    done = 0
    worst = 0.0
    for k, s in enumerate(g["steps"]):
        o.step(dt, int(s) - done)
        done = int(s)
        rods = align_quat_sign(o.get_rod_state(), g["rods"][k])
        obs = o.get_cable_obs()
        ref = g["cables"][k]   # actuator rest, cable rest, length, tension, cable velocity, actuator velocity, start length, damping
        scale = max(1.0, np.abs(g["rods"][k][:, 0:3]).max())
        worst = max(worst, np.abs(rods - g["rods"][k]).max() / scale)
        assert np.abs(rods - g["rods"][k]).max() / scale < 1e-9, (fixture, s)
        assert np.abs(obs[:, 0] - ref[:, 0]).max() < 1e-9 and np.abs(obs[:, 0] - ref[:, 1]).max() < 1e-9, (fixture, s, "rest length")
        assert np.abs(obs[:, 1] - ref[:, 2]).max() < 1e-9, (fixture, s, "length")
        assert (np.abs(obs[:, 2] - ref[:, 3]) / np.maximum(1.0, np.abs(ref[:, 3]))).max() < 1e-9, (fixture, s, "tension")
    print(fixture, "worst relative rod deviation %.2e over %d steps" % (worst, done))
