#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own sources (oracle/_ref/libntrt_ref.so, built by build_ref.py).

Each trace is a hybrid rollout: the world half of tgSimulation::step (Bullet, absent from the reference tree) comes from
the oracle's restatement, the models half (controllers -> tgBasicActuator / tgKinematicActuator::step ->
tgBulletSpringCable::calculateAndApplyForce, anchors, tgTensionController) is executed by reference code.  The pure oracle
must reproduce these traces (tests/test_golden.py), which pins pieces (1) cables and (2) motors/controllers to the
reference's arithmetic.  Run in the container that has /root/reference; the fixtures travel, the reference does not."""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), HERE, os.path.join(ROOT, "tests")]

import ntrtsim_b200 as nt  # noqa: E402
from ntrtsim_b200 import _capi as capi  # noqa: E402
import oracle_py  # noqa: E402
import ref_py  # noqa: E402
import build_ref  # noqa: E402
from helpers import random_poses  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def hybrid_rollout(model, pretension, rods0, dt, steps, controller=None, keep=None):
    """controller(ref_models, step_index, dt) runs the model-level observer before the actuators step."""
    ow = oracle_py.OracleWorld(model, capi.WorldCfg)
    ow.set_rod_state(rods0)
    ref = ref_py.RefModels(model, pretension, rods0)
    # the reference constructors must reproduce the host builder's rest lengths
    rest_builder = np.array([c.rest_length for c in model.cables])
    assert np.abs(ref.cables()[:, 0] - rest_builder).max() < 1e-9, (ref.cables()[:, 0], rest_builder)
    lib = oracle_py.load()
    keep = set(range(steps)) if keep is None else set(keep)
    trace_steps, rods, cabs = [], [], []
    for s in range(steps):
        lib.tgo_world_step(ow._h, dt)
        st = ow.get_rod_state()
        ref.set_bodies(st)
        if controller is not None:
            controller(ref, s, dt)
        rc = ref.step_actuators(dt)
        assert rc == 0
        st2 = ref.get_bodies()
        ow.set_rod_state(st2)
        if s in keep:
            trace_steps.append(s + 1)
            rods.append(st2.copy())
            cabs.append(ref.cables().copy())
    ref.close()
    return np.array(trace_steps), np.array(rods), np.array(cabs)


def sample_steps(steps):
    return sorted(set(list(range(0, 20)) + list(range(20, steps, 10)) + [steps - 1]))


def flemons_golden():
    """Kinematic tetrahedral-complex spine: CPGEquations + tgImpedanceController + tgPIDController + tgKinematicActuator (all reference
    code) under the tgCPGCableControl timing (tgCPGCableControl.cpp:63-98), for the 3-segment and the 12-segment spine."""
    res = os.path.join(os.environ.get("NTRT_REFERENCE", "/root/reference"), "resources", "src", "learningSpines", "TetrahedralComplex", "logs")

    def nnw(name):
        return [float(x) for x in open(os.path.join(res, name)).read().strip().split(",")]
    node_actions = np.array(nnw("bestParameters-default_node-0.nnw"))
    edge_actions = np.array([nnw("bestParameters-default_edge-%d.nnw" % i) for i in range(100)])
    for name, steps in (("FlemonsSpine3", 400), ("FlemonsSpine", 300)):
        m = nt.build_example(name)
        ctl = nt.flemons_cpg(name)
        row = nt.flemons_param_rows(node_actions[None], edge_actions[None], name)[0]
        d = ctl.desc
        cpg = ref_py.RefCPG(d.max_steps)
        nC = m.n_cables
        for _ in range(nC):
            cpg.add_node([row[0], row[0], row[1], row[1], d.r_const, d.d_min, d.d_max])                    # assignNodeNumber
        for i in range(nC):                                                                                # setConnectivity
            ks = range(ctl.begin[i], ctl.begin[i + 1])
            cpg.define_connections(i, [ctl.target[k] for k in ks], [row[2 + 2 * ctl.pair[k]] for k in ks], [row[3 + 2 * ctl.pair[k]] for k in ks])
        st = {"update": 0.0, "cable": [0.0] * nC, "commanded": [0.0] * nC, "bogus": False, "cpg": [], "made": False}

        def spine(ref, s, dt_, st=st, ctl=ctl, d=d, cpg=cpg, nC=nC):
            if not st["made"]:                                         # tgCPGCableControl::onSetup
                for i in range(nC):
                    assert ref.pid_create(i, 5000.0, 0.0, 10.0, True) == 0
                st["made"] = True
            st["update"] += dt_                                        # BaseSpineCPGControl::onStep
            if st["update"] >= d.control_time:
                assert cpg.update([d.descending_command] * nC, st["update"]) == 0
                st["update"] = 0.0
            h = ref.get_bodies()[d.score_body, 1]
            if h > d.height_hi or h < d.height_lo:
                st["bogus"] = True
            vals = cpg.values()
            for i in range(nC):                                        # tgCPGCableControl::onStep (usePID)
                st["cable"][i] += dt_
                if st["cable"][i] >= d.cable_control_step:
                    g = ctl.impedance[i]
                    st["commanded"][i] = ref.impedance_control_pid(i, dt_, g[0], g[1], g[2], g[3], vals[i])
                    st["cable"][i] = 0.0
                else:
                    ref.pid_control(i, dt_, st["commanded"][i])
            st["cpg"].append(cpg.state())
        keep = sample_steps(steps)
        ts, rods, cabs = hybrid_rollout(m, 0.0, m.initial_rod_state(), 1e-3, steps, spine, keep=keep)
        np.savez_compressed(os.path.join(OUT, "%s_kinematic_cpg.npz" % name.lower()), rods0=m.initial_rod_state(), dt=1e-3, steps=ts, rods=rods,
                            cables=cabs, node_actions=node_actions, edge_actions=edge_actions, param_row=row,
                            cpg_states=np.array(st["cpg"])[keep], bogus=st["bogus"])
        print(name, "hybrid rollout done: bogus", st["bogus"], "max tension", cabs[..., 3].max(), "rest range", cabs[..., 0].min(), cabs[..., 0].max())


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "flemons":
        if not ref_py.available():
            build_ref.build()
        flemons_golden()
        return
    if not ref_py.available():
        build_ref.build()
    os.makedirs(OUT, exist_ok=True)

    # ---- 1. CPG known answer (test/util/CPGEquations_test.cpp:72-169) + a trace of small updates ---------------
    def cpg_system():
        cpg = ref_py.RefCPG(5000)
        for _ in range(3):
            cpg.add_node([1.0, 0.0, 1.0, 0.0, 20.0, 0.0, 5.0])
        cpg.define_connections(0, [1, 2], [1.0, 1.0], [math.pi / 2.0, 0.0])
        cpg.define_connections(1, [0, 2], [1.0, 1.0], [math.pi / 2.0, math.pi / 2.0])
        cpg.define_connections(2, [0, 1], [1.0, 1.0], [0.0, math.pi / 2.0])
        return cpg
    a, b = cpg_system(), cpg_system()
    assert a.update([0.0] * 3, 20.0) == 0
    va = a.values()
    trace = []
    for _ in range(100):
        assert b.update([0.0] * 3, 0.2) == 0
        trace.append(b.state())
    vb = b.values()
    print("CPG known answer (reference code + odeint restatement):", va, vb)
    assert abs(va[0]) < 1e-7 and abs(va[1] + 1.0) < 1e-7 and abs(va[2]) < 1e-7       # the reference test's own assertions
    assert np.abs(va - vb).max() < 1e-7
    # a 12-node ring with random weights, stepped like BaseSpineCPGControl does (update(descCom = 2, dt = 1e-3))
    rng = np.random.default_rng(4)
    c = ref_py.RefCPG(200)
    n = 12
    node_params = [0.18588, 0.18588, 0.992645, 0.992645, 20.0, 0.0, 5.0]
    for _ in range(n):
        c.add_node(node_params)
    conn, wts, phs = [], [], []
    for i in range(n):
        nb = [(i + 1) % n, (i - 1) % n, (i + 5) % n]
        w, p = rng.uniform(0, 1, 3), rng.uniform(-math.pi, math.pi, 3)
        c.define_connections(i, nb, w, p)
        conn.append(nb); wts.append(w); phs.append(p)
    small = []
    for _ in range(400):
        assert c.update([2.0] * n, 1e-3) == 0
        small.append(c.state())
    np.savez_compressed(os.path.join(OUT, "cpg.npz"), known_values_20s=va, known_values_100x02=vb, trace_100x02=np.array(trace),
                        ring_node_params=np.array(node_params), ring_conn=np.array(conn), ring_weights=np.array(wts),
                        ring_phases=np.array(phs), ring_trace=np.array(small)[::10], ring_final=small[-1])

    # ---- 2. cable / motor traces -------------------------------------------------------------------------------------
    m = nt.build_example("3_prism_noground")
    rods0 = random_poses(m, 1, seed=1234)[0]
    steps = 300
    ts, rods, cabs = hybrid_rollout(m, 500.0, rods0, 1e-3, steps, keep=sample_steps(steps))
    np.savez_compressed(os.path.join(OUT, "prism_drop.npz"), rods0=rods0, dt=1e-3, steps=ts, rods=rods, cables=cabs)

    for kind, cmd in (("basic", 5.0), ("kinematic", -580.0)):
        m = nt.build_example("tsTestRig_" + kind)
        for dt in (1e-3, 2e-4):
            steps = int(round(0.3 / dt))

            def ctl(ref, s, dt_, kind=kind, cmd=cmd):      # tsTestRig::step, tsTestRig.cpp:181-188
                if kind == "basic":
                    assert ref.set_control_input_dt(0, cmd, dt_) == 0
                else:
                    ref.set_control_input(0, cmd)
            ts, rods, cabs = hybrid_rollout(m, 0.0, m.initial_rod_state(), dt, steps, ctl, keep=sample_steps(steps))
            np.savez_compressed(os.path.join(OUT, "motor_%s_dt%g.npz" % (kind, dt)), rods0=m.initial_rod_state(), dt=dt, steps=ts,
                                rods=rods, cables=cabs, command=cmd)

    m = nt.build_example("SUPERball")
    rods0 = m.initial_rod_state()
    steps = 300

    def t6(ref, s, dt_):                                    # T6TensionController::onStep, T6TensionController.cpp:69-83
        for c in range(ref.C):
            assert ref.tension_control(c, dt_, 1000.0) == 0
    ts, rods, cabs = hybrid_rollout(m, 2452.0, rods0, 1e-3, steps, t6, keep=sample_steps(steps))
    np.savez_compressed(os.path.join(OUT, "superball_tension.npz"), rods0=rods0, dt=1e-3, steps=ts, rods=rods, cables=cabs, set_point=1000.0)

    m = nt.build_example("Escape_c3")
    rods0 = random_poses(m, 1, seed=2345, translate=1.0, lift=(0.0, 5.0), yaw_only=True)[0]
    params = nt.escape_transform_actions(np.random.default_rng(2346).uniform(size=32), 9.0)
    steps = 300
    state = {"t": 0.0}

    def sine(ref, s, dt_):          # EscapeController::onStep + setPreferredMuscleLengths (EscapeController.cpp:100-117,245-291)
        state["t"] += dt_
        phase, cluster = 0.0, 0
        for i in range(24):
            old = cluster
            cluster = 0 if i < 4 else 1 if i < 8 else 2 if i < 12 else 3 if i < 16 else 4 if i < 18 else 5 if i < 20 else 6 if i < 22 else 7
            new_len = params[cluster * 4 + 0] * math.sin(params[cluster * 4 + 1] * state["t"] + phase) + params[cluster * 4 + 3]
            lo, hi = params[32] * (1 - 0.50), params[32] * (1 + 0.50)
            new_len = lo if new_len <= lo else hi if new_len >= hi else new_len
            assert ref.set_control_input_dt(i, new_len, dt_) == 0
            if old != cluster:
                phase += params[cluster * 4 + 2]
        for i in range(ref.C):
            ref.move_motors(i, dt_)
    ts, rods, cabs = hybrid_rollout(m, 0.0, rods0, 1e-3, steps, sine, keep=sample_steps(steps))
    np.savez_compressed(os.path.join(OUT, "escape_sine.npz"), rods0=rods0, dt=1e-3, steps=ts, rods=rods, cables=cabs, params=params)
    # ---- 3. TetraSpine: CPGEquations + tgImpedanceController (reference code) + tgCPGActuatorControl timing --------------
    m = nt.build_example("TetraSpine")
    ctl = nt.tetraspine_cpg()
    res = os.path.join(os.environ.get("NTRT_REFERENCE", "/root/reference"), "resources", "src", "learningSpines", "TetraSpine", "logs")

    def nnw(name):
        return [float(x) for x in open(os.path.join(res, name)).read().strip().split(",")]
    node_actions = np.array(nnw("bestParameters-default_node-0.nnw"))
    edge_actions = np.array([nnw("bestParameters-default_edge-%d.nnw" % i) for i in range(57)])
    row = nt.tetraspine_param_rows(node_actions[None], edge_actions[None])[0]
    cpg = ref_py.RefCPG(ctl.desc.max_steps)
    nC = m.n_cables
    for _ in range(nC):
        cpg.add_node([row[0], row[0], row[1], row[1], ctl.desc.r_const, ctl.desc.d_min, ctl.desc.d_max])  # assignNodeNumber
    for i in range(nC):                                                                                   # setConnectivity
        ks = range(ctl.begin[i], ctl.begin[i + 1])
        cpg.define_connections(i, [ctl.target[k] for k in ks], [row[2 + 2 * ctl.pair[k]] for k in ks], [row[3 + 2 * ctl.pair[k]] for k in ks])
    st = {"update": 0.0, "cable": [0.0] * nC, "bogus": False, "cpg": []}

    def spine(ref, s, dt_):
        st["update"] += dt_                                        # BaseSpineCPGControl::onStep
        if st["update"] >= ctl.desc.control_time:
            assert cpg.update([ctl.desc.descending_command] * nC, st["update"]) == 0
            st["update"] = 0.0
        h = ref.get_bodies()[ctl.desc.score_body, 1]
        if h > ctl.desc.height_hi or h < ctl.desc.height_lo:
            st["bogus"] = True
        vals = cpg.values()
        for i in range(nC):                                        # tgCPGActuatorControl::onStep
            st["cable"][i] += dt_
            if st["cable"][i] >= ctl.desc.cable_control_step:
                g = ctl.impedance[i]
                ref.impedance_control(i, st["cable"][i], g[0], g[1], g[2], g[3], vals[i])
                st["cable"][i] = 0.0
            else:
                ref.move_motors(i, dt_)
        st["cpg"].append(cpg.state())
    steps = 400
    keep = sample_steps(steps)
    ts, rods, cabs = hybrid_rollout(m, 0.0, m.initial_rod_state(), 1e-3, steps, spine, keep=keep)
    np.savez_compressed(os.path.join(OUT, "tetraspine_cpg.npz"), rods0=m.initial_rod_state(), dt=1e-3, steps=ts, rods=rods, cables=cabs,
                        node_actions=node_actions, edge_actions=edge_actions, param_row=row, cpg_states=np.array(st["cpg"])[keep],
                        bogus=st["bogus"])
    flemons_golden()
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-28s %6d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
