#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched tensegrity step on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], SURVEY.md section 8d "C3"): 4096 environments per GPU of the 24-cable SUPERball
(craterEscape EscapeModel: 6 rods, 24 tgBasicActuator cables), flat tgBoxGround, g = 98.1, dt = 1e-3, the EscapeController
sine controller with per-env parameters U(0,1)^32 mapped by transformActions (L0 = 9), start pose = model pose + U(0,5)
extra height + random yaw, seed 2345.  One bench "step" = one pass of the hot path over the batch = ONE call
tgb_step(world, dt, H) that advances every environment H = --sim-steps simulation steps (env state stays in
registers/shared memory for the H steps; HBM is touched at the launch boundaries).

  value  env-steps/s with the state resident in HBM, CUDA events on the world's stream, L2 flushed between steps
  e2e    the same episodes through tgb_rollout_host with pinned HOST buffers: H2D of rod + cable state, H steps, D2H of rod
         state + cable observations, all inside the timed region; call i+1 starts from the host copy of call i's results
  roofline  algorithmic bytes (2*4*(13R+2C)+4*P_ctrl = 1136 B per env-step) x env-steps per launch /
            launch duration against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (FP64 restatement of the NTRTsim/Bullet-2.82 path; Bullet is not in the reference tree,
                so the reference itself cannot be built: kind "port") on all host cores, bounded sample

python bench.py --impl reference times that CPU port alone, all host threads, same workload/metric.

Other workloads (--workload prism | tetraspine | flemons | crater | superball_noground, --envs, --precision f64) are side lines for
DESIGN.md / profiles/; under torchrun (--gpus N) every rank runs its own 4096 environments and the per-rollout scores are all-gathered
once per bench step, overlapping the next step (see the N > 1 branch below).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "env-steps/sec (SUPERball, 4096 envs) at 1/2/4/8 B200; % HBM roofline"
UNIT = "env-steps/s"
SEED = 2345


def build_workload(name, n_env, seed):
    import ntrtsim_b200 as nt
    from ntrtsim_b200 import _capi as capi
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import random_poses
    if name == "superball":
        model = nt.build_example("Escape_c3")
        rods = random_poses(model, n_env, seed=seed, translate=1.0, lift=(0.0, 5.0), yaw_only=True)
        rng = np.random.default_rng(seed + 1)
        params = nt.escape_transform_actions(rng.uniform(size=(n_env, 32)), 9.0)
        ctrl = capi.TGB_CTRL_ESCAPE_SINE
        p_ctrl = 32
    elif name == "superball_noground":   # tuning aid: the same workload without the ground (upper bound for contact-side work)
        model = nt.build_example("Escape_c3")
        model.ground = capi.GroundDesc(capi.TGB_GROUND_EMPTY, (0.0, 1.0, 0.0), 0.0, 0.5, 0.0, 0.0, 0.04)
        model.gravity = 0.0
        rods = random_poses(model, n_env, seed=seed, translate=1.0, lift=(0.0, 5.0), yaw_only=True)
        rng = np.random.default_rng(seed + 1)
        params = nt.escape_transform_actions(rng.uniform(size=(n_env, 32)), 9.0)
        ctrl = capi.TGB_CTRL_ESCAPE_SINE
        p_ctrl = 32
    elif name == "crater":
        # AppEscape complete: the SUPERball dropped onto the rim of tgCraterDeep (four static box walls), sine controller
        model = nt.build_example("Escape_crater")
        dyn = [i for i, b in enumerate(model.bodies) if b.mass != 0.0]
        rng = np.random.default_rng(seed)
        rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
        rods[:, dyn, 0:3] += rng.uniform(-0.3, 0.3, size=(n_env, 1, 3)) + np.array([0.0, 14.0, 0.0])
        params = nt.escape_transform_actions(rng.uniform(size=(n_env, 32)), 9.0)
        ctrl = capi.TGB_CTRL_ESCAPE_SINE
        p_ctrl = 32
    elif name == "contact_cable":
        # ContactCableDemo as test_integration/MuscleNP runs it (launch impulse (0.4, 0, 0.4) on the two end bodies),
        # an ensemble of launches (+- 15 % and a small turn); the cable meets the free rod between steps 3000 and 5000 and bends around it
        # over sliding anchors.  TGB_CC_BENCH_SPEED=4 launches four times as fast: the cable then goes slack and winds itself around its
        # own end rod with up to 12 anchors (a harder case for the one-thread-per-environment kernel: DESIGN.md section 3.7)
        model = nt.build_example("ContactCableDemo")
        rng = np.random.default_rng(seed)
        rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
        speed = float(os.environ.get("TGB_CC_BENCH_SPEED", "1"))
        imp = speed * (np.array([0.4, 0.0, 0.4]) * (1.0 + 0.15 * rng.uniform(-1, 1, size=(n_env, 1))) + 0.06 * rng.uniform(-1, 1, size=(n_env, 3)))
        for b in (0, 1):
            rods[:, b, 7:10] = imp / model.bodies[b].mass
        params, ctrl, p_ctrl = None, capi.TGB_CTRL_NONE, 2 * (5 + 7 * 12)   # the cable's state incl. its sliding anchors, read and written
    elif name == "prism":
        model = nt.build_example("3_prism_noground")
        rods = random_poses(model, n_env, seed=seed)
        params, ctrl, p_ctrl = None, capi.TGB_CTRL_NONE, 0
    elif name in ("tetraspine", "flemons"):
        # NeuroEvo population evaluation (BASELINE config 4 / SURVEY 8f rank 1): every environment is one population member, the
        # shipped parameters mutated like NeuroEvoMember::mutate (N(0, 0.03), clipped to [0, 1]); all members start from the model pose
        example = "TetraSpine" if name == "tetraspine" else "FlemonsSpine"
        model = nt.build_example(example)
        g = np.load(os.path.join(ROOT, "tests", "golden", "tetraspine_cpg.npz" if name == "tetraspine" else "flemonsspine_kinematic_cpg.npz"))
        rng = np.random.default_rng(seed + 1)
        n_edge = g["edge_actions"].shape[0]
        node = np.clip(g["node_actions"][None] + rng.normal(0, 0.03, size=(n_env, 2)), 0, 1)
        edge = np.clip(g["edge_actions"][None] + rng.normal(0, 0.03, size=(n_env, n_edge, 2)), 0, 1)
        if name == "tetraspine":
            model.cpg, params = nt.tetraspine_cpg(), nt.tetraspine_param_rows(node, edge)
        else:
            model.cpg, params = nt.flemons_cpg(example), nt.flemons_param_rows(node, edge, example)
        rods = np.tile(model.initial_rod_state(), (n_env, 1, 1))
        ctrl = capi.TGB_CTRL_CPG_IMPEDANCE
        p_ctrl = 2 * 3 * model.n_cables   # CPG node state (phase, radius, radius rate) read and written every step
    else:
        raise ValueError(name)
    cab = np.tile(model.initial_cable_state(), (n_env, 1, 1))
    b_alg = 2 * 4 * (13 * model.n_bodies + 2 * model.n_cables) + 4 * p_ctrl
    return model, rods, cab, ctrl, params, b_alg


def install_controller(w, model, ctrl, params):
    from ntrtsim_b200 import _capi as capi
    if ctrl == capi.TGB_CTRL_CPG_IMPEDANCE:
        w.set_cpg_controller(model.cpg, params)
    elif ctrl:
        w.set_controller(ctrl, params)


def workload_desc(name, n_env, H, n_gpus):
    what = {"superball_noground": "SUPERball + sine controller, NO ground, zero gravity (tuning aid)",
            "superball": "SUPERball (EscapeModel: 6 rods, 24 basic-actuator cables) + EscapeController sine, flat box ground, g=98.1",
            "prism": "3-prism (3 rods, 9 cables), no ground, g=981 (contact-free drop)",
            "contact_cable": "ContactCableDemo (rod + box joined by one tgBulletContactSpringCable that runs into a free rod), no ground, g=0 "
                             "(MuscleNP_test); one step-kernel launch + one contact-cable launch per sim step",
            "crater": "SUPERball (EscapeModel) + EscapeController sine on tgCraterDeep (4 static box walls), box ground, g=98.1 (AppEscape)",
            "tetraspine": "TetraSpine (3 compound bodies, 12 basic-actuator cables) + CPG/impedance control, population evaluation, box ground, g=981",
            "flemons": "kinematic tetrahedral-complex spine (12 compound bodies, 88 kinematic-actuator cables) + CPG/impedance/PID control, "
                       "population evaluation, box ground, g=981"}[name]
    return "%s; %d envs/GPU x %d GPU; dt=%s; %d sim steps per launch" % (what, n_env, n_gpus, "1/600" if name == "crater" else "1e-3", H)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            self.thread = threading.Thread(target=self._read)
            self.thread.daemon = True
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(workload, H):
    """dram bytes per launch of the step kernel from the committed ncu capture, if one matches this configuration."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            key = "%s_H%s" % (workload, H) if not str(H).startswith("_") else workload + str(H)
            if key in d:
                return d[key]
        except Exception:
            pass
    return None


def ncu_warp_instructions_per_env_step(workload):
    """warp instructions of the step kernel per env-step (committed ncu capture of this workload at 4096 environments), or None"""
    return ncu_traffic_per_launch(workload, "_warp_instructions_per_env_step")


def cpu_port_rate(model, rods, cab, ctrl, params, H, n_env_sample, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    from ntrtsim_b200 import _capi as capi
    p = None if params is None else params[:n_env_sample]
    t0 = time.perf_counter()
    oracle_py.rollout_batch(model, capi.WorldCfg, rods[:n_env_sample], cab[:n_env_sample], model.dt, H, ctrl, p, n_threads=threads,
                            cpg_desc=model.cpg.desc if ctrl == capi.TGB_CTRL_CPG_IMPEDANCE else None)
    dt = time.perf_counter() - t0
    return n_env_sample * H / dt, dt


def run_reference(args):
    """The reference arm: the CPU port of the path (oracle) with every host thread, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ntrtsim_b200.build import build_all
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        build_all()
    cores = os.cpu_count() or 1
    H = args.sim_steps
    sample_env = min(args.envs, cores * 32)
    model, rods, cab, ctrl, params, _ = build_workload(args.workload, sample_env, SEED)
    for _ in range(args.warmup):
        cpu_port_rate(model, rods, cab, ctrl, params, H, sample_env, cores)
    total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_port_rate(model, rods, cab, ctrl, params, H, sample_env, cores)
        total += dt
    value = sample_env * H * args.steps / total
    sample = "%d envs x %d sim steps per step, %d threads (first envs of the same seeded workload)" % (sample_env, H, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_desc(args.workload, args.envs, H, args.gpus)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """Libraries (NCCL prints its version banner) write to stdout: hand them stderr and keep the real stdout for the one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--sim-steps", type=int, default=1000, help="simulation steps per launch (H)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--workload", default="superball", choices=["superball", "prism", "superball_noground", "tetraspine", "flemons", "crater", "contact_cable"])
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import ntrtsim_b200 as nt
    from ntrtsim_b200 import _capi as capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available() or capi.load().tgb_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the step has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_env, H = args.envs, args.sim_steps
    # env-axis sharding: rank r owns envs [r*n_env, (r+1)*n_env) of the global seeded batch; no exchange inside an episode
    model, rods_all, cab_all, ctrl, params_all, b_alg = build_workload(args.workload, n_env * world_size, SEED)
    sl = slice(rank * n_env, (rank + 1) * n_env)
    rods, cab = np.ascontiguousarray(rods_all[sl]), np.ascontiguousarray(cab_all[sl])
    params = None if params_all is None else np.ascontiguousarray(params_all[sl])

    w = nt.BatchedWorld(model, n_env, precision=args.precision, device=local_rank, lanes_per_env=args.lanes, block_threads=args.threads)
    w.set_rod_state(rods)
    w.set_cable_state(cab)
    install_controller(w, model, ctrl, params)
    lanes, block, grid = w.launch_geometry()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.fill_(rank & 0xff)
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # the path's only collective: allgather of per-rollout scores after an episode chunk (N > 1)
    scores_all = None
    if dist is not None:
        scores_all = torch.empty(world_size * n_env * 2, dtype=torch.float32, device="cuda")

    def gather_scores():
        ptr, nbytes = w.scores_device_ptr()

        class _Wrap(object):
            __cuda_array_interface__ = {"shape": (n_env * 2,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        local = torch.as_tensor(_Wrap(), device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(scores_all, local)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    # ---- warm-up ---------------------------------------------------------------------------------------------
    flush_l2()  # the first use of torch's fill kernel loads it (~20 ms on the host): not inside the timed bracket
    for _ in range(args.warmup):
        w.step_timed(model.dt, H, 1)
        if dist is not None:
            gather_scores()

    # ---- timed region: device-resident state ---------------------------------------------------------------------
    # The same method at every N: ONE CUDA-event bracket on the world's stream around the K steps, L2 flushes included; the host
    # queues everything and never blocks inside it.  Stream order per step k: [torch stream] wait for step k-1, flush L2;
    # [world stream] wait for the flush, H sim steps (events around the launch give the kernel's own duration), and for N > 1:
    # wait for all-gather k-1 (its source buffer is about to be rewritten), refresh the score buffer; [collective stream] wait
    # for the refresh, all-gather k.  The all-gather of step k thus overlaps step k+1, ranks only have to stay within one step of
    # each other, and every collective completes before the bracket closes.
    launches0 = w.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t_wall0 = time.perf_counter()
    ws = torch.cuda.ExternalStream(w.stream_handle())
    cs = torch.cuda.Stream() if dist is not None else None
    cur = torch.cuda.current_stream()
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_g0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_g1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    gather_done, step_done, keep = None, None, []
    ev_start.record(ws)
    for k in range(args.steps):
        if step_done is not None:
            cur.wait_event(step_done)
        flush.fill_(rank & 0xff)
        fl = torch.cuda.Event()
        fl.record(cur)
        ws.wait_event(fl)
        ev_k0[k].record(ws)
        w.step(model.dt, H)                       # asynchronous on the world's stream
        ev_k1[k].record(ws)
        if dist is None:
            step_done = torch.cuda.Event()
            step_done.record(ws)
            keep.append(fl)
            continue
        if gather_done is not None:
            ws.wait_event(gather_done)
        ptr, nbytes = w.scores_refresh_async()    # scores of step k, queued on the world's stream
        step_done = torch.cuda.Event()
        step_done.record(ws)

        class _Wrap(object):
            __cuda_array_interface__ = {"shape": (n_env * 2,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        with torch.cuda.stream(cs):
            cs.wait_event(step_done)
            local = torch.as_tensor(_Wrap(), device="cuda")
            ev_g0[k].record(cs)
            work = dist.all_gather_into_tensor(scores_all, local, async_op=True)
            work.wait()                           # orders the collective stream after NCCL's; does not block the host
            ev_g1[k].record(cs)
            gather_done = torch.cuda.Event()
            gather_done.record(cs)
        keep.append((local, work, fl))
    if gather_done is not None:
        ws.wait_event(gather_done)
    ev_end.record(ws)
    ev_end.synchronize()
    torch.cuda.synchronize()
    ms_kernel = [a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)]
    ms_gather = [a.elapsed_time(b) for a, b in zip(ev_g0, ev_g1)] if dist is not None else []
    total_ms = ev_start.elapsed_time(ev_end)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = w.launch_count() - launches0
    status_bad = int((w.status() & 1).sum())
    # contact between bodies (rod x rod): the smallest surface distance between two rods every environment has seen over the warm-up and
    # timed launches, measured by the step kernel itself on every step
    clr = w.min_clearance()
    radius = float(model.shapes[0].radius)
    body_body = {"pair_contacts": "on (tgb_pairs.cuh: every pair of bodies tested every step)", "min_clearance": float(clr.min()),
                 "envs_touching": int((clr <= 0.02).sum()), "envs_deeper_than_solver_slop": int((clr < -0.1).sum()),
                 "envs_deeper_than_one_radius": int((clr < -radius).sum()), "envs": int(n_env),
                 "sim_steps_seen": int(H * (args.steps + args.warmup))}

    # ---- end to end through the C ABI with pinned host buffers ---------------------------------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = a
        return t, v
    rods_t, rods_p = pinned(rods)
    cab_t, cab_p = pinned(cab)
    rod_out_t, rod_out = pinned(np.zeros((n_env, w.R, capi.ROD_STATE_DOUBLES)))
    cab_out_t, cab_out = pinned(np.zeros((n_env, w.C, capi.CABLE_OBS_DOUBLES)))
    h2d = rods_p.nbytes + cab_p.nbytes
    d2h = rod_out.nbytes + cab_out.nbytes
    # The end-to-end arm replays the same episodes as the device-resident arm: it restarts from the seeded initial state and
    # controller time, and every call's inputs are the previous call's outputs (rod records; cable rest length and length),
    # handed over on the host.  Warm-up calls = the warm-up episodes above, so timed call i simulates the same second of
    # each rollout as timed step i above.  Each timed call moves that step's inputs host->device and its results back.
    install_controller(w, model, ctrl, params)

    def hand_over():
        rods_p[...] = rod_out
        cab_p[..., 0] = cab_out[..., 0]  # rest length
        cab_p[..., 1] = cab_out[..., 1]  # current length = the next step's previous length
    for _ in range(args.warmup):
        w.rollout_host(rods_p, cab_p, model.dt, H, rod_out, cab_out)
        hand_over()
    barrier()
    e2e_s = []
    for _ in range(args.steps):
        flush_l2()
        t0 = time.perf_counter()
        w.rollout_host(rods_p, cab_p, model.dt, H, rod_out, cab_out)  # synchronises its stream before returning
        e2e_s.append(time.perf_counter() - t0)
        hand_over()
    barrier()
    e2e_total = sum(e2e_s)
    e2e_checksum = float(np.abs(cab_out[..., 2]).mean()) if w.C else float(np.abs(rod_out[..., 7:10]).mean())

    # max over ranks; the per-rank kernel durations travel too, so that the line names what limits the scaling
    per_rank_kernel = [ms_kernel]
    per_rank_total = [total_ms]
    if dist is not None:
        mine = torch.tensor(ms_kernel + [total_ms, float(np.mean(ms_gather))], dtype=torch.float64, device="cuda")
        everyone = [torch.empty_like(mine) for _ in range(world_size)]
        dist.all_gather(everyone, mine)
        per_rank_kernel = [[float(x) for x in t[:args.steps]] for t in everyone]
        per_rank_total = [float(t[args.steps]) for t in everyone]
        per_rank_gather = [float(t[args.steps + 1]) for t in everyone]
        t = torch.tensor([total_ms, e2e_total, float(launches), float(status_bad)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(t[0]), float(t[1])

    if rank == 0:
        env_steps = float(world_size) * n_env * H * args.steps
        value = env_steps / (total_ms * 1e-3)
        e2e_value = env_steps / e2e_total
        peak, peak_src = measured_hbm_peak()
        k_ms = float(np.mean(ms_kernel))
        achieved = b_alg * n_env * H / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_desc(args.workload, n_env, H, world_size), "envs_per_gpu": n_env, "sim_steps_per_launch": H,
                       "lanes_per_env": lanes, "block_threads": block, "grid": grid, "l2": "flushed between timed steps (256 MiB write)",
                       "seed": SEED, "wall_s_timed_region": t_wall, "nonfinite_envs": status_bad},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_total / args.steps, "checksum_mean_tension": e2e_checksum},
            "gpu_launches": int(launches),
            "body_body": body_body,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(args.workload, H), "kernel": "tgb_step_kernel<%s,%d>" % (args.precision, lanes),
                         "kernel_ms": k_ms, "algorithmic_bytes_per_env_step": b_alg, "env_steps_per_launch": n_env * H,
                         "peak_source": peak_src,
                         "note": "effective figure: env state stays on chip for the H fused steps, real DRAM traffic is ~1/H of it"},
        }
        # what actually binds this kernel: instruction issue.  warp instructions per env-step from the committed ncu capture of this
        # configuration x measured env-steps/s against the SMs' issue rate (4 schedulers x 1 warp instruction per cycle per SM)
        per_env_step = ncu_warp_instructions_per_env_step(args.workload)
        if per_env_step is not None and clocks and clocks.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            issue_peak = sms * 4 * clocks["sm_mhz"] * 1e6
            rate = n_env * H / (k_ms * 1e-3)
            line["roofline"]["issue"] = {"warp_instr_per_env_step": per_env_step, "issue_peak_warp_instr_per_s": issue_peak,
                                         "achieved_warp_instr_per_s": per_env_step * rate, "frac": per_env_step * rate / issue_peak,
                                         "note": "the binding limit (HBM carries ~1/H of the algorithmic bytes): fraction of the SMs' issue slots; "
                                                 "instruction count from profiles/ (ncu, same command), rate measured here"}
        line["config"]["timing"] = "one CUDA-event bracket on the world's stream around the K steps incl. the L2 flushes, same at every N; max over ranks"
        allk = np.array(per_rank_kernel)
        line["per_rank_kernel_ms"] = {"min": float(allk.min()), "median": float(np.median(allk)), "max": float(allk.max()),
                                      "mean_by_rank": [float(x) for x in allk.mean(axis=1)],
                                      "sum_of_step_max_over_ranks": float(allk.max(axis=0).sum()),
                                      "bracket_ms_by_rank": per_rank_total}
        if world_size > 1:
            line["config"]["allgather"] = "scores all-gather of step k issued asynchronously, overlapping step k+1, completed inside the timed bracket"
            line["allgather_ms"] = {"mean_by_rank": per_rank_gather, "bytes_per_rank": int(n_env * 2 * 4),
                                    "note": "device time of the collective on its own stream; overlaps the next step's kernel"}
        if world_size == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            # calibrate, then size the sample (whole batch, H x reps sim steps) for ~10-15 s of CPU work
            cal_env = min(n_env, cores * 2)
            rate, _ = cpu_port_rate(model, rods, cab, ctrl, params, min(H, 200), cal_env, cores)
            sample_env = int(min(n_env, max(cores, rate * 20.0 / H)))
            sample_env = max(cores, sample_env // cores * cores)
            reps = int(max(1, min(50, rate * 20.0 / (sample_env * H))))
            rate, secs = cpu_port_rate(model, rods, cab, ctrl, params, H * reps, sample_env, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d envs x %d sim steps of the same workload, %d threads, %.1f s (FP64 CPU restatement of the "
                                              "NTRTsim/Bullet-2.82 path; Bullet is absent so the reference cannot be built)"
                                              % (sample_env, H * reps, cores, secs)}
        emit(line)
    w.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
