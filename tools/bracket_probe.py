"""Where does the time between the steps of bench.py's bracket go?  (tuning aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import numpy as np
import ntrtsim_b200 as nt
import bench

n_env, H, K = 4096, 1000, 5
model, rods, cab, ctrl, params, b_alg = bench.build_workload("superball", n_env, bench.SEED)
w = nt.BatchedWorld(model, n_env, precision="f32")
w.set_rod_state(rods); w.set_cable_state(cab); w.set_controller(ctrl, params)
for _ in range(3):
    w.step_timed(model.dt, H, 1)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ws = torch.cuda.ExternalStream(w.stream_handle())
cur = torch.cuda.current_stream()
for mode in ("steps only", "flush on torch stream + waits", "flush, host sync per step"):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t0 = time.perf_counter()
    e0.record(ws)
    done = None
    host = []
    for k in range(K):
        th = time.perf_counter()
        if mode != "steps only":
            if done is not None:
                cur.wait_event(done)
            evs[k][2].record(cur)
            flush.fill_(1)
            evs[k][3].record(cur)
            fl = torch.cuda.Event(); fl.record(cur); ws.wait_event(fl)
        evs[k][0].record(ws)
        w.step(model.dt, H)
        evs[k][1].record(ws)
        done = torch.cuda.Event(); done.record(ws)
        if mode.endswith("per step"):
            torch.cuda.synchronize()
        host.append(time.perf_counter() - th)
    e1.record(ws)
    e1.synchronize(); torch.cuda.synchronize()
    print(mode, "bracket %.3f ms, wall %.3f ms" % (e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)),
          "kernels", ["%.2f" % a.elapsed_time(b) for a, b, _, _ in evs],
          "flushes", ["%.2f" % c.elapsed_time(d) for _, _, c, d in evs] if mode != "steps only" else "",
          "host ms per iteration", ["%.2f" % (1e3 * h) for h in host])
w.close()
