#!/usr/bin/env python
"""Aggregate an ncu SASS source page per phase of tgb_step_kernel. Phase boundaries are found from the marker comments in
tgb_step.cuh. usage: ncu_by_phase.py sass.csv disasm.txt kernel_substring n_warps n_steps"""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKERS = [("world step (tgSimulation", "gravity+gyro"), ("ground contact manifold (analytic", "contact gen+rows"),
           ("projected Gauss-Seidel: split", "PGS"), ("write back, split-impulse", "writeback+integrate"),
           ("env frame: advance", "frame+publish"), ("models step (tgSimulation", "cable: pose+state loads"),
           ("model-level controllers (run before", "cable: controllers+motors"), ("tgKinematicActuator::integrateRestLength, tg", "cable: kinematic"),
           ("tgBulletSpringCable::calculateAndApplyForce, tg", "cable: force"), ("applyImpulse of every incident", "gather"),
           ("store: bodies back into the record", "epilogue")]


def main():
    sass_csv, disasm, kname = sys.argv[1:4]
    n_warps, n_steps = float(sys.argv[4]), float(sys.argv[5])
    src = open(os.path.join(ROOT, "ntrtsim_b200", "csrc", "tgb_step.cuh")).read().split("\n")
    bounds = []
    for i, l in enumerate(src, 1):
        for key, name in MARKERS:
            if key in l and "/*" in l and not l.strip().startswith("*"):
                bounds.append((i, name))
    bounds.sort()
    loop_start = [i for i, l in enumerate(src, 1) if "for (int step = 0; step < " in l][0]

    def phase_of(n):
        if n < loop_start:
            return "prologue"
        cur = "prologue"
        for b, name in bounds:
            if n >= b:
                cur = name
        return cur
    lines = open(disasm, errors="replace").read().split("\n")
    in_k, cur, off2 = False, ("?", 0), {}
    for l in lines:
        if l.startswith("//---") and ".text." in l:
            in_k = kname in l
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            for ff, nn in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
                if ff.endswith("tgb_step.cuh"):
                    cur = ("tgb_step.cuh", int(nn))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            off2[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hi]
    ia, iex, ith, ismp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    base, agg, tot = None, defaultdict(lambda: [0, 0, 0]), [0, 0]
    for r in rows[hi + 1:]:
        if len(r) <= ismp or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        f, n = off2.get(a - base, ("?", 0))
        key = phase_of(n) if f == "tgb_step.cuh" else "inlined:" + f
        ex, th, sm = int(r[iex] or 0), int(r[ith] or 0), int(r[ismp] or 0)
        agg[key][0] += ex; agg[key][1] += th; agg[key][2] += sm
        tot[0] += ex; tot[1] += sm
    print("%-30s %7s %8s %6s %s" % ("phase", "instr%", "samples%", "lanes", "warp-instr per warp per step"))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-30s %6.1f%% %7.1f%% %6.1f %8.0f" % (k, 100.0 * v[0] / tot[0], 100.0 * v[2] / max(1, tot[1]), v[1] / max(1, v[0]), v[0] / n_warps / n_steps))
    print("total %.0f warp-instr per warp per step" % (tot[0] / n_warps / n_steps))


if __name__ == "__main__":
    main()
