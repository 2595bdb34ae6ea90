#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -c -gi line
info of the same cubin and aggregate executed instructions / stall samples per source line.
usage: ncu_by_line.py sass.csv disasm.txt kernel_mangled_substring [top_n]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, disasm, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # nvdisasm: offset -> (file:line of the outermost non-inlined location in tgb_step.cuh when available)
    lines = open(disasm, errors="replace").read().split("\n")
    in_k, cur, cur_outer = False, "?", "?"
    off2line = {}
    for i, l in enumerate(lines):
        if l.startswith("//---") and ".text." in l:
            in_k = kname in l
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            f = m.group(1).split("/")[-1]
            cur = "%s:%s" % (f, m.group(2))
            rest = m.group(3)
            m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
            cur_outer = cur
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+) inlined at', l)
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
    rows = list(csv.reader(open(sass_csv)))
    # find header row
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hi]
    ia, isrc = h.index("Address"), h.index("Source")
    iex, ithr = h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    ismp = h.index("Warp Stall Sampling (All Samples)")
    stall_cols = [(j, n) for j, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    iwf = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
    iwx = h.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in h else None
    smem = defaultdict(lambda: [0, 0, ""])
    base = None
    agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot = [0, 0, 0]
    stall_tot = defaultdict(int)
    opagg = defaultdict(lambda: [0, 0])
    for r in rows[hi + 1:]:
        if len(r) <= ismp or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        off = a - base
        line, _ = off2line.get(off, ("?", ""))
        ex, th, sm = int(r[iex] or 0), int(r[ithr] or 0), int(r[ismp] or 0)
        e = agg[line]
        e[0] += ex; e[1] += th; e[2] += sm
        tot[0] += ex; tot[1] += th; tot[2] += sm
        for j, n in stall_cols:
            v = int(r[j] or 0)
            e[3][n] += v
            stall_tot[n] += v
        if iwf is not None and (r[iwf] or "0") not in ("0", "-"):
            w = smem[(line, r[isrc].strip().split(";")[0][:48])]
            w[0] += int(r[iwf] or 0); w[1] += int(r[iwx] or 0)
        op = r[isrc].strip().split()[0] if r[isrc].strip() else "?"
        if op.startswith("@"):
            op = r[isrc].strip().split()[1]
        op = op.split(".")[0]
        opagg[op][0] += ex; opagg[op][1] += sm
    print("total warp-instructions %d, thread-instr/instr %.2f, samples %d" % (tot[0], tot[1] / max(1, tot[0]), tot[2]))
    print("stall totals:", ", ".join("%s=%.1f%%" % (n[6:], 100.0 * v / max(1, tot[2])) for n, v in sorted(stall_tot.items(), key=lambda x: -x[1])[:10]))
    print("\n-- top opcodes by executed instructions")
    for op, (ex, sm) in sorted(opagg.items(), key=lambda x: -x[1][0])[:25]:
        print("  %-10s %6.2f%% instr  %6.2f%% samples" % (op, 100.0 * ex / tot[0], 100.0 * sm / max(1, tot[2])))
    print("\n-- top source lines by stall samples")
    for line, e in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
        st = ", ".join("%s %d" % (n[6:], v) for n, v in sorted(e[3].items(), key=lambda x: -x[1])[:3] if v)
        print("  %-22s %6.2f%% instr  %6.2f%% samples  lanes %.1f   [%s]" % (line, 100.0 * e[0] / tot[0], 100.0 * e[2] / max(1, tot[2]), e[1] / max(1, e[0]), st))
    if smem:
        tw, tx = sum(v[0] for v in smem.values()), sum(v[1] for v in smem.values())
        print("\n-- shared-memory wavefronts: %d total, %d excessive (bank conflicts); top instructions by excessive wavefronts" % (tw, tx))
        for (line, src), v in sorted(smem.items(), key=lambda x: -x[1][1])[:top]:
            if v[1]:
                print("  %-22s %-50s wavefronts %10d  excessive %10d (%.0f%%)" % (line, src, v[0], v[1], 100.0 * v[1] / max(1, v[0])))


if __name__ == "__main__":
    main()
