#!/usr/bin/env python
"""Split the stall samples / executed instructions of an ncu SASS source page between the kernel body and the device functions
ptxas placed behind it (noinline callees, library slow paths).
usage: ncu_by_function.py sass.csv cubin kernel_mangled_substring"""
import csv
import re
import subprocess
import sys


def main():
    sass_csv, cubin, kname = sys.argv[1:4]
    dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    in_k = False
    labels = []  # (offset, name)
    last_off = -16
    pending = None
    for l in dis:
        if l.startswith("//---") and ".text." in l:
            in_k = kname in l
            continue
        if not in_k:
            continue
        m = re.match(r"^(\$?[_A-Za-z][^\s:]*):\s*$", l)
        if m and not m.group(1).startswith(".L"):
            pending = m.group(1)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            last_off = int(m.group(1), 16)
            if pending:
                labels.append((last_off, pending))
                pending = None
    rows = list(csv.reader(open(sass_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hi]
    ia, iex, ithr, ismp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    base = None
    agg = {}
    tot = [0, 0]
    for r in rows[hi + 1:]:
        if len(r) <= ismp or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        off = a - base
        name = "?"
        for o, n in labels:
            if o <= off:
                name = n
        short = name.split("$")[-1] if "$" in name else "kernel body"
        short = re.sub(r"^_ZN3tgb\d+", "", short)[:40]
        e = agg.setdefault(short, [0, 0, 0])
        e[0] += int(r[iex] or 0); e[1] += int(r[ithr] or 0); e[2] += int(r[ismp] or 0)
        tot[0] += int(r[iex] or 0); tot[1] += int(r[ismp] or 0)
    for n, (ex, th, sm) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print("%-42s instr %6.2f%%  samples %6.2f%%  lanes %.1f" % (n, 100.0 * ex / max(1, tot[0]), 100.0 * sm / max(1, tot[1]), th / max(1, ex)))


if __name__ == "__main__":
    main()
