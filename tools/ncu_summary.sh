#!/bin/bash
# Summarise an ncu --set full capture of the step kernel into a text file for profiles/:
#   tools/ncu_summary.sh gpurun_out/prof.ncu-rep IfLi8ELi1ELb1 tgb_step_f32_sine profiles/ncu_step_kernel_rNN.txt "header line"
# (kernel mangled-name substring, the translation unit whose cubin holds it). Needs the same build in ntrtsim_b200/lib.
set -e
rep=$1; kname=$2; tu=$3; out=$4; header=$5
root=$(cd "$(dirname "$0")/.." && pwd)
work=$(mktemp -d)
ncu -i $rep --page raw --csv 2>/dev/null > $work/raw.csv
ncu -i $rep --page source --csv --print-source sass 2>/dev/null > $work/sass.csv
(cd $work && cuobjdump -xelf all $root/ntrtsim_b200/lib/libntrtsim_b200.so > /dev/null 2>&1 && nvdisasm -c -gi $tu.sm_100a.cubin > disasm.txt 2>/dev/null)
{
  echo "$header"
  echo
  python - $work/raw.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct", "sm__throughput.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared")
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) or k in h for k in keep):
        print("%-90s %s %s" % (h, v, u))
PY
  echo
  echo "== per phase (tools/ncu_by_phase.py) =="
  python $root/tools/ncu_by_phase.py $work/sass.csv $work/disasm.txt $kname ${NWARPS:-1024} ${NSTEPS:-1000}
  echo
  echo "== per opcode / source line (tools/ncu_by_line.py) =="
  python $root/tools/ncu_by_line.py $work/sass.csv $work/disasm.txt $kname 25
} > $out
rm -rf $work
echo wrote $out
