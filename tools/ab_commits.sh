#!/bin/bash
# same-box A/B of bench.py across library builds of earlier commits (variants/wt_<commit>/, built by hand from git worktrees):
#   tools/ab_commits.sh "<bench args>" wt_f318cd1 wt_d55c9f1 now
args=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = now ]; then dir=$PWD; else dir=$PWD/variants/$v; fi
    for cap in 0 6; do
      (cd $dir && TGB_PAIR_CAP=$cap python bench.py --no-cpu-baseline --steps 4 --warmup 3 $args 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v cap$cap rep$rep value %.4g kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))")
    done
  done
done
