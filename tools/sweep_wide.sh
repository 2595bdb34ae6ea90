#!/bin/bash
# narrow vs four-warp layout over environment counts: tools/sweep_wide.sh [workload]   (same box, interleaved)
wl=${1:-superball}
for n in 256 512 1024 1536 1776 2048 3072 4096; do
  for w in 0 1; do
    TGB_WIDE=$w python bench.py --no-cpu-baseline --steps 5 --envs $n --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl envs $n wide=$w value %.4g e2e %.4g block %s grid %s' % (d['value'], d['e2e']['value'], d['config'].get('block_threads'), d['config'].get('grid')))"
  done
done
