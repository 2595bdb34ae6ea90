#!/bin/bash
# the side workloads of profiles/bench_r01_*.json in one go (value / e2e per line)
run() { python bench.py --no-cpu-baseline --steps 5 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', 'value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"; }
run
run --workload prism
run --precision f64
run --envs 65536
run --workload tetraspine
run --workload flemons --precision f64 --envs 1024
run --envs 1024
