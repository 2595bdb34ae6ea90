(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; python -c "import json; d=json.load(open('gpurun_out/bench_r02f.json')); print('superball', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline'].get('issue',{}).get('frac'), d['per_rank_kernel_ms'], d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 3 > gpurun_out/bench_r02f_reference.json 2>> gpurun_out/bench_r02f.err; cat gpurun_out/bench_r02f_reference.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
