(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
timeout 100 python bench.py --workload tetraspine --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('tetraspine', d['value'], d['roofline']['kernel_ms'])"
timeout 100 python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('superball', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline'].get('issue'))"
timeout 100 python bench.py --impl reference --workload prism --steps 3 > gpurun_out/bench_r02e_prism_reference.json; cat gpurun_out/bench_r02e_prism_reference.json | python -c "import json,sys; d=json.load(sys.stdin); print('prism reference', d['value'], d['cpu_baseline'])"
