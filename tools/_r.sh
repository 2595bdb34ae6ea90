(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
timeout 100 python bench.py --workload tetraspine --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('tetraspine', d['value'], d['roofline']['kernel_ms'])"
timeout 100 python bench.py --workload flemons --precision f64 --envs 1024 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('flemons f64 1024', d['value'], d['roofline']['kernel_ms'])"
timeout 100 python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('superball', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])"
