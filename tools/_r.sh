(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
timeout 100 python bench.py --envs 1024 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('1024', '%.4g'%d['value'], d['roofline']['kernel_ms'], d['config']['block_threads'], d['config']['grid'])"
timeout 100 python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('superball', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])"
