for e in 256 512 1024 1184; do
for v in default wide2; do
if [ $v = default ]; then unset TGB_LIB_DIR; else export TGB_LIB_DIR=$PWD/variants/$v; fi
timeout 100 python bench.py --envs $e --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('$e $v', '%.4g'%d['value'], d['roofline']['kernel_ms'], d['config']['block_threads'], d['config']['grid'])"
TGB_WIDE=0 timeout 100 python bench.py --envs $e --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('$e $v narrow', '%.4g'%d['value'], d['roofline']['kernel_ms'], d['config']['block_threads'], d['config']['grid'])"
done; done
