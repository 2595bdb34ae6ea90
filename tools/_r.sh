for c in 125 250 500; do echo chunk $c; TGB_SPLIT_CHUNK=$c timeout 200 python tools/split_probe.py 2>&1 | grep "split -1"; done
TGB_SPLIT_DEBUG=1 TGB_SPLIT_CHUNK=250 timeout 100 python tools/split_probe.py 2>&1 | grep "tail split" | awk '{print $4}' | tr '\n' ' '
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "tail_split" 2>&1 | tail -5
