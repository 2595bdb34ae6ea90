TGB_SPLIT_DEBUG=1 timeout 60 python tools/split_probe.py 2>&1 | tail -22
