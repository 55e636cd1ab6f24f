#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "err median|max rel|second-order|passed|failed|assert" | head -20
python tools/quick_perf.py 1000000 2 2>&1 | tail -1
bash tools/gpu_call20.sh
