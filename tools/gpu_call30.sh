#!/bin/bash
python tools/quick_perf.py 1000000 3 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_call20.sh
