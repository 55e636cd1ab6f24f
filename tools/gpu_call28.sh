#!/bin/bash
python tools/quick_perf.py 1000000 3 2>&1 | tail -2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_packed.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_call20.sh
