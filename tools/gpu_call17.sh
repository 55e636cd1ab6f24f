#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_packed.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/quick_perf.py 1000000 3 2>&1 | grep -E "store=False|rror" | sed "s#^#tree #"
