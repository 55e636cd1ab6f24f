#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_packed.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest7.log
tail -3 gpurun_out/pytest7.log
timeout 300 python tools/quick_perf.py 1000000 2 2>&1 | grep -E "store=False|rror" | sed "s#^#tree #" | tee gpurun_out/ab7.log
