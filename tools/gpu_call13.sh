#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest13.log; cat gpurun_out/pytest13.log
python tools/adaptive_perf.py 2>&1 | tail -3 | sed 's/^/kshared /'
RB_LIB_PATH=$PWD/variants/var_klocal.so python tools/adaptive_perf.py 2>&1 | tail -3 | sed 's/^/klocal  /'
