#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_parity_sweep.py tests/test_gpu_parity.py -m gpu -q -s -k "parity_within or golden_orbits or golden_reentry" 2>&1 | grep -E "err median|max rel|passed|failed" > gpurun_out/r2q_parity_lines.txt
cat gpurun_out/r2q_parity_lines.txt
