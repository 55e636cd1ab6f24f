#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_packed.py tests/test_parity_sweep.py tests/test_adaptive.py tests/test_thermal.py tests/test_ground_track.py -m gpu -x -q 2>&1 | tail -15
python tools/quick_perf.py 1000000 3 2>&1 | tail -3
