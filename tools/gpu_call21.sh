#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_adaptive.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
