#!/bin/bash
# DOP853: parity tests of the adaptive mode + ensemble timing of the three explicit methods
set -x
timeout 600 python -m pytest tests/test_adaptive.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/adaptive_perf.py 100000 RK45,RK23,DOP853 2>&1 | tail -8
