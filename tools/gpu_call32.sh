#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/quick_perf.py 1000000 2 2>&1 | tail -1
timeout 300 python tools/adaptive_perf.py 100000 RK45,DOP853 2>&1 | tail -4
