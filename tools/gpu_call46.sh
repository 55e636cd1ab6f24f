#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_perf.py 1000000 2 2>&1 | tail -1
