#!/bin/bash
timeout 900 python -m pytest tests/test_parity_sweep.py -m gpu -q -s 2>&1 | grep -E "err median|assert|Error|passed|failed" | head -20
