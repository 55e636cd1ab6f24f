#!/bin/bash
timeout 600 python -m pytest tests/test_adaptive.py -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head -20
