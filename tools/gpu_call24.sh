#!/bin/bash
timeout 900 python -m pytest tests/test_packed.py -m gpu -x -q 2>&1 | tail -15
