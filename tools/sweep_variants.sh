#!/bin/bash
# Run tools/quick_perf.py against every kernel-variant library under variants/ (GPU box).
for so in variants/*.so; do
  echo "=== $so"
  RB_LIB_PATH=$PWD/$so timeout 200 python tools/quick_perf.py ${1:-1000000} 1 2>&1 | grep -E "store=False|Error|error" | sed 's/traj-steps.*wall, //'
done
