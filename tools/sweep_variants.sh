#!/bin/bash
# Run tools/quick_perf.py against every kernel-variant library under variants/ (GPU box).
# usage: sweep_variants.sh [n_traj] [spl list, e.g. "0 24 16"]
for so in variants/*.so; do
  for spl in ${2:-0}; do
    echo "=== $so spl=$spl"
    RB_SPL=$spl RB_LIB_PATH=$PWD/$so timeout 200 python tools/quick_perf.py ${1:-1000000} 1 2>&1 | grep -E "store=False|Error|error" | sed 's/traj-steps.*wall, //'
  done
done
