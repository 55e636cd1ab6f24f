#!/bin/bash
mkdir -p gpurun_out
for so in variants/var_norot.so ""; do
  tag=${so:-tree}
  echo "== $tag"
  RB_LIB_PATH=${so:+$PWD/$so} timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
  RB_LIB_PATH=${so:+$PWD/$so} timeout 300 python tools/quick_perf.py 1000000 2 2>&1 | grep -E "store=False|rror" | sed "s#^#$(basename $tag .so) #"
done 2>&1 | tee gpurun_out/ab8.log
