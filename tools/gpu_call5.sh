#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_packed.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest5.log
tail -3 gpurun_out/pytest5.log
for so in "" variants/var_nodual.so variants/var_r2a_inline.so; do
  tag=${so:-tree}
  RB_LIB_PATH=${so:+$PWD/$so} timeout 300 python tools/quick_perf.py 1000000 2 2>&1 | grep -E "store=False|rror" | sed "s#^#$(basename $tag .so) #"
done | tee gpurun_out/ab5.log
