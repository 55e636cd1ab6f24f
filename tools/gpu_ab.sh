#!/bin/bash
# GPU box: A/B of step-kernel variants on the C2-shaped device-resident run (tools/quick_perf.py).
#   variants/orig_tree   the previous round's kernel (flattened stage loop, inlined RHS behind a switch)
#   in-tree library      RHS as a subroutine called from the seven evaluation sites
#   variants/var_*.so    other builds of the current source (RB_LIB_PATH)
mkdir -p gpurun_out
( cd variants/orig_tree && timeout 300 python tools/quick_perf.py 1000000 2 2>&1 | grep -E "store=False|peak|rror" | sed 's/^/orig   /' )
timeout 300 python tools/quick_perf.py 1000000 2 2>&1 | grep -E "store=False|rror" | sed 's/^/call   /'
for so in variants/var_*.so; do
  RB_LIB_PATH=$PWD/$so timeout 300 python tools/quick_perf.py 1000000 2 2>&1 | grep -E "store=False|rror" | sed "s#^#$(basename $so .so) #"
done
