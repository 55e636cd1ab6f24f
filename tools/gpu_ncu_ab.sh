#!/bin/bash
# ncu captures of the step kernel for two builds (in-tree = RHS subroutine, variants/var_inline.so = inlined RHS):
# a vacuum-phase launch pair (-s 40) and a dense-atmosphere pair (-s 84) of the C2-shaped run of tools/quick_perf.py
mkdir -p gpurun_out
python tools/quick_perf.py 1000000 1 > gpurun_out/plain_call.log 2>&1 || exit 1
for s in 40 84; do
  ncu --set full --clock-control none --import-source on -k regex:k_propagate -s $s -c 2 -f -o gpurun_out/prof_call_s$s python tools/quick_perf.py 1000000 1 > gpurun_out/ncu_call_s$s.log 2>&1
done
RB_LIB_PATH=$PWD/variants/var_inline.so python tools/quick_perf.py 1000000 1 > gpurun_out/plain_inline.log 2>&1 || exit 1
for s in 40 84; do
  RB_LIB_PATH=$PWD/variants/var_inline.so ncu --set full --clock-control none --import-source on -k regex:k_propagate -s $s -c 2 -f -o gpurun_out/prof_inline_s$s python tools/quick_perf.py 1000000 1 > gpurun_out/ncu_inline_s$s.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
