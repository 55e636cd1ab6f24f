#!/bin/bash
# Build kernel-variant libraries into variants/ : each argument is "BLOCK MIN_BLOCKS TILE_STEPS [extra -D flags]"
mkdir -p variants; rm -f variants/*.so
for v in "$@"; do
  set -- $v; b=$1; m=$2; t=$3; shift 3; extra="$*"
  tag=$(echo "${b}_${m}_${t}_${extra}" | tr -d ' =-' | tr -c 'A-Za-z0-9_\n' '_')
  env -u CC -u CXX nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fopenmp,-O2 -shared \
    -Iinclude -Ireentry_old_b200/csrc -DRB_BLOCK=$b -DRB_MIN_BLOCKS=$m -DRB_TILE_STEPS=$t $extra -Xptxas -v \
    -o variants/var_${tag}.so reentry_old_b200/csrc/rb_api.cu 2>&1 | grep -A2 "k_propagate" | grep -E "spill|Used" | tr '\n' ' '
  echo " <= $v"
done
