#!/bin/bash
# first GPU call of round 2: tests on both kernel builds, kernel A/B, one bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_call.log
RB_LIB_PATH=$PWD/variants/var_inline.so python -m pytest tests/test_packed.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_inline.log
tools/gpu_ab.sh > gpurun_out/ab.log 2>&1
RB_HOST_TRACE=1 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_call.json 2> gpurun_out/bench_call.err
RB_LIB_PATH=$PWD/variants/var_inline.so python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_inline.json 2> gpurun_out/bench_inline.err
tail -3 gpurun_out/pytest_call.log gpurun_out/pytest_inline.log; cat gpurun_out/ab.log; tail -c 1500 gpurun_out/bench_call.err
