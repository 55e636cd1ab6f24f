#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_n1.err ) 2>&1 | tail -3
python tools/bench_digest.py gpurun_out/r2n_bench_n1.json | tail -9
