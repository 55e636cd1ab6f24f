#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-c5 > gpurun_out/trim_bench.json 2> gpurun_out/trim_bench.err; python tools/bench_digest.py gpurun_out/trim_bench.json | tail -8
