#!/bin/bash
# the driver's own round-end sequence on one GPU: GPU tests, smoke, reference arm, bench (20 steps / 5 warm-up)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench_reference_arm.json 2> gpurun_out/r2h_ref.err ) 2>&1 | tail -4
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_n1.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2h_n1.err
python tools/bench_digest.py gpurun_out/r2h_bench_n1.json 2>&1 | tail -30
