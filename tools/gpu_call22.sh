#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
echo "rc=$?"; tail -c 400 gpurun_out/r2i_bench_n2.err
python tools/bench_digest.py gpurun_out/r2i_bench_n2.json | tail -12
