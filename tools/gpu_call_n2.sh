#!/bin/bash
mkdir -p gpurun_out
RB_HOST_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_n2.err; head -c 300 gpurun_out/bench_n2.json
