#!/bin/bash
# final tree on one 8-GPU box: bench.py at N = 8, 4, 2, 1 (the driver's SCALE sequence, shorter runs)
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_n$n.json 2> gpurun_out/r2j_n$n.err
  echo "N=$n rc=$?"
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_n1.err; echo "N=1 rc=$?"
python tools/bench_digest.py gpurun_out/r2j_bench_n8.json gpurun_out/r2j_bench_n4.json gpurun_out/r2j_bench_n2.json gpurun_out/r2j_bench_n1.json
