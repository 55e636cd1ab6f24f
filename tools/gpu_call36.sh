#!/bin/bash
# last kernel of the round on one 8-GPU box: bench.py at N = 8, 4, 2 (short runs)
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu > gpurun_out/r2m_bench_n$n.json 2> gpurun_out/r2m_n$n.err
  echo "N=$n rc=$?"
done
python tools/bench_digest.py gpurun_out/r2m_bench_n8.json gpurun_out/r2m_bench_n4.json gpurun_out/r2m_bench_n2.json | grep -E "^==|e2e|c5"
