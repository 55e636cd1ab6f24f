#!/bin/bash
# one 8-GPU box: bench.py at N = 8, 4, 2, 1 (the driver's SCALE sequence) + the reference arm; host-path traces at N = 8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
nproc >> gpurun_out/topo_n8.txt; free -g | head -2 >> gpurun_out/topo_n8.txt
for n in 8 4 2; do
  extra="--no-cpu"; [ $n = 8 ] && extra=""
  RB_HOST_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 2 --warmup 3 $extra > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  echo "N=$n rc=$?"
done
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "N=1 rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/scale_ref.json 2> gpurun_out/scale_ref.err; echo "ref rc=$?"
python tools/bench_digest.py gpurun_out/scale_n8.json gpurun_out/scale_n4.json gpurun_out/scale_n2.json gpurun_out/scale_n1.json
