#!/bin/bash
# final kernel of round 2: the driver's bench command, then ncu full-set captures of a vacuum-phase and a dense-atmosphere launch
# pair of the step kernel and the launch list of the bench command (each after its own command exited 0 without ncu)
mkdir -p gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_n1.err ) 2>&1 | tail -3
python tools/bench_digest.py gpurun_out/r2l_bench_n1.json | tail -9
python tools/quick_perf.py 1000000 1 > gpurun_out/plain_1m.log 2>&1 || exit 1
for s in 40 84; do
ncu --set full --clock-control none --import-source on -k regex:k_propagate -s $s -c 2 -f -o gpurun_out/prof_r2l_s$s python tools/quick_perf.py 1000000 1 > gpurun_out/ncu_r2l_s$s.log 2>&1
done
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-c5 --no-parity > gpurun_out/plain_bench_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:. -c 600 --csv --log-file gpurun_out/launches_r2l.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-c5 --no-parity > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/prof_r2l* gpurun_out/launches_r2l.csv
