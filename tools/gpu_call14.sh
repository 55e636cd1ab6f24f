#!/bin/bash
mkdir -p gpurun_out
python tools/quick_perf.py 1000000 1 > gpurun_out/plain_1m.log 2>&1 || exit 1
for s in 40 84; do
ncu --set full --clock-control none --import-source on -k regex:k_propagate -s $s -c 2 -f -o gpurun_out/prof_r2f_s$s python tools/quick_perf.py 1000000 1 > gpurun_out/ncu_r2f_s$s.log 2>&1
done
python tools/adaptive_perf.py > gpurun_out/plain_adaptive.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_adaptive -c 2 -f -o gpurun_out/prof_r2f_adaptive python tools/adaptive_perf.py > gpurun_out/ncu_r2f_adaptive.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:. -c 600 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-c5 --no-parity > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/prof_r2f* gpurun_out/launches_r2f.csv; cat gpurun_out/plain_adaptive.log | tail -2
