#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest10.log; tail -3 gpurun_out/pytest10.log
python tools/quick_perf.py 262144 1 > gpurun_out/plain_pass.log 2>&1 || exit 1
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_propagate -c 64 --csv --log-file gpurun_out/pass_counts.csv python tools/quick_perf.py 262144 1 > gpurun_out/ncu_pass.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench rc=$?"
python tools/bench_digest.py gpurun_out/bench10.json
