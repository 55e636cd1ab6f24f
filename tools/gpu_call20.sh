#!/bin/bash
# final kernel (K rows as pairs): whole-pass instruction counts + one full-set capture of a vacuum-phase and a dense-atmosphere launch pair
mkdir -p gpurun_out
python tools/quick_perf.py 262144 1 > gpurun_out/plain_pass.log 2>&1 || exit 1
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_propagate -c 64 --csv --log-file gpurun_out/pass_counts.csv python tools/quick_perf.py 262144 1 > gpurun_out/ncu_pass.log 2>&1
grep -c k_propagate gpurun_out/pass_counts.csv
grep "store=" gpurun_out/plain_pass.log | head -3
