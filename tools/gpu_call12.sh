#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest12.log; cat gpurun_out/pytest12.log
RB_HOST_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-c5 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench rc=$?"
grep "rb_propagate_host_packed" gpurun_out/bench12.err | tail -2
python tools/bench_digest.py gpurun_out/bench12.json
python tools/adaptive_perf.py 2>&1 | tail -6
