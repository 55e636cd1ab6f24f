#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/run_configs.py c1 c3 c4 > gpurun_out/r2m_configs.jsonl 2> gpurun_out/r2m_configs.err; echo "configs rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2m_configs.jsonl"):
    d = json.loads(l)
    print(d["config"][:100], "| kernel %.4g G/s wall %.4g G/s wall_s %.3f | parity %s | drift %s" % (d.get("traj_steps_per_s_kernel",0)/1e9, d.get("traj_steps_per_s_wall",0)/1e9, d.get("wall_s",0), d.get("parity",{}).get("pos_rel_max"), d.get("full_horizon_drift", {}).get("pos_rel_max")))
PY
