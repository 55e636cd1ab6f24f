#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|Error|err median|max rel position" | tail -16 > gpurun_out/pytest11.log; cat gpurun_out/pytest11.log
python tests/run_configs.py c3 c4 --wide > gpurun_out/configs_c3_c4.jsonl 2> gpurun_out/configs_c3_c4.err; echo "configs rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/configs_c3_c4.jsonl"):
    d = json.loads(l)
    print(d["config"][:90], "| kernel %.3g G/s wall %.3g G/s | parity %s | drift %s" % (d["traj_steps_per_s_kernel"]/1e9, d["traj_steps_per_s_wall"]/1e9, d["parity"]["pos_rel_max"], d.get("full_horizon_drift", {}).get("pos_rel_max")))
PY
