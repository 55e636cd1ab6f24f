"""Turn the whole-pass ncu counts of tools/gpu_ncu_pass.sh (gpurun_out/pass_counts.csv + plain_pass.log) into
profiles/flop_per_traj_step.json: executed FP64 flop (2 DFMA + DMUL + DADD thread instructions) and thread
instructions per trajectory-step over ALL step-kernel launches of one pass, plus the static FP64 mix of the kernel
that was measured (tests/test_sass_properties.py compares the tree's kernel with it).

    python tools/calibrate_flop.py gpurun_out/pass_counts.csv gpurun_out/plain_pass.log <tag>
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
counts_csv, plain_log, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(counts_csv)) if len(r) > 10]
hdr = rows[0]
iN, iM, iV, iI = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
m = re.search(r"store=True rep=0:.*?traj-steps ([0-9.e+]+).*?launches (\d+)", open(plain_log).read())
steps, n_pass = float(m.group(1)), int(m.group(2))     # the FIRST pass of tools/quick_perf.py and its launch count
tot = collections.Counter()
per = collections.OrderedDict()
for r in rows[1:]:
    if "k_propagate" in r[iN]:
        if r[iI] not in per and len(per) == n_pass:
            break                                       # (the capture may run into the next pass)
        v = float(r[iV].replace(",", ""))
        per.setdefault(r[iI], {})[r[iM]] = v
for v in per.values():
    for k, x in v.items():
        tot[k] += x
n_launch = len(per)
assert n_launch == n_pass, (n_launch, n_pass)
dfma, dmul, dadd = (tot[f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum"] for k in ("dfma", "dmul", "dadd"))
flop = 2 * dfma + dmul + dadd
tinst, winst, f64w = tot["smsp__thread_inst_executed.sum"], tot["smsp__inst_executed.sum"], tot["sm__inst_executed_pipe_fp64.sum"]
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "reentry_old_b200", "libreentry_b200.so")], capture_output=True, text=True).stdout
body = re.search(r"Function : _ZN2rb11k_propagateILi128ELi6ELb0EEEvNS_10StepParamsE\b(.*?)(?=\n\s*(?:Fatbin|Function : )|\Z)", txt, re.S).group(1)
ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", body)
static = {k: ops.count(k) for k in ("DFMA", "DMUL", "DADD")}
path = os.path.join(ROOT, "profiles", "flop_per_traj_step.json")
d = json.load(open(path))
d["history"][tag + "_whole_pass"] = round(flop / steps, 1)
d.update({
    "flop_per_traj_step": round(flop / steps, 1),
    "source": f"ncu {tag} (profiles/{tag}_pass_counts.csv): (2*DFMA + DMUL + DADD thread instructions, predicated on) summed over ALL {n_launch} "
              f"k_propagate launches of one pass of 262144 trajectories / its {steps:.4g} trajectory-steps",
    "thread_instr_per_traj_step_ncu": round(tinst / steps, 1),
    "fp64_share_of_issued_pct_ncu": round(100 * f64w / winst, 1),
    "fp64_per_traj_step": {"DFMA": round(dfma / steps, 1), "DMUL": round(dmul / steps, 1), "DADD": round(dadd / steps, 1)},
    "static_fp64_of_calibrated_kernel": static,
})
ia = [v.get("smsp__issue_active.avg.pct_of_peak_sustained_active") for v in per.values()]
fp = [v.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") for v in per.values()]
dur = [v.get("gpu__time_duration.sum") for v in per.values()]
w = [x / sum(dur) for x in dur]
d["issue_active_pct_ncu"] = round(sum(a * b for a, b in zip(ia, w)), 1)
d["fp64_pipe_busy_pct_ncu"] = round(sum(a * b for a, b in zip(fp, w)), 1)
d["issue_port_busy_pct_model"] = round(d["issue_active_pct_ncu"] + d["fp64_pipe_busy_pct_ncu"] / 2, 1)
d["issue_port_note"] = "issue active + half the FP64 pipe busy figure (an FP64 instruction holds the SMSP issue port 2 cycles, profiles/r1m_issue_port.txt), duration-weighted over the launches of the pass"
json.dump(d, open(path, "w"), indent=1)
print(json.dumps({k: d[k] for k in ("flop_per_traj_step", "thread_instr_per_traj_step_ncu", "fp64_share_of_issued_pct_ncu", "fp64_per_traj_step",
                                    "static_fp64_of_calibrated_kernel", "issue_active_pct_ncu", "fp64_pipe_busy_pct_ncu", "issue_port_busy_pct_model")}, indent=1))
