"""Condense an .ncu-rep (read on the CPU box) into the handful of numbers this project tracks.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [traj_steps_per_launch]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return hdr.index(name) if name in hdr else None


def val(r, name, default=None):
    i = col(name)
    if i is None:
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return r[i]


for r in data:
    print("==", val(r, "Kernel Name"), "grid", val(r, "launch__grid_size"), "block", val(r, "launch__block_size"))
    dur = val(r, "gpu__time_duration.sum")
    du = units[col("gpu__time_duration.sum")]
    print(f"   duration {dur} {du}; regs {val(r, 'launch__registers_per_thread')}; occupancy limit regs/smem/warps "
          f"{val(r, 'launch__occupancy_limit_registers')}/{val(r, 'launch__occupancy_limit_shared_mem')}/"
          f"{val(r, 'launch__occupancy_limit_warps')}; warps active {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}%")
    dfma = val(r, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0)
    dmul = val(r, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0)
    dadd = val(r, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0)
    flop = 2 * dfma + dmul + dadd
    tinst = val(r, "smsp__thread_inst_executed.sum", 0)
    winst = val(r, "smsp__inst_executed.sum", 0)
    f64w = val(r, "sm__inst_executed_pipe_fp64.sum", 0)
    sec = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(du, 1e-9)
    print(f"   DFMA {dfma:.4g} DMUL {dmul:.4g} DADD {dadd:.4g} -> {flop:.4g} flop, {flop / sec / 1e12:.2f} TFLOP/s"
          + (f", {flop / steps:.0f} flop/traj-step, {tinst / steps:.0f} thread-instr/traj-step" if steps else ""))
    print(f"   fp64 pipe {val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f}% of peak; fp64 share of "
          f"issued {100 * f64w / max(winst, 1):.1f}%; issue active {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}%; "
          f"eligible/cycle {val(r, 'smsp__warps_eligible.avg.per_cycle_active'):.2f}")
    pipes = {p: val(r, f"sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active") for p in ("alu", "fma", "xu", "lsu", "cbu", "adu", "uniform")}
    print("   pipes " + " ".join(f"{k} {v:.1f}%" for k, v in pipes.items() if isinstance(v, float)))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("   stalls " + " ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    print(f"   dram read {val(r, 'dram__bytes_read.sum')} {units[col('dram__bytes_read.sum')]} write {val(r, 'dram__bytes_write.sum')} "
          f"{units[col('dram__bytes_write.sum')]}; branch uniform {val(r, 'smsp__sass_average_branch_targets_threads_uniform.pct')}%")
