#!/usr/bin/env python
"""bench.py -- headline benchmark of the Reentry-OLD hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # GPU arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port)

Workload (BASELINE.json configs[1]): Monte Carlo reentry dispersion, 1 M perturbed initial
states per GPU (pos/vel/azimuth), dt = 0.5 s, until impact (cap 2048 steps), stride-16 samples
plus summaries.  One "step" = one full pass of the hot path over that batch.  Metric: FP64
trajectory-steps/s (a trajectory-step = one fixed step of one live trajectory), whole job.
N > 1: each rank propagates its own 1 M-state shard (seed + rank; weak scaling), then ONE
NCCL gather of the per-trajectory summaries (impact time, step, final state, status) to rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64_trajectory_steps_per_s"
UNIT = "traj-steps/s"
# Algorithmic FP64 work per trajectory-step (DESIGN.md "Roofline"): SURVEY.md 8(d) cap of 6.5 kflop
# (HW-weighted, DFMA = 2) lowered to the ncu-measured executed flop of the kernel when that is smaller.
def _flop_per_traj_step():
    if "RB_FLOP_PER_TRAJ_STEP" in os.environ:
        return float(os.environ["RB_FLOP_PER_TRAJ_STEP"]), "env"
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "flop_per_traj_step.json")))
        return min(float(d["flop_per_traj_step"]), float(d.get("cap_survey_8d", 6500.0))), d.get("source", "profiles")
    except Exception:
        return 6500.0, "SURVEY.md 8(d) cap (no ncu calibration found)"


FLOP_PER_TRAJ_STEP, FLOP_SOURCE = _flop_per_traj_step()


def _ncu_pipe_numbers():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "flop_per_traj_step.json")))
        return {k: d[k] for k in ("fp64_pipe_busy_pct_ncu", "fp64_share_of_issued_pct_ncu", "issue_active_pct_ncu",
                                  "issue_port_busy_pct_model", "issue_port_note", "dram_bytes_per_traj_step_ncu") if k in d}
    except Exception:
        return {}
SAMPLE_BYTES = 64
SUMMARY_BYTES = 8 * 6 + 4 + 8 + 1 + 4


def workload(n_traj, rank):
    from reentry_old_b200.configs import c2_reentry_mc

    return c2_reentry_mc(n=n_traj, seed=1234 + rank, n_steps=2048, out_stride=16)


def config_dict(w, n_gpus):
    return {"workload": "C2 Monte Carlo reentry dispersion: 1M perturbed initial states (pos/vel/azimuth), "
                        "dt=0.5 s, until impact, J2+drag+Sun/Moon",
            "n_trajectories_per_gpu": w.n, "n_gpus": n_gpus, "dt": w.dt, "max_steps": w.n_steps,
            "max_steps_note": "every trajectory impacts within ~1.9 k steps (mean 1362): BASELINE.json's 6000-step cap is "
                              "never reached; 2048 sizes the time table and the sample buffer",
            "out_stride": w.out_stride, "integrator": "fixed-step Dormand-Prince (scipy RK45 tableau, FSAL)",
            "l2": "working set exceeds L2: ~5.4 GB of sample stores + 48 MB state per step (no flush needed)",
            "sharding": "independent shards, final NCCL gather of summaries" if n_gpus > 1 else "single GPU"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        self.idx = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t_begin <= ts <= t_end + 0.2] or [r for _, r in self.rows]
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for k, name in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(rows[0][2]) if rows and len(rows[0]) >= 3 else None,
                "power_w_max": max((float(r[3]) for r in rows if len(r) >= 9), default=None),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(w, target_seconds=12.0, n_threads=0):
    """The oracle port (C, OpenMP, the reference's per-RHS ephemeris and all) on a bounded
    sample of the same workload.  Returns (traj-steps/s, threads, sample description)."""
    from oracle import oracle as O

    O.build()
    if n_threads <= 0:   # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
        n_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    p = w.params
    probe = 64
    y0 = O.get_initial_states(*p[:probe].T, w.gmst0)
    t = time.perf_counter()
    r = O.propagate(y0, w.Cd, w.A, w.m, w.epoch_jd, w.gmst0, 0.0, w.dt, w.n_steps, w.out_stride, want_samples=False,
                    n_threads=n_threads)
    el = time.perf_counter() - t
    steps = int(np.where(r["impact_step"] > 0, r["impact_step"], w.n_steps).sum())
    rate = steps / el
    n = int(min(w.n, max(probe, rate * target_seconds / (steps / probe))))
    n = max(probe, (n // r["threads"]) * r["threads"])
    y0 = O.get_initial_states(*p[:n].T, w.gmst0)
    t = time.perf_counter()
    r = O.propagate(y0, w.Cd, w.A, w.m, w.epoch_jd, w.gmst0, 0.0, w.dt, w.n_steps, w.out_stride, want_samples=False,
                    n_threads=n_threads)
    el = time.perf_counter() - t
    steps = int(np.where(r["impact_step"] > 0, r["impact_step"], w.n_steps).sum())
    return steps / el, r["threads"], f"first {n} trajectories of the workload, {steps} traj-steps in {el:.1f} s", steps, el


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.n_traj, 0)
    per_step_seconds = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, threads, sample = [], 1, ""
    for k in range(args.warmup + args.steps):
        rate, threads, sample, steps, el = cpu_oracle_rate(w, per_step_seconds)
        if k >= args.warmup:
            vals.append((steps, el))
    tot_steps = sum(s for s, _ in vals)
    tot_t = sum(e for _, e in vals)
    v = tot_steps / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(w, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "each step: " + sample +
                                       " (C port of the reference's numba path, OpenMP over trajectories; the "
                                       "reference itself is Python+numba and cannot travel to the box)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout: point fd 1 at stderr for everything the libraries print (NCCL writes
    its version banner to stdout) and keep the real stdout for emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-traj", type=int, default=1_000_000, help="trajectories per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import ctypes as C

    import torch
    import torch.distributed as dist

    from reentry_old_b200 import SpacecraftModel
    from reentry_old_b200 import _lib as L
    from reentry_old_b200.model import DeviceContext
    from reentry_old_b200.sharding import gather_summaries

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    W = max(args.warmup, 3)
    K = args.steps
    w = workload(args.n_traj, rank)
    model = SpacecraftModel(Cd=w.Cd, A=w.A, m=w.m, epoch=w.epoch_jd, gmst0=w.gmst0, device=local_rank)
    ctx = DeviceContext.get(local_rank)
    params = torch.as_tensor(w.params, device=f"cuda:{local_rank}")
    y0 = model.get_initial_states(*params.t().contiguous(), gmst=w.gmst0)   # resident in HBM before timing
    torch.cuda.synchronize()

    def one_step(profile=False):
        r = model.run_ensemble(y0, t0=0.0, dt=w.dt, n_steps=w.n_steps, out_stride=w.out_stride,
                               steps_per_launch=w.steps_per_launch, profile=profile)
        g = gather_summaries(r, dst=0) if world > 1 else None
        return r, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        r, _ = one_step()
    torch.cuda.synchronize()
    steps_per_pass = int(torch.where(r.impact_step > 0, r.impact_step, torch.full_like(r.impact_step, w.n_steps)).sum())
    launches_per_pass = r.stats["kernel_launches"]
    n_rows_used = int(r.n_samples.max())
    del r

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    ev0.record()
    for _ in range(K):
        r, _ = one_step()
    ev1.record()
    barrier()
    t_end = time.time()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms, float(steps_per_pass)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, total_steps_pass = float(tmax[0]), float(tsum[1])
    else:
        total_steps_pass = float(steps_per_pass)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = total_steps_pass * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (k_propagate), timed live with CUDA events around every launch
    del r
    r, _ = one_step(profile=True)
    torch.cuda.synchronize()
    kern_ms = r.stats["propagate_ms"]
    n_prop = r.stats["propagate_launches"]
    peak = ctx.fp64_peak_tflops(300.0)
    achieved = steps_per_pass * FLOP_PER_TRAJ_STEP / (kern_ms * 1e-3) / 1e12
    sample_bytes = float(r.n_samples.to(torch.float64).sum()) * SAMPLE_BYTES + w.n * SUMMARY_BYTES
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ncu_nums = _ncu_pipe_numbers()
    traffic = (ncu_nums["dram_bytes_per_traj_step_ncu"] * steps_per_pass / max(1, n_prop)
               if "dram_bytes_per_traj_step_ncu" in ncu_nums else None)   # DRAM bytes per k_propagate launch (state only)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_propagate", "kernel_ms_per_pass": kern_ms, "kernel_launches_per_pass": n_prop,
                "kernel_share_of_step": kern_ms / (ms / K), "flop_per_traj_step": FLOP_PER_TRAJ_STEP, "flop_source": FLOP_SOURCE,
                "ncu": ncu_nums,
                "peak_source": "DFMA probe kernel run in this process (MEASURED_PEAKS.json has no FP64 figure); "
                               "nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
                "output_write": {"bytes_per_pass": sample_bytes, "achieved_gbs": sample_bytes / (kern_ms * 1e-3) / 1e9,
                                 "hbm_peak_gbs": hbm_peak,
                                 "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                                 "frac": sample_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak}}
    del r
    torch.cuda.empty_cache()

    # ---- end to end through the C ABI with HOST buffers (H2D of y0, D2H of samples + summaries every step)
    e2e = None
    if not args.no_e2e:
        n = w.n
        n_out = w.n_steps // w.out_stride + 1
        lib = ctx.lib
        y0_h = torch.empty((n, 6), dtype=torch.float64).pin_memory()
        y0_h.copy_(y0)
        rows_cap = min(n_out, n_rows_used + 2)
        # pinned result buffers (the library streams finished sample rows into them while later launches run)
        samples_h = torch.empty((n_out, 8, n), dtype=torch.float64, pin_memory=True) if n_out * 8 * n * 8 < 12e9 else None
        fs_h = torch.empty((n, 6), dtype=torch.float64).pin_memory()
        is_h = torch.empty(n, dtype=torch.int32).pin_memory()
        it_h = torch.empty(n, dtype=torch.float64).pin_memory()
        st_h = torch.empty(n, dtype=torch.uint8).pin_memory()
        ns_h = torch.empty(n, dtype=torch.int32).pin_memory()
        used = C.c_int64(0)
        run = L.RbRun(first_step=0, n_steps=w.n_steps, out_stride=w.out_stride, steps_per_launch=w.steps_per_launch,
                      event_altitude=1000.0,
                      flags=L.RB_FLAG_COMPACT | (L.RB_FLAG_STAGED_SAMPLES if os.environ.get("RB_E2E_STAGED") else 0))

        def host_pass():
            rc = lib.rb_propagate_host(ctx.handle, n, y0_h.data_ptr(), None, None, None, float(w.Cd), float(w.A),
                                       float(w.m), w.epoch_jd, w.gmst0, 0.0, w.dt, C.byref(run),
                                       samples_h.data_ptr() if samples_h is not None else None, n_out,
                                       fs_h.data_ptr(), is_h.data_ptr(), it_h.data_ptr(), st_h.data_ptr(),
                                       ns_h.data_ptr(), C.byref(used))
            ctx.check(rc, "rb_propagate_host")

        for _ in range(2):
            host_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            host_pass()
        barrier()
        el = time.perf_counter() - t0
        ctx.table_key = None
        e2e_steps = int(np.where(is_h.numpy() > 0, is_h.numpy(), w.n_steps).sum())
        tt = torch.tensor([el, float(e2e_steps)], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            a = tt.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = tt.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            el, e2e_total = float(a[0]), float(b[1])
        else:
            e2e_total = float(e2e_steps)
        # pinned sample buffer -> the kernel stores the valid samples straight into host memory (zero-copy)
        d2h = (int(ns_h.to(torch.int64).sum()) * SAMPLE_BYTES if samples_h is not None else 0) + n * (48 + 4 + 8 + 1 + 4)
        e2e = {"value": e2e_total * K / el, "unit": UNIT, "h2d_bytes_per_step": n * 48 + (5 * w.n_steps + 1) * 128,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * el / K,
               "api": "rb_propagate_host (C ABI, pinned host buffers; time table built on the host inside the call; sample "
                      "rows reach the pinned host buffer while later launches run: dense rows by copy engine from a device "
                      "staging buffer, the sparse rows of the retiring phase as direct stores of the step kernel)",
               "d2h_note": "d2h_bytes_per_step counts the valid samples and the summaries the caller receives; rows sent "
                           "by copy engine also carry NaN padding for trajectories that have already retired"}
        del samples_h

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        rate, threads, sample, _, _ = cpu_oracle_rate(w, 12.0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config_dict(w, world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches_per_pass * K), "roofline": roofline, "cpu_baseline": cpu,
                "traj_steps_per_pass": total_steps_pass}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
