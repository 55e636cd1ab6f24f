#!/usr/bin/env python
"""bench.py -- headline benchmark of the Reentry-OLD hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # GPU arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (the reference's path on the host cores)

Workload (BASELINE.json configs[1]): Monte Carlo reentry dispersion, 1 M perturbed initial
states per GPU (pos/vel/azimuth), dt = 0.5 s, until impact (cap 6000 steps), stride-16 samples plus summaries.
One "step" = one full pass of the hot path over that batch.  Metric: FP64 trajectory-steps/s (a trajectory-step =
one fixed step of one live trajectory), whole job.
N > 1: each rank propagates its own 1 M-state shard (seed + rank; weak scaling), then ONE NCCL gather of the
per-trajectory summaries (impact time, step, final state, status) to rank 0.

Blocks of the JSON line beyond the base contract:
  roofline      FP64 roofline of k_propagate (kernel time from a separate profiled pass, see `timing`)
  e2e           the same workload through the host-buffer API (pinned host memory in and out, copies inside)
  e2e_c5        BASELINE.json configs[4] (C5): device-generated dispersions, summaries only (cap 6000), impact points,
                NCCL gather to rank 0, download to the host; 12.5 M trajectories per GPU (100 M at 8 GPUs)
  parity        the oracle (oracle/, test infrastructure) on a strided subsample of the TIMED batch; at N > 1 every rank
                also recomputes 1000 trajectories of its neighbour's shard and rank 0 compares them bit for bit with
                the gathered summaries
  cpu_baseline  the C port of the reference's path and, where oracle/_ref holds it, the reference's own numba code
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
N_STEPS_CAP = 6000          # "until impact, cap 6000 steps" (mean ~1362, max ~1900: early exit makes the cap free)
OUT_STRIDE = 16
SAMPLE_BYTES = 64
SUMMARY_BYTES = 8 * 6 + 4 + 8 + 1 + 4
C5_NOMINAL = [7800.0, 20.0, 0.0, 120e3, 90.0, -1.0]
C5_SIGMA = [10.0, 0.5, 0.5, 500.0, 0.5, 0.1]
RB_FIELDS = 8


# Algorithmic FP64 work per trajectory-step (DESIGN.md "Roofline"): SURVEY.md 8(d) cap of 6.5 kflop (HW-weighted,
# DFMA = 2) lowered to the ncu-measured executed flop of the kernel when that is smaller.  The calibration file names
# the static FP64 instruction mix of the kernel it was measured on; tests/test_sass_properties.py fails when the built
# kernel drifts from it, so a stale number cannot pass silently.
def _calibration():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "flop_per_traj_step.json")))
    except Exception:
        return {}


def _flop_per_traj_step():
    if "RB_FLOP_PER_TRAJ_STEP" in os.environ:
        return float(os.environ["RB_FLOP_PER_TRAJ_STEP"]), "env"
    d = _calibration()
    if "flop_per_traj_step" in d:
        return min(float(d["flop_per_traj_step"]), float(d.get("cap_survey_8d", 6500.0))), d.get("source", "profiles")
    return 6500.0, "SURVEY.md 8(d) cap (no ncu calibration found)"


FLOP_PER_TRAJ_STEP, FLOP_SOURCE = _flop_per_traj_step()


def _ncu_pipe_numbers():
    d = _calibration()
    return {k: d[k] for k in ("fp64_pipe_busy_pct_ncu", "fp64_share_of_issued_pct_ncu", "issue_active_pct_ncu",
                              "issue_port_busy_pct_model", "issue_port_note", "dram_bytes_per_traj_step_ncu",
                              "thread_instr_per_traj_step_ncu") if k in d}


def workload(n_traj, rank):
    from reentry_old_b200.configs import c2_reentry_mc

    return c2_reentry_mc(n=n_traj, seed=1234 + rank, n_steps=N_STEPS_CAP, out_stride=OUT_STRIDE)


def config_dict(w, n_gpus):
    return {"workload": "C2 Monte Carlo reentry dispersion: 1M perturbed initial states (pos/vel/azimuth), "
                        "dt=0.5 s, until impact, J2+drag+Sun/Moon",
            "n_trajectories_per_gpu": w.n, "n_gpus": n_gpus, "dt": w.dt, "max_steps": w.n_steps,
            "max_steps_note": "BASELINE's cap; every trajectory impacts within ~1.9 k steps (mean 1362) and the launch "
                              "sequence stops when the live count reaches 0",
            "out_stride": w.out_stride, "samples": "packed slabs (rows as wide as the live ensemble)",
            "integrator": "fixed-step Dormand-Prince (scipy RK45 tableau, FSAL)",
            "l2": "working set exceeds L2: ~5.6 GB of sample stores + 48 MB state per step (no flush needed)",
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


# ------------------------------------------------------------------------------------ CPU arms (oracle = checker / baseline)
def _host_threads():
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_oracle_rate(w, target_seconds=12.0, n_threads=0):
    """The oracle port (C, OpenMP, the reference's per-RHS ephemeris and all) on a bounded
    sample of the same workload.  Returns (traj-steps/s, threads, sample description, steps, seconds)."""
    from oracle import oracle as O

    O.build()
    if n_threads <= 0:
        n_threads = _host_threads()
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


def cpu_numba_rate(w, target_seconds=10.0):
    """The reference's OWN numba-compiled functions (oracle/_ref, see oracle/build_ref.py) driving the same fixed-step
    Dormand-Prince loop, all host threads.  None where oracle/_ref does not hold them (they are built from
    /root/reference in the build container and travel as machine code, never as source)."""
    try:
        from oracle import ref_native
    except Exception:
        return None
    try:
        return ref_native.rate(w, target_seconds)
    except Exception as e:   # the baseline is informational: never fail the bench on it
        return {"unavailable": f"{type(e).__name__}: {e}"}


def cpu_baseline_block(w, seconds=12.0):
    rate, threads, sample, _, _ = cpu_oracle_rate(w, seconds)
    cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "per_core": rate / max(1, threads), "sample": sample}
    nb = cpu_numba_rate(w, min(seconds, 10.0))
    if nb is not None:
        cpu["reference_numba"] = nb
    return cpu


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on the box's host cores: oracle/_ref (the reference's numba
    functions, compiled from /root/reference by oracle/build_ref.py) when it is there, else the C port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.n_traj, 0)
    per_step_seconds = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    use_ref = False
    try:
        from oracle import ref_native
        use_ref = ref_native.available()
    except Exception:
        use_ref = False
    vals, cores, sample, per_core = [], 1, "", None
    for k in range(args.warmup + args.steps):
        if use_ref:
            r = ref_native.rate(w, per_step_seconds)
            steps = int(r["sample"].split(", ")[1].split()[0])
            el = steps / r["value"]
            cores, sample, per_core = r["cores"], r["sample"], r["per_core"]
        else:
            _, cores, sample, steps, el = cpu_oracle_rate(w, per_step_seconds)
        if k >= args.warmup:
            vals.append((steps, el))
    tot_steps = sum(s for s, _ in vals)
    tot_t = sum(e for _, e in vals)
    v = tot_steps / tot_t
    if use_ref:
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "per_core": per_core,
               "sample": "each step: " + sample + " (the reference's own numba-compiled functions, oracle/_ref, under the "
                                                  "fixed-step Dormand-Prince driver; one process per core)"}
        rate, threads, psample, _, _ = cpu_oracle_rate(w, 6.0)
        cpu["port"] = {"value": rate, "cores": threads, "per_core": rate / max(1, threads), "kind": "port", "sample": psample}
    else:
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "per_core": v / max(1, cores),
               "sample": "each step: " + sample + " (C port of the reference's numba path, OpenMP over trajectories; "
                                                  "oracle/_ref was not built)"}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(w, args.gpus), "cpu_baseline": cpu,
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


# ------------------------------------------------------------------------------------ GPU arm helpers
def steps_of(impact_step, n_steps):
    """Trajectory-steps of a pass from the summaries: the step of the event, or the cap."""
    import torch

    if isinstance(impact_step, torch.Tensor):
        return int(torch.where(impact_step > 0, impact_step, torch.full_like(impact_step, n_steps)).to(torch.int64).sum())
    return int(np.where(impact_step > 0, impact_step, n_steps).astype(np.int64).sum())


def ingest_probe(torch, dist, dev, world, rank, pinned, gib=2.0):
    """Host ingest of this box: every rank copies `gib` GiB device -> its pinned host buffer in 64 MiB pieces at the
    same time (aggregate = what N concurrent result downloads can get), and rank 0 alone (one link)."""
    n = int(gib * 2**30 / 8)
    n = min(n, pinned.numel())
    x = torch.empty(n, dtype=torch.float64, device=dev)
    h = pinned[:n]
    cs = torch.cuda.Stream()
    out = {}
    for label, active in (("alone", rank == 0), ("concurrent", True)):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if active:
            with torch.cuda.stream(cs):
                for a, b in zip(h.split(8 << 20), x.split(8 << 20)):
                    a.copy_(b, non_blocking=True)
            cs.synchronize()
        el = time.perf_counter() - t0
        t = torch.tensor([el if active else 0.0, float(n * 8 if active else 0)], dtype=torch.float64, device=dev)
        if world > 1:
            a = t.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = t.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            el, total = float(a[0]), float(b[1])
        else:
            total = float(t[1])
        out[label] = total / el / 1e9
    del x
    return out


def parity_block(torch, dist, model, w, res, rank, world, y0, gathered):
    """Oracle on a 64-trajectory strided subsample of the TIMED batch (positions / velocities at every stored sample,
    impact step and time); at N > 1 also the cross-rank bit-identity check."""
    from oracle import oracle as O

    from reentry_old_b200.sharding import gather_packed, pack_summary

    O.build()
    n = w.n
    sub = np.arange(0, n, max(1, n // 64))[:64]
    ref = O.propagate(y0[torch.as_tensor(sub, device=y0.device)].cpu().numpy(), w.Cd, w.A, w.m, w.epoch_jd, w.gmst0, 0.0, w.dt,
                      w.n_steps, w.out_stride, n_threads=_host_threads())
    pos_rel = vel_rel = pos_abs = 0.0
    validity_equal = True
    for j, i in enumerate(sub):
        got = res.packed.trajectory(int(i)).cpu().numpy()                   # [n_samples_i, 8]
        r = ref["samples"][j]
        ok = ~np.isnan(r[:, 0])
        r = r[ok]
        if r.shape[0] != got.shape[0]:
            validity_equal = False
            continue
        d = np.linalg.norm(got[:, :3] - r[:, :3], axis=1)
        pos_abs = max(pos_abs, float(d.max()))
        pos_rel = max(pos_rel, float((d / np.linalg.norm(r[:, :3], axis=1)).max()))
        vel_rel = max(vel_rel, float((np.linalg.norm(got[:, 3:6] - r[:, 3:6], axis=1) / np.linalg.norm(r[:, 3:6], axis=1)).max()))
    istep = res.impact_step[torch.as_tensor(sub, device=res.impact_step.device)].cpu().numpy()
    itime = res.impact_time[torch.as_tensor(sub, device=res.impact_time.device)].cpu().numpy()
    me = {"pos_rel_max": pos_rel, "vel_rel_max": vel_rel, "pos_abs_max_m": pos_abs,
          "impact_step_equal": bool(np.array_equal(istep, ref["impact_step"])),
          "impact_time_abs_max_s": float(np.nanmax(np.abs(itime - ref["impact_time"]))),
          "sample_validity_equal": validity_equal}
    t = torch.tensor([pos_rel, vel_rel, pos_abs, me["impact_time_abs_max_s"], 0.0 if me["impact_step_equal"] and validity_equal else 1.0],
                     dtype=torch.float64, device=y0.device)
    out = {"checker": "oracle/ (C restatement of the reference, pinned on golden vectors of the unmodified reference)",
           "subsample": f"{len(sub)} trajectories (stride {max(1, n // 64)}) of every rank's timed batch, all stored samples",
           "bars": {"pos_rel": 1e-9, "vel_rel": 1e-9, "pos_abs_m": 1.0, "impact_time_s": 1e-7}}
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # cross-rank identity: this rank recomputes the first 1000 trajectories of its neighbour's shard
        nb = (rank + 1) % world
        from reentry_old_b200.configs import reentry_dispersion_params
        pn = reentry_dispersion_params(n, seed=1234 + nb)[:1000]   # same generator and seed: the first rows of the neighbour's matrix
        yn = model.get_initial_states(*torch.as_tensor(pn, device=y0.device).t().contiguous(), gmst=w.gmst0)
        rn = model.run_ensemble(yn, t0=0.0, dt=w.dt, n_steps=w.n_steps, out_stride=w.out_stride, store_samples=False)
        chk = gather_packed(pack_summary(rn.final_state, rn.impact_time, rn.impact_step, rn.status, rn.n_samples), dst=0)
        if rank == 0:
            ident = True
            for r in range(world):
                mine = chk[r * 1000:(r + 1) * 1000]
                theirs = gathered[((r + 1) % world) * n:((r + 1) % world) * n + 1000]
                ident = ident and bool(torch.equal(torch.nan_to_num(mine, nan=-1.0), torch.nan_to_num(theirs, nan=-1.0)))
            out["cross_rank_bit_identical"] = ident
            out["cross_rank_note"] = "rank r recomputed trajectories 0..999 of rank (r+1)%N's shard in a 1000-trajectory run; " \
                                     "compared with the NCCL-gathered summaries of the timed pass, all 10 columns, bitwise"
    out.update({"pos_rel_max": float(t[0]), "vel_rel_max": float(t[1]), "pos_abs_max_m": float(t[2]),
                "impact_time_abs_max_s": float(t[3]), "impact_step_and_validity_equal": bool(float(t[4]) == 0.0)})
    out["pass"] = bool(out["pos_rel_max"] <= 1e-9 and out["vel_rel_max"] <= 1e-9 and out["pos_abs_max_m"] < 1.0 and
                       out["impact_time_abs_max_s"] <= 1e-7 and out["impact_step_and_validity_equal"] and
                       out.get("cross_rank_bit_identical", True))
    return out


def c5_block(torch, dist, model, dev, rank, world, n_rank, passes, sub_batches=4):
    """BASELINE.json configs[4]: reentry dispersion, `n_rank` trajectories per GPU, states generated on the device
    (counter-based: trajectory i is the same whatever the GPU count), propagated until impact (cap 6000), impact points
    (geodetic) on the device, NCCL gather of 4 doubles per trajectory to rank 0, download to pinned host memory.
    The shard runs in `sub_batches` pieces so that the gather + download of one piece overlaps the propagation of the
    next; the timed region ends when rank 0's host buffer is complete."""
    lo = rank * n_rank
    bsz = (n_rank + sub_batches - 1) // sub_batches
    host = torch.empty((world, n_rank, 4), dtype=torch.float64).pin_memory() if rank == 0 else None
    s_copy = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def one_pass():
        steps = 0
        keep = []
        for b in range(sub_batches):
            blo, bhi = b * bsz, min(n_rank, (b + 1) * bsz)
            y0 = model.get_dispersed_initial_states(bhi - blo, 1234, C5_NOMINAL, C5_SIGMA, first_index=lo + blo)
            res = model.run_ensemble(y0, dt=0.5, n_steps=N_STEPS_CAP, out_stride=N_STEPS_CAP, store_samples=False)
            pts = model.impact_points(res)                                   # [3, n]: lat, lon, h
            pk = torch.empty((bhi - blo, 4), dtype=torch.float64, device=dev)
            pk[:, 0] = pts[0]; pk[:, 1] = pts[1]; pk[:, 2] = res.impact_time
            pk[:, 3] = res.impact_step.to(torch.float64) + 4294967296.0 * res.status.to(torch.float64)   # exact in f64
            if world > 1:
                parts = [torch.empty_like(pk) for _ in range(world)] if rank == 0 else None
                work = dist.gather(pk, parts, dst=0, async_op=True)
            else:
                parts, work = [pk], None
            if rank == 0:
                with torch.cuda.stream(s_copy):
                    if work is not None:
                        work.wait()                                          # s_copy waits for the collective, not the host
                    else:
                        s_copy.wait_stream(main)
                    for r in range(world):
                        host[r, blo:bhi].copy_(parts[r], non_blocking=True)
            keep.append((res, pk, parts, work))
            steps += steps_of(res.impact_step, N_STEPS_CAP)                 # (a host sync per sub-batch; the queue is already drained by early exit)
        for _, _, _, work in keep:
            if work is not None and rank != 0:
                work.wait()
        main.wait_stream(s_copy)
        return steps

    one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 0
    for _ in range(passes):
        steps = one_pass()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / passes
    t = torch.tensor([ms, float(steps)], dtype=torch.float64, device=dev)
    if world > 1:
        a = t.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = t.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        ms, steps_total = float(a[0]), float(b[1])
    else:
        steps_total = float(steps)
    if rank != 0:
        return None
    flat = host.reshape(-1, 4)
    st = (flat[:, 3] // 4294967296.0).to(torch.int64)
    hit = st == 1
    return {"value": steps_total / (ms * 1e-3), "unit": UNIT, "ms_per_pass": ms, "passes": passes,
            "n_trajectories_total": int(world * n_rank), "n_trajectories_per_gpu": int(n_rank), "sub_batches": sub_batches,
            "max_steps": N_STEPS_CAP, "traj_steps_per_pass": steps_total,
            "gather": "NCCL gather of [lat, lon, impact_time, step | status] (4 f64) per trajectory to rank 0, then D2H" if world > 1
                      else "single GPU: D2H of the same 4 f64 per trajectory",
            "gather_bytes_per_pass": int((world - 1) * n_rank * 32), "d2h_bytes_per_pass": int(world * n_rank * 32),
            "h2d_bytes_per_pass": 0, "impacted": int(hit.sum()), "not_impacted_within_cap": int((st == 0).sum()),
            "nonfinite": int((st == 2).sum()),
            "impact_lat_mean_std": [float(flat[hit, 0].mean()), float(flat[hit, 0].std())],
            "impact_lon_mean_std": [float(flat[hit, 1].mean()), float(flat[hit, 1].std())],
            "checksum_impact_time": float(flat[hit, 2].sum()),
            "config": "BASELINE.json configs[4] (C5), weak-scaled: 12.5 M trajectories per GPU = 100 M at 8 GPUs"}


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
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--c5-traj", type=int, default=12_500_000, help="C5 trajectories per GPU")
    ap.add_argument("--c5-passes", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from reentry_old_b200 import SpacecraftModel
    from reentry_old_b200.model import DeviceContext
    from reentry_old_b200.sharding import gather_summaries

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device(dev))
    W = max(args.warmup, 3)
    K = args.steps
    w = workload(args.n_traj, rank)
    n = w.n
    model = SpacecraftModel(Cd=w.Cd, A=w.A, m=w.m, epoch=w.epoch_jd, gmst0=w.gmst0, device=local_rank)
    ctx = DeviceContext.get(local_rank)
    params = torch.as_tensor(w.params, device=dev)
    y0 = model.get_initial_states(*params.t().contiguous(), gmst=w.gmst0)   # resident in HBM before timing
    torch.cuda.synchronize()
    # packed sample arena: ~86 valid samples per trajectory (+ the columns the lagged live count leaves as padding)
    pack_cap = int(RB_FIELDS * n * 112)

    def one_step(profile=False):
        r = model.run_ensemble(y0, t0=0.0, dt=w.dt, n_steps=w.n_steps, out_stride=w.out_stride,
                               steps_per_launch=w.steps_per_launch, profile=profile, samples="packed", packed_capacity=pack_cap)
        g = gather_summaries(r, dst=0, counts=[n] * world) if world > 1 else None   # (known shard sizes: no count exchange)
        return r, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        r, _ = one_step()
    torch.cuda.synchronize()
    steps_per_pass = steps_of(r.impact_step, w.n_steps)
    launches_per_pass = r.stats["kernel_launches"]
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
        r, gathered = one_step()
    ev1.record()
    barrier()
    t_end = time.time()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms, float(steps_per_pass)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        by_rank = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(by_rank, t)
        ms_by_rank = [float(x[0]) / K for x in by_rank]
        ms, total_steps_pass = float(tmax[0]), float(tsum[1])
    else:
        total_steps_pass = float(steps_per_pass)
        ms_by_rank = [ms / K]
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = total_steps_pass * K / (ms * 1e-3)

    # ---- parity of the TIMED batch (the last timed pass's results) against the oracle; cross-rank identity at N > 1
    parity = None
    if not args.no_parity:
        parity = parity_block(torch, dist, model, w, r, rank, world, y0, gathered)
    valid_samples = float(r.n_samples.to(torch.float64).sum())
    packed_doubles = int(r.packed.used)
    del r, gathered

    # ---- roofline of the dominant kernel (k_propagate): a SEPARATE pass with RB_FLAG_PROFILE, which brackets the
    #      step-kernel launch sequence with CUDA events on its stream (the timed passes above run without it)
    r, _ = one_step(profile=True)
    torch.cuda.synchronize()
    kern_ms = r.stats["propagate_ms"]
    n_prop = r.stats["propagate_launches"]
    peak = ctx.fp64_peak_tflops(300.0)
    achieved = steps_per_pass * FLOP_PER_TRAJ_STEP / (kern_ms * 1e-3) / 1e12
    sample_bytes = valid_samples * SAMPLE_BYTES + n * SUMMARY_BYTES
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ncu_nums = _ncu_pipe_numbers()
    traffic = (ncu_nums["dram_bytes_per_traj_step_ncu"] * steps_per_pass / max(1, n_prop)
               if "dram_bytes_per_traj_step_ncu" in ncu_nums else None)   # DRAM bytes per k_propagate launch
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_propagate", "kernel_ms_per_pass": kern_ms, "kernel_launches_per_pass": n_prop,
                "kernel_share_of_step": kern_ms / (ms / K),
                "timing": "kernel_ms_per_pass is the CUDA-event span of the step-kernel launch sequence in a separate pass run "
                          "with RB_FLAG_PROFILE after the timed region (same inputs); its share of a timed step can therefore "
                          "read a little above 1",
                "flop_per_traj_step": FLOP_PER_TRAJ_STEP, "flop_source": FLOP_SOURCE, "ncu": ncu_nums,
                "peak_source": "DFMA probe kernel run in this process (MEASURED_PEAKS.json has no FP64 figure); "
                               "nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
                "output_write": {"bytes_per_pass": sample_bytes, "packed_arena_bytes": packed_doubles * 8,
                                 "achieved_gbs": sample_bytes / (kern_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
                                 "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                                 "frac": sample_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak}}
    del r
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer API: pinned host memory in (y0) and out (packed samples + summaries)
    e2e = None
    if not args.no_e2e:
        bufs = model.host_buffers(n, w.n_steps, w.out_stride, packed_capacity=pack_cap, steps_per_launch=w.steps_per_launch)
        bufs["y0"].copy_(y0)
        for _ in range(2):
            rh = model.run_ensemble_host(bufs, t0=0.0, dt=w.dt)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            rh = model.run_ensemble_host(bufs, t0=0.0, dt=w.dt)
        barrier()
        el = time.perf_counter() - t0
        e2e_steps = steps_of(rh.impact_step.numpy(), w.n_steps)
        sent = int(rh.stats["sample_bytes_sent"])        # what the copy engine moved: live columns of every slab
        tt = torch.tensor([el, float(e2e_steps), float(sent)], dtype=torch.float64, device=dev)
        if world > 1:
            a = tt.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = tt.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            el, e2e_total, sent_total = float(a[0]), float(b[1]), float(b[2])
        else:
            e2e_total, sent_total = float(e2e_steps), float(sent)
        table_steps = min(w.n_steps, -(-rh.stats["steps_enqueued"] // 512) * 512)
        d2h = sent + n * SUMMARY_BYTES
        ingest = ingest_probe(torch, dist, dev, world, rank, bufs["packed"])
        bound_ms = max(kern_ms, 1e3 * (sent_total + world * n * SUMMARY_BYTES) / (ingest["concurrent"] * 1e9))
        e2e = {"value": e2e_total * K / el, "unit": UNIT, "h2d_bytes_per_step": n * 48 + (5 * table_steps + 1) * 96,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * el / K,
               "valid_sample_bytes_per_step": int(valid_samples * SAMPLE_BYTES),
               "api": "SpacecraftModel.run_ensemble_host -> rb_propagate_host_packed (C ABI): pinned host buffers; the time "
                      "table is built on the host in 512-step chunks while the GPU runs; every packed sample slab (rows as "
                      "wide as the live ensemble) leaves as ONE copy-engine transfer when its launch has finished",
               "host_ingest_gbs": {"one_rank_alone": ingest["alone"], "all_ranks_concurrent": ingest["concurrent"],
                                   "how": "2 GiB device -> pinned host per rank in 64 MiB copies, all ranks at once / rank 0 alone"},
               "bound_ms": bound_ms,
               "bound": "max(step-kernel time, all ranks' D2H bytes / measured concurrent host ingest)",
               "ms_over_bound": (1e3 * el / K) / bound_ms}
        del bufs, rh
        torch.cuda.empty_cache()

    # ---- C5: 100 M-trajectory dispersion sharded over the GPUs (12.5 M per GPU), NCCL gather of the impact points
    c5 = None
    if not args.no_c5:
        c5 = c5_block(torch, dist, model, dev, rank, world, args.c5_traj, args.c5_passes)

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline_block(w, 12.0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config_dict(w, world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches_per_pass * K), "roofline": roofline, "cpu_baseline": cpu,
                "traj_steps_per_pass": total_steps_pass, "parity": parity, "e2e_c5": c5,
                "ms_per_step_by_rank": ms_by_rank,
                "ms_per_step_note": "device time (CUDA events) of each rank's K timed passes / K; ms_per_step is their maximum; at "
                                    "N > 1 a pass includes the NCCL gather of the [n, 10] summaries to rank 0"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
