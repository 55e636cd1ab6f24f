"""TEST INFRASTRUCTURE (uses the CPU oracle as the checker; not collected by pytest).  Run the BASELINE.json configurations C1..C4 on the GPU box: throughput + parity of a subsample
against the CPU oracle (compared horizon of SURVEY.md 7.2(5)) + full-horizon drift.

    python tests/run_configs.py [c1 c2 c3 c4] [--scale 1.0]   -> one JSON line per config
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import oracle as O
from reentry_old_b200 import SpacecraftModel
from reentry_old_b200 import configs as CF


def rel(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)))


def run(w, sub, horizon_steps, drift_n=0, skip=False):
    m = SpacecraftModel(Cd=float(np.atleast_1d(w.Cd)[0]), A=float(np.atleast_1d(w.A)[0]), m=float(np.atleast_1d(w.m)[0]),
                        epoch=w.epoch_jd, gmst0=w.gmst0)
    per = {k: (v if np.ndim(v) else None) for k, v in (("Cd", w.Cd), ("A", w.A), ("m", w.m))}
    y0 = m.get_initial_states(*w.params.T, gmst=w.gmst0) if w.params is not None else torch.as_tensor(w.y0, device="cuda")
    t_tab = time.perf_counter()
    r = m.run_ensemble(y0, dt=w.dt, n_steps=w.n_steps, out_stride=w.out_stride, steps_per_launch=w.steps_per_launch, skip_negligible_drag=skip, **per)
    torch.cuda.synchronize()
    first = time.perf_counter() - t_tab           # includes the host time-table build
    t0 = time.perf_counter()
    r = m.run_ensemble(y0, dt=w.dt, n_steps=w.n_steps, out_stride=w.out_stride, steps_per_launch=w.steps_per_launch,
                       profile=True, skip_negligible_drag=skip, **per)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    steps = int(torch.where(r.impact_step > 0, r.impact_step, torch.full_like(r.impact_step, w.n_steps)).sum())
    out = {"config": w.name + (" [skip negligible drag > 400 km]" if skip else ""), "n": w.n, "dt": w.dt, "n_steps": w.n_steps, "out_stride": w.out_stride,
           "traj_steps": steps, "wall_s": el, "first_call_s_incl_table": first, "kernel_s": r.stats["propagate_ms"] * 1e-3,
           "traj_steps_per_s_wall": steps / el, "traj_steps_per_s_kernel": steps / (r.stats["propagate_ms"] * 1e-3),
           "impacted": int((r.status == 1).sum()), "nonfinite": int((r.status == 2).sum()),
           "launches": r.stats["propagate_launches"]}
    # parity of a subsample over the compared horizon
    idx = np.linspace(0, w.n - 1, sub).astype(int) if w.n > 1 else np.array([0])
    y0h = y0[idx].cpu().numpy()
    pk = lambda v: (np.asarray(v)[idx] if np.ndim(v) else v)
    hs = min(horizon_steps, w.n_steps)
    ref = O.propagate(y0h, pk(w.Cd), pk(w.A), pk(w.m), w.epoch_jd, w.gmst0, 0.0, w.dt, hs, w.out_stride)
    k = hs // w.out_stride + 1
    got = r.samples[:k][:, :, idx].permute(2, 0, 1).cpu().numpy()
    ok = ~np.isnan(ref["samples"][:, :, 0])
    same_valid = bool(np.array_equal(ok, ~np.isnan(got[:, :, 0]))) if hs == w.n_steps else bool(np.all(~np.isnan(got[:, :, 0][ok])))
    out["parity"] = {"subsample": len(idx), "horizon_steps": hs, "valid_pattern_equal": same_valid,
                     "pos_rel_max": rel(got[ok][:, :3], ref["samples"][ok][:, :3]),
                     "vel_rel_max": rel(got[ok][:, 3:6], ref["samples"][ok][:, 3:6]),
                     "pos_abs_max_m": float(np.max(np.linalg.norm(got[ok][:, :3] - ref["samples"][ok][:, :3], axis=1)))}
    if hs == w.n_steps:
        out["parity"]["impact_step_equal"] = bool(np.array_equal(r.impact_step[idx].cpu().numpy(), ref["impact_step"]))
        hit = ref["impact_step"] > 0
        if hit.any():
            out["parity"]["impact_time_abs_max_s"] = float(np.max(np.abs(r.impact_time[idx].cpu().numpy()[hit] - ref["impact_time"][hit])))
    if drift_n:   # full-horizon drift vs the oracle on a few trajectories
        j = idx[:drift_n]
        pk2 = lambda v: (np.asarray(v)[j] if np.ndim(v) else v)
        t1 = time.perf_counter()
        ref2 = O.propagate(y0[j].cpu().numpy(), pk2(w.Cd), pk2(w.A), pk2(w.m), w.epoch_jd, w.gmst0, 0.0, w.dt, w.n_steps,
                           w.out_stride)
        got2 = r.samples[:, :, j].permute(2, 0, 1).cpu().numpy()
        out["full_horizon_drift"] = {"trajectories": int(drift_n), "oracle_s": time.perf_counter() - t1,
                                     "pos_rel_end": rel(got2[:, -1, :3], ref2["samples"][:, -1, :3]),
                                     "pos_rel_max": rel(got2[:, :, :3], ref2["samples"][:, :, :3])}
    print(json.dumps(out), flush=True)


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c3", "c4"]
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    if "c1" in which:
        run(CF.c1_single(), 1, 800)
    if "c2" in which:
        run(CF.c2_reentry_mc(int(1_000_000 * scale)), 64, 2048)
    if "c3" in which:
        run(CF.c3_sso_decay(int(100_000 * scale)), 16, 8640, drift_n=16)
        run(CF.c3_sso_decay(int(100_000 * scale)), 16, 8640, skip=True)
    if "c4" in which:
        run(CF.c4_geo_heo(int(50_000 * scale)), 16, 8640, drift_n=16)
        run(CF.c4_geo_heo(int(50_000 * scale)), 16, 8640, skip=True)


if __name__ == "__main__":
    main()
