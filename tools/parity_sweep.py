"""TEST INFRASTRUCTURE (uses the CPU oracle as the checker).  Wide randomized parity sweep on the GPU box: the CUDA
path against the C oracle on ensembles no golden file contains -- several epochs / GMST / step sizes, reentry, LEO,
HEO and GEO regimes, per-trajectory Cd / A / m.

    python tools/parity_sweep.py [n_per_case]
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(4242)
MU, R = 3.986004418e14, 6378137.0
cases = []
# (name, dt, n_steps, stride, generator of launch parameters v, lat, lon, alt, az, gamma)
def reentry(k):
    return np.column_stack([rng.uniform(6500, 8200, k), rng.uniform(-85, 85, k), rng.uniform(-180, 180, k),
                            rng.uniform(70e3, 140e3, k), rng.uniform(0, 360, k), rng.uniform(-8, -0.3, k)])
def leo(k):
    alt = rng.uniform(200e3, 1500e3, k)
    return np.column_stack([np.sqrt(MU / (R + alt)) * rng.uniform(0.98, 1.05, k), rng.uniform(-80, 80, k), rng.uniform(-180, 180, k),
                            alt, rng.uniform(0, 360, k), rng.normal(0, 0.5, k)])
def high(k):
    alt = 10 ** rng.uniform(6.5, 7.7, k)
    return np.column_stack([np.sqrt(MU / (R + alt)) * rng.uniform(0.8, 1.2, k), rng.uniform(-70, 70, k), rng.uniform(-180, 180, k),
                            alt, rng.uniform(0, 360, k), rng.normal(0, 5, k)])
cases = [("reentry dt 0.5", 0.5, 3000, 50, reentry, 2460310.5, 0.0), ("reentry dt 1", 1.0, 1500, 25, reentry, 2461000.25, 2.5),
         ("LEO dt 10", 10.0, 2000, 40, leo, 2459000.5, 4.0), ("LEO dt 1", 1.0, 3000, 60, leo, 2460700.75, 1.0),
         ("HEO/GEO dt 60", 60.0, 2000, 40, high, 2460310.5, 5.5), ("HEO/GEO dt 10", 10.0, 2500, 50, high, 2462000.0, 0.3)]
worst = {}
for name, dt, n_steps, stride, gen, epoch, gmst0 in cases:
    p = gen(n)
    Cd, A, m = rng.uniform(0.8, 2.5, n), rng.uniform(1, 40, n), rng.uniform(100, 9000, n)
    mdl = SpacecraftModel(Cd=1.0, A=1.0, m=1.0, epoch=epoch, gmst0=gmst0)
    y0 = mdl.get_initial_states(*p.T, gmst=gmst0)
    r = mdl.run_ensemble(y0, dt=dt, n_steps=n_steps, out_stride=stride, Cd=Cd, A=A, m=m)
    torch.cuda.synchronize()
    t = time.perf_counter()
    ref = O.propagate(y0.cpu().numpy(), Cd, A, m, epoch, gmst0, 0.0, dt, n_steps, stride)
    el = time.perf_counter() - t
    got = r.samples.permute(2, 0, 1).cpu().numpy()          # [n, n_out, 8]
    a, b = got[:, :, :3], ref["samples"][:, :, :3]
    same_pattern = np.array_equal(np.isnan(a[..., 0]), np.isnan(b[..., 0]))
    ok = ~np.isnan(b[..., 0]) & ~np.isnan(a[..., 0])
    rel = np.linalg.norm(a[ok] - b[ok], axis=-1) / np.linalg.norm(b[ok], axis=-1)
    va, vb = got[:, :, 3:6], ref["samples"][:, :, 3:6]
    vrel = np.linalg.norm(va[ok] - vb[ok], axis=-1) / np.linalg.norm(vb[ok], axis=-1)
    # conditioning of each trajectory, two ways:
    #  (a) the oracle re-run from initial states perturbed by ONE ulp-sized relative change (round 1's yardstick: a
    #      single kick at t = 0 -- an implementation difference kicks at every stage of every step, so errors of
    #      10^2..10^3 times this are expected and say nothing);
    #  (b) the IMPLEMENTATION NOISE: the same oracle source compiled with FMA contraction (oracle/Makefile), i.e. a second,
    #      equally valid rounding of the same algorithm at every operation -- how far two correct implementations
    #      drift apart on this trajectory.  This is the yardstick tests/test_parity_sweep.py gates on.
    y0p = y0.cpu().numpy() * (1.0 + 2.2e-16 * rng.choice([-1.0, 1.0], size=(n, 6)))
    ref2 = O.propagate(y0p, Cd, A, m, epoch, gmst0, 0.0, dt, n_steps, stride)
    ref3 = O.propagate(y0.cpu().numpy(), Cd, A, m, epoch, gmst0, 0.0, dt, n_steps, stride, fma=True)
    b2, b3 = ref2["samples"][:, :, :3], ref3["samples"][:, :, :3]
    with np.errstate(invalid="ignore"):
        e_traj = np.nanmax(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1), axis=1)
        s_traj = np.nanmax(np.linalg.norm(b2 - b, axis=-1) / np.linalg.norm(b, axis=-1), axis=1)
        n_traj = np.nanmax(np.linalg.norm(b3 - b, axis=-1) / np.linalg.norm(b, axis=-1), axis=1)
    ratio = e_traj / np.maximum(s_traj, 1e-16)
    ratio_n = e_traj / np.maximum(n_traj, 1e-15)
    min_alt = np.nanmin(ref["samples"][:, :, 6], axis=1)
    worst_i = np.argsort(-ratio_n)[:3]
    worst[name] = [dict(i=int(i), error=float(e_traj[i]), noise_two_cpu_roundings=float(n_traj[i]), one_ulp_kick=float(s_traj[i]),
                        min_altitude_km=float(min_alt[i] / 1e3), start_altitude_km=float(p[i, 3] / 1e3), gamma_deg=float(p[i, 5]),
                        v0=float(p[i, 0]), impact_step=int(ref["impact_step"][i])) for i in worst_i]
    big = e_traj > 1e-9
    step_equal = int((r.impact_step.cpu().numpy() == ref["impact_step"]).sum())
    hit = ref["impact_step"] > 0
    dte = float(np.nanmax(np.abs(r.impact_time.cpu().numpy()[hit] - ref["impact_time"][hit]))) if hit.any() else 0.0
    out = dict(case=name, n=n, dt=dt, n_steps=n_steps, pos_rel_max=float(rel.max()), vel_rel_max=float(vrel.max()),
               same_valid_pattern=bool(same_pattern), impact_step_equal=f"{step_equal}/{n}", impacted=int(hit.sum()),
               impact_time_abs_max_s=dte, nonfinite=int((r.status == 2).sum()), oracle_s=round(el, 1),
               pos_rel_median=float(np.median(e_traj)), pos_rel_p99=float(np.percentile(e_traj, 99)),
               one_ulp_sensitivity_median=float(np.median(s_traj)), one_ulp_sensitivity_max=float(s_traj.max()),
               error_over_sensitivity_median=float(np.median(ratio)), error_over_sensitivity_max=float(ratio.max()),
               well_conditioned_pos_rel_max=float(e_traj[s_traj < 1e-12].max()) if (s_traj < 1e-12).any() else None,
               well_conditioned=int((s_traj < 1e-12).sum()),
               noise_median=float(np.median(n_traj)), noise_max=float(n_traj.max()),
               error_over_noise_median=float(np.median(ratio_n)), error_over_noise_p99=float(np.percentile(ratio_n, 99)),
               error_over_noise_max=float(ratio_n.max()), above_1e9=int(big.sum()),
               above_1e9_error_over_noise_max=float(ratio_n[big].max()) if big.any() else None,
               above_1e9_min_altitude_km=[float(np.min(min_alt[big]) / 1e3), float(np.max(min_alt[big]) / 1e3)] if big.any() else None,
               worst_by_error_over_noise=worst[name])
    print(json.dumps(out), flush=True)
