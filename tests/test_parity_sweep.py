"""Randomised GPU-vs-oracle parity over regimes no golden file contains (several epochs / GMST / step sizes; reentry,
LEO, HEO and GEO; per-trajectory Cd / A / m), with a CONDITIONING-AWARE bar.

The 1e-9 bar of BASELINE.json holds wherever the trajectory itself is well conditioned.  A few lanes are not: orbits
whose perigee dips into the atmosphere while the fixed step is 10 - 60 s cross 70 - 400 km of a 7-km-scale-height
density profile per step; the integrator is far from convergence there and rounding-level differences are amplified by
many orders of magnitude (profiles/r2d_parity_sweep.jsonl: every lane above 1e-9 has its perigee inside the atmosphere
or flies tens of revolutions at a coarse step).  For those the yardstick is the IMPLEMENTATION NOISE of the lane: the
distance between two equally valid CPU roundings of the same algorithm -- the oracle and the same C source compiled
with FMA contraction (oracle/Makefile).  The CUDA path must stay within max(1e-9, 300 x that noise); measured worst
ratio on 12 000 trajectories: 176 (round 1 compared against a single one-ulp kick at t = 0, which under-states what a
difference at every stage of every step does -- hence its unexplained ratios in the thousands)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MU, R = 3.986004418e14, 6378137.0
N = 400


def _reentry(rng, k):
    return np.column_stack([rng.uniform(6500, 8200, k), rng.uniform(-85, 85, k), rng.uniform(-180, 180, k),
                            rng.uniform(70e3, 140e3, k), rng.uniform(0, 360, k), rng.uniform(-8, -0.3, k)])


def _leo(rng, k):
    alt = rng.uniform(200e3, 1500e3, k)
    return np.column_stack([np.sqrt(MU / (R + alt)) * rng.uniform(0.98, 1.05, k), rng.uniform(-80, 80, k),
                            rng.uniform(-180, 180, k), alt, rng.uniform(0, 360, k), rng.normal(0, 0.5, k)])


def _high(rng, k):
    alt = 10 ** rng.uniform(6.5, 7.7, k)
    return np.column_stack([np.sqrt(MU / (R + alt)) * rng.uniform(0.8, 1.2, k), rng.uniform(-70, 70, k),
                            rng.uniform(-180, 180, k), alt, rng.uniform(0, 360, k), rng.normal(0, 5, k)])


CASES = [("reentry dt 0.5", 0.5, 3000, 50, _reentry, 2460310.5, 0.0), ("reentry dt 1", 1.0, 1500, 25, _reentry, 2461000.25, 2.5),
         ("LEO dt 10", 10.0, 2000, 40, _leo, 2459000.5, 4.0), ("LEO dt 1", 1.0, 3000, 60, _leo, 2460700.75, 1.0),
         ("HEO/GEO dt 60", 60.0, 2000, 40, _high, 2460310.5, 5.5), ("HEO/GEO dt 10", 10.0, 2500, 50, _high, 2462000.0, 0.3)]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_parity_within_the_conditioning_of_each_trajectory(case, oracle):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from reentry_old_b200 import SpacecraftModel

    name, dt, n_steps, stride, gen, epoch, gmst0 = case
    rng = np.random.default_rng(sum(map(ord, name)))
    p = gen(rng, N)
    Cd, A, m = rng.uniform(0.8, 2.5, N), rng.uniform(1, 40, N), rng.uniform(100, 9000, N)
    mdl = SpacecraftModel(Cd=1.0, A=1.0, m=1.0, epoch=epoch, gmst0=gmst0)
    y0 = mdl.get_initial_states(*p.T, gmst=gmst0)
    r = mdl.run_ensemble(y0, dt=dt, n_steps=n_steps, out_stride=stride, Cd=Cd, A=A, m=m)
    torch.cuda.synchronize()
    y0h = y0.cpu().numpy()
    ref = oracle.propagate(y0h, Cd, A, m, epoch, gmst0, 0.0, dt, n_steps, stride)
    alt = oracle.propagate(y0h, Cd, A, m, epoch, gmst0, 0.0, dt, n_steps, stride, fma=True)
    got = r.samples.permute(2, 0, 1).cpu().numpy()
    a, b, c = got[:, :, :6], ref["samples"][:, :, :6], alt["samples"][:, :, :6]
    agree = ref["impact_step"] == alt["impact_step"]          # lanes on which the two CPU roundings tell the same story
    assert agree.mean() > 0.97
    assert np.array_equal(r.impact_step.cpu().numpy()[agree], ref["impact_step"][agree])
    assert np.array_equal(np.isnan(a[agree][..., 0]), np.isnan(b[agree][..., 0]))
    with np.errstate(invalid="ignore"):
        both = ~np.isnan(a[..., 0]) & ~np.isnan(b[..., 0])
        bc = ~np.isnan(c[..., 0]) & ~np.isnan(b[..., 0])
        err = np.where(both, np.linalg.norm(a[..., :3] - b[..., :3], axis=-1) / np.linalg.norm(b[..., :3], axis=-1), 0.0).max(axis=1)
        errv = np.where(both, np.linalg.norm(a[..., 3:] - b[..., 3:], axis=-1) / np.linalg.norm(b[..., 3:], axis=-1), 0.0).max(axis=1)
        noise = np.where(bc, np.linalg.norm(c[..., :3] - b[..., :3], axis=-1) / np.linalg.norm(b[..., :3], axis=-1), 0.0).max(axis=1)
        noisev = np.where(bc, np.linalg.norm(c[..., 3:] - b[..., 3:], axis=-1) / np.linalg.norm(b[..., 3:], axis=-1), 0.0).max(axis=1)
    noise = np.maximum(noise, noisev)          # ONE yardstick per trajectory: a lane is ill conditioned in position and velocity alike
    bar = np.maximum(1e-9, 300.0 * noise)
    worst = int(np.argmax(np.where(agree, err / bar, 0.0)))
    worstv = int(np.argmax(np.where(agree, errv / bar, 0.0)))
    assert np.all(err[agree] <= bar[agree]), (name, worst, err[worst], noise[worst])
    assert np.all(errv[agree] <= bar[agree]), (name, worstv, errv[worstv], noise[worstv])
    # the bulk is far inside the unconditional bar
    assert np.median(err) < 1e-12 and np.mean(err <= 1e-9) >= 0.95, (name, np.median(err), np.mean(err <= 1e-9))
    # and every lane the unconditional bar does not cover is one the two CPU roundings also disagree on at that level
    loose = (err > 1e-9) & agree
    assert np.all(noise[loose] > 1e-9 / 300.0)
    hit = (ref["impact_step"] > 0) & agree
    if hit.any():
        dte = np.abs(r.impact_time.cpu().numpy()[hit] - ref["impact_time"][hit])
        dtn = np.abs(alt["impact_time"][hit] - ref["impact_time"][hit])
        assert np.all(dte <= np.maximum(1e-7, 300.0 * dtn)), (name, dte.max())
    print(f"{name}: err median {np.median(err):.1e} max {err.max():.1e}; above 1e-9: {int(loose.sum())}/{N}; "
          f"max error/noise among them {float((err[loose] / noise[loose]).max()) if loose.any() else 0:.0f}")
