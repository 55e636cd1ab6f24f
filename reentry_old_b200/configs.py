"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d) (host-side input generation only).

Each generator returns launch-geometry parameters (v m/s, lat deg, lon deg, alt m, azimuth deg,
gamma deg) for `SpacecraftModel.get_initial_states`, or direct ECI states, plus the run
parameters.  numpy.random.default_rng(seed) with the draw order documented here is part of
the workload definition (the golden fixtures in tests/golden were generated with it).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

EPOCH_JD = 2460310.5
EARTH_MU = 398600441800000.0
EARTH_R = 6378137.0


@dataclass
class Workload:
    name: str
    n: int
    dt: float
    n_steps: int
    out_stride: int
    Cd: object
    A: object
    m: object
    epoch_jd: float = EPOCH_JD
    gmst0: float = 0.0
    params: np.ndarray = None   # [n, 6] launch geometry, or None
    y0: np.ndarray = None       # [n, 6] direct ECI states, or None
    steps_per_launch: int = 32


def reentry_dispersion_params(n: int, seed: int = 1234) -> np.ndarray:
    """C2/C5: nominal (7800, 20, 0, 120e3, 90, -1), 1-sigma (10, .5, .5, 500, .5, .1); one
    standard-normal matrix z[n, 6] drawn row-major."""
    rng = np.random.default_rng(seed)
    z = rng.normal(size=(n, 6))
    nominal = np.array([7800.0, 20.0, 0.0, 120e3, 90.0, -1.0])
    sigma = np.array([10.0, 0.5, 0.5, 500.0, 0.5, 0.1])
    return nominal + sigma * z


def c1_single() -> Workload:
    p = np.array([[7800.0, 0.0, 0.0, 120e3, 90.0, -1.0]])
    return Workload("C1 single capsule reentry from 120 km, dt=1 s", 1, 1.0, 800, 1, 1.3, 14.0, 5000.0, params=p)


def c2_reentry_mc(n: int = 1_000_000, seed: int = 1234, n_steps: int = 2048, out_stride: int = 16) -> Workload:
    return Workload(f"C2 Monte Carlo reentry dispersion, {n} states, dt=0.5 s, until impact", n, 0.5, n_steps,
                    out_stride, 1.3, 14.0, 5000.0, params=reentry_dispersion_params(n, seed))


def c3_sso_decay(n: int = 100_000, seed: int = 2345, days: float = 30.0, out_stride: int = 360) -> Workload:
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.5, 2, n)
    m = rng.uniform(50, 500, n)
    vcirc = np.sqrt(EARTH_MU / (EARTH_R + 550e3))
    lat = rng.uniform(-80, 80, n)
    lon = rng.uniform(-180, 180, n)
    p = np.stack([vcirc + rng.normal(0, 1, n), lat, lon, 550e3 + rng.normal(0, 5e3, n),
                  np.full(n, 352.4), np.zeros(n)], axis=1)
    return Workload(f"C3 sun-synchronous 550 km ensemble, {n} sats, {days:g} d, dt=10 s", n, 10.0,
                    int(round(days * 8640)), out_stride, np.full(n, 2.2), A, m, params=p, steps_per_launch=1024)


def kepler_to_state(mu, a, e, inc, raan, argp, nu):
    p = a * (1 - e * e)
    r = p / (1 + e * np.cos(nu))
    rp = np.stack([r * np.cos(nu), r * np.sin(nu), np.zeros_like(nu)], axis=-1)
    k = np.sqrt(mu / p)
    vp = np.stack([-k * np.sin(nu), k * (e + np.cos(nu)), np.zeros_like(nu)], axis=-1)
    cO, sO, cw, sw, ci, si = np.cos(raan), np.sin(raan), np.cos(argp), np.sin(argp), np.cos(inc), np.sin(inc)
    R = np.stack([np.stack([cO * cw - sO * sw * ci, -cO * sw - sO * cw * ci, sO * si], -1),
                  np.stack([sO * cw + cO * sw * ci, -sO * sw + cO * cw * ci, -cO * si], -1),
                  np.stack([sw * si, cw * si, ci], -1)], -2)
    return np.concatenate([np.einsum("...ij,...j->...i", R, rp), np.einsum("...ij,...j->...i", R, vp)], axis=-1)


def c4_geo_heo(n: int = 50_000, seed: int = 3456, days: float = 365.0, out_stride: int = 1440) -> Workload:
    rng = np.random.default_rng(seed)
    h = n // 2
    geo = kepler_to_state(EARTH_MU, np.full(h, 42164e3), np.zeros(h), np.radians(np.abs(rng.normal(0, 0.1, h))),
                          rng.uniform(0, 2 * np.pi, h), np.zeros(h), rng.uniform(0, 2 * np.pi, h))
    rp, ra = EARTH_R + 600e3, EARTH_R + 39750e3
    k = n - h
    heo = kepler_to_state(EARTH_MU, np.full(k, (rp + ra) / 2), np.full(k, (ra - rp) / (ra + rp)),
                          np.full(k, np.radians(63.4)), rng.uniform(0, 2 * np.pi, k), np.full(k, np.radians(270.0)),
                          rng.uniform(0, 2 * np.pi, k))
    return Workload(f"C4 GEO/HEO luni-solar ensemble, {n} sats, {days:g} d, dt=60 s", n, 60.0,
                    int(round(days * 1440)), out_stride, 2.2, 10.0, 2000.0, y0=np.concatenate([geo, heo]),
                    steps_per_launch=1024)


def c5_reentry_sharded(n_total: int, rank: int, world: int, seed: int = 1234, n_steps: int = 2048) -> Workload:
    """C5: the C2 dispersion sharded over ranks; seed + rank per shard; summary output only."""
    lo, hi = shard_range(n_total, rank, world)
    w = c2_reentry_mc(hi - lo, seed + rank, n_steps, out_stride=n_steps)
    w.name = f"C5 reentry dispersion, {n_total} states over {world} GPUs (rank {rank}: {hi - lo})"
    return w


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
