"""TEST / MEASUREMENT INFRASTRUCTURE -- the reference's own numba-compiled path, from oracle/_ref.

oracle/build_ref.py compiles (numba.pycc, ahead of time) the O3 driver -- a fixed-step Dormand-Prince loop whose RHS
is the reference's unmodified @jit leaves and the glue of spacecraft_model.py:437-470 -- from /root/reference in the
build container into oracle/_ref/ref_o3_native.*.so.  This module loads that extension (no reference source, no JIT
at run time) and drives it over all host cores with one worker PROCESS per core (an AOT module has no prange and holds
the GIL).  Only bench.py's cpu_baseline leg and tests/ may import it.
"""
from __future__ import annotations

import glob
import importlib.util
import multiprocessing as mp
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def module_path():
    c = sorted(glob.glob(os.path.join(HERE, "_ref", "ref_o3_native*.so")))
    return c[0] if c else None


def available() -> bool:
    return module_path() is not None


def load():
    global _mod
    if _mod is None:
        p = module_path()
        if p is None:
            raise RuntimeError("oracle/_ref holds no ref_o3_native module (python oracle/build_ref.py, build container only)")
        spec = importlib.util.spec_from_file_location("ref_o3_native", p)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def propagate(y0, Cd, A, m, epoch_jd, gmst0, t0, dt, n_steps, stride):
    """Same contract as ref_shim.o3_run's core: returns dict(ys [n, n_out, 6] NaN after the event, impact_step, n_done)."""
    mod = load()
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1, 6)
    n = y0.shape[0]
    n_out = n_steps // stride + 1
    ys = np.full((n, n_out, 6), np.nan)
    impact_step = np.zeros(n, dtype=np.int64)
    n_done = np.zeros(n, dtype=np.int64)
    K_event = np.zeros((n, 8, 6))
    bc = lambda v: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))  # noqa: E731
    mod.propagate_batch(y0, float(t0), float(dt), int(n_steps), int(stride), float(epoch_jd), float(gmst0), bc(Cd), bc(A), bc(m),
                        ys, impact_step, n_done, K_event)
    return dict(ys=ys, impact_step=impact_step, n_done=n_done, K_event=K_event)


def _worker(args):
    y0, Cd, A, m, epoch_jd, gmst0, dt, n_steps = args
    t = time.perf_counter()
    r = propagate(y0, Cd, A, m, epoch_jd, gmst0, 0.0, dt, n_steps, n_steps)
    el = time.perf_counter() - t
    steps = int(np.where(r["impact_step"] > 0, r["impact_step"], n_steps).sum())
    return steps, el


def rate(w, target_seconds=10.0, n_procs=0):
    """traj-steps/s of the reference's numba code on a bounded sample of workload `w` (reentry_old_b200.configs.Workload
    with launch-geometry params), all host cores.  Returns the dict bench.py reports."""
    from oracle import oracle as O   # only for get_initial_states of the sample (not timed)

    if n_procs <= 0:
        n_procs = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load()
    p = w.params
    probe = O.get_initial_states(*p[:8].T, w.gmst0)
    s, el = _worker((probe, w.Cd, w.A, w.m, w.epoch_jd, w.gmst0, w.dt, w.n_steps))
    per_traj = el / 8
    n = int(max(n_procs, min(w.n, target_seconds / per_traj * n_procs)))
    n = (n // n_procs) * n_procs
    y0 = O.get_initial_states(*p[:n].T, w.gmst0)
    chunks = [(c, w.Cd, w.A, w.m, w.epoch_jd, w.gmst0, w.dt, w.n_steps) for c in np.array_split(y0, n_procs)]
    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs) as pool:
        pool.map(_worker, [(probe[:1], w.Cd, w.A, w.m, w.epoch_jd, w.gmst0, w.dt, w.n_steps)] * n_procs)   # workers up, module loaded
        t = time.perf_counter()
        res = pool.map(_worker, chunks, chunksize=1)
        wall = time.perf_counter() - t
    steps = sum(s for s, _ in res)
    return {"value": steps / wall, "unit": "traj-steps/s", "cores": n_procs, "per_core": steps / sum(e for _, e in res),
            "kind": "reference",
            "what": "the reference's own @jit functions (spacecraft_model.py:71-388, coordinate_converter.py:29-116) compiled "
                    "ahead of time by numba (oracle/build_ref.py -> oracle/_ref, x86-64-v3) under the O3 fixed-step "
                    "Dormand-Prince driver, one process per core",
            "sample": f"first {n} trajectories of the workload, {steps} traj-steps in {wall:.1f} s"}
