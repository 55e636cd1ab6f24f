"""Adaptive mode (N4) ensemble throughput on the GPU box: exact per-stage ephemeris vs the interpolated grid.

    python tools/adaptive_perf.py [n_traj] [RK45,RK23,DOP853]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
for sim_type in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["RK45"]):
    m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0, sim_type=sim_type)
    y0 = m.get_initial_states(*reentry_dispersion_params(n, 7).T)
    for eph in (0.0, 16.0):
        for rep in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            r = m.run_ensemble_adaptive(y0, (0.0, 1024.0), 1.0, ephemeris_dt=eph, store_samples=False)
            torch.cuda.synchronize(); el = time.perf_counter() - t
        att = int(r.n_accepted.sum() + r.n_rejected.sum())
        print(f"{sim_type} n={n} ephemeris_dt={eph}: {el*1e3:.1f} ms, {att} step attempts, {att/el/1e6:.1f} M attempts/s, impacted {(r.status==1).sum().item()}")
