"""Quick device-side throughput probe (GPU box): C2-shaped ensemble, prints traj-steps/s.

    python tools/quick_perf.py [n_traj] [reps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params
from reentry_old_b200.model import DeviceContext

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spl = int(os.environ.get("RB_SPL", "0"))
m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0)
p = reentry_dispersion_params(n, 1234)
y0 = m.get_initial_states(*p.T)
peak = DeviceContext.get().fp64_peak_tflops(200.0)
print(f"DFMA peak probe: {peak:.2f} TFLOP/s")
for stride, store in ((16, True), (2048, False)):
    for rep in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=stride, store_samples=store, profile=True, steps_per_launch=spl)
        torch.cuda.synchronize()
        el = time.perf_counter() - t
        steps = int(r.impact_step.clamp(min=0).sum())
        ms = r.stats["propagate_ms"]
        print(f"n={n} stride={stride} store={store} rep={rep}: wall {el*1e3:.1f} ms, k_propagate {ms:.1f} ms, "
              f"traj-steps {steps:.3e}, {steps/el/1e9:.3f} G steps/s wall, {steps/(ms*1e-3)/1e9:.3f} G steps/s kernel, "
              f"launches {r.stats['propagate_launches']}, hit {(r.status==1).sum().item()}")
