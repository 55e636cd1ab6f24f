"""How much does lane coherence buy?  C2 ensemble in caller order vs physically sorted by a key (GPU box).

    python tools/sort_probe.py [n_traj]
Keys: the impact step of a previous run (near-ideal static ordering), initial flight-path angle, initial speed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0)
p = reentry_dispersion_params(n, 1234)
y0 = m.get_initial_states(*p.T)


def run(y, tag):
    best = None
    for _ in range(3):
        r = m.run_ensemble(y, dt=0.5, n_steps=2048, out_stride=2048, store_samples=False, profile=True)
        torch.cuda.synchronize()
        ms = r.stats["propagate_ms"]
        best = ms if best is None else min(best, ms)
    steps = int(r.impact_step.clamp(min=0).sum())
    print(f"{tag:28s} k_propagate {best:7.2f} ms  {steps / best / 1e6:.3f} G traj-steps/s")
    return r


r = run(y0, "caller order")
for tag, key in (("sorted by impact step", r.impact_step.cpu().numpy()), ("sorted by gamma", p[:, 5]), ("sorted by speed", p[:, 0]),
                 ("sorted by altitude", p[:, 3])):
    order = torch.as_tensor(np.argsort(key, kind="stable"), device=y0.device)
    run(y0[order].contiguous(), tag)
