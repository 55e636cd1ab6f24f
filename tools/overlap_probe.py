"""Experiment: does splitting the ensemble over P concurrent streams (separate contexts) hide the
per-launch tails?  python tools/overlap_probe.py [n] [P]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params
from reentry_old_b200.model import DeviceContext

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0)
y0 = m.get_initial_states(*reentry_dispersion_params(n, 1234).T)
for P in (1, 2, 3, 4):
    ctxs = [DeviceContext(0) for _ in range(P)]
    streams = [torch.cuda.Stream() for _ in range(P)]
    parts = torch.chunk(y0, P)
    def go():
        out = []
        for c, s, part in zip(ctxs, streams, parts):
            DeviceContext._cache[0] = c
            with torch.cuda.stream(s):
                out.append(m.run_ensemble(part, dt=0.5, n_steps=1728, out_stride=2048, store_samples=False, early_exit=False))
        return out
    go(); torch.cuda.synchronize()
    t = time.perf_counter(); r = go(); torch.cuda.synchronize(); el = time.perf_counter() - t
    steps = sum(int(x.impact_step.clamp(min=0).sum()) for x in r)
    print(f"P={P}: {el*1e3:.1f} ms, {steps/el/1e9:.3f} G steps/s, hit {sum(int((x.status==1).sum()) for x in r)}")
