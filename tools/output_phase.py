"""Output-write phase on a sized-down FULL-output run (out_stride = 1): achieved HBM write rate of the step
kernel and its effect on throughput (full per-step output of C2 would be 87 GB)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0)
y0 = m.get_initial_states(*reentry_dispersion_params(n, 1234).T)
for stride in (1, 4, 16, 1024):
    r = m.run_ensemble(y0, dt=0.5, n_steps=1024, out_stride=stride, profile=True)      # nobody impacts before step 1196
    r = m.run_ensemble(y0, dt=0.5, n_steps=1024, out_stride=stride, profile=True)
    torch.cuda.synchronize()
    ms = r.stats["propagate_ms"]
    steps = n * 1024
    out_bytes = float(r.n_samples.sum()) * 64
    print(f"n={n} out_stride={stride}: kernel {ms:.1f} ms, {steps/(ms*1e-3)/1e9:.3f} G traj-steps/s, sample bytes {out_bytes/1e9:.2f} GB "
          f"-> {out_bytes/(ms*1e-3)/1e9:.1f} GB/s written")
    del r
    torch.cuda.empty_cache()
