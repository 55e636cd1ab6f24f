"""Does a concurrent copy-engine D2H stream slow the step kernel down?  (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params

m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0)
y0 = m.get_initial_states(*reentry_dispersion_params(1_000_000, 1234).T)
src = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")      # 512 MB
dst = torch.empty_like(src, device="cpu").pin_memory()
side = torch.cuda.Stream()

def run(with_copy):
    torch.cuda.synchronize()
    n_copies = 0
    t = time.perf_counter()
    if with_copy:
        with torch.cuda.stream(side):
            for _ in range(10):                      # 5 GB queued on the copy engine (~90 ms at 56 GB/s)
                dst.copy_(src, non_blocking=True); n_copies += 1
    r = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=2048, store_samples=False, profile=True)
    torch.cuda.synchronize()
    el = time.perf_counter() - t
    return r.stats["propagate_ms"], el * 1e3, n_copies

for _ in range(2): run(False)
for wc in (False, True, False, True):
    k, wall, nc = run(wc)
    print(f"concurrent D2H copies {nc:2d} x 512 MB: k_propagate span {k:.1f} ms, wall {wall:.1f} ms")
