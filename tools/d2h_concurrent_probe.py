import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import c2_reentry_mc
w = c2_reentry_mc(1_000_000); m = SpacecraftModel(Cd=w.Cd, A=w.A, m=w.m)
y0 = m.get_initial_states(*torch.as_tensor(w.params, device="cuda").t().contiguous())
x = torch.empty(4 * 1024**3 // 8, dtype=torch.float64, device="cuda"); h = torch.empty_like(x, device="cpu").pin_memory()
cs = torch.cuda.Stream()
def d2h(chunks):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record()
        for a, b in zip(h.chunk(chunks), x.chunk(chunks)): a.copy_(b, non_blocking=True)
        e1.record()
    return e0, e1
for chunks in (1, 32):
    e0, e1 = d2h(chunks); torch.cuda.synchronize(); print(f"idle GPU, {chunks} chunk(s): {4.295/ (e0.elapsed_time(e1)*1e-3):.1f} GB/s")
for chunks in (1, 32):
    r = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=2048, store_samples=False, early_exit=False)  # async
    e0, e1 = d2h(chunks); torch.cuda.synchronize(); print(f"during propagate, {chunks} chunk(s): {4.295/ (e0.elapsed_time(e1)*1e-3):.1f} GB/s")
