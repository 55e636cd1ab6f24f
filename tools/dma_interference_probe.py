"""D2H copy-engine bandwidth alone / beside a register-resident DFMA kernel / beside the step kernel (GPU box)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params
from reentry_old_b200.model import DeviceContext

m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0)
y0 = m.get_initial_states(*reentry_dispersion_params(1_000_000, 1234).T)
ctx = DeviceContext.get()
src = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
dst = torch.empty_like(src, device="cpu").pin_memory()
side = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def copies(n=10):
    with torch.cuda.stream(side):
        e0.record(side)
        for _ in range(n):
            dst.copy_(src, non_blocking=True)
        e1.record(side)

def report(tag):
    torch.cuda.synchronize()
    print(f"{tag:42s} D2H 5 GB in {e0.elapsed_time(e1):6.1f} ms = {5.37e3 / e0.elapsed_time(e1):5.1f} GB/s")

copies(); report("warm-up")
copies(); report("copy engine alone")
copies(); ctx.fp64_peak_tflops(150.0); report("beside the DFMA peak kernel (150 ms)")
copies(); m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=2048, store_samples=False); report("beside k_propagate (no samples)")
copies(); m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=2048, store_samples=False, overlap=False); report("beside k_propagate, one partition")
