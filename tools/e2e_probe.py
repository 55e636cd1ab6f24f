"""Where does the end-to-end time go?  D2H bandwidth of the box, rb_propagate_host with / without samples."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reentry_old_b200 import SpacecraftModel, _lib as L
from reentry_old_b200.configs import c2_reentry_mc
from reentry_old_b200.model import DeviceContext

x = torch.empty(2 * 1024**3 // 8, dtype=torch.float64, device="cuda"); h = torch.empty_like(x, device="cpu").pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); el = time.perf_counter() - t
print(f"D2H pinned 2 GiB: {2.147/el:.1f} GB/s")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize(); el = time.perf_counter() - t
print(f"H2D pinned 2 GiB: {2.147/el:.1f} GB/s")
del x, h
w = c2_reentry_mc(1_000_000)
m = SpacecraftModel(Cd=w.Cd, A=w.A, m=w.m); ctx = DeviceContext.get(); lib = ctx.lib
y0 = m.get_initial_states(*torch.as_tensor(w.params, device="cuda").t().contiguous())
n, n_out = w.n, w.n_steps // w.out_stride + 1
y0_h = torch.empty((n, 6), dtype=torch.float64).pin_memory(); y0_h.copy_(y0)
fs = torch.empty((n, 6), dtype=torch.float64).pin_memory(); ist = torch.empty(n, dtype=torch.int32).pin_memory()
it = torch.empty(n, dtype=torch.float64).pin_memory(); st = torch.empty(n, dtype=torch.uint8).pin_memory(); ns = torch.empty(n, dtype=torch.int32).pin_memory()
used = C.c_int64(0)
for stride, with_samples in ((16, True), (16, False)):
    no = w.n_steps // stride + 1
    sm = torch.empty((no, 8, n), dtype=torch.float64, pin_memory=True) if with_samples else None
    run = L.RbRun(first_step=0, n_steps=w.n_steps, out_stride=stride, steps_per_launch=int(os.environ.get('RB_SPL', '32')), event_altitude=1000.0,
                  flags=L.RB_FLAG_COMPACT | (L.RB_FLAG_STAGED_SAMPLES if os.environ.get('RB_E2E_STAGED') else 0)
                  | (L.RB_FLAG_DIRECT_SAMPLES if os.environ.get('RB_E2E_DIRECT') else 0) | (L.RB_FLAG_PROFILE if os.environ.get('RB_E2E_PROFILE') else 0))
    def go():
        ctx.check(lib.rb_propagate_host(ctx.handle, n, y0_h.data_ptr(), None, None, None, float(w.Cd), float(w.A), float(w.m), w.epoch_jd, w.gmst0, 0.0, w.dt, C.byref(run), sm.data_ptr() if sm is not None else None, no, fs.data_ptr(), ist.data_ptr(), it.data_ptr(), st.data_ptr(), ns.data_ptr(), C.byref(used)))
    go(); go()
    t = time.perf_counter(); go(); go(); el = (time.perf_counter() - t) / 2
    if os.environ.get('RB_E2E_PROFILE'):
        print("   k_propagate span (first launch .. join):", round(ctx.stats()["propagate_ms"], 1), "ms")
    print(f"stride {stride} samples {with_samples}: {el*1e3:.1f} ms/pass, rows {used.value}, D2H {used.value*64e-3*n/1e6:.2f} GB -> {used.value*64*n/el/1e9:.1f} GB/s")
    del sm
