import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from reentry_old_b200.model import DeviceContext
ctx = DeviceContext.get(); rng = np.random.default_rng(0); n = 1_000_000
def run(op, x, y=None):
    xd = torch.as_tensor(x, device="cuda"); yd = torch.as_tensor(y, device="cuda") if y is not None else None
    out = torch.empty_like(xd)
    ctx.check(ctx.lib.rb_math_probe(ctx.handle, op, len(x), xd.data_ptr(), yd.data_ptr() if yd is not None else None, out.data_ptr(), None))
    torch.cuda.synchronize(); return out.cpu().numpy()
def ulps(got, ref): return np.abs(got - ref) / np.spacing(np.abs(ref))
x = 10 ** rng.uniform(-30, 30, n)
u = ulps(run(0, x), 1.0/x); print("rcp max ulp", u.max(), "mean", u.mean())
u = ulps(run(1, x), 1.0/np.sqrt(x)); print("rsqrt max ulp", u.max(), "mean", u.mean(), "frac>1", (u>1).mean())
x = rng.uniform(1, 4, n); u = ulps(run(1, x), 1.0/np.sqrt(x)); print("rsqrt[1,4] max ulp", u.max(), "mean", u.mean())
x = rng.uniform(-700, 5, n); u = ulps(run(2, x), np.exp(x)); print("exp max ulp", u.max(), "mean", u.mean())
x = rng.uniform(0.7, 1.3, n); print("log abs err", np.max(np.abs(run(3, x) - np.log(x))))
th = rng.uniform(0, np.pi/2, n); s, c = np.sin(th), np.cos(th); nr = np.hypot(s,c); s/=nr; c/=nr
print("atan2 abs err", np.max(np.abs(run(4, s, c) - np.arctan2(s, c))))
