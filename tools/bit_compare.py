"""Bit-compare two builds of the library on a C2-shaped ensemble (GPU box):
    python tools/bit_compare.py variants/old.so [n_traj]
Runs the ensemble with the in-tree library and with the given one (separate processes) and compares
final states, impact steps and impact times bit for bit."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import reentry_dispersion_params
n = int(sys.argv[2])
m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0)
p = reentry_dispersion_params(n, 99)
p[:, 1] = np.linspace(-89.0, 89.0, n)          # every latitude band
y0 = m.get_initial_states(*p.T)
r = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=64)
np.savez(sys.argv[1], fs=r.final_state.cpu().numpy(), st=r.impact_step.cpu().numpy(), ti=r.impact_time.cpu().numpy(),
         sm=torch.nan_to_num(r.samples, nan=-1.0).cpu().numpy())
''' % ROOT


def run(lib, n, out):
    env = dict(os.environ)
    if lib:
        env["RB_LIB_PATH"] = os.path.abspath(lib)
    subprocess.run([sys.executable, "-c", WORKER, out, str(n)], check=True, env=env)


if __name__ == "__main__":
    import numpy as np
    other = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    with tempfile.TemporaryDirectory() as d:
        run(None, n, os.path.join(d, "a.npz"))
        run(other, n, os.path.join(d, "b.npz"))
        a, b = np.load(os.path.join(d, "a.npz")), np.load(os.path.join(d, "b.npz"))
        for k in a.files:
            same = np.array_equal(a[k], b[k], equal_nan=True)
            print(k, "bit-identical" if same else f"DIFFERENT ({(a[k] != b[k]).sum()} elements)")
