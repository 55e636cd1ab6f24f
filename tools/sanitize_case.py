"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel of the library once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reentry_old_b200 import SpacecraftModel

m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0, gmst0=0.4)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
y0 = m.get_dispersed_initial_states(n, 7, [7800.0, 20.0, 0.0, 120e3, 90.0, -1.5], [10.0, 0.5, 0.5, 500.0, 0.5, 0.1], gmst=0.4)
r = m.run_ensemble(y0, dt=0.5, n_steps=1500, out_stride=16, steps_per_launch=40)      # 5 tiles per launch, partial last tile
gt = m.ground_track(r)
ip = m.impact_points(r)
d = m.equations_of_motion_batch(np.zeros(64), y0[:64])
sol = m.run_simulation((0, 300), y0[0].cpu().numpy(), np.arange(0, 300, 1.0))
torch.cuda.synchronize()
print("ok", int((r.status == 1).sum()), float(gt["downrange"][5, 0]), float(ip[0, 0]), sol.y.shape)
