"""C5 (BASELINE.json configs[4]): reentry dispersion sharded over the GPUs of one box, initial states
generated on the device (counter-based, independent of the GPU count), summary output only, ONE NCCL
gather of the impact points.

    torchrun --nproc-per-node G tools/run_c5.py [n_total]          (G = 1 works without torchrun)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from reentry_old_b200 import SpacecraftModel
from reentry_old_b200.configs import shard_range
from reentry_old_b200.sharding import gather_packed

NOMINAL = [7800.0, 20.0, 0.0, 120e3, 90.0, -1.0]
SIGMA = [10.0, 0.5, 0.5, 500.0, 0.5, 0.1]


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lo, hi = shard_range(n_total, rank, world)
    m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0, device=local)

    def one_pass():
        y0 = m.get_dispersed_initial_states(hi - lo, 1234, NOMINAL, SIGMA, first_index=lo)
        res = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=2048, store_samples=False)
        pts = m.impact_points(res)                                       # [3, n] lat, lon, h
        packed = torch.cat([pts.t(), res.impact_time[:, None], res.impact_step.to(torch.float64)[:, None]], dim=1).contiguous()
        allp = gather_packed(packed, dst=0) if world > 1 else packed
        return res, allp

    one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res, allp = one_pass()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}", dtype=torch.float64)
    steps = res.impact_step.clamp(min=0).sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps, op=dist.ReduceOp.SUM)
    if rank == 0:
        hit = ~torch.isnan(allp[:, 3])
        lat, lon = allp[hit, 0], allp[hit, 1]
        print(json.dumps({"config": f"C5 reentry dispersion, {n_total} trajectories over {world} GPU(s)", "n_total": n_total,
                          "n_gpus": world, "seconds": ms.item() * 1e-3, "traj_steps": steps.item(),
                          "traj_steps_per_s": steps.item() / (ms.item() * 1e-3), "gathered_rows": int(allp.shape[0]),
                          "gather_bytes": int(allp.numel() * 8),
                          "impact_lat_mean_std": [lat.mean().item(), lat.std().item()],
                          "impact_lon_mean_std": [lon.mean().item(), lon.std().item()],
                          "not_impacted_within_cap": int(torch.isnan(allp[:, 3]).sum().item()),
                          "checksum": float(torch.nansum(allp[:, 3]).item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
