"""Multi-GPU plumbing: trajectories are independent, so an ensemble shards
into contiguous index ranges with NO data-path collective; the only exchange is ONE gather of
the per-trajectory summaries at the end (NCCL on GPUs; the same code runs over gloo on CPU
tensors, which is how tests/test_sharding.py covers it with world_size 2).
"""
from __future__ import annotations

import numpy as np

from .configs import shard_range  # noqa: F401  (re-export)

SUMMARY_COLUMNS = ("x", "y", "z", "vx", "vy", "vz", "impact_time", "impact_step", "status", "n_samples")


def pack_summary(final_state, impact_time, impact_step, status, n_samples):
    """[N, 10] float64: final state (6), impact time, impact step, status, n_samples (ints are exact in f64)."""
    import torch

    n = impact_time.shape[0]
    out = torch.empty((n, len(SUMMARY_COLUMNS)), dtype=torch.float64, device=impact_time.device)
    out[:, 0:6] = final_state.t() if final_state.shape[0] == 6 and final_state.shape[-1] == n else final_state
    out[:, 6] = impact_time
    out[:, 7] = impact_step.to(torch.float64)
    out[:, 8] = status.to(torch.float64)
    out[:, 9] = n_samples.to(torch.float64)
    return out


def gather_packed(packed, dst=0, group=None, counts=None):
    """One collective: gather variable-length [n_r, 10] shards to rank `dst` in rank order.
    Returns the concatenated [sum n_r, 10] tensor on dst, None elsewhere.
    counts: the shard sizes of all ranks when the caller knows them (shard_range gives them): without it they are
    exchanged first, which costs a small all_gather and a host synchronisation per call."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if counts is None:
        counts = [torch.zeros(1, dtype=torch.int64, device=packed.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([packed.shape[0]], dtype=torch.int64, device=packed.device), group=group)
        counts = [int(c.item()) for c in counts]
    else:
        counts = [int(c) for c in counts]
        if len(counts) != world or counts[rank] != packed.shape[0]:
            raise ValueError(f"counts {counts} do not describe this gather (rank {rank} holds {packed.shape[0]} rows)")
    nmax = max(counts)
    buf = packed
    if packed.shape[0] != nmax:  # pad so that every rank contributes the same shape
        buf = torch.zeros((nmax, packed.shape[1]), dtype=packed.dtype, device=packed.device)
        buf[: packed.shape[0]] = packed
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf.contiguous(), parts, dst=dst, group=group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    dist.gather(buf.contiguous(), None, dst=dst, group=group)
    return None


def gather_summaries(result, dst=0, group=None, counts=None):
    """EnsembleResult of this rank's shard -> [N_total, 10] on rank dst (None elsewhere)."""
    return gather_packed(pack_summary(result.final_state, result.impact_time, result.impact_step, result.status,
                                      result.n_samples), dst=dst, group=group, counts=counts)


def unpack_summary(packed):
    """[N, 10] -> dict of numpy arrays with the original dtypes."""
    a = packed.cpu().numpy()
    return dict(final_state=a[:, 0:6].copy(), impact_time=a[:, 6].copy(), impact_step=a[:, 7].astype(np.int32),
                status=a[:, 8].astype(np.uint8), n_samples=a[:, 9].astype(np.int32))


def run_sharded(model, workload_fn, n_total, group=None, **run_kw):
    """Convenience: build this rank's shard with workload_fn(lo, hi) -> (y0 [n,6], per-trajectory kwargs),
    propagate it on this rank's GPU and gather the summaries on rank 0."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_total, rank, world)
    y0, kw = workload_fn(lo, hi)
    res = model.run_ensemble(y0, **kw, **run_kw)
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    return res, gather_summaries(res, dst=0, group=group, counts=counts)
