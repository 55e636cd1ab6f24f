"""Host-side multi-rank logic on CPU: shard arithmetic and the single summary gather,
exercised with world_size 2 and 3 over gloo (the GPU run uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reentry_old_b200.configs import shard_range
from reentry_old_b200.sharding import gather_packed, pack_summary, unpack_summary


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 1000, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def fake_shard(lo, hi):
    """Deterministic per-trajectory summary as a function of the GLOBAL index."""
    idx = torch.arange(lo, hi, dtype=torch.float64)
    fs = torch.stack([idx * (c + 1) for c in range(6)])            # [6, n] as the library returns it
    return pack_summary(fs, idx * 0.5, (idx % 1000).to(torch.int32), (idx % 3).to(torch.uint8),
                        (idx % 7).to(torch.int32))


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    out = gather_packed(fake_shard(lo, hi), dst=0)
    # the same gather with the shard sizes given by the caller (no count exchange, no host synchronisation)
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    out2 = gather_packed(fake_shard(lo, hi), dst=0, counts=counts)
    try:
        gather_packed(fake_shard(lo, hi), dst=0, counts=[c + 1 for c in counts])
        raise AssertionError("wrong counts accepted")
    except ValueError:
        pass
    if rank == 0:
        assert torch.equal(out, out2)
        q.put(out.numpy())
    else:
        assert out is None and out2 is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 1001), (3, 10)])
def test_gather_is_rank_count_independent(world, n_total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = fake_shard(0, n_total).numpy()            # what a single rank would have produced
    assert np.array_equal(got, ref)
    d = unpack_summary(torch.from_numpy(got))
    assert d["impact_step"].dtype == np.int32 and d["status"].dtype == np.uint8
    assert np.array_equal(d["final_state"][:, 1], 2.0 * np.arange(n_total))
