"""GPU tests of the PACKED sample slabs (RB_FLAG_PACKED_SAMPLES, include/reentry_b200.h rb_sample_slab): the same
samples as the dense [n_out][8][N] rows, bit for bit, in rows only as wide as the live ensemble -- on the device
path (rb_propagate) and on the host-buffer path (rb_propagate_host_packed: one contiguous copy-engine transfer per
slab, time table built chunk-wise while the GPU runs), plus the host-side unpack helper.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t

    if not t.cuda.is_available():
        pytest.skip("no GPU")
    return t


@pytest.fixture(scope="module")
def Model(torch):
    from reentry_old_b200 import SpacecraftModel

    return SpacecraftModel


def _same(torch, a, b):
    return torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))


def _ensemble(Model, n, seed, steeper=0.5):
    from reentry_old_b200.configs import reentry_dispersion_params

    p = reentry_dispersion_params(n, seed=seed)
    p[::3, 5] -= steeper          # impacts spread over many launches
    m = Model(Cd=1.3, A=14.0, m=5000.0)
    return m, m.get_initial_states(*p.T)


@pytest.mark.parametrize("kw", [dict(), dict(steps_per_launch=16), dict(steps_per_launch=50), dict(compact=False),
                                dict(early_exit=False), dict(steps_per_launch=100, compact=False, early_exit=False)])
def test_packed_equals_dense_on_device(Model, torch, kw):
    n, n_steps, stride = 3000, 1800, 7
    m, y0 = _ensemble(Model, n, 77)
    dense = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride, **kw)
    pk = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride, samples="packed", **kw)
    torch.cuda.synchronize()
    assert pk.samples is None and pk.packed is not None
    for k in ("final_state", "impact_step", "status", "n_samples"):
        assert torch.equal(getattr(pk, k), getattr(dense, k)), k
    assert _same(torch, pk.impact_time, dense.impact_time)
    assert _same(torch, pk.packed.to_dense(), dense.samples), kw
    # slabs never exceed the arena, rows are contiguous, widths shrink with the live ensemble when compaction + polling run
    slabs = pk.packed.slabs
    assert slabs[0]["first_row"] == 0 and slabs[0]["n_rows"] == 1 and slabs[0]["n_traj"] == n
    live = pk.packed.slab_live.cpu().numpy()
    for i, s in enumerate(slabs[1:], start=1):
        cols = pk.packed.columns(s)
        assert cols.numel() == live[i] <= s["width"] and s["width"] % 16 == 0
        assert bool((cols[1:] > cols[:-1]).all())
    if not kw or "steps_per_launch" in kw and len(kw) == 1:
        assert slabs[-1]["width"] < slabs[1]["width"]                      # rows narrowed
        assert pk.packed.used < 0.9 * dense.samples.numel()                  # and the arena holds fewer doubles than the dense rows


def test_packed_two_partitions_and_accessors(Model, torch):
    n, n_steps, stride = 640_001, 2048, 128
    m, y0 = _ensemble(Model, n, 21, steeper=0.0)
    dense = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride)
    pk = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride, samples="packed")
    torch.cuda.synchronize()
    assert any(s["first_traj"] > 0 for s in pk.packed.slabs)               # two partitions -> two slabs per launch
    assert _same(torch, pk.packed.to_dense(), dense.samples)
    ids, vals = pk.packed.row(5)
    assert torch.equal(torch.nan_to_num(vals, nan=-1.0), torch.nan_to_num(dense.samples[5][:, ids], nan=-1.0))
    for i in (0, 17, n // 2, n - 1):
        tr = pk.packed.trajectory(i)
        ns = int(dense.n_samples[i])
        assert tr.shape[0] == ns and torch.equal(tr, dense.samples[:ns, :, i])


def test_packed_capacity_error(Model, torch):
    from reentry_old_b200._lib import ReentryB200Error

    m, y0 = _ensemble(Model, 2000, 5)
    with pytest.raises(ReentryB200Error, match="packed sample arena"):
        m.run_ensemble(y0, dt=0.5, n_steps=1800, out_stride=7, samples="packed", packed_capacity=8 * 2016 * 3)
    r = m.run_ensemble(y0, dt=0.5, n_steps=1800, out_stride=7)             # the context is usable afterwards
    assert int((r.status == 1).sum()) == 2000


@pytest.mark.parametrize("n", [5000, 700_000])
def test_host_packed_matches_device_path(Model, torch, n):
    """rb_propagate_host_packed (pinned host buffers in and out) == the device path, bit for bit; the time table it
    builds chunk-wise on the host only covers what the ensemble flies; rb_unpack_samples restores the dense rows."""
    from reentry_old_b200.model import DeviceContext

    n_steps, stride = 6000, 16                                           # BASELINE's cap: far beyond the last impact
    m, y0 = _ensemble(Model, n, 11, steeper=0.0)
    dense = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=stride, store_samples=(n <= 100_000))
    bufs = m.host_buffers(n, n_steps, stride, packed_capacity=int(1.3 * 8 * n * 110))
    bufs["y0"].copy_(y0)
    for rep in range(2):                                                   # second call: buffers and context reused
        bufs["packed"].fill_(-7.0)
        r = m.run_ensemble_host(bufs, dt=0.5)
        assert torch.equal(r.impact_step, dense.impact_step.cpu()) and torch.equal(r.status, dense.status.cpu())
        assert torch.equal(r.n_samples, dense.n_samples.cpu())
        assert torch.equal(r.impact_time, dense.impact_time.cpu())
        assert torch.equal(r.final_state, dense.final_state.cpu())
        assert r.stats["steps_enqueued"] < 2048                            # early exit: the 6000-step cap costs nothing
        live = [s["n_live"] for s in r.packed.slabs]
        assert live[0] == n and all(a >= 0 for a in live) and live[-1] < n
        if dense.samples is not None:
            got = r.packed.to_dense(n_out=2048 // stride + 1)
            assert _same(torch, got, dense.samples.cpu())
            # the C helper does the same scatter on the host
            ctx = DeviceContext.get()
            out = np.empty((2048 // stride + 1, 8, n))
            slabs = bufs["slabs"]
            rc = ctx.lib.rb_unpack_samples(bufs["packed"].data_ptr(), slabs, len(r.packed.slabs), n,
                                           bufs["impact_step"].data_ptr(), bufs["status"].data_ptr(), out.shape[0],
                                           out.ctypes.data, 0)
            assert rc == 0
            assert np.array_equal(np.nan_to_num(out, nan=-1.0), np.nan_to_num(dense.samples.cpu().numpy(), nan=-1.0))
    # after the host call restaged the context's table, the device path must notice (no stale-table reuse)
    again = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=stride, store_samples=False)
    assert torch.equal(again.impact_step, dense.impact_step)


def test_host_packed_without_samples_and_errors(Model, torch):
    from reentry_old_b200 import _lib as L
    from reentry_old_b200.model import DeviceContext

    n = 3000
    m, y0 = _ensemble(Model, n, 3)
    dense = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=2048, store_samples=False)
    bufs = m.host_buffers(n, 2048, 2048, packed_capacity=0)
    bufs["y0"].copy_(y0)
    r = m.run_ensemble_host(bufs, dt=0.5)
    assert r.packed is None and torch.equal(r.impact_step, dense.impact_step.cpu())
    ctx = DeviceContext.get()
    run = L.RbRun(first_step=0, n_steps=100, out_stride=0, steps_per_launch=0, event_altitude=1000.0, flags=0)
    args = [ctx.handle, n, bufs["y0"].data_ptr(), None, None, None, 1.3, 14.0, 5000.0, m.epoch, 0.0, 0.0, 0.5, C.byref(run),
            None, 0, bufs["slabs"], len(bufs["slabs"]), None, None, bufs["final_state"].data_ptr(),
            bufs["impact_step"].data_ptr(), bufs["impact_time"].data_ptr(), bufs["status"].data_ptr(), bufs["n_samples"].data_ptr()]
    assert ctx.lib.rb_propagate_host_packed(*args) == -1   # out_stride 0: invalid argument, not SIGFPE
    # the dense host call validates out_stride too (it used to divide first)
    used = C.c_int64(0)
    rc = ctx.lib.rb_propagate_host(ctx.handle, n, bufs["y0"].data_ptr(), None, None, None, 1.3, 14.0, 5000.0, m.epoch, 0.0, 0.0,
                                   0.5, C.byref(run), None, 0, bufs["final_state"].data_ptr(), bufs["impact_step"].data_ptr(),
                                   bufs["impact_time"].data_ptr(), bufs["status"].data_ptr(), bufs["n_samples"].data_ptr(), C.byref(used))
    assert rc == -1
    # too small an arena: clean error, context usable afterwards
    small = m.host_buffers(n, 2048, 16, packed_capacity=8 * 3008 * 4)
    small["y0"].copy_(y0)
    with pytest.raises(L.ReentryB200Error, match="packed sample arena"):
        m.run_ensemble_host(small, dt=0.5)
    r = m.run_ensemble_host(bufs, dt=0.5)
    assert torch.equal(r.impact_step, dense.impact_step.cpu())


def test_blocks_that_start_fully_retired(Model, torch):
    """Without compaction the live list keeps naming retired lanes: blocks whose lanes are all dead race through
    their tiles and leave through the warp vote with a bulk copy still in flight unless the exit waits for it
    (ADVICE r1).  Small steps_per_launch, clustered impacts, many launches: results must equal the compacted run."""
    n, n_steps = 40_000, 2048
    m, y0 = _ensemble(Model, n, 9, steeper=0.0)
    base = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=64)
    for spl in (8, 24):
        r = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=64, compact=False, early_exit=False, steps_per_launch=spl)
        torch.cuda.synchronize()
        assert torch.equal(r.impact_step, base.impact_step) and torch.equal(r.final_state, base.final_state)
        assert _same(torch, r.samples, base.samples)


@pytest.mark.parametrize("n", [1, 15, 17, 129, 1025])
def test_packed_ragged_sizes(Model, torch, n):
    """Ensemble sizes that are not a multiple of the slab-width granule (16), of a warp or of a block -- one trajectory
    included: packed == dense on the device path and on the host-buffer path."""
    n_steps, stride = 2048, 5
    m, y0 = _ensemble(Model, n, 5)
    dense = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride)
    pk = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride, samples="packed")
    assert _same(torch, pk.packed.to_dense(), dense.samples)
    assert torch.equal(pk.impact_step, dense.impact_step) and torch.equal(pk.n_samples, dense.n_samples)
    for i in {0, n // 2, n - 1}:
        tr = pk.packed.trajectory(i)
        ns = int(dense.n_samples[i])
        assert tr.shape == (ns, 8) and torch.equal(tr, dense.samples[:ns, :, i])
    bufs = m.host_buffers(n, n_steps, stride, packed_capacity=8 * (n + 16) * (n_steps // stride + 2))
    bufs["y0"].copy_(y0)
    r = m.run_ensemble_host(bufs, dt=0.5)
    assert torch.equal(r.impact_step, dense.impact_step.cpu()) and torch.equal(r.final_state, dense.final_state.cpu())
    assert _same(torch, r.packed.to_dense(n_out=n_steps // stride + 1), dense.samples.cpu())
