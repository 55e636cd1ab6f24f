"""CPU tests of the PACKED sample-slab FORMAT (include/reentry_b200.h, rb_sample_slab): a synthetic arena laid out exactly as
the header documents it -- per launch and partition one slab [rows][8][width], column j = the j-th trajectory of the partition
(ascending index) that was still alive when the launch started, no index array stored -- is read back by the Python accessors
(model.PackedSamples) and by the library's host-side converter rb_unpack_samples (pure format conversion: no GPU involved).
The GPU tests (tests/test_packed.py) check that the kernels WRITE this format; these check that both readers agree on it.
"""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def synthetic(n, n_steps, stride, spl, parts, seed, pad=0):
    """Dense samples [n_out, 8, n] of a made-up ensemble plus the packed arena / slab list that holds the same samples."""
    from reentry_old_b200 import _lib as L

    rng = np.random.default_rng(seed)
    n_out = n_steps // stride + 1
    impact_step = rng.integers(1, n_steps + 1, n).astype(np.int32)
    impact_step[rng.random(n) < 0.2] = -1                                  # still alive at the end
    status = np.where(impact_step > 0, L.TRAJ_IMPACTED, L.TRAJ_ALIVE).astype(np.uint8)
    n_samples = np.where(impact_step > 0, (impact_step - 1) // stride + 1, n_out).astype(np.int32)
    dense = rng.normal(size=(n_out, 8, n))
    dense = np.where((np.arange(n_out)[:, None, None] >= n_samples[None, None, :]), np.nan, dense)
    bounds = np.linspace(0, n, parts + 1).astype(int)
    slabs, chunks, offset = [], [], 0

    def add(first_row, n_rows, a, b, first_step, cols, width):
        nonlocal offset
        block = np.full((n_rows, 8, width), -123.0)                         # columns beyond the live count: never read
        block[:, :, : len(cols)] = dense[first_row: first_row + n_rows][:, :, cols]
        slabs.append(dict(offset=offset, first_row=first_row, n_rows=n_rows, first_traj=a, n_traj=b - a, width=width,
                          first_step=first_step, n_live=len(cols)))
        chunks.append(block.ravel())
        offset += block.size

    for a, b in zip(bounds[:-1], bounds[1:]):                              # sample 0: every trajectory of the partition
        add(0, 1, a, b, 0, np.arange(a, b), b - a)
    for first in range(0, n_steps, spl):                                   # one slab per launch and partition
        rows = [k for k in range(1, n_out) if first < k * stride <= first + spl]
        if not rows:
            continue
        for a, b in zip(bounds[:-1], bounds[1:]):
            live = np.nonzero((status[a:b] == L.TRAJ_ALIVE) | (impact_step[a:b] > first))[0] + a
            width = ((len(live) + 15) & ~15) + pad
            add(rows[0], len(rows), a, b, first, live, width)
    return dense, np.concatenate(chunks), slabs, impact_step, status, n_samples, n_out


def same(a, b):
    return np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))


@pytest.mark.parametrize("n,parts,spl,stride,pad", [(1000, 1, 32, 16, 0), (257, 2, 32, 5, 16), (33, 2, 50, 7, 0), (1, 1, 32, 16, 0)])
def test_python_accessors_read_the_documented_layout(n, parts, spl, stride, pad):
    from reentry_old_b200.model import PackedSamples

    dense, arena, slabs, impact_step, status, n_samples, n_out = synthetic(n, 256, stride, spl, parts, seed=n, pad=pad)
    pk = PackedSamples(torch.from_numpy(arena), [dict(s) for s in slabs], torch.from_numpy(impact_step), torch.from_numpy(status),
                       torch.from_numpy(n_samples), n, n_out)
    assert pk.used == arena.size
    assert same(pk.to_dense().numpy(), dense)
    for k in (0, 1, n_out // 2, n_out - 1):
        ids, vals = pk.row(k)
        assert same(vals.numpy(), dense[k][:, ids.numpy()])
        alive_then = np.nonzero(n_samples > k)[0]
        assert np.isin(alive_then, ids.numpy()).all()                      # every trajectory with a valid sample k is in the row
    for i in {0, n // 3, n - 1}:
        tr = pk.trajectory(i).numpy()
        assert tr.shape == (n_samples[i], 8) and same(tr, dense[: n_samples[i], :, i])
    with pytest.raises(IndexError):
        pk.row(n_out)


@pytest.mark.parametrize("n,parts,spl,stride,pad", [(1000, 1, 32, 16, 0), (257, 2, 32, 5, 16), (33, 2, 50, 7, 0)])
def test_library_unpack_agrees_with_the_python_reader(n, parts, spl, stride, pad):
    from reentry_old_b200 import _lib as L

    lib = L.load()
    dense, arena, slabs, impact_step, status, n_samples, n_out = synthetic(n, 256, stride, spl, parts, seed=7 * n, pad=pad)
    arr = (L.RbSampleSlab * len(slabs))(*[L.RbSampleSlab(**s) for s in slabs])
    out = np.full((n_out, 8, n), 5.0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.rb_unpack_samples(P(arena), arr, len(slabs), n, P(impact_step), P(status), n_out, P(out), 2)
    assert rc == 0 and same(out, dense)
    # a slab list that does not fit the summaries is refused, not scattered
    arr[len(slabs) - 1].n_live += 1
    assert lib.rb_unpack_samples(P(arena), arr, len(slabs), n, P(impact_step), P(status), n_out, P(out), 1) == -1   # RB_ERR_INVALID_ARG
    arr[len(slabs) - 1].n_live -= 1
    arr[0].first_row = n_out
    assert lib.rb_unpack_samples(P(arena), arr, len(slabs), n, P(impact_step), P(status), n_out, P(out), 1) == -1   # RB_ERR_INVALID_ARG
    assert lib.rb_unpack_samples(None, arr, len(slabs), n, P(impact_step), P(status), n_out, P(out), 1) == -1   # RB_ERR_INVALID_ARG
