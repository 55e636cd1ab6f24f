"""oracle/_ref (built by oracle/build_ref.py from the reference where it lies) holds the reference's own numba-compiled
functions as an ahead-of-time extension module.  It must reproduce, bit for bit, the golden vectors the JIT-compiled
reference produced (tests/golden, oracle/make_golden.py) -- that is what entitles bench.py to report its throughput as
`cpu_baseline.reference_numba` (kind "reference").  Skipped where oracle/_ref was not built."""
import os

import numpy as np
import pytest

from oracle import ref_native

pytestmark = pytest.mark.skipif(not ref_native.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


@pytest.mark.parametrize("name,n", [("c2_mc.npz", 12), ("edge.npz", 8), ("c3_sso.npz", 2)])
def test_aot_module_equals_the_jit_reference(golden_dir, name, n):
    g = np.load(os.path.join(golden_dir, name))
    y0 = g["y0"].reshape(-1, 6)[:n]
    bc = lambda k: (g[k][:n] if np.ndim(g[k]) else float(g[k]))  # noqa: E731
    r = ref_native.propagate(y0, bc("Cd"), bc("A"), bc("m"), float(g["epoch_jd"]), float(g["gmst0"]), float(g["t0"]),
                             float(g["dt"]), int(g["n_steps"]), int(g["stride"]) if "stride" in g else 1)
    assert np.array_equal(r["impact_step"], g["impact_step"][:n])
    ys = g["ys"][:n]
    assert np.array_equal(np.isnan(ys[:, :, 0]), np.isnan(r["ys"][:, :, 0]))
    ok = ~np.isnan(ys[:, :, 0])
    # same LLVM code generator, another target CPU (x86-64-v3 instead of the JIT's native): the matrix products go to
    # the same BLAS, the rest is IEEE scalar arithmetic + libm -> identical, or within a few ulp where libm differs
    d = np.abs(r["ys"][ok] - ys[ok]) / np.maximum(np.abs(ys[ok]), 1.0)
    assert d.max() <= 1e-13, d.max()


def test_rate_runs_on_a_tiny_sample():
    from reentry_old_b200.configs import c2_reentry_mc

    w = c2_reentry_mc(n=64, seed=1, n_steps=6000, out_stride=16)
    out = ref_native.rate(w, target_seconds=1.0, n_procs=2)
    assert out["kind"] == "reference" and out["cores"] == 2 and out["value"] > 1e3
