"""TEST / MEASUREMENT INFRASTRUCTURE -- build oracle/_ref from the reference sources WHERE THEY LIE.

The reference's hot path is Python + numba: there is no C to compile.  What CAN be built from it is what numba
itself makes of it: this recipe imports the unmodified reference from /root/reference (oracle/ref_shim.py: stubs for
astropy / poliastro / matplotlib, which the path never touches), wraps the O3 driver (a fixed-step Dormand-Prince loop
over the reference's own @jit leaves -- euclidean_norm, eci_to_ecef, ecef_to_eci, ecef_to_geodetic,
J2_perturbation_numba, moon / sun_position_vector, third_body_acceleration, simplified_nrlmsise_00,
atmospheric_drag) in numba's ahead-of-time compiler (numba.pycc) and writes ONE extension module,

    oracle/_ref/ref_o3_native.<abi>.so      (git-ignored; travels to the GPU box like any built .so)

i.e. the machine code numba generates for the reference's functions -- no reference source is copied anywhere.
oracle/ref_native.py loads it (bench.py's cpu_baseline leg; tests/test_ref_native.py pins it bit for bit on the
golden vectors the JIT-compiled reference produced).

    python oracle/build_ref.py          # only where /root/reference exists (the build container)
"""
from __future__ import annotations

import glob
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODULE = "ref_o3_native"


def built_path():
    c = sorted(glob.glob(os.path.join(OUT, MODULE + "*.so")))
    return c[0] if c else None


def build(force: bool = False):
    """Returns the path of the extension module, or None where the reference is not available."""
    sys.path.insert(0, HERE)
    import ref_shim as R

    if not R.available():
        return built_path()
    srcs = [os.path.join(R.REFERENCE_DIR, f) for f in ("spacecraft_model.py", "coordinate_converter.py", "constants.py")] + \
        [os.path.join(HERE, "ref_shim.py"), os.path.abspath(__file__)]
    have = built_path()
    if have and not force and all(os.path.getmtime(s) <= os.path.getmtime(have) for s in srcs if os.path.exists(s)):
        return have
    warnings.filterwarnings("ignore")
    import numpy as np  # noqa: F401
    from numba.pycc import CC

    o3 = R.o3_build()
    propagate_one = o3["propagate_one"]
    os.makedirs(OUT, exist_ok=True)
    cc = CC(MODULE)
    cc.output_dir = OUT
    cc.verbose = False
    # AVX2 / FMA class x86-64: every host a B200 sits in has it; the code is scalar FP64 + libm, so the JIT's
    # -march=native would not produce anything different in kind
    cc.target_cpu = "x86-64-v3"

    @cc.export("propagate_batch", "void(f8[:,:], f8, f8, i8, i8, f8, f8, f8[:], f8[:], f8[:], f8[:,:,:], i8[:], i8[:], f8[:,:,:])")
    def propagate_batch(y0, t0, dt, n_steps, stride, epoch0, gmst0, Cd, A, m, ys, impact_step, n_done, K_event):
        # serial over trajectories: the callers parallelise over processes (an AOT module has no prange)
        for i in range(y0.shape[0]):
            a, b = propagate_one(y0[i], t0, dt, n_steps, stride, epoch0, gmst0, Cd[i], A[i], m[i], ys[i], K_event[i])
            impact_step[i] = a
            n_done[i] = b

    cc.compile()
    return built_path()


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p or "reference not available: nothing built")
