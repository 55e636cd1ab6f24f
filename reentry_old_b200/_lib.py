"""ctypes binding of the C ABI (include/reentry_b200.h).  Fails loudly: there is no CPU fallback.

This is exactly the stub a maintainer of the reference would add: plain
pointers and sizes, no torch types in the signatures.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RB_LIB_PATH") or os.path.join(PKG, "libreentry_b200.so")  # RB_LIB_PATH: kernel-variant sweeps

RB_FLAG_COMPACT = 1
RB_FLAG_EARLY_EXIT = 2
RB_FLAG_NO_REFINE = 4
RB_FLAG_PROFILE = 8
RB_FLAG_SKIP_NEGLIGIBLE_DRAG = 16
RB_FLAG_NO_OVERLAP = 32
RB_FLAG_STAGED_SAMPLES = 64
RB_FLAG_DIRECT_SAMPLES = 128
RB_FLAG_RESUME = 256
RB_FLAG_PACKED_SAMPLES = 512
RB_ABI_VERSION = 3
RB_METHOD_RK45, RB_METHOD_RK23, RB_METHOD_DOP853 = 0, 1, 2
RB_TABLE_ENTRY_DOUBLES = 16
RB_SAMPLE_FIELDS = 8
RB_DIAG_FIELDS = 23
RB_THERMAL_FIELDS = 6
TRAJ_ALIVE, TRAJ_IMPACTED, TRAJ_NONFINITE, TRAJ_STEP_FAILED = 0, 1, 2, 3

vp = C.c_void_p
i64 = C.c_int64
i32 = C.c_int32
f64 = C.c_double


class RbIn(C.Structure):
    _fields_ = [("n", i64), ("y0", vp), ("y0_traj_stride", i64), ("y0_comp_stride", i64),
                ("cd", vp), ("area", vp), ("mass", vp),
                ("cd_scalar", f64), ("area_scalar", f64), ("mass_scalar", f64)]


class RbRun(C.Structure):
    _fields_ = [("first_step", i64), ("n_steps", i64), ("out_stride", i64), ("steps_per_launch", i64),
                ("event_altitude", f64), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class RbSampleSlab(C.Structure):
    _fields_ = [("offset", i64), ("first_row", i64), ("n_rows", i64), ("first_traj", i64), ("n_traj", i64),
                ("width", i64), ("first_step", i64), ("n_live", i64)]


class RbOut(C.Structure):
    _fields_ = [("samples", vp), ("sample_stride", i64), ("field_stride", i64), ("n_out_capacity", i64),
                ("final_state", vp), ("impact_step", vp), ("impact_time", vp), ("status", vp), ("n_samples", vp),
                ("packed_capacity", i64), ("slabs", C.POINTER(RbSampleSlab)), ("slab_capacity", i64),
                ("n_slabs", C.POINTER(i64)), ("slab_live", vp)]


class RbThermal(C.Structure):
    _fields_ = [("capsule_length", f64), ("dt", f64), ("iter_fact", f64), ("thermal_conductivity", f64),
                ("specific_heat_capacity", f64), ("emissivity", f64), ("ablation_efficiency", f64)]


class RbAdaptiveRun(C.Structure):
    _fields_ = [("t0", f64), ("t_bound", f64), ("rtol", f64), ("atol", f64), ("t_eval0", f64), ("t_eval_dt", f64),
                ("n_eval", i64), ("max_steps", i64), ("event_altitude", f64), ("epoch_jd", f64), ("gmst0", f64), ("ephemeris_dt", f64),
                ("method", i32), ("reserved", i32)]


class RbAdaptiveOut(C.Structure):
    _fields_ = [("samples", vp), ("sample_stride", i64), ("field_stride", i64), ("final_state", vp), ("impact_time", vp),
                ("status", vp), ("n_samples", vp), ("n_accepted", vp), ("n_rejected", vp), ("step_log", vp), ("step_capacity", i64)]


class RbStats(C.Structure):
    _fields_ = [("kernel_launches", i64), ("propagate_launches", i64), ("steps_enqueued", i64),
                ("last_polled_live", i64), ("propagate_ms", f64), ("sample_bytes_sent", i64)]


# every symbol include/reentry_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rb_abi_version": (i32, []),
    "rb_strerror": (C.c_char_p, [i32]),
    "rb_ctx_last_error": (C.c_char_p, [vp]),
    "rb_ctx_create": (i32, [C.POINTER(vp), i32]),
    "rb_ctx_destroy": (i32, [vp]),
    "rb_time_table_entries": (i64, [i64]),
    "rb_build_time_table": (i32, [f64, f64, f64, f64, i64, vp, i32]),
    "rb_ctx_set_time_table": (i32, [vp, vp, i64, f64, f64, vp]),
    "rb_initial_states": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, f64, vp, i64, i64, vp]),
    "rb_dispersed_initial_states": (i32, [vp, i64, i64, C.c_uint64, vp, vp, f64, vp, vp, i64, i64, vp]),
    "rb_propagate": (i32, [vp, C.POINTER(RbIn), C.POINTER(RbRun), C.POINTER(RbOut), vp]),
    "rb_ctx_stats": (i32, [vp, C.POINTER(RbStats)]),
    "rb_equations_of_motion": (i32, [vp, i64, vp, vp, f64, f64, vp, vp, vp, f64, f64, f64, C.POINTER(RbThermal), vp, vp]),
    "rb_propagate_host": (i32, [vp, i64, vp, vp, vp, vp, f64, f64, f64, f64, f64, f64, f64, C.POINTER(RbRun),
                                vp, i64, vp, vp, vp, vp, vp, C.POINTER(i64)]),
    "rb_propagate_host_packed": (i32, [vp, i64, vp, vp, vp, vp, f64, f64, f64, f64, f64, f64, f64, C.POINTER(RbRun),
                                       vp, i64, C.POINTER(RbSampleSlab), i64, C.POINTER(i64), C.POINTER(i64),
                                       vp, vp, vp, vp, vp]),
    "rb_unpack_samples": (i32, [vp, C.POINTER(RbSampleSlab), i64, i64, vp, vp, i64, vp, i32]),
    "rb_ctx_table_generation": (i64, [vp]),
    "rb_host_alloc": (vp, [i64]),
    "rb_host_free": (None, [vp]),
    "rb_propagate_adaptive": (i32, [vp, C.POINTER(RbIn), C.POINTER(RbAdaptiveRun), C.POINTER(RbAdaptiveOut), vp]),
    "rb_ground_track": (i32, [vp, i64, i64, vp, i64, i64, vp, f64, f64, f64, f64, vp, vp, vp, vp, vp, vp]),
    "rb_impact_points": (i32, [vp, i64, vp, vp, vp, f64, f64, vp, vp]),
    "rb_math_probe": (i32, [vp, i32, i64, vp, vp, vp, vp]),
    "rb_fp64_peak_probe": (i32, [vp, f64, C.POINTER(f64), vp]),
}

_lib = None


class ReentryB200Error(RuntimeError):
    pass


def load():
    """Load libreentry_b200.so (built in-tree by reentry_old_b200.build).  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ReentryB200Error(
            f"{LIB_PATH} is missing: build it with `python -m reentry_old_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.rb_abi_version() != RB_ABI_VERSION:
        raise ReentryB200Error(f"{LIB_PATH} has ABI version {lib.rb_abi_version()}, this binding expects {RB_ABI_VERSION}: rebuild it")
    _lib = lib
    return lib


def check(rc: int, ctx=None, what: str = ""):
    if rc != 0:
        lib = load()
        msg = lib.rb_strerror(rc).decode()
        if ctx is not None and rc == -2:
            msg += ": " + lib.rb_ctx_last_error(ctx).decode()
        raise ReentryB200Error(f"{what or 'reentry_b200'} failed ({rc}): {msg}")
