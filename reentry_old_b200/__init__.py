"""B200-native ensemble trajectory propagator for the Reentry-OLD hot path.

Public surface: `SpacecraftModel` (the reference's call surface, spacecraft_model.py:392-519,
plus `run_ensemble`), `EnsembleResult`, the workload definitions in `configs`, and the
multi-GPU helpers in `sharding`.  All numerics run in libreentry_b200.so (hand-written
sm_100a kernels behind the C ABI of include/reentry_b200.h); importing this package does
not need a GPU, calling it does.
"""
from ._lib import ReentryB200Error, load as load_library  # noqa: F401
from .model import DeviceContext, EnsembleResult, Epoch, SpacecraftModel  # noqa: F401
from .timeutil import epoch_and_gmst  # noqa: F401

__all__ = ["SpacecraftModel", "EnsembleResult", "Epoch", "DeviceContext", "ReentryB200Error", "load_library", "epoch_and_gmst"]
