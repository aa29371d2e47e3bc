"""cup3d_b200 -- B200-native (sm_100a) implementation of CUP3D's data-parallel
hot path behind a C ABI (include/cup3d_b200.h).  The Python here is a thin
ctypes binding plus synthetic-mesh helpers for tests and bench.py."""
from . import capi, mesh  # noqa: F401
from .capi import Context, CupError  # noqa: F401
