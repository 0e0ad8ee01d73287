"""square_wheel_b200 — B200-native (sm_100a) implementation of Square-Wheel's per-frame world-drawing hot path:
surface building, span rasterisation and tonemapping behind the C-ABI declared in include/sqw_cuda.h.

The CUDA library is the only compute path; there is no CPU fallback (see _native.load_cuda_library)."""
from . import _native  # noqa: F401

__all__ = ["_native", "device"]
