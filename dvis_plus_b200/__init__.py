"""dvis_plus_b200 -- B200-native (sm_100a) implementation of the DVIS++ per-frame dense hot path.

Host side: Python / PyTorch mirrors of the reference's op and module interfaces.
Device side: hand-written CUDA kernels behind the C ABI in include/dvis_b200.h (libdvis_b200.so).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
