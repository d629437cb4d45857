"""metalens_b200 -- B200-native near-field -> far-field engine behind the metalens API.

Host side in Python (mirrors the reference's call surface), compute in hand-written
sm_100a CUDA kernels reached through the ctypes C-ABI of ``libmetalens_b200.so``.
"""
from ._lib import MetalensB200Error, load as load_library  # noqa: F401

__all__ = ["MetalensB200Error", "load_library"]
__version__ = "0.1.0"
