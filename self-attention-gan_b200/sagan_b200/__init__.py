"""sagan_b200 -- B200-native SAGAN generator/discriminator hot path.

Python host (this package) -> ctypes -> C ABI (include/sagan_b200.h) -> hand-written sm_100a CUDA
(csrc/).  torch is used for device memory, streams, autograd bookkeeping and torch.distributed.
There is no CPU / PyTorch fallback: importing works anywhere, computing needs the built
libsagan_b200.so and a B200.
"""
from . import _lib  # noqa: F401
from ._lib import MATH_BF16_TC, MATH_FP32_STRICT, SaganError  # noqa: F401

__all__ = ["MATH_BF16_TC", "MATH_FP32_STRICT", "SaganError"]
