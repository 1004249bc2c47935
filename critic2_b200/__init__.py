"""critic2_b200 -- B200-native on-grid QTAIM hot path for critic2 (BADER / YT / INTEGRABLE / NCIPLOT).

The product is the CUDA shared library critic2_b200/libcritic2_gpu.so behind the C ABI of
include/critic2_gpu.h; this package only holds its ctypes binding (capi) and the host-side mirror of
the reference's keyword surface (host).  There is no CPU fallback.
"""
from . import capi  # noqa: F401
