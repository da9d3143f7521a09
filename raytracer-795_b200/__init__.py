"""raytracer-795_b200: B200-native rendering core for the CENG795 ray tracer hot path.

The product is the C ABI in include/rt795.h implemented by lib/librt795.so (hand-written sm_100a CUDA kernels in
csrc/, host shell in host/).  This Python package is only the ctypes driver used by tests and bench.py plus the
synthetic scene generators.  The directory name carries a hyphen (it mirrors the reference repository's name), so it
is loaded through __graft_entry__.load_package() under the module name `raytracer_795_b200`.
"""
from . import binding, multigpu, scenes, tiling  # noqa: F401
from .binding import *  # noqa: F401,F403
