"""tlab_b200: B200-native (sm_100a) hot path of turbulencia/tlab's incompressible/Boussinesq DNS.

The product is libtlab_gpu.so (hand-written CUDA behind a C ABI, include/tlab_gpu.h); this package is
the thin host-side mirror of the reference's Fortran operator interface used by tests and benchmarks.
"""
from . import lib  # noqa: F401
