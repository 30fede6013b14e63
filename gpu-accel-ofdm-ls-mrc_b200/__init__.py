"""B200-native uplink OFDM receiver hot path: CP strip -> FFT -> LS -> MRC -> hard demap.

The product is libofdm_lsmrc.so (hand-written sm_100a kernels behind the C ABI of
include/ofdm_lsmrc.h) plus the C++ facade in host/ that keeps the reference's names.
This Python package is a thin ctypes binding of the same ABI for tests and benchmarks.
"""
from .configs import CONFIGS, RxConfig  # noqa: F401
from .binding import ABI, LsMrcReceiver, LsmrcError, load_library  # noqa: F401
from . import build, sharding, synth  # noqa: F401
