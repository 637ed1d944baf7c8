"""B200-native FM receive chain: Python face of libfmrx_b200.so.

The product is the CUDA library behind the C ABI in ``include/fmrx.h`` plus the
C++ host code under ``host/``; this package only binds it (ctypes) for the
tests and the benchmark, and fabricates synthetic IQ.
"""
from . import binding, parallel, synth  # noqa: F401
from .binding import (FMDemod, FmrxError, LRExtraction, PLL, Pipeline, impulseResponseBPF,  # noqa: F401
                      impulseResponseLPF, mixer, mode_table, pcm_pack, readBlockData, resample)
