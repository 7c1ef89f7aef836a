"""B200-native time-step engine behind kspaceFirstOrder-CUDA's solver (hot path only).

The compute path is the CUDA library ``libkwave_b200.so`` (C ABI in ``include/kwave_b200.h``); this package is a thin
ctypes binding used by the tests and the benchmark, plus the synthetic input generator.  There is no CPU fallback:
every compute entry point raises when the library or a CUDA device is missing.
"""
from .capi import (  # noqa: F401
    ARRAY_IDS,
    STREAM_IDS,
    KwError,
    Simulation,
    c40_decode,
    c40_encode,
    fft_c2r_3d,
    intensity_avg_block,
    fft_r2c_3d,
    fft_zmid,
    length_supported,
    library_path,
    load_library,
    nccl_unique_id,
)
from . import slab, synth  # noqa: F401
