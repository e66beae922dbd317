"""starch3_b200 -- B200-native (sm_100a) implementation of starch3's compression hot path.

The product is the C-ABI shared library ``libstarch3_b200.so`` (include/starch3_b200.h);
this package is a thin ctypes binding over it plus the synthetic-input generators.
There is no CPU fallback: if the CUDA library is missing, or no GPU is present when a
context is created, the call fails loudly.
"""
from .api import (Starch3Error, Context, Result, lib, lib_path, have_library, build_library, multi_compress_bed, batch_chunks, chain_layout, C_ABI_SYMBOLS)  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["Starch3Error", "Context", "Result", "lib", "lib_path", "have_library", "build_library", "synth",
           "multi_compress_bed", "C_ABI_SYMBOLS"]
