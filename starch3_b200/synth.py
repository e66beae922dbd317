"""Synthetic sorted-BED inputs of the BASELINE.json shapes (csrc/synth.c)."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_path = os.path.join(HERE, "libs3synth.so")
_lib = None

CONFIGS = {
    1: "synthetic sorted BED3, single chromosome (chr1)",
    2: "synthetic hg38-shaped BED6 across 24 chromosomes with ids and scores",
    3: "dense DNase-footprint-like BED3, short uniform-length intervals",
    4: "sparse wide-interval BED6 with high-entropy names and scores",
    5: "whole-genome mix of the above across 24 chromosomes",
}


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_path):
            raise RuntimeError(f"{_path} missing: run `make -C starch3_b200/csrc`")
        _lib = C.CDLL(_path)
        _lib.s3synth_bed.restype = C.c_uint64
        _lib.s3synth_bed.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
    return _lib


def bed(cfg, n_lines, seed=42, variant=0, out=None):
    """Returns a numpy uint8 array holding n_lines lines of config `cfg` (1..5)."""
    cap = n_lines * 96 + 64
    buf = out if out is not None else np.empty(cap, dtype=np.uint8)
    assert buf.nbytes >= cap
    n = _load().s3synth_bed(cfg, variant, n_lines, seed, buf.ctypes.data_as(C.c_void_p), buf.nbytes)
    assert n > 0 or n_lines == 0
    return buf[:n]
