"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, exports
exactly the symbols include/starch3_b200.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

import starch3_b200 as s3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    """every function include/*.h declares: starch3_b200.h (S3G_API) and the libbz2-shaped front s3g_bzlib.h (S3G_BZ_API)"""
    txt = open(os.path.join(ROOT, "include", "starch3_b200.h")).read()
    syms = set(re.findall(r"S3G_API[^;(]*?\b(s3g_\w+)\s*\(", txt))
    txt = open(os.path.join(ROOT, "include", "s3g_bzlib.h")).read()
    syms |= set(re.findall(r"S3G_BZ_API[^;(]*?\b(s3g_\w+)\s*\(", txt))
    return sorted(syms)


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(s3.C_ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not s3.have_library():
        s3.build_library()
    L = ctypes.CDLL(s3.lib_path)
    for sym in _header_symbols():
        assert hasattr(L, sym), sym


def test_struct_layouts_match_header():
    from starch3_b200.api import CChrom, CResult, CBlockDesc
    assert ctypes.sizeof(CChrom) == 72
    assert ctypes.sizeof(CBlockDesc) == 8 + 8 + 4 + 4 + 256
    assert ctypes.sizeof(CResult) == (12 + 2 + 8 + 3) * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(s3.Starch3Error) as e:
        s3.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "starch3_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "api.py" and "oracle" not in txt, (dirpath, f)


def test_synth_is_deterministic():
    a = s3.synth.bed(2, 5000, seed=42)
    b = s3.synth.bed(2, 5000, seed=42)
    assert a.tobytes() == b.tobytes()
    assert a.tobytes().count(b"\n") == 5000
    assert s3.synth.bed(2, 5000, seed=43).tobytes() != a.tobytes()
