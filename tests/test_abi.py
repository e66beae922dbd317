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


def test_batches_are_dealt_in_chunks_that_tile_them_and_fit_the_gpu():
    """host logic of mtf_huff.cu (plan_chunks): for every batch size the chunks are consecutive, cover the batch exactly, and a
    chunk's CTAs (blocks x CTAs per block) fit one wave -- 296 CTAs for the MTF kernels, 148 for the 1024-thread Huffman
    form -- unless it is made of whole waves of one CTA per block"""
    import starch3_b200 as s3
    sm = 148
    for stage, wave in ((3, 2 * sm), (4, sm)):
        for nb in list(range(1, 700)) + [893, 1024, 5000]:
            chunks = s3.batch_chunks(nb, stage)
            at = 0
            for first, count, ctas in chunks:
                assert first == at and count > 0
                at += count
                if stage == 4 and ctas == 0:
                    assert count > sm and len(chunks) == 1      # the 512-thread form takes a whole batch
                elif ctas == 1:
                    assert count % wave == 0 or count <= wave
                else:
                    assert ctas in (2, 4, 8) and ctas * count <= wave
            assert at == nb
            assert len(chunks) <= 6
    assert s3.batch_chunks(0, 3) == []
    assert s3.batch_chunks(296, 3) == [(0, 296, 1)] and s3.batch_chunks(178, 3) == [(0, 148, 2), (148, 30, 8)]
    assert s3.batch_chunks(893, 3) == [(0, 888, 1), (888, 5, 8)] and s3.batch_chunks(893, 4) == [(0, 893, 0)]
    assert s3.batch_chunks(178, 4) == [(0, 148, 1), (148, 30, 4)] and s3.batch_chunks(37, 4) == [(0, 37, 4)]
