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


def test_chained_steps_lay_the_bits_out_like_one_step():
    """host logic of the chained entries (api.cu, chain_layout_step): streams of random block bit lengths and CRCs, laid out in one
    step and in random chains of steps -- a step ends inside a stream, its last one or two blocks are not final and come back in
    the next step -- give the same bytes (every block's bits, the "BZh" headers, the trailers with the folded CRC), the same
    stream offsets and lengths"""
    import numpy as np
    import starch3_b200 as s3
    rng = np.random.default_rng(99)

    def render(total_bytes, placed):
        bits = np.zeros(total_bytes * 8 + 64, dtype=np.uint8)
        for pos, payload in placed:
            assert not bits[pos:pos + len(payload)].any()          # nothing is written twice
            bits[pos:pos + len(payload)] = payload
        return np.packbits(bits[:total_bytes * 8]).tobytes()

    def word_bits(w):
        return np.array([(w >> (31 - k)) & 1 for k in range(32)], dtype=np.uint8)

    for case in range(40):
        level = int(rng.integers(1, 10))
        n_streams = int(rng.integers(1, 7))
        blocks = []                                               # (stream, n_bits, crc, payload)
        for s in range(n_streams):
            for _ in range(int(rng.integers(1, 9))):
                nb_ = int(rng.integers(120, 4000))
                blocks.append((s, nb_, int(rng.integers(0, 2 ** 32)), np.ones(nb_, dtype=np.uint8)))
        def run(steps):
            """steps: list of (first block, end block, final end) over the global block list; returns bytes, offsets, lengths"""
            state = np.zeros(4, dtype=np.uint64)
            placed, off, ln = [], {}, {}
            for (b0, b1, bf) in steps:
                streams = sorted(set(b[0] for b in blocks[b0:b1]))
                local = {g: i for i, g in enumerate(streams)}
                cont = bool(state[0]) and blocks[b0][0] == run.open_stream
                pos, patches, start, lens = s3.chain_layout(state, level, len(streams), cont, [local[b[0]] for b in blocks[b0:b1]],
                                                            [b[1] for b in blocks[b0:b1]], [b[2] for b in blocks[b0:b1]], bf - b0)
                for k in range(bf - b0):
                    placed.append((int(pos[k]), blocks[b0 + k][3]))
                for p_, w in patches:
                    placed.append((p_, word_bits(w)))
                for g, i in local.items():
                    if int(start[i]) != 2 ** 64 - 1:
                        off[g] = int(start[i])
                    if int(lens[i]):
                        ln[g] = int(lens[i])
                run.open_stream = blocks[b1 - 1][0] if state[0] else None
            assert not state[0]
            return render(int(state[3]), placed), off, ln
        run.open_stream = None
        one = run([(0, len(blocks), len(blocks))])
        # a chain: a step covers the blocks up to a random point; the last one or two blocks of its last stream are not final
        # (unless the step ends the input) and open the next step
        fixed, b0 = [], 0
        while b0 < len(blocks):
            b1 = min(len(blocks), b0 + int(rng.integers(2, 9)))
            if b1 == len(blocks):
                fixed.append((b0, b1, b1)); break
            last_stream = blocks[b1 - 1][0]
            first_of_last = next(k for k in range(b0, b1) if blocks[k][0] == last_stream)
            bf = max(first_of_last, b1 - int(rng.integers(1, 3)))
            if bf == b0:                                          # a step that finishes nothing is not a step
                continue
            fixed.append((b0, b1, bf))
            b0 = bf
        run.open_stream = None
        chained = run(fixed)
        assert chained == one, case


def test_host_only_entry_points_refuse_bad_arguments():
    import numpy as np
    import starch3_b200 as s3
    with pytest.raises(s3.Starch3Error) as e:
        s3.batch_chunks(10, 7)                                   # stages 3 (MTF) and 4 (Huffman) only
    assert e.value.code == -2
    state = np.zeros(4, dtype=np.uint64)
    with pytest.raises(s3.Starch3Error) as e:                    # a step that continues a stream when none is open
        s3.chain_layout(state, 9, 1, True, [0], [100], [1], 1)
    assert e.value.code == -2
    with pytest.raises(s3.Starch3Error) as e:                    # blocks that name a stream the step does not have
        s3.chain_layout(state, 9, 1, False, [0, 3], [100, 100], [1, 2], 2)
    assert e.value.code == -2
    pos, patches, start, ln = s3.chain_layout(state, 9, 1, False, [0], [100], [0xdeadbeef], 1)
    assert list(pos) == [32] and patches[0] == (0, 0x425a6839) and int(start[0]) == 0
    assert int(ln[0]) == (32 + 100 + 80 + 7) // 8 and list(state) == [0, 0, 0, int(ln[0])]
    # the trailer: 0x177245385090 then the combined CRC (one block: its CRC)
    assert patches[1:] == [(132, 0x17724538), (164, 0x5090dead), (196, 0xbeef0000)]
