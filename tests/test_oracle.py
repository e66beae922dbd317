"""CPU tests: pin the oracle (oracle/s3_oracle.c) against every fixed point available:
bzip2's own golden vectors, the reference-compiled libbz2, CPython's bz2, the
transformed stream dumped by the reference starch3 binary, and the worked examples
recorded in SURVEY.md section 8(c)."""
import bz2

import numpy as np
import pytest

from conftest import golden
from starch3_b200 import synth


# ---- bzip2 golden vectors (bz/Makefile:56-69: sampleN.ref -N -> sampleN.bz2) ----
@pytest.mark.parametrize("idx,level", [(1, 1), (2, 2), (3, 3)])
def test_bzip2_golden_samples(oracle, idx, level):
    gold = golden(f"sample{idx}.bz2")
    data = bz2.decompress(gold)
    assert oracle.bz_compress(data, level) == gold
    if oracle.have_ref():
        assert oracle.ref_bz_compress(data, level) == gold


def _corpus():
    rng = np.random.default_rng(7)
    yield "empty", b""
    yield "one", b"a"
    yield "two_same", b"aa"
    yield "run4", b"aaaa"
    yield "run5", b"aaaaa"
    yield "run255", b"b" * 255
    yield "run256", b"b" * 256
    yield "run259", b"b" * 259 + b"c"
    yield "run_1000", b"z" * 1000
    yield "periodic", b"5\n" * 3000
    yield "periodic3", b"abc" * 5000
    yield "text", bytes(rng.integers(97, 101, 50000, dtype=np.uint8))
    yield "random", bytes(rng.integers(0, 256, 30000, dtype=np.uint8))
    yield "runs", b"".join(bytes([int(c)]) * int(n) for c, n in zip(rng.integers(48, 52, 3000), rng.integers(1, 12, 3000)))
    yield "allbytes", bytes(range(256)) * 40


@pytest.mark.parametrize("name,data", list(_corpus()), ids=[n for n, _ in _corpus()])
def test_oracle_vs_reference_libbz2(oracle, name, data):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for level in (1, 9):
        assert oracle.bz_compress(data, level) == oracle.ref_bz_compress(data, level), (name, level)


@pytest.mark.parametrize("name,data", list(_corpus()), ids=[n for n, _ in _corpus()])
def test_oracle_vs_cpython_bz2(oracle, name, data):
    # CPython feeds BZ_RUN then BZ_FINISH; identical to a single BZ_FINISH feed except when a
    # block fills exactly as the input ends -- none of these cases does.
    assert oracle.bz_compress(data, 9) == bz2.compress(data, 9)
    assert bz2.decompress(oracle.bz_compress(data, 1)) == data


def test_oracle_multiblock_level1(oracle):
    rng = np.random.default_rng(3)
    data = bytes(rng.integers(48, 58, 350000, dtype=np.uint8))
    mine = oracle.bz_compress(data, 1)
    assert mine == bz2.compress(data, 1)
    if oracle.have_ref():
        assert mine == oracle.ref_bz_compress(data, 1)
    blocks, rle = oracle.rle1_blocks(data, 1)
    assert len(blocks) == 4 and sum(b["nblock"] for b in blocks) == len(rle)
    assert blocks[0]["in_start"] == 0 and blocks[-1]["in_end"] == len(data)
    for a, b in zip(blocks, blocks[1:]):
        assert a["in_end"] == b["in_start"]
    for b in blocks:
        assert b["crc"] == oracle.crc32(data[b["in_start"]:b["in_end"]])


def test_oracle_exact_fill_single_trailing_byte(oracle):
    """A block that fills exactly when one input byte is left absorbs that byte
    (handle_compress tests the finish condition first, bz/bzlib.c:393-396)."""
    nmax = 100000 - 19
    body = bytes(97 + (i * 7 + i // 26) % 26 for i in range(nmax))     # no two equal neighbours: RLE1 leaves it unchanged
    assert all(x != y for x, y in zip(body, body[1:]))
    data = body + (b"\n" if body[-1:] != b"\n" else b"!")
    blocks, _ = oracle.rle1_blocks(data, 1)
    assert [b["nblock"] for b in blocks] == [nmax + 1]
    if oracle.have_ref():
        assert oracle.bz_compress(data, 1) == oracle.ref_bz_compress(data, 1)
    # two trailing bytes: the block closes and a second block holds the rest
    data2 = data + b"#"
    blocks2, _ = oracle.rle1_blocks(data2, 1)
    assert [b["nblock"] for b in blocks2] == [nmax, 2]
    if oracle.have_ref():
        assert oracle.bz_compress(data2, 1) == oracle.ref_bz_compress(data2, 1)


def test_oracle_bwt_vs_reference_blocksort(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    cases = [bytes(rng.integers(48, 58, 20000, dtype=np.uint8)), b"ab" * 6000, b"5\n" * 2500, b"x" * 3000,
             bytes(rng.integers(0, 256, 12000, dtype=np.uint8)), (b"abcabd" * 2000)[:11999]]
    for blk in cases:
        p1, o1 = oracle.bwt(blk)
        p2, o2 = oracle.ref_bwt(blk)
        assert o1 == o2
        assert np.array_equal(p1, p2)


def test_oracle_mtf_huff_vs_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    tf, _, _ = oracle.transform(synth.bed(2, 20000))
    blocks, rle = oracle.rle1_blocks(tf, 9)
    blk = rle[:blocks[0]["nblock"]]
    ptr, _ = oracle.ref_bwt(blk)
    ref = oracle.ref_mtf_huff(blk, ptr, blocks[0]["in_use"])
    mtfv, freq, nu = oracle.mtf(blk, ptr, blocks[0]["in_use"])
    assert np.array_equal(mtfv, ref["mtfv"])
    assert np.array_equal(freq[:nu + 2], ref["freq"][:nu + 2])
    h = oracle.huff_select(mtfv, freq, nu)
    assert np.array_equal(h["selector"], ref["selector"])
    assert np.array_equal(h["len"][:h["n_groups"], :nu + 2], ref["len"][:h["n_groups"], :nu + 2])


# ---- transform --------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg3const", "cfg4", "overlap", "zerolen"])
def test_transform_vs_reference_binary_golden(oracle, name):
    bed = golden(f"transform_{name}.bed")
    tf, chroms, dropped = oracle.transform(bed)
    assert tf == golden(f"transform_{name}.tf")
    assert len(chroms) == 1 and chroms[0]["name"] == b"chr1" and dropped == 0
    assert chroms[0]["line_count"] == bed.count(b"\n")


def test_transform_worked_examples(oracle):
    # outputs of the reference binary, and the reset at a chromosome change
    tf, ch, _ = oracle.transform(b"chr1\t100\t200\tid1\t5\t+\nchr1\t150\t250\tid2\t7\t-\nchr1\t300\t400\tid3\t1\t+\nchr1\t400\t450\n")
    assert tf == b"p100\n100\tid1\t5\t+\n-50\tid2\t7\t-\n50\tid3\t1\t+\np50\n0\n"
    tf, ch, _ = oracle.transform(b"chr1\t0\t0\nchr1\t0\t10\nchr1\t20\t30\nchr1\t30\t40\nchr1\t35\t50\n")
    assert tf == b"0\np10\n0\n10\n0\np15\n-5\n"
    assert ch[0]["bases_nonunique"] == 45 and ch[0]["bases_unique"] == 40
    tf, ch, _ = oracle.transform(b"chr1\t5\t9\nchr2\t10\t20\tx\nchr2\t30\t40\ty\nchr1\t1\t2\n")
    assert tf == b"p4\n5\n" + b"p10\n10\tx\n10\ty\n" + b"p1\n1\n"
    assert [c["name"] for c in ch] == [b"chr1", b"chr2", b"chr1"]        # a reappearing name opens a new stream
    assert [c["line_count"] for c in ch] == [1, 2, 1]
    assert [(c["tf_off"], c["tf_len"]) for c in ch] == [(0, 5), (5, 14), (19, 5)]


def test_transform_edges(oracle):
    assert oracle.transform(b"") == (b"", [], 0)
    tf, ch, dropped = oracle.transform(b"chr1\t1\t2\nchr1\t5\t6")       # unterminated tail dropped (hpp:181-190)
    assert tf == b"p1\n1\n" and dropped == 8 and ch[0]["line_count"] == 1
    tf, ch, _ = oracle.transform(b"chr1\t1\t2\t\n")                     # empty remainder: no tab is written (hpp:470)
    assert tf == b"p1\n1\n"
    tf, ch, _ = oracle.transform(b"c\t10\t5\n")                         # negative length prints its sign
    assert tf == b"p-5\n10\n"
    with pytest.raises(ValueError):
        oracle.transform(b"chr1\t5\n")
    big = 9000000000000000000
    tf, ch, _ = oracle.transform(f"c\t{big}\t{big + 7}\n".encode())
    assert tf == f"p7\n{big}\n".encode()


def test_transform_then_bzip2_roundtrip(oracle):
    bed = synth.bed(2, 30000)
    tf, ch, _ = oracle.transform(bed)
    assert len(ch) == 24
    for c in ch[:3]:
        s = tf[c["tf_off"]:c["tf_off"] + c["tf_len"]]
        z = oracle.bz_compress(s, 9)
        assert bz2.decompress(z) == s


def test_archive_container(oracle):
    import json
    bed = b"chr1\t5\t9\nchr2\t10\t20\tx\nchr2\t30\t40\ty\n"
    arc = oracle.archive(bed, 9, 'a "note"')
    assert arc[:4] == bytes([0xca, 0x5c, 0xad, 0x1a])                   # starch3api.hpp:907-910
    nl = arc.index(b"\n", 4)
    meta = json.loads(arc[4:nl])
    assert meta["archive"]["note"] == 'a "note"' and meta["archive"]["compression"] == "bzip2"
    payload = arc[nl + 1:]
    texts = [bz2.decompress(payload[s["offset"]:s["offset"] + s["size"]]) for s in meta["streams"]]
    assert texts == [b"p4\n5\n", b"p10\n10\tx\n10\ty\n"]
    assert [s["lines"] for s in meta["streams"]] == [1, 2]
    # compact jansson-style text: no spaces, insertion order
    assert arc[4:nl].startswith(b'{"archive":{"type":"starch","version":{"major":3,"minor":0,"revision":0},')


def test_metadata_text_is_what_the_reference_jansson_prints(oracle):
    """ARCHIVE_FORMAT.md says the metadata is the text jansson 2.9 json_dumps(JSON_COMPACT) prints; here the
    reference's own vendored jansson (oracle/_ref/libs3jansson.so, built from its tarball) prints it."""
    import json
    if not oracle.have_jansson():
        pytest.skip("oracle/_ref/libs3jansson.so not built (no /root/reference here)")
    cases = [
        (synth.bed(2, 3000).tobytes(), 9, ""),
        (synth.bed(5, 2000).tobytes(), 3, 'note with "quotes"\tand\\slashes/\x01\x1f and caf\u00e9'),
        (b'we"ird\\n\x07me\t1\t5\n' + "chr\u00e9\t2\t9\tz\n".encode() + b"a/b\t1\t2\n\x7f\t3\t4\n", 1, "x"),
        (b"", 9, "empty"),
    ]
    for bed, level, note in cases:
        arc = oracle.archive(bed, level, note)
        nl = arc.index(b"\n", 4)
        meta = json.loads(arc[4:nl])
        streams = [(s["chromosome"].encode(), s["offset"], s["size"], s["lines"], s["blocks"], s["transformedBytes"],
                    s["nonUniqueBases"], s["uniqueBases"]) for s in meta["streams"]]
        assert oracle.jansson_header(level, note, streams) == arc[4:nl]
    # jansson refuses text that is not UTF-8; this writer passes the bytes through (documented in ARCHIVE_FORMAT.md)
    assert oracle.jansson_header(9, "", [(b"chr\xff", 0, 1, 1, 1, 1, 1, 1)]) is None


@pytest.mark.parametrize("cfg,lines,level", [(1, 200000, 1), (2, 200000, 2), (4, 60000, 3), (3, 400000, 1), (5, 60000, 9)])
def test_block_parallel_oracle_equals_serial(oracle, cfg, lines, level):
    """archive_mt / bz_compress_blockwise (blocks of one stream compressed on a thread pool and re-joined bit by
    bit) are the checker at the BASELINE sizes; here they are pinned against the serial reference libbz2."""
    bed = synth.bed(cfg, lines).tobytes()
    assert oracle.archive_mt(bed, level, "mt") == oracle.archive(bed, level, "mt")
    tf, ch, _ = oracle.transform(bed)
    s = tf[ch[0]["tf_off"]:ch[0]["tf_off"] + ch[0]["tf_len"]]
    comp = oracle.ref_bz_compress if oracle.have_ref() else oracle.bz_compress
    assert oracle.bz_compress_blockwise(s, level) == comp(s, level)


def test_oracle_unarchive_roundtrip(oracle):
    """the decoder path's checker: unarchive(archive(bed)) == bed on every synthetic shape (CPU only)"""
    from starch3_b200 import synth
    for cfg in (1, 2, 3, 4, 5):
        bed = synth.bed(cfg, 20000).tobytes()
        assert oracle.unarchive(oracle.archive(bed)) == bed, cfg
    assert oracle.inverse_transform(b"chr1", b"p100\n100\tid1\t5\t+\n-50\tid2\t7\t-\n50\tid3\t1\t+\np50\n0\n") == \
        b"chr1\t100\t200\tid1\t5\t+\nchr1\t150\t250\tid2\t7\t-\nchr1\t300\t400\tid3\t1\t+\nchr1\t400\t450\n"
