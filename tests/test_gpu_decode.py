"""GPU parity tests of the decoder path (decode.cu) through the C ABI: bzip2 stream decode against the reference's
BZ2_bzDecompress / CPython bz2 and bzip2's own golden vectors, the inverse transform against the checker's, whole
archives back to the BED text they were made from, and what happens to damaged archives."""
import bz2

import numpy as np
import pytest

from conftest import golden
from starch3_b200 import synth

pytestmark = pytest.mark.gpu


def _plain():
    rng = np.random.default_rng(17)
    yield "empty", b"", 9
    yield "one", b"a", 9
    yield "run4", b"aaaa", 9
    yield "run5", b"aaaaab", 9
    yield "run8", b"aaaaaaaa", 9
    yield "run9_then_same_count", b"a" * 4 + b"b" * 4 + b"aaaa" * 3 + b"\x04" * 9, 9
    yield "run255", b"b" * 255, 9
    yield "run256", b"b" * 256 + b"c", 9
    yield "run1000", b"z" * 1000, 9
    yield "run_long", b"q" * 70000 + b"r" * 3, 1
    yield "count_bytes", b"".join(bytes([c]) * (4 + c) for c in range(0, 252, 7)), 9
    yield "runs", b"".join(bytes([int(c)]) * int(n) for c, n in zip(rng.integers(48, 52, 30000), rng.integers(1, 12, 30000))), 1
    yield "digits_l1", bytes(rng.integers(48, 58, 350000, dtype=np.uint8)), 1
    yield "random", bytes(rng.integers(0, 256, 250000, dtype=np.uint8)), 1
    yield "random9", bytes(rng.integers(0, 256, 1200000, dtype=np.uint8)), 9
    yield "allbytes", bytes(range(256)) * 40, 9
    yield "periodic2", b"5\n" * 3000, 9
    yield "periodic2_big", b"5\n" * 400000, 9
    yield "periodic_ab", b"ab" * 6000, 9
    yield "periodic_text", bytes(rng.integers(97, 123, 1000, dtype=np.uint8)) * 300, 9
    yield "two_symbols", bytes(rng.integers(0, 2, 120000, dtype=np.uint8) + 48), 9
    yield "long_codes", b"".join(bytes([i]) * (1 << min(i, 14)) for i in range(20)) + bytes(rng.integers(0, 256, 3000, dtype=np.uint8)), 9


PLAIN = list(_plain())


@pytest.mark.parametrize("name,data,level", PLAIN, ids=[p[0] for p in PLAIN])
def test_bz_decompress(ctx, oracle, name, data, level):
    z = bz2.compress(data, level)
    assert ctx.bz_decompress(z) == data
    if oracle.have_ref():
        assert oracle.ref_bz_decompress(z, len(data) + 16) == data


@pytest.mark.parametrize("k", [1, 2, 3])
def test_bz_decompress_bzip2_golden_vectors(ctx, k):
    z = golden(f"sample{k}.bz2")
    assert ctx.bz_decompress(z) == bz2.decompress(z)


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_bz_decompress_transformed_streams(ctx, oracle, cfg):
    tf, chroms, _ = oracle.transform(synth.bed(cfg, 60000).tobytes())
    c = max(chroms, key=lambda c: c["tf_len"])
    s = tf[c["tf_off"]:c["tf_off"] + c["tf_len"]]
    for level in (1, 9):
        z = oracle.bz_compress(s, level)
        assert ctx.bz_decompress(z) == s
    assert ctx.bz_decompress(ctx.bz_compress(s, 9)) == s


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_inverse_transform(ctx, oracle, cfg):
    tf, chroms, _ = oracle.transform(synth.bed(cfg, 30000).tobytes())
    for c in chroms[:3]:
        s = tf[c["tf_off"]:c["tf_off"] + c["tf_len"]]
        assert ctx.inverse_transform(s, c["name"]) == oracle.inverse_transform(c["name"], s)


def test_inverse_transform_edges(ctx, oracle):
    cases = [b"", b"5\n", b"p7\n3\n", b"0\n0\n0\n", b"p100\n100\tid1\t5\t+\n-50\tid2\t7\t-\n50\tid3\t1\t+\np50\n0\n",
             b"p9000000000000000000\n9000000000000000\tx\n-8999999999999999999\n", b"p3\np4\n1\t\tq\n2\t\n"]
    for tf in cases:
        assert ctx.inverse_transform(tf, b"chrQ") == oracle.inverse_transform(b"chrQ", tf), tf
    assert ctx.inverse_transform(b"p2\n1\n", b"") == b"\t1\t3\n"


@pytest.mark.parametrize("cfg,lines", [(1, 60000), (2, 60000), (3, 150000), (4, 30000), (5, 60000)])
def test_archive_roundtrip(ctx, oracle, cfg, lines):
    bed = synth.bed(cfg, lines).tobytes()
    arc = ctx.compress_bed(bed, 9, note="rt").archive
    got, info = ctx.decompress_archive(arc)
    assert got == bed
    assert got == oracle.unarchive(arc)
    assert info["n_blocks"] >= info["n_streams"] >= 1
    # an archive written by the CPU checker decodes to the same text
    assert ctx.decompress_archive(oracle.archive(bed, 9, "rt"))[0] == bed


def test_archive_roundtrip_odd_inputs(ctx, oracle):
    for bed in (b"", b"chrZ\t0\t1\n", b"chr1\t5\t9\nchr2\t10\t20\tx\nchr2\t30\t40\ty\nchr1\t1\t2\n",
                b"".join(f"s{i}\t{i}\t{i + 3}\tn{i}\n".encode() for i in range(5000)),
                b'we"ird\\n\x07me\t1\t5\n' + "chr\u00e9\t2\t9\tz\n".encode(),
                b"c\t10\t5\nc\t3\t4\tq\t\tr\n"):
        arc = ctx.compress_bed(bed, 9).archive
        assert ctx.decompress_archive(arc)[0] == bed
    # level 1: many small blocks per stream, joined at arbitrary bit offsets
    bed = synth.bed(2, 40000).tobytes()
    assert ctx.decompress_archive(ctx.compress_bed(bed, 1).archive)[0] == bed


def test_damaged_archives_are_refused(ctx):
    import starch3_b200 as s3
    bed = synth.bed(2, 20000).tobytes()
    arc = bytearray(ctx.compress_bed(bed, 9).archive)
    nl = arc.index(b"\n", 4)
    for where in (nl + 1 + 5, nl + 1 + 2000, len(arc) - 3, len(arc) - 40):
        bad = bytearray(arc)
        bad[where] ^= 0x10
        with pytest.raises(s3.Starch3Error):
            ctx.decompress_archive(bytes(bad))
    with pytest.raises(s3.Starch3Error):
        ctx.decompress_archive(bytes(arc[:len(arc) - 100]))
    with pytest.raises(s3.Starch3Error):
        ctx.decompress_archive(b"\x00" * 64)
    # the context is still usable
    assert ctx.decompress_archive(bytes(arc))[0] == bed
