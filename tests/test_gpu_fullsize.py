"""GPU parity at the BASELINE.json sizes and across batch boundaries.

The checker is oracle.archive_mt: restated transform + the reference's vendored libbz2 (oracle/_ref), chromosomes
and the blocks of a chromosome spread over the host threads and re-joined bit by bit (pinned against the serial
reference in tests/test_oracle.py).  Every comparison is of whole archives, byte for byte."""
import json

import numpy as np
import pytest

from starch3_b200 import synth

pytestmark = pytest.mark.gpu


def _same(res_archive, expect):
    if res_archive == expect:
        return True
    n = min(len(res_archive), len(expect))
    a = np.frombuffer(res_archive[:n], dtype=np.uint8)
    b = np.frombuffer(expect[:n], dtype=np.uint8)
    d = np.flatnonzero(a != b)
    raise AssertionError(f"archives differ: sizes {len(res_archive)} / {len(expect)}, first difference at byte "
                         f"{int(d[0]) if len(d) else n}, {len(d)} differing bytes")


@pytest.mark.parametrize("cfg,lines", [(2, 10_000_000), (4, 20_000_000), (1, 1_000_000)])
def test_archive_byte_parity_at_baseline_size(ctx, oracle, cfg, lines):
    """cfg2 at its 10 M lines (296 blocks, 24 chromosomes), cfg4 at its 20 M lines (one chromosome, ~920 blocks of a
    70-symbol alphabet), cfg1 at its 1 M lines: the archive of the host entry (pipelined upload, several batches where
    the blocks do not fit one) equals the reference-libbz2 oracle's."""
    bed = synth.bed(cfg, lines)
    res = ctx.compress_bed(bed, 9, note="baseline size", lazy=True)
    got = bytes(res.archive_view)
    assert res.n_lines == lines
    del res
    _same(got, oracle.archive_mt(bed, 9, "baseline size"))


def test_archive_byte_parity_cfg3_sample_device_entry(ctx, oracle):
    """30 M lines of cfg3 (dense BED3, 660 MB, one chromosome) through the device-resident entry."""
    import torch
    bed = synth.bed(3, 30_000_000)
    t = torch.from_numpy(bed).cuda()
    res = ctx.compress_bed_device(t.data_ptr(), t.numel(), 9, note="", want_archive=True, bed_bytes=None)
    _same(res.archive, oracle.archive_mt(bed, 9, ""))
    assert res.rle_bytes > 0 and res.mtf_symbols > res.n_blocks and abs(sum(res.stage_ms.values()) - res.device_ms) < 0.25 * res.device_ms + 1.0


@pytest.mark.parametrize("batch", [1, 3, 7])
def test_batches_of_blocks(ctx, oracle, batch, monkeypatch):
    """S3G_BATCH forces stages 3b..3d through several batches (run_pool_append with b0 > 0): cfg1 at 1.2 M lines has
    12 blocks in one stream, the cfg5 mix has streams that end inside a batch."""
    monkeypatch.setenv("S3G_PARTS", "1")
    monkeypatch.setenv("S3G_BATCH", str(batch))
    for cfg, lines in ((1, 1_200_000), (5, 600_000)):
        bed = synth.bed(cfg, lines)
        res = ctx.compress_bed(bed, 9, note="b")
        assert res.n_blocks >= 10
        _same(res.archive, oracle.archive_mt(bed, 9, "b"))
    monkeypatch.setenv("S3G_PARTS", "3")           # batches inside the ranges of the pipelined entry
    bed = synth.bed(2, 400_000)
    _same(ctx.compress_bed(bed, 1, note="b").archive, oracle.archive_mt(bed, 1, "b"))


def test_many_small_chromosomes(ctx, oracle):
    """Thousands of contigs: one small block each.  Block bytes are packed (BlockInfo.blk_off), not one 900 kB slot per
    block, and the bit-level concatenation indexes blocks in grid.x."""
    rng = np.random.default_rng(5)
    lines = []
    for c in range(5000):
        pos = 0
        for _ in range(int(rng.integers(1, 6))):
            pos += int(rng.integers(1, 1000))
            ln = int(rng.integers(1, 400))
            lines.append(b"contig_%05d\t%d\t%d\tf%d\n" % (c, pos, pos + ln, int(rng.integers(0, 99))))
            pos += ln
    bed = b"".join(lines)
    res = ctx.compress_bed(bed, 9)
    assert len(res.chroms) == 5000 and res.n_blocks == 5000
    _same(res.archive, oracle.archive_mt(bed, 9, ""))


def test_read_streams_after_either_host_entry(ctx, monkeypatch):
    bed = synth.bed(2, 80_000).tobytes()
    for parts in ("1", "3"):
        monkeypatch.setenv("S3G_PARTS", parts)
        res = ctx.compress_bed(bed, 9)
        assert ctx.read_streams(res.streams_size) == res.archive[res.streams_off:]


def test_two_contexts_in_one_process(oracle):
    """Contexts on the same device and, when the box has more than one GPU, on different devices, from one process:
    the dynamic shared-memory attributes are set per context (they are per device), so the second device's kernels launch."""
    import torch
    import starch3_b200 as s3
    bed = synth.bed(2, 60_000).tobytes()
    expect = oracle.archive(bed, 9, "")
    devs = [0, 0] + ([1] if torch.cuda.device_count() > 1 else [])
    ctxs = [s3.Context(d) for d in devs]
    try:
        for c in ctxs:
            assert c.compress_bed(bed, 9).archive == expect
        for c in reversed(ctxs):
            assert c.bz_compress(b"abc" * 1000, 9) == oracle.bz_compress(b"abc" * 1000, 9)
    finally:
        for c in ctxs:
            c.close()


def test_header_text_is_janssons(ctx, oracle):
    if not oracle.have_jansson():
        pytest.skip("oracle/_ref/libs3jansson.so not built")
    bed = synth.bed(5, 20_000).tobytes()
    note = 'n "q"\t\\ / \x02 café'
    arc = ctx.compress_bed(bed, 9, note=note).archive
    nl = arc.index(b"\n", 4)
    meta = json.loads(arc[4:nl])
    streams = [(s["chromosome"].encode(), s["offset"], s["size"], s["lines"], s["blocks"], s["transformedBytes"],
                s["nonUniqueBases"], s["uniqueBases"]) for s in meta["streams"]]
    assert oracle.jansson_header(9, note, streams) == arc[4:nl]


def test_archive_byte_parity_bucket_form_of_the_block_sort(ctx, oracle, monkeypatch):
    """S3G_SORT=bucket: every block of at least 8192 bytes through bwt_bucket.cu (sample-sort buckets finished in shared
    memory) -- 3 M lines of cfg2 (89 blocks, 24 chromosomes) and 1 M lines of cfg4 (70 symbols)."""
    monkeypatch.setenv("S3G_SORT", "bucket")
    monkeypatch.setenv("S3G_PARTS", "1")
    for cfg, lines in ((2, 3_000_000), (4, 1_000_000)):
        bed = synth.bed(cfg, lines)
        before = ctx.sort_stats
        res = ctx.compress_bed(bed, 9, note="bucket")
        assert ctx.sort_stats[0] - before[0] >= res.n_blocks - 24
        _same(res.archive, oracle.archive_mt(bed, 9, "bucket"))


@pytest.mark.parametrize("cfg,lines", [(2, 10_000_000), (3, 30_000_000)])
def test_decoder_roundtrip_at_size(ctx, cfg, lines):
    """the decoder path at size: cfg2 at its 10 M lines (296 blocks over 24 streams), 30 M lines of cfg3 (one stream of
    75 blocks): the archive decodes to the BED text it was made from, every block and stream CRC checked on the way"""
    bed = synth.bed(cfg, lines)
    res = ctx.compress_bed(bed, 9, note="", lazy=True)
    arc = bytes(res.archive_view)
    n_blocks = res.n_blocks
    del res
    got, info = ctx.decompress_archive(arc)
    assert info["n_blocks"] == n_blocks
    assert len(got) == bed.nbytes and np.array_equal(np.frombuffer(got, dtype=np.uint8), bed)


@pytest.mark.parametrize("batch", [75, 178, 230, 300, 450, 0])
def test_batches_that_leave_partial_waves(ctx, oracle, batch, monkeypatch):
    """MTF and Huffman cut a batch into whole waves of one CTA per block plus a remainder spread over clusters of 2, 4 and 8
    CTAs per block (mtf_huff.cu, plan_chunks).  700 small streams and a few multi-block ones, in batches whose sizes leave every
    kind of remainder (0 = one batch of all blocks)."""
    monkeypatch.setenv("S3G_PARTS", "1")
    if batch:
        monkeypatch.setenv("S3G_BATCH", str(batch))
    rng = np.random.default_rng(11)
    lines = []
    for c in range(700):
        pos = 0
        n = 30000 if c % 97 == 5 else int(rng.integers(1, 40))
        for _ in range(n):
            pos += int(rng.integers(1, 1000))
            ln = int(rng.integers(1, 400))
            lines.append(b"ctg%04d\t%d\t%d\tf%d\n" % (c, pos, pos + ln, int(rng.integers(0, 99))))
            pos += ln
    bed = b"".join(lines)
    res = ctx.compress_bed(bed, 1, note="w")
    assert res.n_blocks > 720
    _same(res.archive, oracle.archive_mt(bed, 1, "w"))
