"""GPU parity tests, stage by stage, through the C ABI (include/starch3_b200.h) against the
CPU oracle on the same seeded inputs.  Bit-exact everywhere: this path is integer/byte work."""
import bz2

import numpy as np
import pytest

from conftest import golden
from starch3_b200 import synth

pytestmark = pytest.mark.gpu


def _beds():
    yield "cfg1", synth.bed(1, 20000).tobytes()
    yield "cfg2", synth.bed(2, 20000).tobytes()
    yield "cfg3", synth.bed(3, 20000).tobytes()
    yield "cfg3const", synth.bed(3, 20000, variant=1).tobytes()
    yield "cfg4", synth.bed(4, 20000).tobytes()
    yield "cfg5", synth.bed(5, 20000).tobytes()
    yield "reappear", b"chr1\t5\t9\nchr2\t10\t20\tx\nchr2\t30\t40\ty\nchr1\t1\t2\n"
    yield "overlap", golden("transform_overlap.bed")
    yield "zerolen", golden("transform_zerolen.bed")
    yield "negative", b"c\t10\t5\nc\t3\t4\tq\t\tr\nc\t-7\t+9\n"
    yield "emptyrem", b"chr1\t1\t2\t\nchr1\t3\t4\t\t\n"
    yield "big", f"c\t{9 * 10**18}\t{9 * 10**18 + 7}\nc\t{9 * 10**18 + 1}\t{9 * 10**18 + 8}\n".encode()
    yield "tail", b"chr1\t1\t2\nchr1\t5\t6"
    yield "one", b"chrZ\t0\t1\n"
    yield "manychroms", b"".join(f"s{i}\t{i}\t{i + 3}\tn{i}\n".encode() for i in range(5000))
    yield "longrem", b"chr1\t1\t2\t" + b"x" * 5000 + b"\nchr1\t2\t3\t" + b"y\tz" * 700 + b"\n"
    # the fused front end walks the input in 8 KiB tiles with the 512 bytes before them staged: lines that end exactly
    # at, one before and one after a tile boundary; a line longer than two tiles followed by short ones; the shortest
    # valid lines (2731 of them per tile); chromosome changes at tile boundaries
    for d in (-1, 0, 1):
        head = b"chrA\t10\t20\t" + b"q" * (8192 + d - 12 - 1) + b"\n"
        yield f"tile_edge{d:+d}", head + b"chrA\t15\t30\nchrB\t1\t2\tz\n" + b"chrB\t7\t9\n" * 900
    yield "verylong", b"c\t1\t5\n" + b"c\t2\t9\t" + b"L" * 20000 + b"\n" + b"".join(f"c\t{i}\t{i + 4}\n".encode() for i in range(10, 3000))
    yield "tiny_lines", b"\t\t\n" * 9000
    yield "tiny_changes", b"".join((b"a" if (i // 7) % 2 else b"") + b"\t%d\t%d\n" % (i, i + 1) for i in range(8000))
    yield "edge_changes", b"".join(f"k{i // 315}\t{i}\t{i + 2}\tabcdefgh\n".encode() for i in range(4000))


BEDS = list(_beds())


def _num(field):
    """[sign]digits up to the first other byte, 0 if there are none (the in-domain behaviour of sscanf("%lld"), hpp:306-307)"""
    import re
    m = re.match(rb"[+-]?[0-9]*", field)
    txt = m.group(0)
    return int(txt) if txt.strip(b"+-") else 0


@pytest.mark.parametrize("name,bed", BEDS, ids=[n for n, _ in BEDS])
def test_tokenize(ctx, oracle, name, bed):
    t = ctx.tokenize(bed)
    lines = bed.split(b"\n")[:-1]
    assert t["n_lines"] == len(lines)
    pos = 0
    prev_chr = None
    for i, ln in enumerate(lines):
        f = ln.split(b"\t", 3)
        assert t["line_start"][i] == pos
        assert t["start"][i] == _num(f[1]) and t["stop"][i] == _num(f[2])
        assert t["rem_off"][i] == (len(ln) - len(f[3]) if len(f) == 4 else len(ln))
        assert t["chrom_change"][i] == (1 if f[0] != prev_chr else 0)
        prev_chr = f[0]
        pos += len(ln) + 1
    assert t["line_start"][len(lines)] == pos


@pytest.mark.parametrize("front", ["one", "two"])
@pytest.mark.parametrize("name,bed", BEDS, ids=[n for n, _ in BEDS])
def test_transform(ctx, oracle, name, bed, front, monkeypatch):
    """the front end as one launch (k_front_fused: decoupled look-back over the chunk aggregates) and as its two passes"""
    monkeypatch.setenv("S3G_FRONT", front)
    tf, chroms, dropped = ctx.transform(bed)
    otf, ochroms, odropped = oracle.transform(bed)
    assert tf == otf
    assert dropped == odropped
    assert chroms == ochroms


def test_transform_one_launch_large_and_many_chromosomes(ctx, oracle, monkeypatch):
    """thousands of chunks in flight (the look-back folds up to 32 predecessors a step), more chromosomes than the one-launch
    form has table room for (it reports the overflow and the two passes take over), and the same answer from both forms"""
    big = synth.bed(5, 1_500_000).tobytes()
    many = b"".join(b"s%d\t%d\t%d\n" % (i, i, i + 3) for i in range(6000))
    for bed in (big, many):
        monkeypatch.setenv("S3G_FRONT", "one")
        a = ctx.transform(bed)
        monkeypatch.setenv("S3G_FRONT", "two")
        b = ctx.transform(bed)
        assert a == b
        otf, ochroms, odropped = oracle.transform(bed)
        assert a[0] == otf and a[1] == ochroms and a[2] == odropped


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg3const", "cfg4", "overlap", "zerolen"])
def test_transform_vs_reference_binary_golden(ctx, name):
    tf, chroms, _ = ctx.transform(golden(f"transform_{name}.bed"))
    assert tf == golden(f"transform_{name}.tf")


def test_transform_errors(ctx):
    import starch3_b200 as s3
    with pytest.raises(s3.Starch3Error) as e:
        ctx.transform(b"chr1\t5\t6\nchr1\t7\n")
    assert e.value.code == -4
    assert ctx.transform(b"") == (b"", [], 0)
    assert ctx.transform(b"no newline") == (b"", [], 10)
    for bad in (b"\n" * 10000, b"chr1\t1\t2\n" * 2000 + b"chr1\t5\n" + b"chr1\t7\t8\n" * 2000, b"a\tb\n"):
        with pytest.raises(s3.Starch3Error) as e:
            ctx.transform(bad)
        assert e.value.code == -4


def _streams():
    rng = np.random.default_rng(7)
    yield "empty", b"", 9
    yield "one", b"a", 9
    yield "run4", b"aaaa", 9
    yield "run5", b"aaaaab", 9
    yield "run255", b"b" * 255, 9
    yield "run256", b"b" * 256 + b"c", 9
    yield "run1000", b"z" * 1000, 9
    yield "run_long", b"q" * 70000 + b"r" * 3, 1
    yield "runs", b"".join(bytes([int(c)]) * int(n) for c, n in zip(rng.integers(48, 52, 30000), rng.integers(1, 12, 30000))), 1
    yield "digits_l1", bytes(rng.integers(48, 58, 350000, dtype=np.uint8)), 1
    yield "random", bytes(rng.integers(0, 256, 250000, dtype=np.uint8)), 1
    yield "allbytes", bytes(range(256)) * 40, 9
    nmax = 100000 - 19
    body = bytes(97 + (i * 7 + i // 26) % 26 for i in range(nmax))
    yield "exactfill1", body + b"\n", 1
    yield "exactfill2", body + b"\n#", 1
    yield "exactfill_runs", b"k" * 255 * 4 + body[: nmax - 20] + b"\n", 1


STREAMS = list(_streams())


@pytest.mark.parametrize("name,data,level", STREAMS, ids=[s[0] for s in STREAMS])
def test_rle1_cut_crc(ctx, oracle, name, data, level):
    blocks, rle = ctx.rle1(data, level)
    oblocks, orle = oracle.rle1_blocks(data, level)
    oblocks = [b for b in oblocks if b["nblock"]]
    assert [(b["in_start"], b["in_end"], b["nblock"]) for b in blocks] == [(b["in_start"], b["in_end"], b["nblock"]) for b in oblocks]
    assert np.array_equal(rle, orle)
    for b, ob in zip(blocks, oblocks):
        assert b["crc"] == ob["crc"]
        assert np.array_equal(b["in_use"], ob["in_use"])


def _blocks():
    rng = np.random.default_rng(5)
    yield "digits", bytes(rng.integers(48, 58, 20000, dtype=np.uint8))
    yield "tiny1", b"a"
    yield "tiny2", b"ba"
    yield "tiny5", b"hello"
    yield "binary", bytes(rng.integers(0, 2, 5000, dtype=np.uint8) + 65)
    yield "random256", bytes(rng.integers(0, 256, 40000, dtype=np.uint8))
    yield "repeats", (b"abcabd" * 3000)[:17999]
    yield "deep", bytes(rng.integers(97, 100, 300, dtype=np.uint8)) * 97 + b"!"
    yield "tf_cfg2", None
    yield "tf_cfg4", None


BLOCKS = list(_blocks())


def _resolve_block(oracle, name, blk):
    if blk is not None:
        return blk
    cfg = int(name[-1])
    tf, _, _ = oracle.transform(synth.bed(cfg, 40000))
    return tf[:900000]


def test_bwt_batch(ctx, oracle):
    blks = [_resolve_block(oracle, n, b) for n, b in BLOCKS]
    got = ctx.bwt(blks)
    for (name, _), blk, (ptr, orig) in zip(BLOCKS, blks, got):
        optr, oorig = oracle.bwt(blk)
        assert np.array_equal(ptr, optr), name
        assert orig == oorig, name


def _finisher_blocks():
    """Blocks aimed at the group finisher (bwt.cu): groups at and just past the size it sorts in shared
    memory (2048), ties that need several deeper levels, ties that outlast them (doubling rounds), a
    group spilling over tile boundaries, and a large block of long repeated fields."""
    rng = np.random.default_rng(11)
    ctx9 = b"ABCDEFGHIJKL"                                    # longer than any initial key of a 30-symbol alphabet
    def tagged(count, width, filler=3000):
        # `count` occurrences of the same context, each followed by a distinct number, in random filler
        parts = [bytes(rng.integers(97, 123, filler, dtype=np.uint8))]
        for i in rng.permutation(count):
            parts.append(ctx9 + (b"%0*d" % (width, int(i))) + bytes(rng.integers(97, 123, 5, dtype=np.uint8)))
        return b"".join(parts)
    yield "group2048", tagged(2048, 6)
    yield "group2049", tagged(2049, 6)
    yield "group5000", tagged(5000, 6)
    # equal for 40 symbols after the context: level 0 ties, deeper levels resolve
    yield "deep_ties", b"".join(ctx9 + b"x" * 40 + (b"%05d" % int(i)) + b"\n" for i in rng.permutation(700))
    # equal for 300 symbols: the levels give up, the doubling rounds finish
    yield "very_deep_ties", b"".join(ctx9 + b"y" * 300 + (b"%04d" % int(i)) + b"\n" for i in rng.permutation(60))
    # long constant fields on every line of a big block (a BED file with a repeated annotation)
    lines = [b"%d\tENSG%011d\tprotein_coding\tKNOWN\t+\n" % (int(rng.integers(1, 500)), int(rng.integers(0, 10**9))) for _ in range(16000)]
    yield "annotated", b"".join(lines)[:899000]
    yield "two_symbols", bytes(rng.integers(0, 2, 120000, dtype=np.uint8) + 48)
    yield "all_bytes_big", bytes(rng.integers(0, 256, 300000, dtype=np.uint8))


@pytest.mark.parametrize("name,blk", list(_finisher_blocks()), ids=[n for n, _ in _finisher_blocks()])
def test_bwt_finisher_limits_and_fallback(ctx, oracle, name, blk):
    (ptr, orig), = ctx.bwt([blk])
    optr, oorig = oracle.bwt(blk)
    assert orig == oorig, name
    assert np.array_equal(ptr, optr), name


def _periodic_blocks():
    """Blocks that are a power of a shorter string: equal rotations tie and origPtr (and the order inside the ties) is
    whatever fallbackSort leaves (bz/blocksort.c:212-329), replayed by k_fallback_exact -- one thread per non-uniform
    bucket of a round.  Short periods (all buckets uniform after the first round), long periods over small and large
    alphabets (rounds that really sort), a full-size block."""
    rng = np.random.default_rng(3)
    yield "p2", b"5\n" * 3000
    yield "p2_big", b"5\n" * 400000
    yield "p7", b"p20\n5\n\n" * 20000
    yield "p1", b"x" * 100000
    yield "ab", b"ab" * 6000
    yield "allbytes", bytes(range(256)) * 40
    yield "text1000", bytes(rng.integers(97, 123, 1000, dtype=np.uint8)) * 300
    yield "bits5000", bytes(rng.integers(48, 50, 5000, dtype=np.uint8)) * 60
    yield "bedlike", b"".join(b"%d\tid-%d\t%d\t+\n" % (i * 7 % 50, i, i * 13 % 1000) for i in range(40)) * 200
    yield "small", b"abcab" * 7


@pytest.mark.parametrize("name,blk", list(_periodic_blocks()), ids=[n for n, _ in _periodic_blocks()])
def test_bwt_periodic_blocks(ctx, oracle, name, blk):
    (ptr, orig), = ctx.bwt([blk])
    optr, oorig = oracle.bwt(blk)
    assert orig == oorig, name
    assert np.array_equal(ptr, optr), name


def _bucket_blocks():
    """Blocks for the bucket form (bwt_bucket.cu, blocks of 8192 bytes and more): real shapes, groups of equal keys up to
    its sub-bucket capacity (256) with ties 1 .. 4 deeper levels long and ties that outlast its 12 levels (left to the
    doubling rounds), an alphabet of 2 and of 256 symbols, and sizes around the bucket-count steps."""
    rng = np.random.default_rng(23)
    def tagged(groups, depth, filler=20000):
        parts = [bytes(rng.integers(97, 123, filler, dtype=np.uint8))]
        for gi, count in enumerate(groups):
            # the group's own random context (no key is shared between groups), then `depth` symbols shared by the group
            tag = bytes(rng.integers(65, 91, 12, dtype=np.uint8)) + b"%03d" % gi + bytes(rng.integers(97, 123, depth, dtype=np.uint8))
            for i in rng.permutation(count):
                parts.append(tag + (b"%04d" % int(i)) + bytes(rng.integers(97, 123, int(rng.integers(1, 7)), dtype=np.uint8)))
        return b"".join(parts)
    yield "groups_shallow", tagged([2, 3, 17, 31, 32, 33, 63, 64, 65, 100, 128, 129, 200, 255], 0)
    yield "groups_deep20", tagged([5, 40, 90, 150, 250], 20)
    yield "groups_deep60", tagged([7, 64, 130, 240], 60)
    yield "groups_beyond_levels", tagged([3, 50, 120], 400)
    yield "two_symbols", bytes(rng.integers(0, 2, 60000, dtype=np.uint8) + 65)
    yield "all_bytes", bytes(rng.integers(0, 256, 70000, dtype=np.uint8))
    for n in (8192, 8193, 14081, 28160, 28161, 56321, 450000):
        yield "random%d" % n, bytes(rng.integers(97, 110, n, dtype=np.uint8))
    yield "skewed", bytes(rng.choice(np.frombuffer(b"aaaaaaaaaaaaaaaaaaaaaaaab\ncd", dtype=np.uint8), 300000))


BUCKET_BLOCKS = list(_bucket_blocks())


@pytest.mark.parametrize("name,blk", BUCKET_BLOCKS, ids=[n for n, _ in BUCKET_BLOCKS])
def test_bwt_bucket_form(ctx, oracle, name, blk, monkeypatch):
    monkeypatch.setenv("S3G_SORT", "bucket")
    before = ctx.sort_stats
    (ptr, orig), = ctx.bwt([blk])
    after = ctx.sort_stats
    optr, oorig = oracle.bwt(blk)
    assert orig == oorig, name
    assert np.array_equal(ptr, optr), name
    assert after[0] == before[0] + 1, "the bucket form did not run"
    if name != "skewed":
        assert after[1] == before[1], "handed back to the radix form"


def test_bwt_bucket_and_radix_forms_in_one_batch(ctx, oracle, monkeypatch):
    """Small blocks (radix form), large ones (bucket form) and one the bucket form hands back (2048 equal keys) in the
    same batch (S3G_SORT=bucket); then every block through the radix form alone (the default): same order."""
    rng = np.random.default_rng(3)
    blks = [b"abc" * 100, _resolve_block(oracle, "cfg2", None), bytes(rng.integers(97, 123, 5000, dtype=np.uint8)),
            dict(_finisher_blocks())["group2049"], _resolve_block(oracle, "cfg4", None)[:200000], b"ab" * 3000 + b"c",
            _resolve_block(oracle, "cfg3", None)[:120000]]
    expect = [oracle.bwt(b) for b in blks]
    for mode in ("bucket", None):
        if mode:
            monkeypatch.setenv("S3G_SORT", mode)
        else:
            monkeypatch.delenv("S3G_SORT")
        before = ctx.sort_stats
        got = ctx.bwt(blks)
        for (ptr, orig), (optr, oorig) in zip(got, expect):
            assert orig == oorig and np.array_equal(ptr, optr)
        grew = ctx.sort_stats[0] - before[0]
        assert grew == (4 if mode else 0)


@pytest.mark.parametrize("mode", ["safe", "broken"])
def test_bwt_sort_safety_net(ctx, oracle, monkeypatch, mode):
    """The radix passes rank with ordered shared-memory atomics and the finisher checks that the keys it
    receives ascend (bwt.cu, k_sweep).  `safe` runs the peer-mask passes directly; `broken` makes the first
    attempt rank without any order, so the check must trip and the second attempt must give the right answer."""
    monkeypatch.setenv("S3G_SORT", mode)
    before = ctx.sort_retries
    blks = [_resolve_block(oracle, "cfg2", None), _resolve_block(oracle, "cfg4", None)[:300000], b"ab" * 3000 + b"c"]
    got = ctx.bwt(blks)
    for blk, (ptr, orig) in zip(blks, got):
        optr, oorig = oracle.bwt(blk)
        assert np.array_equal(ptr, optr)
        assert orig == oorig
    assert (ctx.sort_retries > before) == (mode == "broken")


def test_bwt_window_limits(ctx, oracle):
    """Groups around the sizes the warp finisher handles itself (a window of four rows: up to 97..128 rotations
    depending on where the group starts), the warp-per-group kernel (up to 256) and the CTA-per-group kernel
    (counting up to 256, bitonic level 0 above, up to 2048) take over."""
    rng = np.random.default_rng(5)
    ctx9 = b"QRSTUVWXYZ012"
    parts = [bytes(rng.integers(97, 123, 2000, dtype=np.uint8))]
    for count in (31, 32, 33, 64, 95, 96, 97, 98, 127, 128, 129, 130, 200, 255, 256, 257, 300, 511, 512, 513, 1000, 1500):
        tag = ctx9 + b"%03d" % count
        for i in rng.permutation(count):
            parts.append(tag + (b"%04d" % int(i)) + bytes(rng.integers(97, 123, int(rng.integers(1, 9)), dtype=np.uint8)))
    # the same with ties at level 0 (equal for 12 symbols after the context, then a distinct number)
    for count in (120, 250, 400):
        tag = ctx9 + b"T%03d" % count + b"=" * 12
        for i in rng.permutation(count):
            parts.append(tag + (b"%04d" % int(i)) + b"\n")
    blk = b"".join(parts)
    (ptr, orig), = ctx.bwt([blk])
    optr, oorig = oracle.bwt(blk)
    assert orig == oorig
    assert np.array_equal(ptr, optr)


def test_bwt_full_block_vs_reference(ctx, oracle):
    tf, _, _ = oracle.transform(synth.bed(1, 250000))
    blocks, rle = oracle.rle1_blocks(tf, 9)
    blk = rle[:blocks[0]["nblock"]]
    assert len(blk) >= 899981
    (ptr, orig), = ctx.bwt([blk])
    optr, oorig = (oracle.ref_bwt if oracle.have_ref() else oracle.bwt)(blk)
    assert orig == oorig
    assert np.array_equal(ptr, optr)


@pytest.mark.parametrize("name,blk", BLOCKS, ids=[n for n, _ in BLOCKS])
def test_mtf_and_huffman(ctx, oracle, name, blk):
    blk = _resolve_block(oracle, name, blk)
    ptr, _ = oracle.bwt(blk)
    in_use = np.zeros(256, dtype=np.uint8)
    in_use[np.frombuffer(blk, dtype=np.uint8)] = 1
    omtfv, ofreq, nu = oracle.mtf(blk, ptr, in_use)
    mtfv, freq = ctx.mtf(blk, ptr, in_use)
    assert np.array_equal(mtfv, omtfv)
    assert np.array_equal(freq[:nu + 2], ofreq[:nu + 2])
    oh = oracle.huff_select(omtfv, ofreq, nu)
    h = ctx.huff(omtfv, ofreq, in_use)
    assert h["n_groups"] == oh["n_groups"] and h["n_selectors"] == oh["n_selectors"]
    assert np.array_equal(h["selector"], oh["selector"])
    assert np.array_equal(h["len"][:oh["n_groups"], :nu + 2], oh["len"][:oh["n_groups"], :nu + 2])
    if oracle.have_ref():
        ref = oracle.ref_mtf_huff(blk, ptr, in_use)
        assert h["nbits"] == ref["nbits"]
        assert np.array_equal(h["bits"], ref["bits"])


@pytest.mark.parametrize("zrun", ["fused", "split"])
@pytest.mark.parametrize("name,blk", BLOCKS, ids=[n for n, _ in BLOCKS])
def test_mtf_zero_run_forms(ctx, oracle, name, blk, zrun, monkeypatch):
    """Zero-run coding runs inside the MTF kernels for batches that fill the SMs and as tile-parallel kernels for
    smaller ones (mtf_huff.cu, k_zrun_*): both forms on every block, whatever the batch size would pick."""
    monkeypatch.setenv("S3G_ZRUN", zrun)
    blk = _resolve_block(oracle, name, blk)
    ptr, _ = oracle.bwt(blk)
    in_use = np.zeros(256, dtype=np.uint8)
    in_use[np.frombuffer(blk, dtype=np.uint8)] = 1
    omtfv, ofreq, nu = oracle.mtf(blk, ptr, in_use)
    mtfv, freq = ctx.mtf(blk, ptr, in_use)
    assert np.array_equal(mtfv, omtfv)
    assert np.array_equal(freq[:nu + 2], ofreq[:nu + 2])


@pytest.mark.parametrize("cs", [1, 2, 4, 8])
@pytest.mark.parametrize("name", ["digits", "tiny5", "random256", "repeats", "tf_cfg2", "tf_cfg4"])
def test_mtf_and_huffman_cluster_sizes(ctx, oracle, name, cs, monkeypatch):
    """A batch smaller than the GPU spreads each bzip2 block over a thread-block cluster of 2, 4 or 8 CTAs (mtf_huff.cu:
    last occurrences, symbol frequencies, selector history and bit counts go through distributed shared memory).  Every
    size, forced, gives the ranks, frequencies, selectors, code lengths and bits of the one-CTA form and of the oracle."""
    monkeypatch.setenv("S3G_CLUSTER", str(cs))
    blk = _resolve_block(oracle, name, dict(BLOCKS)[name])
    ptr, _ = oracle.bwt(blk)
    in_use = np.zeros(256, dtype=np.uint8)
    in_use[np.frombuffer(blk, dtype=np.uint8)] = 1
    omtfv, ofreq, nu = oracle.mtf(blk, ptr, in_use)
    mtfv, freq = ctx.mtf(blk, ptr, in_use)
    assert np.array_equal(mtfv, omtfv)
    assert np.array_equal(freq[:nu + 2], ofreq[:nu + 2])
    oh = oracle.huff_select(omtfv, ofreq, nu)
    h = ctx.huff(omtfv, ofreq, in_use)
    assert np.array_equal(h["selector"], oh["selector"])
    assert np.array_equal(h["len"][:oh["n_groups"], :nu + 2], oh["len"][:oh["n_groups"], :nu + 2])
    if oracle.have_ref():
        ref = oracle.ref_mtf_huff(blk, ptr, in_use)
        assert h["nbits"] == ref["nbits"]
        assert np.array_equal(h["bits"], ref["bits"])


@pytest.mark.parametrize("cs", [1, 2, 4, 8])
def test_archive_with_every_cluster_size(ctx, oracle, cs, monkeypatch):
    monkeypatch.setenv("S3G_CLUSTER", str(cs))
    monkeypatch.setenv("S3G_PARTS", "1")
    bed = synth.bed(5, 60000).tobytes()
    assert ctx.compress_bed(bed, 9, note="c").archive == oracle.archive(bed, 9, "c")
    bed = synth.bed(4, 40000).tobytes()          # 70 symbols: the 512-thread form of the MTF kernel
    assert ctx.compress_bed(bed, 9, note="c").archive == oracle.archive(bed, 9, "c")


@pytest.mark.parametrize("zrun", ["fused", "split"])
def test_archive_with_either_zero_run_form(ctx, oracle, zrun, monkeypatch):
    monkeypatch.setenv("S3G_ZRUN", zrun)
    bed = synth.bed(5, 60000).tobytes()
    assert ctx.compress_bed(bed, 9, note="z").archive == oracle.archive(bed, 9, "z")
    # long zero runs across tile boundaries, and a block that is one run
    for data in (b"a" * 70000 + b"b" + b"a" * 9000 + b"cab" * 5000, b"\n".join(b"7" for _ in range(40000)) + b"\n", b"q" * 300000):
        z = ctx.bz_compress(data, 9)
        assert z == oracle.bz_compress(data, 9)


@pytest.mark.parametrize("name,data,level", STREAMS, ids=[s[0] for s in STREAMS])
def test_bz_compress_stream(ctx, oracle, name, data, level):
    z = ctx.bz_compress(data, level)
    assert z == oracle.bz_compress(data, level)
    assert bz2.decompress(z) == data


@pytest.mark.parametrize("idx,level", [(1, 1), (2, 2), (3, 3)])
def test_bz_compress_golden_samples(ctx, idx, level):
    gold = golden(f"sample{idx}.bz2")
    data = bz2.decompress(gold)
    assert ctx.bz_compress(data, level) == gold


def test_libbz2_shaped_front(ctx, oracle):
    """include/s3g_bzlib.h: the bz_stream layout of the reference's patched libbz2 (bz/bzlib.h:48-69) and its three compress
    calls (bz/bzlib.h:103-117) over the GPU compressor: input in pieces with BZ_RUN, BZ_FINISH into a small output buffer,
    the stream-end functor called once (bz/bzlib.c:470), the bytes those of libbz2."""
    import ctypes as C
    import starch3_b200 as s3
    L = s3.lib()

    class BzStream(C.Structure):
        _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_uint), ("total_in_lo32", C.c_uint), ("total_in_hi32", C.c_uint),
                    ("next_out", C.c_void_p), ("avail_out", C.c_uint), ("total_out_lo32", C.c_uint), ("total_out_hi32", C.c_uint),
                    ("state", C.c_void_p), ("bzalloc", C.c_void_p), ("bzfree", C.c_void_p), ("opaque", C.c_void_p),
                    ("handler", C.c_void_p), ("block_close_functor", C.c_void_p)]
    assert C.sizeof(BzStream) == 96          # the stock 80 bytes plus the two pointers of the patch
    FUNC = C.CFUNCTYPE(None, C.c_void_p)
    calls = []
    cb = FUNC(lambda h: calls.append(h))
    L.s3g_BZ2_bzCompressInit.argtypes = [C.POINTER(BzStream), C.c_int, C.c_int, C.c_int]
    L.s3g_BZ2_bzCompress.argtypes = [C.POINTER(BzStream), C.c_int]
    L.s3g_BZ2_bzCompressEnd.argtypes = [C.POINTER(BzStream)]
    tf, chroms, _ = oracle.transform(synth.bed(2, 60000).tobytes())
    data = tf[:1500000]
    for level, piece, outcap in ((9, 100000, 4096), (1, 333333, 1 << 20)):
        z = BzStream()
        assert L.s3g_BZ2_bzCompressInit(C.byref(z), level, 0, 30) == 0
        assert z.block_close_functor is None
        z.block_close_functor = C.cast(cb, C.c_void_p); z.handler = 0x1234
        src = C.create_string_buffer(data, len(data))
        out = C.create_string_buffer(outcap)
        got = bytearray()
        for off in range(0, len(data), piece):
            z.next_in = C.addressof(src) + off; z.avail_in = min(piece, len(data) - off)
            z.next_out = C.addressof(out); z.avail_out = outcap
            assert L.s3g_BZ2_bzCompress(C.byref(z), 0) == 1                  # BZ_RUN -> BZ_RUN_OK
            assert z.avail_in == 0
        assert L.s3g_BZ2_bzCompress(C.byref(z), 0) == -2                     # BZ_RUN without input: no progress (bz/bzlib.c:432)
        while True:
            z.next_out = C.addressof(out); z.avail_out = outcap
            rc = L.s3g_BZ2_bzCompress(C.byref(z), 2)                         # BZ_FINISH
            got += out.raw[:outcap - z.avail_out]
            if rc == 4:                                                      # BZ_STREAM_END
                break
            assert rc == 3                                                   # BZ_FINISH_OK
        assert L.s3g_BZ2_bzCompress(C.byref(z), 2) == -1                     # BZ_SEQUENCE_ERROR after the end
        assert z.total_in_lo32 == len(data) and z.total_out_lo32 == len(got)
        assert L.s3g_BZ2_bzCompressEnd(C.byref(z)) == 0 and z.state is None
        assert bytes(got) == (oracle.ref_bz_compress if oracle.have_ref() else oracle.bz_compress)(data, level)
    assert calls == [0x1234, 0x1234]
    bad = BzStream()
    assert L.s3g_BZ2_bzCompressInit(C.byref(bad), 0, 0, 30) == -2 and L.s3g_BZ2_bzCompress(C.byref(bad), 0) == -2
