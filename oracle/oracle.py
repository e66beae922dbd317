"""ctypes loader for the CPU oracle (oracle/s3_oracle.c) and the reference-built
libbz2 harness (oracle/_ref/libs3ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline and
`--impl reference` legs and __graft_entry__.smoke().  The product package
starch3_b200 must never import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "libs3oracle.so")
_REF_SO = os.path.join(HERE, "_ref", "libs3ref.so")
_JANSSON_SO = os.path.join(HERE, "_ref", "libs3jansson.so")
REF_BINARY = os.path.join(HERE, "_ref", "starch3_ref")

u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(HERE, "s3_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
                               "-o", _ORACLE_SO, src])
    if os.path.isdir("/root/reference") and (force or not os.path.exists(_REF_SO) or not os.path.exists(_JANSSON_SO)):
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")])


class ChromInfo(C.Structure):
    _fields_ = [("name_off", C.c_uint64), ("name_len", C.c_uint32), ("pad", C.c_uint32),
                ("tf_off", C.c_uint64), ("tf_len", C.c_uint64), ("line_count", C.c_int64),
                ("bases_nonunique", C.c_int64), ("bases_unique", C.c_int64)]


class BlockDesc(C.Structure):
    _fields_ = [("in_start", C.c_uint64), ("in_end", C.c_uint64), ("nblock", C.c_uint32),
                ("crc", C.c_uint32), ("in_use", C.c_uint8 * 256)]


class HuffSel(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_selectors", C.c_int32), ("alpha", C.c_int32),
                ("len", (C.c_uint8 * 258) * 6), ("selector", C.c_uint8 * (18002 + 8))]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_ORACLE_SO)
        _lib.s3o_bz_compress.restype = C.c_int64
        _lib.s3o_bz_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]
        _lib.s3o_crc32.restype = C.c_uint32
        _lib.s3o_crc32.argtypes = [C.c_void_p, C.c_uint64]
        _lib.s3o_bwt.restype = C.c_int32
        _lib.s3o_bwt.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _lib.s3o_mtf.restype = C.c_int32
        _lib.s3o_mtf.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_int32)]
        _lib.s3o_huff_select.restype = None
        _lib.s3o_huff_select.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(HuffSel)]
        _lib.s3o_rle1_blocks.restype = C.c_int64
        _lib.s3o_rle1_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_uint64]
        _lib.s3o_transform.restype = C.c_int
        _lib.s3o_transform.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                       C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return _lib


def have_ref():
    return os.path.exists(_REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        _ref = C.CDLL(_REF_SO)
        _ref.s3ref_bz_compress.restype = C.c_int64
        _ref.s3ref_bz_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]
        _ref.s3ref_bz_decompress.restype = C.c_int64
        _ref.s3ref_bz_decompress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _ref.s3ref_block_sort.restype = C.c_int32
        _ref.s3ref_block_sort.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _ref.s3ref_mtf_huff.restype = C.c_int32
        _ref.s3ref_mtf_huff.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    return _ref


def _u8(a):
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- whole-stream bzip2 ----------------------------------------------------
def bz_compress(data, level=9):
    """Restated bzip2 stream compress (one BZ_FINISH feed)."""
    a = _u8(data)
    cap = len(a) + len(a) // 50 + 4096
    out = np.empty(cap, dtype=np.uint8)
    n = lib().s3o_bz_compress(_ptr(a), len(a), level, _ptr(out), cap)
    assert n >= 0, n
    return out[:n].tobytes()


def ref_bz_compress(data, level=9):
    """The reference's vendored libbz2 (streaming API + no-op functor)."""
    a = _u8(data)
    cap = len(a) + len(a) // 50 + 4096
    out = np.empty(cap, dtype=np.uint8)
    n = ref().s3ref_bz_compress(_ptr(a), len(a), level, _ptr(out), cap)
    assert n >= 0, n
    return out[:n].tobytes()


def ref_bz_decompress(data, cap):
    a = _u8(data)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = ref().s3ref_bz_decompress(_ptr(a), len(a), _ptr(out), cap)
    assert n >= 0, n
    return out[:n].tobytes()


def crc32(data):
    a = _u8(data)
    return lib().s3o_crc32(_ptr(a), len(a))


# ---- stage-wise ------------------------------------------------------------
def rle1_blocks(data, level=9):
    """-> (list of dict(in_start,in_end,nblock,crc,in_use ndarray), concatenated post-RLE1 bytes)"""
    a = _u8(data)
    cap = len(a) // (100000 * level - 19 - 260) + 4
    descs = (BlockDesc * cap)()
    rle = np.empty(len(a) + len(a) // 4 + 64, dtype=np.uint8)
    n = lib().s3o_rle1_blocks(_ptr(a), len(a), level, descs, cap, _ptr(rle), len(rle))
    assert n >= 0, n
    out = []
    for d in descs[:n]:
        out.append(dict(in_start=d.in_start, in_end=d.in_end, nblock=d.nblock, crc=d.crc,
                        in_use=np.frombuffer(bytes(d.in_use), dtype=np.uint8).copy()))
    tot = sum(d["nblock"] for d in out)
    return out, rle[:tot].copy()


def bwt(block):
    a = _u8(block)
    ptr = np.empty(len(a), dtype=np.uint32)
    orig = lib().s3o_bwt(_ptr(a), len(a), _ptr(ptr))
    return ptr, orig


def ref_bwt(block):
    a = _u8(block)
    ptr = np.empty(len(a), dtype=np.uint32)
    orig = ref().s3ref_block_sort(_ptr(a), len(a), _ptr(ptr))
    return ptr, orig


def mtf(block, ptr, in_use):
    a = _u8(block)
    ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
    iu = _u8(in_use)
    mtfv = np.empty(len(a) + 2, dtype=np.uint16)
    freq = np.zeros(258, dtype=np.int32)
    nu = C.c_int32(0)
    n = lib().s3o_mtf(_ptr(a), len(a), _ptr(ptr), _ptr(iu), _ptr(mtfv), _ptr(freq), C.byref(nu))
    return mtfv[:n].copy(), freq, nu.value


def huff_select(mtfv, freq, n_in_use):
    m = np.ascontiguousarray(mtfv, dtype=np.uint16)
    f = np.ascontiguousarray(freq, dtype=np.int32)
    h = HuffSel()
    lib().s3o_huff_select(_ptr(m), len(m), _ptr(f), n_in_use, C.byref(h))
    lens = np.frombuffer(bytes(h.len), dtype=np.uint8).reshape(6, 258).copy()
    sel = np.frombuffer(bytes(h.selector), dtype=np.uint8)[:h.n_selectors].copy()
    return dict(n_groups=h.n_groups, n_selectors=h.n_selectors, alpha=h.alpha, len=lens, selector=sel)


def ref_mtf_huff(block, ptr, in_use):
    a = _u8(block)
    ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
    iu = _u8(in_use)
    mtfv = np.empty(len(a) + 2, dtype=np.uint16)
    freq = np.zeros(258, dtype=np.int32)
    sel = np.zeros(18002 + 8, dtype=np.uint8)
    lens = np.zeros((6, 258), dtype=np.uint8)
    bits = np.zeros(len(a) * 2 + 4096, dtype=np.uint8)
    nbits = C.c_uint64(0)
    n = ref().s3ref_mtf_huff(_ptr(a), len(a), _ptr(ptr), _ptr(iu), _ptr(mtfv), _ptr(freq), _ptr(sel),
                             _ptr(lens), _ptr(bits), len(bits), C.byref(nbits))
    nsel = (n + 49) // 50
    return dict(mtfv=mtfv[:n].copy(), freq=freq, selector=sel[:nsel].copy(), len=lens,
                bits=bits[:(nbits.value + 7) // 8].copy(), nbits=nbits.value)


# ---- transform ---------------------------------------------------------------
def transform(bed):
    """-> (tf bytes, list of dict per chromosome, dropped_tail_bytes)"""
    a = _u8(bed)
    cap = 2 * len(a) + 4096
    tf = np.empty(cap, dtype=np.uint8)
    nlines = int(np.count_nonzero(a == 10))
    ccap = max(16, min(nlines, 1 << 22))
    chroms = (ChromInfo * ccap)()
    tf_len = C.c_uint64(0); nc = C.c_uint64(0); dropped = C.c_uint64(0)
    rc = lib().s3o_transform(_ptr(a), len(a), _ptr(tf), cap, C.byref(tf_len), chroms, ccap,
                             C.byref(nc), C.byref(dropped))
    if rc == -2:
        raise ValueError("malformed BED line (fewer than three fields)")
    assert rc == 0, rc
    raw = a.tobytes()
    out = []
    for c in chroms[:nc.value]:
        out.append(dict(name=raw[c.name_off:c.name_off + c.name_len], tf_off=c.tf_off, tf_len=c.tf_len,
                        line_count=c.line_count, bases_nonunique=c.bases_nonunique,
                        bases_unique=c.bases_unique))
    return tf[:tf_len.value].tobytes(), out, dropped.value


# ---- archive container (ARCHIVE_FORMAT.md; parity unpinned by the reference) -----------------
def _json_string(b: bytes) -> bytes:
    """jansson 2.9 dump_string without JSON_ESCAPE_SLASH / JSON_ENSURE_ASCII (src/dump.c:70-160)."""
    out = bytearray(b'"')
    short = {0x5c: b"\\\\", 0x22: b'\\"', 0x08: b"\\b", 0x0c: b"\\f", 0x0a: b"\\n", 0x0d: b"\\r", 0x09: b"\\t"}
    for c in b:
        if c in short:
            out += short[c]
        elif c < 0x20:
            out += b"\\u%04X" % c
        else:
            out.append(c)
    out += b'"'
    return bytes(out)


def have_jansson():
    return os.path.exists(_JANSSON_SO)


def jansson_header(level, note, streams):
    """The metadata object printed by the reference's vendored jansson 2.9 (oracle/jansson_harness.c):
    streams = [(name bytes, offset, size, lines, blocks, transformedBytes, nonUniqueBases, uniqueBases)].
    Returns None when jansson refuses the object (a name or note that is not valid UTF-8)."""
    L = C.CDLL(_JANSSON_SO)
    L.s3ref_archive_header.restype = C.c_int64
    L.s3ref_archive_header.argtypes = [C.c_int, C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_uint64]
    names = np.frombuffer(b"".join(s[0] for s in streams) + b"\0", dtype=np.uint8)
    lens = np.array([len(s[0]) for s in streams] + [0], dtype=np.uint32)
    fields = np.array([list(s[1:8]) for s in streams] + [[0] * 7], dtype=np.int64)
    note_b = note.encode() if isinstance(note, str) else (note or b"")
    cap = 4096 + len(note_b) * 6 + sum(len(s[0]) * 6 + 256 for s in streams)
    out = np.empty(cap, dtype=np.uint8)
    n = L.s3ref_archive_header(level, note_b, len(note_b), len(streams), _ptr(names), _ptr(lens), _ptr(fields), _ptr(out), cap)
    if n == -1:
        return None
    assert n >= 0, n
    return out[:n].tobytes()


def archive(bed, level=9, note=""):
    """The whole path on the CPU: restated transform, one bzip2 stream per chromosome
    (reference libbz2 when oracle/_ref is built, else the restatement), container."""
    tf, chroms, _ = transform(bed)
    comp = ref_bz_compress if have_ref() else bz_compress
    streams, metas, off = [], [], 0
    for c in chroms:
        s = tf[c["tf_off"]:c["tf_off"] + c["tf_len"]]
        z = comp(s, level)
        nblk = len([b for b in rle1_blocks(s, level)[0] if b["nblock"]])
        metas.append(b'{"chromosome":' + _json_string(c["name"]) +
                     b',"offset":%d,"size":%d,"lines":%d,"blocks":%d,"transformedBytes":%d,"nonUniqueBases":%d,"uniqueBases":%d}'
                     % (off, len(z), c["line_count"], nblk, c["tf_len"], c["bases_nonunique"], c["bases_unique"]))
        streams.append(z)
        off += len(z)
    note_b = note.encode() if isinstance(note, str) else (note or b"")
    hdr = (b'{"archive":{"type":"starch","version":{"major":3,"minor":0,"revision":0},"creator":"starch3_b200",'
           b'"compression":"bzip2","blockSize100k":%d,"note":' % level) + _json_string(note_b) + b'},"streams":[' + b",".join(metas) + b"]}"
    return bytes([0xca, 0x5c, 0xad, 0x1a]) + hdr + b"\n" + b"".join(streams)


# ---- the decoder path's checker --------------------------------------------------------------------------
def inverse_transform(name, tf):
    """ARCHIVE_FORMAT.md "Streams": a line p<L> sets the element length, any other line <d>[\t<rest>] is an element with
    start = previous stop + d, stop = start + L; both start at 0 (the inverse of starch3api.hpp:428-504)."""
    out = []
    prev_stop = 0
    cur_len = 0
    for ln in bytes(tf).split(b"\n")[:-1]:
        if ln[:1] == b"p":
            cur_len = int(ln[1:])
            continue
        f = ln.split(b"\t", 1)
        start = prev_stop + int(f[0])
        stop = start + cur_len
        out.append(name + b"\t%d\t%d" % (start, stop) + (b"\t" + f[1] if len(f) == 2 else b"") + b"\n")
        prev_stop = stop
    return b"".join(out)


def bz_decompress(z, cap=None):
    """One bzip2 stream -> bytes: the reference's vendored BZ2_bzDecompress when oracle/_ref is built, else CPython's bz2."""
    if have_ref():
        return ref_bz_decompress(z, cap if cap is not None else max(len(z) * 60, 1 << 20))
    import bz2
    return bz2.decompress(bytes(z))


def unarchive(arc):
    """archive -> BED text, on the CPU (ARCHIVE_FORMAT.md "Reading")."""
    import json
    arc = bytes(arc)
    assert arc[:4] == bytes([0xca, 0x5c, 0xad, 0x1a])
    nl = arc.index(b"\n", 4)
    meta = json.loads(arc[4:nl].decode("utf-8", "surrogateescape"))
    payload = arc[nl + 1:]
    out = []
    for st in meta["streams"]:
        z = payload[st["offset"]:st["offset"] + st["size"]]
        tf = bz_decompress(z, st["transformedBytes"] + 16)
        assert len(tf) == st["transformedBytes"]
        out.append(inverse_transform(st["chromosome"].encode("utf-8", "surrogateescape"), tf))
    return b"".join(out)


# ---- block-parallel form of the same reference compressor -------------------------------------------------
# libbz2 compresses one stream on one core; the BASELINE-size single-chromosome inputs (cfg1/3/4) would take
# minutes.  A bzip2 stream is "BZh<level>" + the blocks bit-concatenated + trailer (bz/compress.c:602-667), a block
# depends only on its own input range (blocks are whole RLE1 chunks, bz/bzlib.c:225-338), and the ranges come from the
# restated cut (s3o_rle1_blocks).  So: compress every range as a stream of its own with the reference libbz2 on a
# thread pool, lift the block bits out of each, and join them.  tests/test_oracle.py pins this against
# ref_bz_compress / bz_compress of the whole stream.
_TRAILER = 0x177245385090


def block_ranges(data, level=9):
    """-> [(in_start, in_end)] of the blocks the reference cuts `data` into (no RLE bytes returned)."""
    a = _u8(data)
    cap = len(a) // (100000 * level - 19 - 260) + 4
    descs = (BlockDesc * cap)()
    n = lib().s3o_rle1_blocks(_ptr(a), len(a), level, descs, cap, None, 0)
    assert n >= 0, n
    return [(d.in_start, d.in_end) for d in descs[:n] if d.nblock]


def _block_bits(z):
    """(block bits as int, number of bits, block crc) of a ONE-block bzip2 stream z."""
    v = int.from_bytes(z, "big")
    total = len(z) * 8
    for pad in range(8):                      # the trailer is followed by 0..7 padding bits
        if (v >> (pad + 32)) & ((1 << 48) - 1) == _TRAILER:
            nbits = total - pad - 80 - 32     # minus trailer, minus "BZh9"
            bits = (v >> (pad + 80)) & ((1 << nbits) - 1)
            crc = (bits >> (nbits - 80)) & 0xffffffff     # after the 48-bit block magic
            return bits, nbits, crc
    raise AssertionError("no stream trailer found")


def bz_compress_blockwise(data, level=9, threads=None):
    """Same bytes as ref_bz_compress(data, level) (bz_compress when oracle/_ref is absent), block-parallel."""
    from concurrent.futures import ThreadPoolExecutor
    a = _u8(data)
    comp = ref_bz_compress if have_ref() else bz_compress
    ranges = block_ranges(a, level)
    if len(ranges) <= 1:
        return comp(a, level)
    threads = threads or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(lambda r: _block_bits(comp(a[r[0]:r[1]], level)), ranges))
    acc = int.from_bytes(b"BZh" + bytes([48 + level]), "big")
    nacc = 32
    comb = 0
    chunks = []
    for bits, nbits, crc in parts:
        comb = (((comb << 1) | (comb >> 31)) & 0xffffffff) ^ crc          # bz/compress.c:607-608
        acc = (acc << nbits) | bits
        nacc += nbits
        if nacc >= 1 << 23:                    # flush whole bytes so the big integer stays small
            keep = nacc & 7
            chunks.append((acc >> keep).to_bytes((nacc - keep) // 8, "big"))
            acc &= (1 << keep) - 1
            nacc = keep
    acc = (acc << 80) | (_TRAILER << 32) | comb
    nacc += 80
    pad = (-nacc) & 7
    chunks.append((acc << pad).to_bytes((nacc + pad) // 8, "big"))
    return b"".join(chunks)


def archive_mt(bed, level=9, note="", threads=None):
    """oracle.archive(bed, level, note) computed with all host threads: chromosomes and, inside a chromosome,
    bzip2 blocks in parallel.  Used at the BASELINE sizes (10 M .. 100 M lines)."""
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or min(32, os.cpu_count() or 1)
    tf, chroms, _ = transform(bed)
    tfa = np.frombuffer(tf, dtype=np.uint8)
    comp = ref_bz_compress if have_ref() else bz_compress
    # work units: (chromosome, block range); a one-block chromosome is compressed whole
    units = []
    per_chrom = []
    for ci, c in enumerate(chroms):
        s = tfa[c["tf_off"]:c["tf_off"] + c["tf_len"]]
        rng = block_ranges(s, level)
        per_chrom.append((s, rng))
        if len(rng) <= 1:
            units.append((ci, None))
        else:
            units.extend((ci, r) for r in rng)

    def work(u):
        ci, r = u
        s = per_chrom[ci][0]
        return comp(s, level) if r is None else _block_bits(comp(s[r[0]:r[1]], level))

    with ThreadPoolExecutor(max_workers=threads) as ex:
        done = list(ex.map(work, units))
    streams, metas, off, k = [], [], 0, 0
    for ci, c in enumerate(chroms):
        rng = per_chrom[ci][1]
        if len(rng) <= 1:
            z = done[k]; k += 1
        else:
            acc = int.from_bytes(b"BZh" + bytes([48 + level]), "big"); nacc = 32; comb = 0; chunks = []
            for bits, nbits, crc in done[k:k + len(rng)]:
                comb = (((comb << 1) | (comb >> 31)) & 0xffffffff) ^ crc
                acc = (acc << nbits) | bits; nacc += nbits
                if nacc >= 1 << 23:
                    keep = nacc & 7
                    chunks.append((acc >> keep).to_bytes((nacc - keep) // 8, "big"))
                    acc &= (1 << keep) - 1; nacc = keep
            k += len(rng)
            acc = (acc << 80) | (_TRAILER << 32) | comb; nacc += 80
            pad = (-nacc) & 7
            chunks.append((acc << pad).to_bytes((nacc + pad) // 8, "big"))
            z = b"".join(chunks)
        metas.append(b'{"chromosome":' + _json_string(c["name"]) +
                     b',"offset":%d,"size":%d,"lines":%d,"blocks":%d,"transformedBytes":%d,"nonUniqueBases":%d,"uniqueBases":%d}'
                     % (off, len(z), c["line_count"], len(rng), c["tf_len"], c["bases_nonunique"], c["bases_unique"]))
        streams.append(z)
        off += len(z)
    note_b = note.encode() if isinstance(note, str) else (note or b"")
    hdr = (b'{"archive":{"type":"starch","version":{"major":3,"minor":0,"revision":0},"creator":"starch3_b200",'
           b'"compression":"bzip2","blockSize100k":%d,"note":' % level) + _json_string(note_b) + b'},"streams":[' + b",".join(metas) + b"]}"
    return bytes([0xca, 0x5c, 0xad, 0x1a]) + hdr + b"\n" + b"".join(streams)
