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
REF_BINARY = os.path.join(HERE, "_ref", "starch3_ref")

u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(HERE, "s3_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
                               "-o", _ORACLE_SO, src])
    if os.path.isdir("/root/reference") and (force or not os.path.exists(_REF_SO)):
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
