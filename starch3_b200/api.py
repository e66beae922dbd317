"""ctypes binding of include/starch3_b200.h.

Python mirror of the reference's operator surface for this path: a `Context` plays the
role of the `starch3::Starch` object's compression stream (starch3api.hpp:771-888) and
`Context.compress_bed` the role of the produce_line/consume_line/process_tf_buffer
pipeline (starch3api.hpp:158-407).  Errors raise `Starch3Error` carrying the C-ABI code;
the reference prints "Error: ..." and exits with an errno-style code
(starch3api.hpp:175-176, :841-848) -- the CLI wrapper maps the codes back.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(HERE, "libstarch3_b200.so")

S3G_OK, S3G_E_CUDA, S3G_E_PARAM, S3G_E_NOMEM, S3G_E_MALFORMED, S3G_E_CAPACITY, S3G_E_LIMIT = 0, -1, -2, -3, -4, -5, -6

C_ABI_SYMBOLS = [
    "s3g_init", "s3g_destroy", "s3g_last_error", "s3g_set_stream", "s3g_launch_count", "s3g_last_host_entry", "s3g_sort_retries", "s3g_sort_stats", "s3g_profile", "s3g_profile_report", "s3g_profile_filter",
    "s3g_compress_bed", "s3g_compress_bed_device", "s3g_result_free", "s3g_read_streams",
    "s3g_stream_begin", "s3g_stream_write", "s3g_stream_end",
    "s3g_shard_tokenize", "s3g_shard_transform", "s3g_shard_transform_peers", "s3g_shard_plan", "s3g_shard_compress", "s3g_shard_assemble", "s3g_shard_place", "s3g_multi_compress_bed", "s3g_stage_times", "s3g_batch_chunks", "s3g_chain_layout",
    "s3g_tokenize", "s3g_transform", "s3g_rle1", "s3g_bwt", "s3g_mtf", "s3g_huff", "s3g_bz_compress",
    "s3g_decompress_archive", "s3g_bz_decompress", "s3g_inverse_transform",
    "s3g_BZ2_bzCompressInit", "s3g_BZ2_bzCompress", "s3g_BZ2_bzCompressEnd",
]


class Starch3Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"s3g error {code}: {msg}")
        self.code = code


class CChrom(C.Structure):
    _fields_ = [("name_off", C.c_uint64), ("name_len", C.c_uint32), ("n_blocks", C.c_uint32),
                ("tf_off", C.c_uint64), ("tf_len", C.c_uint64), ("line_count", C.c_int64),
                ("bases_nonunique", C.c_int64), ("bases_unique", C.c_int64),
                ("bz_off", C.c_uint64), ("bz_len", C.c_uint64)]


class CResult(C.Structure):
    _fields_ = [("archive", C.POINTER(C.c_uint8)), ("archive_size", C.c_uint64), ("streams_off", C.c_uint64),
                ("chroms", C.POINTER(CChrom)), ("n_chroms", C.c_uint64), ("n_lines", C.c_uint64),
                ("n_blocks", C.c_uint64), ("tf_bytes", C.c_uint64), ("dropped_tail_bytes", C.c_uint64),
                ("d_streams", C.c_void_p), ("streams_size", C.c_uint64), ("device_ms", C.c_double),
                ("rle_bytes", C.c_uint64), ("mtf_symbols", C.c_uint64), ("stage_ms", C.c_double * 8),
                ("unsorted_lines", C.c_uint64), ("crlf_lines", C.c_uint64), ("reappearing_chroms", C.c_uint64)]


STAGE_NAMES = ["tokenise+transform", "rle1+cut+crc", "blocksort", "mtf", "huffman", "assemble"]


class CShardSummary(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("tail_max", C.c_int64), ("continues", C.c_uint32), ("single_piece", C.c_uint32),
                ("dropped_tail_bytes", C.c_uint64), ("tf_bytes", C.c_uint64)]


class CDecodeInfo(C.Structure):
    _fields_ = [("n_streams", C.c_uint64), ("n_blocks", C.c_uint64), ("tf_bytes", C.c_uint64), ("device_ms", C.c_double),
                ("d_bed", C.c_void_p)]


class CBlockDesc(C.Structure):
    _fields_ = [("in_start", C.c_uint64), ("in_end", C.c_uint64), ("nblock", C.c_uint32), ("crc", C.c_uint32),
                ("in_use", C.c_uint8 * 256)]


def have_library():
    return os.path.exists(lib_path)


def build_library(verbose=False):
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], stdout=out)


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not have_library():
            raise Starch3Error(S3G_E_CUDA, f"{lib_path} is missing: build it with `make -C starch3_b200/csrc` "
                                           "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(lib_path)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.s3g_init.argtypes = [i32, C.POINTER(vp)]
        L.s3g_destroy.argtypes = [vp]; L.s3g_destroy.restype = None
        L.s3g_last_error.restype = C.c_char_p
        L.s3g_set_stream.argtypes = [vp, vp]
        L.s3g_launch_count.argtypes = [vp]; L.s3g_launch_count.restype = u64
        L.s3g_sort_retries.argtypes = [vp]; L.s3g_sort_retries.restype = u64
        L.s3g_last_host_entry.argtypes = [vp]; L.s3g_last_host_entry.restype = i32
        L.s3g_batch_chunks.argtypes = [u64, i32, vp, vp, vp, u64, C.POINTER(u64)]
        L.s3g_chain_layout.argtypes = [vp, i32, u64, i32, vp, vp, vp, u64, u64, vp, vp, u64, C.POINTER(u64), vp, vp]
        L.s3g_sort_stats.argtypes = [vp, vp]
        L.s3g_profile.argtypes = [vp, i32]
        L.s3g_profile_report.argtypes = [vp, C.c_char_p, u64]
        L.s3g_profile_filter.argtypes = [vp, C.c_char_p]
        L.s3g_compress_bed.argtypes = [vp, vp, u64, i32, C.c_char_p, C.POINTER(CResult)]
        L.s3g_compress_bed_device.argtypes = [vp, vp, u64, i32, C.c_char_p, i32, C.POINTER(CResult)]
        L.s3g_result_free.argtypes = [C.POINTER(CResult)]; L.s3g_result_free.restype = None
        L.s3g_read_streams.argtypes = [vp, vp, u64, C.POINTER(u64)]
        L.s3g_stream_begin.argtypes = [vp, i32, C.c_char_p, u64]
        L.s3g_stream_write.argtypes = [vp, vp, u64]
        L.s3g_stream_end.argtypes = [vp, C.POINTER(CResult)]
        L.s3g_shard_tokenize.argtypes = [vp, vp, u64, u64, C.POINTER(CShardSummary)]
        L.s3g_shard_transform.argtypes = [vp, C.c_int64, vp, u64, C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
        L.s3g_shard_transform_peers.argtypes = [vp, C.c_int64, vp, u64, C.POINTER(u64), vp, C.c_uint32, u64, u64, C.POINTER(u64)]
        L.s3g_shard_plan.argtypes = [vp, vp, u64, vp, u64, i32, C.POINTER(u64), vp, vp, u64]
        L.s3g_shard_compress.argtypes = [vp, u64, u64, vp, vp, vp]
        L.s3g_shard_assemble.argtypes = [vp, vp, vp, u64, u64, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64), vp, vp]
        L.s3g_shard_place.argtypes = [vp, u64, u64, u64]
        L.s3g_multi_compress_bed.argtypes = [vp, i32, vp, u64, i32, C.c_char_p, C.POINTER(CResult)]
        L.s3g_stage_times.argtypes = [vp, vp]
        L.s3g_tokenize.argtypes = [vp, vp, u64, u64, C.POINTER(u64), vp, vp, vp, vp, vp]
        L.s3g_transform.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), vp, u64, C.POINTER(u64), C.POINTER(u64)]
        L.s3g_rle1.argtypes = [vp, vp, u64, i32, vp, u64, C.POINTER(u64), vp, u64]
        L.s3g_bwt.argtypes = [vp, vp, vp, u64, vp, vp]
        L.s3g_mtf.argtypes = [vp, vp, C.c_uint32, vp, vp, vp, C.POINTER(C.c_uint32), vp]
        L.s3g_huff.argtypes = [vp, vp, C.c_uint32, vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, vp, vp, u64,
                               C.POINTER(u64)]
        L.s3g_bz_compress.argtypes = [vp, vp, u64, i32, vp, u64, C.POINTER(u64)]
        L.s3g_decompress_archive.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(CDecodeInfo)]
        L.s3g_bz_decompress.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64)]
        L.s3g_inverse_transform.argtypes = [vp, vp, u64, vp, C.c_uint32, vp, u64, C.POINTER(u64)]
        _lib = L
    return _lib


def _u8(a):
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Result:
    """Host-side view of an s3g_result (copied out; the C result is freed)."""

    def __init__(self, cres, bed_bytes=None, keep_archive=True):
        self.n_lines = cres.n_lines
        self.n_blocks = cres.n_blocks
        self.tf_bytes = cres.tf_bytes
        self.dropped_tail_bytes = cres.dropped_tail_bytes
        self.device_ms = cres.device_ms
        self.rle_bytes = cres.rle_bytes
        self.mtf_symbols = cres.mtf_symbols
        self.stage_ms = {nm: cres.stage_ms[i] for i, nm in enumerate(STAGE_NAMES)}
        self.unsorted_lines = cres.unsorted_lines
        self.crlf_lines = cres.crlf_lines
        self.reappearing_chroms = cres.reappearing_chroms
        self.streams_size = cres.streams_size
        self.streams_off = cres.streams_off
        self.d_streams = cres.d_streams
        # zero-copy view of the context-owned archive buffer (valid until the next compress call);
        # `.archive` materialises an independent bytes object on first use
        self.archive_size = cres.archive_size if keep_archive else 0
        self._archive_bytes = None
        self.archive_view = None
        if keep_archive and cres.archive_size:
            self.archive_view = memoryview((C.c_uint8 * cres.archive_size).from_address(C.addressof(cres.archive.contents))).cast("B")
        self.chroms = []
        for i in range(cres.n_chroms):
            c = cres.chroms[i]
            d = {k: getattr(c, k) for k, _ in CChrom._fields_}
            if bed_bytes is not None:
                d["name"] = bytes(bed_bytes[c.name_off:c.name_off + c.name_len])
            self.chroms.append(d)

    @property
    def archive(self):
        if self._archive_bytes is None and self.archive_view is not None:
            self._archive_bytes = bytes(self.archive_view)
        return self._archive_bytes

    def stream(self, i):
        """bzip2 stream of chromosome i (bytes), from the archive."""
        c = self.chroms[i]
        o = self.streams_off + c["bz_off"]
        return self.archive[o:o + c["bz_len"]]


def chain_layout(state, level, n_streams, first_continues, stream_of, n_bits, crc, n_final):
    """one step of the chained entries' bit layout (s3g_chain_layout; host logic only): state = np.uint64[4], updated in place
    -> (block positions, [(bit position, word)], stream starts, stream lengths)"""
    stream_of = np.ascontiguousarray(stream_of, dtype=np.uint32); n_bits = np.ascontiguousarray(n_bits, dtype=np.uint64)
    crc = np.ascontiguousarray(crc, dtype=np.uint32)
    nb = len(stream_of)
    pos = np.zeros(max(n_final, 1), dtype=np.uint64); cap = 4 * n_streams + 4
    patch = np.zeros(2 * cap, dtype=np.uint64); npatch = C.c_uint64(0)
    start = np.zeros(max(n_streams, 1), dtype=np.uint64); ln = np.zeros(max(n_streams, 1), dtype=np.uint64)
    rc = lib().s3g_chain_layout(_p(state), level, n_streams, 1 if first_continues else 0, _p(stream_of), _p(n_bits), _p(crc), nb, n_final,
                                _p(pos), _p(patch), cap, C.byref(npatch), _p(start), _p(ln))
    if rc != S3G_OK:
        raise Starch3Error(rc, lib().s3g_last_error().decode("utf-8", "replace"))
    return (pos[:n_final].copy(), [(int(patch[2 * k]), int(patch[2 * k + 1])) for k in range(npatch.value)], start[:n_streams].copy(),
            ln[:n_streams].copy())


def batch_chunks(n_blocks, stage):
    """how a batch of n_blocks bzip2 blocks is dealt to the MTF (stage 3) / Huffman (stage 4) kernels: [(first, count, ctas per block)];
    host logic only, no device needed"""
    cap = 64
    first = (C.c_uint64 * cap)(); count = (C.c_uint64 * cap)(); ctas = (C.c_uint32 * cap)(); n = C.c_uint64(0)
    rc = lib().s3g_batch_chunks(n_blocks, stage, first, count, ctas, cap, C.byref(n))
    if rc != S3G_OK:
        raise Starch3Error(rc, lib().s3g_last_error().decode("utf-8", "replace"))
    return [(int(first[i]), int(count[i]), int(ctas[i])) for i in range(n.value)]


def multi_compress_bed(contexts, bed, block_size_100k=9, note=None):
    """One archive from several GPUs in this process (s3g_multi_compress_bed): `contexts` = one Context per GPU."""
    a = _u8(bed)
    hs = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    r = CResult()
    rc = lib().s3g_multi_compress_bed(hs, len(contexts), _p(a), len(a), block_size_100k, note.encode() if isinstance(note, str) else note, C.byref(r))
    try:
        contexts[0]._check(rc)
        res = Result(r, None)
        res.archive
        return res
    finally:
        lib().s3g_result_free(C.byref(r))


class Context:
    """One device context (s3g_ctx).  Single caller, one CUDA stream."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._lib = lib()
        self._check(self._lib.s3g_init(device, C.byref(self._h)))

    def _check(self, rc):
        if rc != S3G_OK:
            raise Starch3Error(rc, self._lib.s3g_last_error().decode("utf-8", "replace"))

    def close(self):
        if self._h:
            self._lib.s3g_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        self._check(self._lib.s3g_set_stream(self._h, C.c_void_p(cuda_stream)))

    @property
    def launch_count(self):
        return self._lib.s3g_launch_count(self._h)

    @property
    def last_host_entry(self):
        """0: one piece; n > 0: n ranges by chromosome; n < 0: -n ranges chained by block (s3g_last_host_entry)"""
        return self._lib.s3g_last_host_entry(self._h)

    @property
    def sort_retries(self):
        return self._lib.s3g_sort_retries(self._h)

    @property
    def sort_stats(self):
        """(blocks sorted by the bucket form, batches in which it handed blocks back to the radix form, radix retries)"""
        out = (C.c_uint64 * 3)()
        self._check(self._lib.s3g_sort_stats(self._h, out))
        return tuple(out)

    def profile(self, enable=True):
        self._check(self._lib.s3g_profile(self._h, 1 if enable else 0))

    def profile_filter(self, kernel_name=None):
        """Time only launches of `kernel_name` (None = all kernels)."""
        self._check(self._lib.s3g_profile_filter(self._h, kernel_name.encode() if kernel_name else None))

    def profile_report(self):
        """-> {kernel name: (launches, total_ms, algorithmic_bytes)} since profiling was enabled / last report."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self._lib.s3g_profile_report(self._h, buf, len(buf)))
        out = {}
        for ln in buf.value.decode().splitlines():
            name, cnt, ms, by = ln.split("\t")
            out[name] = (int(cnt), float(ms), float(by))
        return out

    # ---- whole path ----
    def compress_bed(self, bed, block_size_100k=9, note=None, lazy=False):
        a = _u8(bed)
        r = CResult()
        rc = self._lib.s3g_compress_bed(self._h, _p(a), len(a), block_size_100k,
                                        note.encode() if isinstance(note, str) else note, C.byref(r))
        try:
            self._check(rc)
            res = Result(r, a)
            if not lazy:
                res.archive          # copy out now: the view dies with the next call
            return res
        finally:
            self._lib.s3g_result_free(C.byref(r))

    def compress_stream(self, pieces, block_size_100k=9, note=None, range_bytes=0):
        """bounded-memory ingestion: `pieces` is an iterable of bytes-like chunks of the BED text (any sizes)"""
        self._check(self._lib.s3g_stream_begin(self._h, block_size_100k, note.encode() if isinstance(note, str) else note, range_bytes))
        for pc in pieces:
            a = _u8(pc)
            self._check(self._lib.s3g_stream_write(self._h, _p(a), len(a)))
        r = CResult()
        rc = self._lib.s3g_stream_end(self._h, C.byref(r))
        try:
            self._check(rc)
            res = Result(r, None)
            res.archive
            return res
        finally:
            self._lib.s3g_result_free(C.byref(r))

    def compress_bed_device(self, dptr, n, block_size_100k=9, note=None, want_archive=False, bed_bytes=None):
        r = CResult()
        rc = self._lib.s3g_compress_bed_device(self._h, C.c_void_p(dptr), n, block_size_100k,
                                               note.encode() if isinstance(note, str) else note,
                                               1 if want_archive else 0, C.byref(r))
        try:
            self._check(rc)
            res = Result(r, bed_bytes, keep_archive=want_archive)
            res.archive
            return res
        finally:
            self._lib.s3g_result_free(C.byref(r))

    def read_streams(self, size):
        """Host copy of the device-resident bzip2 streams left by the last compress call."""
        out = np.empty(max(1, size), dtype=np.uint8)
        n = C.c_uint64(0)
        self._check(self._lib.s3g_read_streams(self._h, _p(out), size, C.byref(n)))
        return out[:n.value].tobytes()

    # ---- one archive from several GPUs: the phases (include/starch3_b200.h, csrc/shard.cu) ----
    def shard_tokenize(self, d_range, n, halo_bytes):
        out = CShardSummary()
        self._check(self._lib.s3g_shard_tokenize(self._h, C.c_void_p(d_range), n, halo_bytes, C.byref(out)))
        return dict(n_lines=out.n_lines, tail_max=out.tail_max, continues=out.continues, single_piece=out.single_piece,
                    dropped_tail_bytes=out.dropped_tail_bytes, tf_bytes=out.tf_bytes)

    def shard_transform_peers(self, carry_max, peer_bufs, dst_off, cap=4096, multicast_buf=0):
        """the transform with its all-gather fused in: the bytes go to dst_off of every buffer in peer_bufs (device
        addresses, this GPU's own first).  -> (pieces, transformed bytes written)"""
        pb = np.ascontiguousarray(peer_bufs, dtype=np.uint64)
        while True:
            pieces = (CChrom * cap)()
            n = C.c_uint64(0); tl = C.c_uint64(0)
            rc = self._lib.s3g_shard_transform_peers(self._h, carry_max, pieces, cap, C.byref(n), _p(pb), len(pb), multicast_buf, dst_off, C.byref(tl))
            if rc == S3G_E_CAPACITY and cap < (1 << 24):
                cap *= 16
                continue
            self._check(rc)
            break
        return [{k: getattr(pieces[i], k) for k, _ in CChrom._fields_} for i in range(n.value)], tl.value

    def shard_transform(self, carry_max, cap=4096):
        """-> (pieces: list of dict, device pointer of the transformed bytes, their length)"""
        while True:
            pieces = (CChrom * cap)()
            n = C.c_uint64(0); d_tf = C.c_void_p(); tl = C.c_uint64(0)
            rc = self._lib.s3g_shard_transform(self._h, carry_max, pieces, cap, C.byref(n), C.byref(d_tf), C.byref(tl))
            if rc == S3G_E_CAPACITY and cap < (1 << 24):
                cap *= 16
                continue
            self._check(rc)
            break
        out = [{k: getattr(pieces[i], k) for k, _ in CChrom._fields_} for i in range(n.value)]
        return out, d_tf.value or 0, tl.value

    def shard_plan(self, d_tf_all, tf_total, soff, level=9):
        """-> (nblock[], stream_of[]) of every block, in archive order"""
        soff = np.ascontiguousarray(soff, dtype=np.uint64)
        n_streams = len(soff) - 1
        cap = int(tf_total // (100000 * level - 19 - 260)) + n_streams + 8
        nblock = np.zeros(cap, dtype=np.uint32); sof = np.zeros(cap, dtype=np.uint32)
        nb = C.c_uint64(0)
        self._check(self._lib.s3g_shard_plan(self._h, C.c_void_p(d_tf_all), tf_total, _p(soff), n_streams, level, C.byref(nb),
                                             _p(nblock), _p(sof), cap))
        return nblock[:nb.value].copy(), sof[:nb.value].copy()

    def shard_compress(self, b_lo, b_hi):
        k = max(1, b_hi - b_lo)
        n_bits = np.zeros(k, dtype=np.uint64); crc = np.zeros(k, dtype=np.uint32); n_mtf = np.zeros(k, dtype=np.uint32)
        self._check(self._lib.s3g_shard_compress(self._h, b_lo, b_hi, _p(n_bits), _p(crc), _p(n_mtf)))
        return n_bits[:b_hi - b_lo], crc[:b_hi - b_lo], n_mtf[:b_hi - b_lo]

    def shard_assemble(self, n_bits_all, crc_all, b_lo, b_hi, n_streams):
        nb = np.ascontiguousarray(n_bits_all, dtype=np.uint64); cr = np.ascontiguousarray(crc_all, dtype=np.uint32)
        so = np.zeros(max(1, n_streams), dtype=np.uint64); sl = np.zeros(max(1, n_streams), dtype=np.uint64)
        d = C.c_void_p(); lo = C.c_uint64(0); hi = C.c_uint64(0)
        self._check(self._lib.s3g_shard_assemble(self._h, _p(nb), _p(cr), b_lo, b_hi, C.byref(d), C.byref(lo), C.byref(hi), _p(so), _p(sl)))
        return d.value or 0, lo.value, hi.value, so[:n_streams], sl[:n_streams]

    def shard_place(self, gather_buf, lo, hi):
        self._check(self._lib.s3g_shard_place(self._h, gather_buf, lo, hi))

    def stage_times(self):
        t = (C.c_double * 8)()
        self._check(self._lib.s3g_stage_times(self._h, t))
        return {nm: t[i] for i, nm in enumerate(STAGE_NAMES)}

    # ---- stages ----
    def tokenize(self, bed):
        a = _u8(bed)
        cap = int(np.count_nonzero(a == 10)) + 1
        n = C.c_uint64(0)
        ls = np.zeros(cap + 1, dtype=np.uint64); st = np.zeros(cap, dtype=np.int64); sp = np.zeros(cap, dtype=np.int64)
        ro = np.zeros(cap, dtype=np.uint32); cc = np.zeros(cap, dtype=np.uint8)
        self._check(self._lib.s3g_tokenize(self._h, _p(a), len(a), cap, C.byref(n), _p(ls), _p(st), _p(sp), _p(ro), _p(cc)))
        m = n.value
        return dict(n_lines=m, line_start=ls[:m + 1], start=st[:m], stop=sp[:m], rem_off=ro[:m], chrom_change=cc[:m])

    def transform(self, bed):
        a = _u8(bed)
        cap = 2 * len(a) + 4096
        tf = np.empty(cap, dtype=np.uint8)
        ccap = int(np.count_nonzero(a == 10)) + 1
        ccap = min(ccap, 1 << 22)
        chroms = (CChrom * ccap)()
        tl = C.c_uint64(0); nc = C.c_uint64(0); dr = C.c_uint64(0)
        self._check(self._lib.s3g_transform(self._h, _p(a), len(a), _p(tf), cap, C.byref(tl), chroms, ccap, C.byref(nc), C.byref(dr)))
        raw = a.tobytes()
        out = []
        for c in chroms[:nc.value]:
            out.append(dict(name=raw[c.name_off:c.name_off + c.name_len], tf_off=c.tf_off, tf_len=c.tf_len,
                            line_count=c.line_count, bases_nonunique=c.bases_nonunique, bases_unique=c.bases_unique))
        return tf[:tl.value].tobytes(), out, dr.value

    def rle1(self, data, block_size_100k=9):
        a = _u8(data)
        cap = len(a) // (100000 * block_size_100k - 19 - 260) + 4
        descs = (CBlockDesc * cap)()
        rle = np.empty(len(a) + len(a) // 4 + 64, dtype=np.uint8)
        nb = C.c_uint64(0)
        self._check(self._lib.s3g_rle1(self._h, _p(a), len(a), block_size_100k, descs, cap, C.byref(nb), _p(rle), len(rle)))
        out = []
        for d in descs[:nb.value]:
            out.append(dict(in_start=d.in_start, in_end=d.in_end, nblock=d.nblock, crc=d.crc,
                            in_use=np.frombuffer(bytes(d.in_use), dtype=np.uint8).copy()))
        tot = sum(d["nblock"] for d in out)
        return out, rle[:tot].copy()

    def bwt(self, blocks):
        """blocks: list of bytes-like (each <= 900000).  -> list of (ptr ndarray, origPtr)."""
        arrs = [_u8(b) for b in blocks]
        off = np.zeros(len(arrs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(x) for x in arrs])
        cat = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
        ptr = np.empty(int(off[-1]), dtype=np.uint32)
        orig = np.empty(len(arrs), dtype=np.int32)
        self._check(self._lib.s3g_bwt(self._h, _p(cat), _p(off), len(arrs), _p(ptr), _p(orig)))
        return [(ptr[int(off[i]):int(off[i + 1])].copy(), int(orig[i])) for i in range(len(arrs))]

    def mtf(self, block, ptr, in_use):
        a = _u8(block)
        ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
        iu = _u8(in_use)
        mtfv = np.empty(len(a) + 2, dtype=np.uint16)
        freq = np.zeros(258, dtype=np.int32)
        n = C.c_uint32(0)
        self._check(self._lib.s3g_mtf(self._h, _p(a), len(a), _p(ptr), _p(iu), _p(mtfv), C.byref(n), _p(freq)))
        return mtfv[:n.value].copy(), freq

    def huff(self, mtfv, freq, in_use):
        m = np.ascontiguousarray(mtfv, dtype=np.uint16)
        f = np.ascontiguousarray(freq, dtype=np.int32)
        iu = _u8(in_use)
        ng = C.c_int32(0); ns = C.c_int32(0); nbits = C.c_uint64(0)
        sel = np.zeros(18004, dtype=np.uint8)
        lens = np.zeros((6, 258), dtype=np.uint8)
        bits = np.zeros(len(m) * 3 + 65536, dtype=np.uint8)
        self._check(self._lib.s3g_huff(self._h, _p(m), len(m), _p(f), _p(iu), C.byref(ng), C.byref(ns), _p(sel), _p(lens),
                                       _p(bits), len(bits), C.byref(nbits)))
        return dict(n_groups=ng.value, n_selectors=ns.value, selector=sel[:ns.value].copy(), len=lens,
                    bits=bits[:(nbits.value + 7) // 8].copy(), nbits=nbits.value)

    # ---- the decoder path ----
    def decompress_archive(self, archive, want_bed=True):
        """archive bytes -> (BED bytes or None, info dict); the whole decoder on the GPU"""
        a = _u8(archive)
        n = C.c_uint64(0)
        info = CDecodeInfo()
        self._check(self._lib.s3g_decompress_archive(self._h, _p(a), len(a), None, 0, C.byref(n), C.byref(info)))
        d = dict(n_streams=info.n_streams, n_blocks=info.n_blocks, tf_bytes=info.tf_bytes, device_ms=info.device_ms, bed_len=n.value)
        if not want_bed:
            return None, d
        out = np.empty(max(n.value, 1), dtype=np.uint8)
        self._check(self._lib.s3g_decompress_archive(self._h, _p(a), len(a), _p(out), len(out), C.byref(n), C.byref(info)))
        d["device_ms"] = info.device_ms
        return out[:n.value].tobytes(), d

    def bz_decompress(self, data, cap=None):
        a = _u8(data)
        cap = cap or (len(a) * 8 + (1 << 20))
        n = C.c_uint64(0)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            rc = self._lib.s3g_bz_decompress(self._h, _p(a), len(a), _p(out), cap, C.byref(n))
            if rc == S3G_E_CAPACITY and n.value > cap:
                cap = n.value
                continue
            self._check(rc)
            return out[:n.value].tobytes()

    def inverse_transform(self, tf, name):
        a = _u8(tf)
        nm = _u8(name)
        cap = 2 * len(a) + (len(nm) + 48) * (int(np.count_nonzero(a == 10)) + 1)
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_uint64(0)
        self._check(self._lib.s3g_inverse_transform(self._h, _p(a), len(a), _p(nm), len(nm), _p(out), cap, C.byref(n)))
        return out[:n.value].tobytes()

    def bz_compress(self, data, block_size_100k=9):
        a = _u8(data)
        cap = len(a) + len(a) // 3 + 65536
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_uint64(0)
        self._check(self._lib.s3g_bz_compress(self._h, _p(a), len(a), block_size_100k, _p(out), cap, C.byref(n)))
        return out[:n.value].tobytes()
