"""One archive from several GPUs, one process per GPU.

Host-side orchestration of the phases of include/starch3_b200.h ("one archive from several GPUs", csrc/shard.cu)
with torch.distributed for the exchanges between them:

    rank r: newline-aligned byte range r of the input, plus the line before it (the halo)
      s3g_shard_tokenize    -> summary            | all_gather: 5 numbers per rank
      s3g_shard_transform   <- carried maximum    | all_gather: piece tables, transformed bytes (NVLink)
      s3g_shard_plan        (every rank, same plan)
      s3g_shard_compress    blocks [b_lo, b_hi)   | all_reduce: bit length + CRC of every block
      s3g_shard_assemble    -> bytes [lo, hi) of the concatenated streams
    rank 0: gathers the byte strings, ORs the seam bytes, writes magic + metadata in front (ARCHIVE_FORMAT.md).

The archive is the same bytes as the 1-GPU archive.  `Phases` is the GPU implementation over the C ABI; the tests
drive the same orchestration with a CPU checker in its place (tests/test_multigpu.py, gloo, world size 2 and 3).
"""
import os

import numpy as np

MAGIC = bytes([0xca, 0x5c, 0xad, 0x1a])
I64_MIN = -(1 << 63)
NAME_WORDS = 8                      # chromosome names travel in the piece table: 64 bytes each
ROW = 8 + NAME_WORDS                # int64 words per piece
OPT_PIECES = 64                     # pieces per rank the first exchange has room for


# ---------------------------------------------------------------------------------------------------------
# host logic (pure functions; unit-tested on the CPU)
# ---------------------------------------------------------------------------------------------------------
def plan_ranges(bed, world):
    """Cut points at line starts, and for every rank the length of the line before its range (0: none).
    bed: numpy uint8 array.  -> (cut[world + 1], halo[world])"""
    n = len(bed)
    cut = [0]
    for r in range(1, world):
        p = max(cut[-1], (r * n) // world)
        while True:                                  # first newline at or after p - 1
            if p >= n:
                p = n
                break
            if p == 0 or bed[p - 1] == 10:
                break
            w = np.flatnonzero(bed[p:p + 65536] == 10)
            if len(w):
                p = p + int(w[0]) + 1
                break
            p = min(n, p + 65536)
        cut.append(p)
    cut.append(n)
    halo = []
    for r in range(world):
        c = cut[r]
        if c == 0:
            halo.append(0)
            continue
        lo, span = c - 1, 4096                       # the newline that ends the halo line is bed[c - 1]
        while True:
            a = max(0, lo - span)
            w = np.flatnonzero(bed[a:lo] == 10)
            if len(w):
                halo.append(c - (a + int(w[-1]) + 1))
                break
            if a == 0:
                halo.append(c)
                break
            span *= 16
    return cut, halo


def carry_chain(summaries):
    """summaries[r] = (tail_max, continues, single_piece, n_lines).  -> carry_max per rank: the largest stop of all
    earlier lines of the chromosome the rank's range continues (I64_MIN if it continues none)."""
    out = []
    run = I64_MIN                                    # largest stop of the open chromosome over the ranks so far
    for tail_max, continues, single, n_lines in summaries:
        c = run if continues else I64_MIN
        out.append(c)
        if n_lines == 0:
            continue                                 # an empty range hands on what it got
        run = max(tail_max, c) if single and continues else tail_max
    return out


def merge_pieces(per_rank):
    """per_rank[r] = (continues, [piece dicts with name, tf_len, line_count, bases_nonunique, bases_unique]) in rank
    order.  A rank's first piece that continues the previous rank's last chromosome is the same stream.
    -> streams: list of dicts (name, tf_off, tf_len, line_count, bases_nonunique, bases_unique)"""
    streams = []
    off = 0
    for continues, pieces in per_rank:
        for k, p in enumerate(pieces):
            if k == 0 and continues and streams:
                s = streams[-1]
                s["tf_len"] += p["tf_len"]; s["line_count"] += p["line_count"]
                s["bases_nonunique"] += p["bases_nonunique"]; s["bases_unique"] += p["bases_unique"]
            else:
                streams.append(dict(name=p["name"], tf_off=off, tf_len=p["tf_len"], line_count=p["line_count"],
                                    bases_nonunique=p["bases_nonunique"], bases_unique=p["bases_unique"]))
            off += p["tf_len"]
    return streams


def block_shares(nblock, world):
    """Contiguous shares of the block list, balanced by bytes, no share larger than ceil(blocks / world) -> bounds[world + 1].
    (MTF and Huffman are one CTA -- or one cluster -- per block: a share of 149 blocks takes a 148-SM GPU as long as one of 296.)"""
    nb = len(nblock)
    cum = np.concatenate(([0], np.cumsum(nblock.astype(np.int64))))
    total = int(cum[-1])
    cap = (nb + world - 1) // world
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cum, (total * r + world // 2) // world, side="left"))
        b = min(b, bounds[-1] + cap)                 # this share within the cap ...
        b = max(b, nb - (world - r) * cap)           # ... and the shares after it too
        bounds.append(min(max(b, bounds[-1]), nb))
    bounds.append(nb)
    return bounds


def _json_string(b):
    o = bytearray(b'"')
    short = {0x5c: b"\\\\", 0x22: b'\\"', 0x08: b"\\b", 0x0c: b"\\f", 0x0a: b"\\n", 0x0d: b"\\r", 0x09: b"\\t"}
    for c in b:
        o += short[c] if c in short else (b"\\u%04X" % c if c < 0x20 else bytes([c]))
    return bytes(o + b'"')


def build_header(streams, blocks_of, stream_off, stream_len, level, note):
    """magic + metadata + line feed (ARCHIVE_FORMAT.md; the text csrc/api.cu build_header writes)."""
    metas = []
    for i, s in enumerate(streams):
        metas.append(b'{"chromosome":' + _json_string(s["name"]) +
                     b',"offset":%d,"size":%d,"lines":%d,"blocks":%d,"transformedBytes":%d,"nonUniqueBases":%d,"uniqueBases":%d}'
                     % (int(stream_off[i]), int(stream_len[i]), s["line_count"], int(blocks_of[i]), s["tf_len"], s["bases_nonunique"],
                        s["bases_unique"]))
    note_b = note.encode() if isinstance(note, str) else (note or b"")
    hdr = (b'{"archive":{"type":"starch","version":{"major":3,"minor":0,"revision":0},"creator":"starch3_b200",'
           b'"compression":"bzip2","blockSize100k":%d,"note":' % level) + _json_string(note_b) + b'},"streams":[' + b",".join(metas) + b"]}"
    return MAGIC + hdr + b"\n"


def pack_pieces(pieces, names, cap):
    """piece table -> int64 array [cap, ROW]; row = tf_len, line_count, nonunique, unique, name_len, valid, 0, 0, name[64]"""
    t = np.zeros((cap, ROW), dtype=np.int64)
    for k, (p, nm) in enumerate(zip(pieces[:cap], names[:cap])):
        t[k, :6] = (p["tf_len"], p["line_count"], p["bases_nonunique"], p["bases_unique"], len(nm), 1)
        t[k, 8:].view(np.uint8)[:min(len(nm), 8 * NAME_WORDS)] = np.frombuffer(nm[:8 * NAME_WORDS], dtype=np.uint8)
    return t


def unpack_pieces(t, count):
    out = []
    for k in range(count):
        r = t[k]
        nl = int(r[4])
        out.append(dict(tf_len=int(r[0]), line_count=int(r[1]), bases_nonunique=int(r[2]), bases_unique=int(r[3]),
                        name=r[8:].view(np.uint8)[:min(nl, 8 * NAME_WORDS)].tobytes(), name_len=nl))
    return out


# ---------------------------------------------------------------------------------------------------------
# the GPU phases over the C ABI
# ---------------------------------------------------------------------------------------------------------
class _DevView:
    """Zero-copy torch view of library-owned device memory."""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (max(n, 0),), "typestr": "|u1", "data": (ptr, False), "version": 2}


class Phases:
    """The phases on one GPU (a starch3_b200.Context); tensors are torch uint8 CUDA tensors."""
    def __init__(self, ctx, device):
        import torch
        self.ctx, self.torch, self.device = ctx, torch, device

    def view(self, ptr, n):
        if n == 0 or not ptr:
            return self.torch.empty(0, dtype=self.torch.uint8, device=self.device)
        return self.torch.as_tensor(_DevView(ptr, n), device=self.device)

    def tokenize(self, d_range, n, halo):
        return self.ctx.shard_tokenize(d_range.data_ptr(), n, halo)

    def transform(self, carry):
        pieces, ptr, n = self.ctx.shard_transform(carry)
        return pieces, self.view(ptr, n)

    # ---- the transform with its all-gather fused in (NVLink peer stores) ----
    def peer_buffer(self, dist, nbytes):
        """A buffer of at least nbytes that every rank of the group can store into: torch symmetric memory (the ranks'
        allocations mapped into each other's address space over NVLink).  Collective: every rank calls it with the same
        nbytes.  -> (local uint8 tensor, [device address of rank r's buffer, as seen from this GPU]) or None when the
        platform has no symmetric memory (the orchestration then exchanges the bytes with NCCL)."""
        if getattr(self, "_symm_off", False):
            return None
        if getattr(self, "_symm_cap", 0) >= nbytes:
            return self._symm_t, self._symm_ptrs
        try:
            import torch.distributed._symmetric_memory as symm
            cap = int(nbytes + nbytes // 4 + (1 << 20)) & ~255
            # second half: where rank 0 collects the finished byte strings (a bzip2 stream is never much larger than its input)
            t = symm.empty(2 * cap + (1 << 20), dtype=self.torch.uint8, device=self.device)
            hdl = symm.rendezvous(t, dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != dist.get_world_size() or any(p == 0 or p % 16 for p in ptrs):
                raise RuntimeError("unexpected symmetric-memory pointers")
            self._symm_t, self._symm_hdl, self._symm_ptrs, self._symm_cap = t, hdl, ptrs, cap
            mc = 0
            try:
                mc = int(hdl.multicast_ptr or 0)      # NVLS: one store, replicated by the switch into every rank's buffer
            except Exception:
                mc = 0
            self._symm_mc = 0 if (mc % 16 or os.environ.get("S3G_NO_MULTICAST")) else mc
            return t, ptrs
        except Exception as e:                     # no NVLink peer mapping here: fall back, once, loudly
            import sys
            print(f"[starch3_b200.multigpu] symmetric memory unavailable ({type(e).__name__}: {e}); the transformed bytes go through NCCL", file=sys.stderr)
            self._symm_off = True
            return None

    def place(self, gather_ptr, lo, hi):
        self.ctx.shard_place(gather_ptr, lo, hi)

    def transform_into(self, carry, peer_ptrs, rank, dst_off):
        order = [peer_ptrs[rank]] + [p for r, p in enumerate(peer_ptrs) if r != rank]      # this GPU's own buffer first
        return self.ctx.shard_transform_peers(carry, order, dst_off, multicast_buf=getattr(self, "_symm_mc", 0))

    def plan(self, tf_all, tf_total, soff, level):
        return self.ctx.shard_plan(tf_all.data_ptr(), tf_total, soff, level)

    def compress(self, b_lo, b_hi):
        return self.ctx.shard_compress(b_lo, b_hi)

    def assemble(self, n_bits_all, crc_all, b_lo, b_hi, n_streams):
        ptr, lo, hi, so, sl = self.ctx.shard_assemble(n_bits_all, crc_all, b_lo, b_hi, n_streams)
        return self.view(ptr, hi - lo), lo, hi, so, sl


# ---------------------------------------------------------------------------------------------------------
# orchestration (every rank runs this; rank 0 returns the streams)
# ---------------------------------------------------------------------------------------------------------
def compress_sharded(ph, dist, rank, world, d_range, n_range, halo, names_src, range_base, level, device, torch):
    """d_range: device tensor holding the rank's range (halo line first).  names_src: host bytes-like the piece names
    are read from (index 0 = first byte of d_range ... via range_base = global offset of d_range[0]).
    -> dict(streams table, n_blocks, rle/mtf totals ...) and, on rank 0, `payload`: device tensor with the
    concatenated bzip2 streams."""
    i64 = torch.int64
    import os, time
    _t = [] if os.environ.get("S3G_SHARD_TIMING") else None      # wall clock per phase (host view, rank 0 prints)
    def _mark(name):
        if _t is not None:
            if device.type == "cuda":
                torch.cuda.synchronize()
            _t.append((name, time.perf_counter()))
    _mark("start")
    # ---- phase 1: tokenizer; exchange the summaries ----
    sm = ph.tokenize(d_range, n_range, halo)
    _mark("tokenize")
    mine = torch.tensor([sm["tail_max"], sm["continues"], sm["single_piece"], sm["n_lines"], sm["dropped_tail_bytes"], sm.get("tf_bytes", -1)],
                        dtype=i64, device=device)
    allsm = torch.empty(world * 6, dtype=i64, device=device)
    dist.all_gather_into_tensor(allsm, mine)
    allsm = allsm.cpu().numpy().reshape(world, 6)
    carries = carry_chain([tuple(int(x) for x in allsm[r, :4]) for r in range(world)])
    _mark("x summaries")
    # ---- phase 2: transform; exchange piece tables and transformed bytes ----
    # With NVLink peer memory the exchange of the bytes is part of the transform kernel: every rank knows from the summaries
    # where its piece goes, and stores it into every rank's copy of the transformed buffer (s3g_shard_transform_peers).
    # The exchange of the piece tables below is then also the point after which all copies are complete: a rank takes part
    # in it only after its transform kernel has finished.  (The exchange of the summaries in the NEXT call is what keeps a
    # fast rank from overwriting a buffer that a slow rank still reads.)
    peer = None
    if world > 1 and hasattr(ph, "peer_buffer") and dist.get_backend() == "nccl" and int(allsm[:, 5].min()) >= 0 and not os.environ.get("S3G_NO_PEER_STORES"):
        peer = ph.peer_buffer(dist, int(allsm[:, 5].sum()) + 64)
    gather = None
    if peer is not None:
        tf_all, peer_ptrs = peer
        gather_at = ph._symm_cap                       # offset of the gather region inside the symmetric buffer
        gather = tf_all[gather_at:]
        if rank == 0:
            # zero before anybody places a string: the others get here only after the block-table exchange below, which
            # this rank enters after this memset has run (s3g_shard_compress synchronises)
            gather[:min(int(gather.numel()), int(allsm[:, 5].sum()) + (1 << 20))].zero_()
        pieces, my_tf_len = ph.transform_into(carries[rank], peer_ptrs, rank, int(allsm[:rank, 5].sum()))
        tf = None
    else:
        pieces, tf = ph.transform(carries[rank])
        my_tf_len = int(tf.numel())
    _mark("transform")
    names = [bytes(names_src[p["name_off"] + range_base:p["name_off"] + range_base + p["name_len"]]) for p in pieces]
    cap = OPT_PIECES
    while True:
        tab = np.zeros((cap + 1, ROW), dtype=np.int64)
        tab[0, 0], tab[0, 1] = len(pieces), my_tf_len
        tab[1:] = pack_pieces(pieces, names, cap)
        alltab = torch.empty(world * (cap + 1) * ROW, dtype=i64, device=device)
        dist.all_gather_into_tensor(alltab, torch.from_numpy(tab.reshape(-1)).to(device))
        alltab = alltab.cpu().numpy().reshape(world, cap + 1, ROW)
        need = int(alltab[:, 0, 0].max())
        if need <= cap:
            break
        cap = need                                  # more pieces than the first exchange has room for: once more, larger
    if any(int(alltab[r, 1 + k, 4]) > 8 * NAME_WORDS for r in range(world) for k in range(int(alltab[r, 0, 0]))):
        raise ValueError("chromosome names longer than %d bytes are not supported by the multi-GPU exchange" % (8 * NAME_WORDS))
    per_rank = [(int(allsm[r, 1]), unpack_pieces(alltab[r, 1:], int(alltab[r, 0, 0]))) for r in range(world)]
    tf_lens = [int(alltab[r, 0, 1]) for r in range(world)]
    streams = merge_pieces(per_rank)
    _mark("x pieces")
    tf_total = sum(tf_lens)
    offs = np.concatenate(([0], np.cumsum(tf_lens))).astype(np.int64)
    if peer is not None:
        if tf_lens != [int(x) for x in allsm[:, 5]]:
            raise RuntimeError("transformed sizes differ from the measured ones")
    else:
        tf_all = torch.empty(tf_total + 64, dtype=torch.uint8, device=device)
    if peer is not None:
        pass                                         # the bytes are already in place
    elif world > 1 and dist.get_backend() == "nccl":
        # every rank sends its piece to every other rank and receives theirs straight into place: one group of
        # point-to-point transfers over NVLink (the list form of all_gather takes a broadcast per rank, 1.1 ms for 257 MB
        # at two ranks against 0.3 ms this way)
        mine_t = tf[:tf_lens[rank]].contiguous()
        tf_all[int(offs[rank]):int(offs[rank + 1])].copy_(mine_t)
        ops = []
        for d in range(1, world):
            to, frm = (rank + d) % world, (rank - d) % world
            if tf_lens[rank]:
                ops.append(dist.P2POp(dist.isend, mine_t, to))
            if tf_lens[frm]:
                ops.append(dist.P2POp(dist.irecv, tf_all[int(offs[frm]):int(offs[frm + 1])], frm))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    elif world > 1:
        # other backends (the CPU tests run gloo): equal-sized slots, then into place
        slot = max(tf_lens) + 1
        mine_p = torch.zeros(slot, dtype=torch.uint8, device=device)
        mine_p[:tf_lens[rank]] = tf[:tf_lens[rank]]
        slots = torch.empty(world * slot, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(slots, mine_p)
        for r in range(world):
            tf_all[int(offs[r]):int(offs[r + 1])] = slots[r * slot:r * slot + tf_lens[r]]
    else:
        tf_all[:tf_total] = tf
    soff = np.array([s["tf_off"] for s in streams] + [tf_total], dtype=np.uint64)
    _mark("x transformed bytes")
    # ---- phase 3: the block plan (every rank computes the same one) ----
    nblock, stream_of = ph.plan(tf_all, tf_total, soff, level)
    _mark("plan")
    nb = len(nblock)
    bounds = block_shares(nblock, world)
    b_lo, b_hi = bounds[rank], bounds[rank + 1]
    # ---- phase 4: the rank's share of the blocks; exchange bit lengths and CRCs ----
    n_bits, crc, n_mtf = ph.compress(b_lo, b_hi)
    _mark("compress")
    tab = np.zeros((nb + 1, 2), dtype=np.int64)
    tab[b_lo:b_hi, 0] = n_bits.astype(np.int64); tab[b_lo:b_hi, 1] = crc.astype(np.int64)
    tab[nb, 0] = int(n_mtf.sum())
    t = torch.from_numpy(tab.reshape(-1)).to(device)
    if world > 1:
        dist.all_reduce(t)
    tab = t.cpu().numpy().reshape(nb + 1, 2)
    n_bits_all = tab[:nb, 0].astype(np.uint64); crc_all = tab[:nb, 1].astype(np.uint32)
    _mark("x block table")
    # ---- phase 5: place the share; gather on rank 0 ----
    piece, lo, hi, stream_off, stream_len = ph.assemble(n_bits_all, crc_all, b_lo, b_hi, len(streams))
    _mark("assemble")
    total = int(stream_off[-1] + stream_len[-1]) if len(streams) else 0
    payload = None
    use_gather = world > 1 and gather is not None and total + 8 <= int(gather.numel())
    if not use_gather:
        span = torch.tensor([lo, hi], dtype=i64, device=device)
        spans = torch.empty(world * 2, dtype=i64, device=device)
        dist.all_gather_into_tensor(spans, span)
        spans = spans.cpu().numpy().reshape(world, 2)
    if use_gather:
        # every rank stores its string into rank 0's gather region through the NVLink-mapped pointer (the end bytes ORed
        # in): no spans to exchange, no receives, one small all-reduce as the "everybody has stored" point
        if hi > lo:
            ph.place(peer_ptrs[0] + gather_at, lo, hi)
        done = torch.zeros(1, dtype=i64, device=device)
        dist.all_reduce(done)                        # a rank contributes only after its place kernel has finished
        done.cpu()
        if rank == 0:
            payload = gather[:total + 8]
    elif world > 1:
        if rank == 0:
            payload = torch.zeros(total + 8, dtype=torch.uint8, device=device)
            payload[lo:hi] = piece[:hi - lo]
            for r in range(1, world):
                rl, rh = int(spans[r, 0]), int(spans[r, 1])
                if rh > rl:
                    buf = torch.empty(rh - rl, dtype=torch.uint8, device=device)
                    dist.recv(buf, src=r)
                    # the two end bytes may be shared with the neighbours: OR them, copy the rest
                    payload[rl:rl + 1] |= buf[:1]
                    if rh - rl > 1:
                        payload[rh - 1:rh] |= buf[-1:]
                    if rh - rl > 2:
                        payload[rl + 1:rh - 1] = buf[1:-1]
        elif hi > lo:
            dist.send(piece[:hi - lo].contiguous(), dst=0)
    else:
        payload = piece[:total]
    _mark("x byte strings")
    if _t is not None and rank == 0:
        print(f"[shard] peer stores {'on' if peer is not None else 'off'}, multicast {'on' if getattr(ph, '_symm_mc', 0) else 'off'}, gather by stores {'on' if use_gather else 'off'}", flush=True)
        print("[shard timing, ms] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(_t, _t[1:])), flush=True)
    blocks_of = np.bincount(stream_of, minlength=len(streams)) if nb else np.zeros(len(streams), dtype=np.int64)
    return dict(streams=streams, blocks_of=blocks_of, stream_off=stream_off, stream_len=stream_len, payload=payload, total=total,
                n_blocks=nb, n_lines=int(allsm[:, 3].sum()), dropped=int(allsm[world - 1, 4]), tf_total=tf_total,
                rle_bytes=int(nblock.astype(np.int64).sum()), mtf_symbols=int(tab[nb, 0]), share=(b_lo, b_hi))


def compress_bed(ctx, bed, level=9, note="", device=None, resident=None):
    """The user-facing call, one process per GPU (torch.distributed initialised): every rank holds the input
    `bed` (numpy uint8; a file every rank can read), uploads only its range, and rank 0 returns the archive bytes
    (None elsewhere).  `resident`: a device tensor that already holds the rank's range (bench.py's device-resident
    timing); otherwise the range is uploaded here."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    device = device or torch.device("cuda", torch.cuda.current_device())
    cut, halo = plan_ranges(bed, world)
    lo, hi = cut[rank] - halo[rank], cut[rank + 1]
    if resident is None:
        resident = torch.from_numpy(bed[lo:hi]).to(device, non_blocking=True) if hi > lo else torch.empty(16, dtype=torch.uint8, device=device)
    ph = Phases(ctx, device)
    out = compress_sharded(ph, dist, rank, world, resident, hi - lo, halo[rank], bed, lo, level, device, torch)
    if rank != 0:
        return None, out
    hdr = build_header(out["streams"], out["blocks_of"], out["stream_off"], out["stream_len"], level, note)
    arc = torch.empty(len(hdr) + out["total"], dtype=torch.uint8).pin_memory() if out["total"] else torch.empty(len(hdr), dtype=torch.uint8)
    arc[:len(hdr)] = torch.frombuffer(bytearray(hdr), dtype=torch.uint8)
    if out["total"]:
        arc[len(hdr):].copy_(out["payload"][:out["total"]])
    return arc.numpy(), out


# ---------------------------------------------------------------------------------------------------------
# one process, several contexts: the same orchestration with threads instead of ranks
# ---------------------------------------------------------------------------------------------------------
class LocalGroup:
    """The handful of torch.distributed calls compress_sharded makes, between the threads of ONE process: rank r is
    thread r with its own context (on its own device, or several contexts on one device).  Tensors are exchanged by
    device-to-device copies (peer copies over NVLink between devices)."""
    def __init__(self, world):
        import threading
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world
        self.mail = {}
        self.cv = threading.Condition()

    def view(self, rank):
        g = self

        class _D:
            @staticmethod
            def get_backend():
                return "local"

            @staticmethod
            def _exchange(t):
                if t.is_cuda:
                    import torch
                    torch.cuda.current_stream(t.device).synchronize()
                g.slots[rank] = t
                g.bar.wait()
                got = list(g.slots)
                return got

            @staticmethod
            def _done():
                g.bar.wait()

            @staticmethod
            def all_gather_into_tensor(out, t):
                got = _D._exchange(t)
                n = t.numel()
                for r, src in enumerate(got):
                    out.view(-1)[r * n:(r + 1) * n].copy_(src.view(-1))
                _D._sync(out)
                _D._done()

            @staticmethod
            def all_gather(outs, t):
                got = _D._exchange(t)
                for o, src in zip(outs, got):
                    o.copy_(src)
                _D._sync(outs[0])
                _D._done()

            @staticmethod
            def all_reduce(t):
                got = _D._exchange(t.clone())
                acc = None
                for src in got:
                    v = src.to(t.device)
                    acc = v if acc is None else acc + v
                t.copy_(acc)
                _D._sync(t)
                _D._done()

            @staticmethod
            def _sync(t):
                if t.is_cuda:
                    import torch
                    torch.cuda.current_stream(t.device).synchronize()

            @staticmethod
            def send(t, dst):
                _D._sync(t)
                with g.cv:
                    g.mail[(rank, dst)] = t
                    g.cv.notify_all()
                    g.cv.wait_for(lambda: (rank, dst) not in g.mail)      # until the receiver has copied it

            @staticmethod
            def recv(buf, src):
                with g.cv:
                    g.cv.wait_for(lambda: (src, rank) in g.mail)
                    t = g.mail[(src, rank)]
                buf.copy_(t)
                _D._sync(buf)
                with g.cv:
                    del g.mail[(src, rank)]
                    g.cv.notify_all()
        return _D


def compress_bed_local(ctxs, devices, bed, level=9, note=""):
    """One process, one context per entry of `devices` (a device may appear more than once): the ranges go to the
    contexts, the archive (bytes) comes back.  The in-process counterpart of compress_bed."""
    import threading
    import torch
    world = len(ctxs)
    group = LocalGroup(world)
    cut, halo = plan_ranges(bed, world)
    results = [None] * world
    errors = []

    def work(r):
        try:
            dev = torch.device("cuda", devices[r])
            torch.cuda.set_device(dev)
            stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(stream):
                ctxs[r].set_stream(stream.cuda_stream)
                lo, hi = cut[r] - halo[r], cut[r + 1]
                d_range = torch.from_numpy(bed[lo:hi]).to(dev) if hi > lo else torch.empty(16, dtype=torch.uint8, device=dev)
                results[r] = compress_sharded(Phases(ctxs[r], dev), group.view(r), r, world, d_range, hi - lo, halo[r], bed, lo, level, dev, torch)
                stream.synchronize()
                ctxs[r].set_stream(0)
        except BaseException as e:          # a dead thread must not leave the others at the barrier
            errors.append(e)
            group.bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errors:
        raise errors[0]
    out = results[0]
    hdr = build_header(out["streams"], out["blocks_of"], out["stream_off"], out["stream_len"], level, note)
    return hdr + bytes(out["payload"][:out["total"]].cpu().numpy()), results
