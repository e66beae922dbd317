"""One archive from several GPUs: host-side partition and gather.

The path has no collective reduction: chromosomes are independent bzip2 streams.  Whole
chromosomes are dealt to ranks (longest first), every rank runs the single-GPU path on its
share, and rank 0 gathers the finished streams in archive order and writes the container
(ARCHIVE_FORMAT.md).  torch.distributed is used for the object gather only (gloo or nccl
process groups both work; no tensor collective touches the data path).

`compress_fn(bed_bytes, block_size_100k) -> list of dict(name, stream, lines, blocks, tf_len,
bases_nonunique, bases_unique)` is the per-rank compressor: `gpu_compress_fn(ctx)` in
production; the tests inject a CPU checker to exercise this host logic without a GPU.
"""
import numpy as np

MAGIC = bytes([0xca, 0x5c, 0xad, 0x1a])


def chrom_segments(bed):
    """[(name, byte_start, byte_end)] of the maximal runs of lines sharing their first field.
    A name that reappears later opens a new segment (starch3api.hpp:331)."""
    a = np.frombuffer(bed, dtype=np.uint8) if not isinstance(bed, np.ndarray) else bed
    nl = np.flatnonzero(a == 10)
    if len(nl) == 0:
        return []
    starts = np.concatenate(([0], nl[:-1] + 1))
    raw = a.tobytes() if isinstance(bed, np.ndarray) else bytes(bed)
    segs = []
    prev = None
    # candidates: lines whose first bytes differ from the previous line's are cheap to spot
    # with a vectorised compare of the first 8 bytes; the rest is confirmed exactly.
    pad = np.concatenate((a, np.zeros(16, dtype=np.uint8)))
    idx = starts[:, None] + np.arange(8)[None, :]
    head = pad[idx]
    tab = (head == 9)
    has_tab = tab.any(axis=1)
    first_tab = np.where(has_tab, tab.argmax(axis=1), 8)
    masked = np.where(np.arange(8)[None, :] < first_tab[:, None], head, 0)
    same_as_prev = np.ones(len(starts), dtype=bool)
    same_as_prev[0] = False
    same_as_prev[1:] = (masked[1:] == masked[:-1]).all(axis=1) & (first_tab[1:] == first_tab[:-1]) & has_tab[1:] & has_tab[:-1]
    for i in np.flatnonzero(~same_as_prev | ~has_tab):
        s = int(starts[i])
        name = raw[s:raw.index(b"\t", s)] if b"\t" in raw[s:int(nl[i])] else raw[s:int(nl[i])]
        if name != prev:
            segs.append([name, s, None])
            prev = name
    for k in range(len(segs)):
        segs[k][2] = segs[k + 1][1] if k + 1 < len(segs) else int(nl[-1]) + 1
    return [tuple(s) for s in segs]


def partition(sizes, world):
    """Longest-processing-time-first assignment of items to `world` ranks -> list of index lists."""
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        out[r].append(i)
        load[r] += sizes[i]
    for r in range(world):
        out[r].sort()
    return out


def _json_string(b):
    o = bytearray(b'"')
    short = {0x5c: b"\\\\", 0x22: b'\\"', 0x08: b"\\b", 0x0c: b"\\f", 0x0a: b"\\n", 0x0d: b"\\r", 0x09: b"\\t"}
    for c in b:
        o += short[c] if c in short else (b"\\u%04X" % c if c < 0x20 else bytes([c]))
    return bytes(o + b'"')


def build_archive(entries, block_size_100k=9, note=""):
    """entries in archive order -> archive bytes (same text as csrc/api.cu build_header)."""
    metas, off = [], 0
    for e in entries:
        metas.append(b'{"chromosome":' + _json_string(e["name"]) +
                     b',"offset":%d,"size":%d,"lines":%d,"blocks":%d,"transformedBytes":%d,"nonUniqueBases":%d,"uniqueBases":%d}'
                     % (off, len(e["stream"]), e["lines"], e["blocks"], e["tf_len"], e["bases_nonunique"], e["bases_unique"]))
        off += len(e["stream"])
    note_b = note.encode() if isinstance(note, str) else (note or b"")
    hdr = (b'{"archive":{"type":"starch","version":{"major":3,"minor":0,"revision":0},"creator":"starch3_b200",'
           b'"compression":"bzip2","blockSize100k":%d,"note":' % block_size_100k) + _json_string(note_b) + b'},"streams":[' + b",".join(metas) + b"]}"
    return MAGIC + hdr + b"\n" + b"".join(e["stream"] for e in entries)


def gpu_compress_fn(ctx):
    """Per-rank compressor over the C ABI (starch3_b200.Context)."""
    def fn(bed_bytes, block_size_100k):
        res = ctx.compress_bed(bed_bytes, block_size_100k)
        return [dict(name=c["name"], stream=res.stream(i), lines=c["line_count"], blocks=c["n_blocks"], tf_len=c["tf_len"],
                     bases_nonunique=c["bases_nonunique"], bases_unique=c["bases_unique"]) for i, c in enumerate(res.chroms)]
    return fn


def compress_sharded(bed, compress_fn, block_size_100k=9, note="", group=None):
    """Every rank holds (or can read) `bed`; returns the archive on rank 0, None elsewhere."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    raw = bytes(bed) if not isinstance(bed, (bytes, bytearray)) else bed
    segs = chrom_segments(raw)
    mine = partition([e - s for _, s, e in segs], world)[rank]
    # adjacent segments with equal names must not be fused by concatenation: compress them apart
    results = {}
    group_start = 0
    while group_start < len(mine):
        group_end = group_start + 1
        while group_end < len(mine) and segs[mine[group_end]][0] != segs[mine[group_end - 1]][0]:
            group_end += 1
        ids = mine[group_start:group_end]
        out = compress_fn(b"".join(raw[segs[i][1]:segs[i][2]] for i in ids), block_size_100k)
        assert len(out) == len(ids), (len(out), len(ids))
        for i, e in zip(ids, out):
            results[i] = e
        group_start = group_end
    if world == 1:
        gathered = [results]
    else:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(results, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for g in gathered:
        merged.update(g)
    assert sorted(merged) == list(range(len(segs)))
    return build_archive([merged[i] for i in range(len(segs))], block_size_100k, note)
