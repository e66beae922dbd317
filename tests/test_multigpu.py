"""CPU tests of the multi-GPU host logic (starch3_b200/multigpu.py): range cuts and halo lines, the carried maximum,
piece merging, block shares, the exchange of tables and bytes over torch.distributed (gloo, world size 2 and 3) and the
seam-byte join -- with a CPU checker standing in for the five GPU phases (this tests the orchestration, not the
kernels; the GPU version is tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from starch3_b200 import multigpu as M, synth

I64_MIN = -(1 << 63)


class CpuPhases:
    """The five phases of csrc/shard.cu restated on the CPU with the oracle (test double)."""
    def __init__(self, oracle):
        self.O = oracle

    def tokenize(self, d_range, n, halo):
        raw = bytes(d_range[:n].numpy())
        self.halo = 1 if halo else 0
        self.lines = []
        pos = 0
        while True:
            e = raw.find(b"\n", pos)
            if e < 0:
                break
            f = raw[pos:e].split(b"\t", 3)
            self.lines.append((f[0], int(f[1]), int(f[2]), f[3] if len(f) == 4 else b"", pos))
            pos = e + 1
        flags = [i == 0 or self.lines[i][0] != self.lines[i - 1][0] for i in range(len(self.lines))]
        last = max([i for i, f in enumerate(flags) if f], default=0)
        lo = 1 if (self.halo and last == 0) else last
        stops = [l[2] for l in self.lines[lo:]]
        return dict(n_lines=len(self.lines) - self.halo, tail_max=max(stops) if stops else I64_MIN,
                    continues=1 if (self.halo and len(self.lines) > 1 and not flags[1]) else 0,
                    single_piece=1 if last == 0 else 0, dropped_tail_bytes=(n - pos) if self.lines else n)

    def transform(self, carry):
        out = bytearray()
        pieces = []
        prev_chr = None
        prev_stop = prev_len = 0
        runmax = I64_MIN
        cur = None
        for i, (ch, s, t, rem, off) in enumerate(self.lines):
            if i == 0 and self.halo:
                prev_chr, prev_stop, prev_len, runmax = ch, t, t - s, carry
                cur = dict(name_off=off, name_len=len(ch), tf_off=0, tf_len=0, line_count=0, bases_nonunique=0, bases_unique=0)
                pieces.append(cur)
                continue
            if ch != prev_chr:
                prev_stop = prev_len = 0
                runmax = I64_MIN
                cur = dict(name_off=off, name_len=len(ch), tf_off=len(out), tf_len=0, line_count=0, bases_nonunique=0, bases_unique=0)
                pieces.append(cur)
                prev_chr = ch
            ln = t - s
            o = b""
            if ln != prev_len:
                o += b"p%d\n" % ln
            o += b"%d" % (s - prev_stop) + (b"\t" + rem if rem else b"") + b"\n"
            out += o
            cur["tf_len"] += len(o); cur["line_count"] += 1; cur["bases_nonunique"] += ln
            cur["bases_unique"] += max(0, t - max(s, runmax))
            runmax = max(runmax, t)
            prev_stop, prev_len = t, ln
        if self.halo and pieces and pieces[0]["line_count"] == 0:
            pieces.pop(0)
        return pieces, torch.frombuffer(bytearray(bytes(out) or b"\0"), dtype=torch.uint8)[:len(out)]

    def plan(self, tf_all, tf_total, soff, level):
        self.tf = bytes(tf_all[:tf_total].numpy())
        self.level = level
        self.ranges, nblock, sof = [], [], []
        for s in range(len(soff) - 1):
            stream = self.tf[int(soff[s]):int(soff[s + 1])]
            for d in self.O.rle1_blocks(stream, level)[0]:
                if d["nblock"]:
                    self.ranges.append((int(soff[s]) + d["in_start"], int(soff[s]) + d["in_end"]))
                    nblock.append(d["nblock"]); sof.append(s)
        return np.array(nblock, dtype=np.uint32), np.array(sof, dtype=np.uint32)

    def compress(self, b_lo, b_hi):
        comp = self.O.ref_bz_compress if self.O.have_ref() else self.O.bz_compress
        self.bits = {b: self.O._block_bits(comp(self.tf[self.ranges[b][0]:self.ranges[b][1]], self.level)) for b in range(b_lo, b_hi)}
        k = b_hi - b_lo
        return (np.array([self.bits[b][1] for b in range(b_lo, b_hi)], dtype=np.uint64).reshape(k),
                np.array([self.bits[b][2] for b in range(b_lo, b_hi)], dtype=np.uint32).reshape(k), np.zeros(k, dtype=np.uint32))

    def assemble(self, n_bits_all, crc_all, b_lo, b_hi, n_streams):
        sof = []
        for s in range(n_streams):
            pass
        # layout of every stream from the full table
        stream_of = self._stream_of
        gpos, so, sl, first, comb = [0] * len(n_bits_all), [], [], {}, {}
        byte_off, b = 0, 0
        for s in range(n_streams):
            first[s] = b
            bits, c = 32, 0
            while b < len(n_bits_all) and stream_of[b] == s:
                gpos[b] = byte_off * 8 + bits
                bits += int(n_bits_all[b])
                c = (((c << 1) | (c >> 31)) & 0xffffffff) ^ int(crc_all[b])
                b += 1
            bits += 80
            so.append(byte_off); sl.append((bits + 7) // 8); comb[s] = c
            byte_off += sl[-1]
        first[n_streams] = len(n_bits_all)
        if b_lo >= b_hi:
            return torch.empty(0, dtype=torch.uint8), 0, 0, np.array(so, dtype=np.uint64), np.array(sl, dtype=np.uint64)
        s_lo, s_hi = stream_of[b_lo], stream_of[b_hi - 1]
        bit_lo = so[s_lo] * 8 if first[s_lo] == b_lo else gpos[b_lo]
        bit_hi = (so[s_hi] + sl[s_hi]) * 8 if first[s_hi + 1] == b_hi else gpos[b_hi - 1] + int(n_bits_all[b_hi - 1])
        lo, hi = bit_lo // 8, (bit_hi + 7) // 8
        acc, nbits_total = 0, (hi - lo) * 8

        def put(pos, val, width):
            nonlocal acc
            acc |= val << (nbits_total - (pos - lo * 8) - width)

        for bb in range(b_lo, b_hi):
            put(gpos[bb], self.bits[bb][0], self.bits[bb][1])
        for s in range(s_lo, s_hi + 1):
            if b_lo <= first[s] < b_hi:
                put(so[s] * 8, int.from_bytes(b"BZh" + bytes([48 + self.level]), "big"), 32)
            last = first[s + 1] - 1
            if b_lo <= last < b_hi:
                put(gpos[last] + int(n_bits_all[last]), (0x177245385090 << 32) | comb[s], 80)
        return (torch.frombuffer(bytearray(acc.to_bytes(hi - lo, "big")), dtype=torch.uint8), lo, hi,
                np.array(so, dtype=np.uint64), np.array(sl, dtype=np.uint64))


def _run_rank(rank, world, bed_bytes, level, note, oracle):
    bed = np.frombuffer(bed_bytes, dtype=np.uint8)
    cut, halo = M.plan_ranges(bed, world)
    lo, hi = cut[rank] - halo[rank], cut[rank + 1]
    d_range = torch.from_numpy(bed[lo:hi].copy()) if hi > lo else torch.zeros(16, dtype=torch.uint8)
    ph = CpuPhases(oracle)
    # the checker's assemble needs stream_of: capture it from plan
    plan0 = ph.plan

    def plan(tf_all, tf_total, soff, lvl):
        nblock, sof = plan0(tf_all, tf_total, soff, lvl)
        ph._stream_of = [int(x) for x in sof]
        return nblock, sof
    ph.plan = plan

    class OneRank:                     # world size 1 without a process group
        @staticmethod
        def all_gather_into_tensor(out, t): out.view(-1)[:] = t.view(-1)
        @staticmethod
        def get_backend(): return "none"

    d = dist if world > 1 else OneRank
    out = M.compress_sharded(ph, d, rank, world, d_range, hi - lo, halo[rank], bed, lo, level, torch.device("cpu"), torch)
    if rank != 0:
        return None
    hdr = M.build_header(out["streams"], out["blocks_of"], out["stream_off"], out["stream_len"], level, note)
    return hdr + bytes(out["payload"][:out["total"]].numpy())


def _worker(rank, world, port, bed, level, note, q):
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        arc = _run_rank(rank, world, bed, level, note, O)
        if rank == 0:
            q.put(arc)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _spawn(world, bed, level, note):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bed, level, note, q)) for r in range(world)]
    for p in procs:
        p.start()
    arc = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return arc


def test_plan_ranges_and_halos():
    bed = np.frombuffer(b"a\t1\t2\nbb\t3\t4\tx\nccc\t5\t6\n", dtype=np.uint8)
    for world in (1, 2, 3, 5, 9):
        cut, halo = M.plan_ranges(bed, world)
        assert cut[0] == 0 and cut[-1] == len(bed) and all(a <= b for a, b in zip(cut, cut[1:]))
        for r in range(world):
            c = cut[r]
            assert c == 0 or bed[c - 1] == 10
            if c:
                assert halo[r] > 0 and (c - halo[r] == 0 or bed[c - halo[r] - 1] == 10) and 10 not in bed[c - halo[r]:c - 1]
            else:
                assert halo[r] == 0
    tail = np.frombuffer(b"a\t1\t2\nb\t3\t4", dtype=np.uint8)           # unterminated last line
    cut, halo = M.plan_ranges(tail, 3)
    assert cut[-1] == len(tail)


def test_carry_chain_and_shares():
    # rank 1 continues rank 0's chromosome and is one piece; rank 2 continues it as well; rank 3 starts a new one
    s = [(100, 0, 1, 5), (90, 1, 1, 5), (300, 1, 0, 5), (7, 0, 0, 5), (I64_MIN, 0, 1, 0), (9, 1, 1, 2)]
    assert M.carry_chain(s) == [I64_MIN, 100, 100, I64_MIN, I64_MIN, 7]
    nb = np.array([900000] * 10 + [5], dtype=np.uint32)
    for world in (1, 2, 3, 8, 16):
        b = M.block_shares(nb, world)
        assert b[0] == 0 and b[-1] == len(nb) and all(x <= y for x, y in zip(b, b[1:]))
    assert M.block_shares(np.zeros(0, dtype=np.uint32), 4) == [0, 0, 0, 0, 0]


@pytest.mark.parametrize("cfg,lines,level", [(5, 5000, 9), (1, 30000, 1), (2, 8000, 1)])
def test_one_rank_orchestration_equals_oracle(oracle, cfg, lines, level):
    bed = synth.bed(cfg, lines).tobytes()
    assert _run_rank(0, 1, bed, level, "m", oracle) == oracle.archive(bed, level, "m")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,cfg,lines,level", [(2, 5, 5000, 9), (3, 1, 40000, 1), (2, 2, 12000, 1), (3, 3, 140000, 1)])
def test_ranks_produce_the_single_gpu_archive(oracle, world, cfg, lines, level):
    """cfg1 / cfg3 are ONE chromosome: ranges, pieces and blocks of the same stream on different ranks, overlapping
    intervals across the range boundary (uniqueBases needs the carried maximum), blocks joined at bit seams."""
    bed = synth.bed(cfg, lines).tobytes()
    if cfg == 5:
        bed += b"chr1\t5\t9\n"                                          # chr1 reappears at the end
    assert _spawn(world, bed, level, "sharded") == oracle.archive(bed, level, "sharded")


@pytest.mark.timeout(600)
def test_more_ranks_than_lines_and_overlaps(oracle):
    bed = b"c1\t10\t500\nc1\t20\t30\nc1\t40\t600\tx\nc1\t100\t200\nc2\t5\t6\n"
    assert _spawn(3, bed, 9, "") == oracle.archive(bed, 9, "")
    assert _run_rank(0, 1, b"", 9, "", oracle) == oracle.archive(b"", 9, "")
