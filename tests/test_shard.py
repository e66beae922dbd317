"""CPU tests (gloo, world size 2) of the multi-rank host logic: partition, per-rank compression,
gather in archive order.  The per-rank compressor is the CPU checker here -- this tests the
sharding, not the kernels; the GPU version of the same test is in test_gpu_e2e.py."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from starch3_b200 import shard, synth


def test_chrom_segments_and_partition():
    bed = b"chr1\t1\t2\nchr1\t3\t4\nchr10\t1\t2\nchr1\t9\t10\nchrX\t5\t6\tq\nchrX\t7\t8\n"
    segs = shard.chrom_segments(bed)
    assert [(n, bed[s:e]) for n, s, e in segs] == [
        (b"chr1", b"chr1\t1\t2\nchr1\t3\t4\n"), (b"chr10", b"chr10\t1\t2\n"), (b"chr1", b"chr1\t9\t10\n"),
        (b"chrX", b"chrX\t5\t6\tq\nchrX\t7\t8\n")]
    big = synth.bed(2, 20000).tobytes()
    segs = shard.chrom_segments(big)
    assert len(segs) == 24 and segs[0][1] == 0 and segs[-1][2] == len(big)
    assert all(a[2] == b[1] for a, b in zip(segs, segs[1:]))
    long_names = b"".join(b"scaffold_%08d\t1\t2\n" % (i // 3) for i in range(30))
    assert len(shard.chrom_segments(long_names)) == 10
    parts = shard.partition([10, 9, 8, 1, 1, 1], 2)
    assert sorted(sum(parts, [])) == list(range(6))
    assert max(sum([10, 9, 8, 1, 1, 1][i] for i in p) for p in parts) <= 17


def _cpu_checker_fn():
    from oracle import oracle as O

    def fn(bed_bytes, level):
        tf, chroms, _ = O.transform(bed_bytes)
        out = []
        for c in chroms:
            s = tf[c["tf_off"]:c["tf_off"] + c["tf_len"]]
            out.append(dict(name=c["name"], stream=O.bz_compress(s, level), lines=c["line_count"],
                            blocks=len([b for b in O.rle1_blocks(s, level)[0] if b["nblock"]]), tf_len=c["tf_len"],
                            bases_nonunique=c["bases_nonunique"], bases_unique=c["bases_unique"]))
        return out
    return fn


def _worker(rank, world, port, bed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arc = shard.compress_sharded(bed, _cpu_checker_fn(), 9, "sharded")
    if rank == 0:
        q.put(arc)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_produce_the_single_rank_archive(oracle):
    bed = synth.bed(5, 6000).tobytes() + b"chr1\t5\t9\n"        # chr1 reappears at the end
    expect = oracle.archive(bed, 9, "sharded")
    assert shard.compress_sharded(bed, _cpu_checker_fn(), 9, "sharded") == expect      # world size 1
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bed, q)) for r in range(2)]
    for p in procs:
        p.start()
    arc = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert arc == expect
