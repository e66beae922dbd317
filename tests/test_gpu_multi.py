"""GPU tests of the one-archive-from-several-GPUs path (csrc/shard.cu phases + starch3_b200/multigpu.py): the archive
of N ranks equals the single-GPU archive and the oracle's, byte for byte.  On a one-GPU box the ranks are N contexts on
the same device driven by N host threads (multigpu.LocalGroup: the same orchestration, exchanges by device copies); with
two or more GPUs the same test also runs one context per device, and as real processes over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from starch3_b200 import multigpu as M, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _local(bed, world, level, note, devices=None):
    import starch3_b200 as s3
    devices = devices or [0] * world
    ctxs = [s3.Context(d) for d in devices]
    try:
        return M.compress_bed_local(ctxs, devices, bed, level, note)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world,cfg,lines,level", [(1, 2, 60000, 9), (2, 2, 120000, 9), (3, 5, 90000, 9), (4, 1, 400000, 9), (2, 3, 900000, 9),
                                                   (3, 4, 60000, 1), (5, 2, 50000, 1), (8, 1, 1200000, 9)])
def test_virtual_ranks_on_one_gpu(ctx, oracle, world, cfg, lines, level):
    bed = synth.bed(cfg, lines)
    arc, res = _local(bed, world, level, "multi")
    assert arc == ctx.compress_bed(bed, level, note="multi").archive
    assert arc == oracle.archive_mt(bed, level, "multi")
    shares = [r["share"] for r in res]
    assert shares[0][0] == 0 and shares[-1][1] == res[0]["n_blocks"] and all(a[1] == b[0] for a, b in zip(shares, shares[1:]))
    assert res[0]["n_lines"] == lines


def test_virtual_ranks_edge_inputs(ctx, oracle):
    for bed in (b"c1\t10\t500\nc1\t20\t30\nc1\t40\t600\tx\nc1\t100\t200\nc2\t5\t6\n", b"chr1\t1\t2\n", b"",
                b"chrA\t10\t20\n" * 3 + b"chrB\t1\t2\n" * 2 + b"chrA\t5\t6\n", b"chr1\t1\t2\nchr1\t5\t9\nchr1\t7\t1"):
        a = np.frombuffer(bed, dtype=np.uint8)
        for world in (2, 3):
            arc, _ = _local(a, world, 9, "")
            assert arc == oracle.archive(bed, 9, ""), (bed, world)


def test_one_context_per_device(oracle):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU")
    bed = synth.bed(2, 400000)
    arc, _ = _local(bed, n, 9, "dev", devices=list(range(n)))
    assert arc == oracle.archive_mt(bed, 9, "dev")


def test_processes_over_nccl(oracle, tmp_path):
    """torchrun, one process per GPU: starch3_b200.multigpu.compress_bed."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = tmp_path / "arc.bin"
    code = ("import os, sys, numpy as np, torch, torch.distributed as dist\n"
            f"sys.path.insert(0, {ROOT!r})\n"
            "import starch3_b200 as s3\nfrom starch3_b200 import multigpu as M, synth\n"
            "lr = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(lr)\n"
            "dist.init_process_group('nccl', device_id=torch.device('cuda', lr))\n"
            "ctx = s3.Context(lr); bed = synth.bed(5, 600000)\n"
            "arc, _ = M.compress_bed(ctx, bed, 9, 'nccl')\n"
            f"if dist.get_rank() == 0: open({str(out)!r}, 'wb').write(arc.tobytes())\n"
            "dist.barrier(); ctx.close(); dist.destroy_process_group()\n")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), "-c", code], capture_output=True, timeout=600)
    if p.returncode != 0 and b"unrecognized arguments: -c" in p.stderr:
        f = tmp_path / "w.py"; f.write_text(code)
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                            "--master-port", str(port), str(f)], capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert out.read_bytes() == oracle.archive_mt(synth.bed(5, 600000), 9, "nccl")


def test_transform_into_several_buffers(ctx, oracle):
    """s3g_shard_transform_peers: the transform kernel stores a range's transformed bytes at an offset into every buffer it
    is given (in the N-GPU path: this GPU's copy of the transformed buffer and the peers' copies through their NVLink-mapped
    pointers; here three buffers on one device).  Every buffer ends up with the bytes s3g_shard_transform leaves in its own,
    and nothing outside [offset, offset + length) is touched."""
    import torch
    for cfg, lines, off in ((2, 60000, 0), (3, 200000, 7), (4, 20000, 4099), (5, 30000, 13)):
        bed = synth.bed(cfg, lines)
        d = torch.from_numpy(bed).cuda()
        sm = ctx.shard_tokenize(d.data_ptr(), d.numel(), 0)
        pieces, ptr, n = ctx.shard_transform(M.I64_MIN)
        want = torch.as_tensor(M._DevView(ptr, n), device="cuda").clone()
        assert sm["tf_bytes"] == n
        bufs = [torch.full((n + off + 4096,), 0xAB, dtype=torch.uint8, device="cuda") for _ in range(3)]
        ctx.shard_tokenize(d.data_ptr(), d.numel(), 0)
        pieces2, n2 = ctx.shard_transform_peers(M.I64_MIN, [b.data_ptr() for b in bufs], off)
        assert n2 == n and [(p["tf_off"], p["tf_len"], p["line_count"]) for p in pieces2] == [(p["tf_off"], p["tf_len"], p["line_count"]) for p in pieces]
        for b in bufs:
            assert torch.equal(b[off:off + n], want)
            assert bool((b[:off] == 0xAB).all()) and bool((b[off + n:] == 0xAB).all())
        tf, _, _ = oracle.transform(bed.tobytes())
        assert bytes(want.cpu().numpy()) == tf


@pytest.mark.parametrize("world,cfg,lines,level", [(1, 2, 60000, 9), (2, 2, 120000, 9), (3, 5, 90000, 9), (4, 1, 400000, 9), (2, 3, 900000, 9),
                                                   (3, 4, 60000, 1), (8, 1, 1200000, 9)])
def test_cpp_host_virtual_ranks(ctx, oracle, world, cfg, lines, level):
    """s3g_multi_compress_bed (multi.cu): the phases driven by one C++ process, a host thread and a context per rank; on a
    one-GPU box the contexts share the device.  Same bytes as the single-GPU archive and the oracle's."""
    import starch3_b200 as s3
    bed = synth.bed(cfg, lines)
    ctxs = [s3.Context(0) for _ in range(world)]
    try:
        res = s3.multi_compress_bed(ctxs, bed, level, note="cpp")
        assert res.archive == ctx.compress_bed(bed, level, note="cpp").archive
        assert res.archive == oracle.archive_mt(bed, level, "cpp")
        assert res.n_lines == lines and res.n_blocks > 0
        again = s3.multi_compress_bed(ctxs, bed, level, note="cpp")            # contexts are reusable
        assert again.archive == res.archive
    finally:
        for c in ctxs:
            c.close()


def test_cpp_host_edge_inputs(oracle):
    import starch3_b200 as s3
    ctxs = [s3.Context(0) for _ in range(3)]
    try:
        for bed in (b"c1\t10\t500\nc1\t20\t30\nc1\t40\t600\tx\nc1\t100\t200\nc2\t5\t6\n", b"chr1\t1\t2\n", b"",
                    b"chrA\t10\t20\n" * 3 + b"chrB\t1\t2\n" * 2 + b"chrA\t5\t6\n", b"chr1\t1\t2\nchr1\t5\t9\nchr1\t7\t1"):
            assert s3.multi_compress_bed(ctxs, bed, 9, note="").archive == oracle.archive(bed, 9, ""), bed
        with pytest.raises(s3.Starch3Error) as e:
            s3.multi_compress_bed(ctxs, b"chr1\t1\t2\n" * 3000 + b"chr1\t5\n" + b"chr2\t1\t2\n" * 3000, 9)
        assert e.value.code == -4
        assert s3.multi_compress_bed(ctxs, b"chr1\t1\t2\n", 9).n_lines == 1       # still usable after the error
    finally:
        for c in ctxs:
            c.close()


def test_cpp_host_one_context_per_device(oracle):
    import torch
    import starch3_b200 as s3
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU")
    bed = synth.bed(2, 400000)
    ctxs = [s3.Context(d) for d in range(n)]
    try:
        assert s3.multi_compress_bed(ctxs, bed, 9, note="dev").archive == oracle.archive_mt(bed, 9, "dev")
    finally:
        for c in ctxs:
            c.close()
