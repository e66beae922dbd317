#!/usr/bin/env python3
"""One archive from N GPUs: whole chromosomes are dealt to the ranks, every rank compresses
its share through the C ABI, rank 0 gathers the streams in archive order and writes the container.  Checks on rank 0
that the archive equals the single-GPU one, and prints the strong-scaling wall time.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/shard_bench.py [lines] [cfg]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import starch3_b200 as s3
from starch3_b200 import shard, synth

lines = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")          # object gather of finished streams only: no tensor collective on the data path
bed = synth.bed(cfg, lines).tobytes()
ctx = s3.Context(local)
fn = shard.gpu_compress_fn(ctx)
times = []
arc = None
for it in range(3):
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    arc = shard.compress_sharded(bed, fn, 9, "s")
    if world > 1: dist.barrier()
    times.append(time.perf_counter() - t0)
if rank == 0:
    one = ctx.compress_bed(bed, 9, note="s").archive
    best = min(times)
    print(json.dumps({"what": "one archive from N GPUs, chromosomes dealt to ranks (host partition + gather included)", "cfg": cfg, "lines": lines,
                      "n_gpus": world, "input_mb": round(len(bed) / 1e6, 1), "wall_ms": round(best * 1e3, 1),
                      "MBps": round(len(bed) / 1e6 / best), "archive_equals_single_gpu": bool(arc == one), "archive_bytes": len(arc)}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
