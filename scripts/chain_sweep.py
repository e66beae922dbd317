#!/usr/bin/env python3
"""Host-to-host call on an input of one chromosome against the range size of the chained entry (S3G_CHAIN_BYTES).
usage: scripts/chain_sweep.py [cfg] [lines] [MiB ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lines = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
sizes = [int(x) for x in sys.argv[3:]] or [64, 96, 128, 160, 224, 320]
bed = synth.bed(cfg, lines)
pinned = torch.empty(bed.nbytes, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = bed
hv = pinned.numpy()
ctx = s3.Context(0)
def timed(n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = ctx.compress_bed(hv, 9, lazy=True); ts.append((time.perf_counter() - t0) * 1e3)
    return ts, r
for _ in range(3):
    t0 = time.perf_counter(); d = pinned.cuda(); torch.cuda.synchronize(); h2d = (time.perf_counter() - t0) * 1e3
print(f"input {bed.nbytes / 1e6:.1f} MB, plain H2D {h2d:.2f} ms")
os.environ["S3G_CHAIN"] = "0"
ts, ref = timed(3)
print(f"not chained: min {min(ts):.2f} median {sorted(ts)[len(ts) // 2]:.2f} ms")
ref_arc = bytes(ref.archive_view)
os.environ["S3G_CHAIN"] = "1"
for mb in sizes:
    os.environ["S3G_CHAIN_BYTES"] = str(mb << 20)
    timed(1)
    ts, r = timed()
    print(f"ranges of {mb} MiB: min {min(ts):.2f} median {sorted(ts)[len(ts) // 2]:.2f} ms same_archive={bytes(r.archive_view) == ref_arc}")
os.environ["S3G_CHAIN_BYTES"] = str(sizes[len(sizes) // 2] << 20)
os.environ["S3G_TIMING"] = "1"
timed(1)
