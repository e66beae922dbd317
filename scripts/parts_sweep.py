#!/usr/bin/env python3
"""Host-to-host call (s3g_compress_bed, pinned input) against the number of ranges of the pipelined entry and the share of
the first range.  usage: scripts/parts_sweep.py [lines]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
lines = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
bed = synth.bed(2, lines)
pinned = torch.empty(bed.nbytes, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = bed
hv = pinned.numpy()
ctx = s3.Context(0)
ref = None
for parts, first in [(1, None), (2, None), (3, None), (3, 0.1), (4, None), (4, 0.08), (5, None), (6, None), (6, 0.05), (8, None), (8, 0.04), (3, None)]:
    os.environ["S3G_PARTS"] = str(parts)
    if first is None: os.environ.pop("S3G_FIRST", None)
    else: os.environ["S3G_FIRST"] = str(first)
    for _ in range(3): r = ctx.compress_bed(hv, 9, lazy=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); r = ctx.compress_bed(hv, 9, lazy=True); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    a = bytes(r.archive_view)
    if ref is None: ref = a
    print(f"parts={parts} first={first} min {min(ts):.2f} median {sorted(ts)[len(ts)//2]:.2f} ms same_archive={a == ref}", flush=True)
