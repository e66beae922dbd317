import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lines = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
bed = synth.bed(cfg, lines)
pinned = torch.empty(bed.nbytes, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = bed
hv = pinned.numpy()
ctx = s3.Context(0)
os.environ["S3G_CHAIN"] = "0"
def timed(n=6):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = ctx.compress_bed(hv, 9, lazy=True); ts.append((time.perf_counter() - t0) * 1e3)
    return ts, r
os.environ["S3G_PARTS"] = "1"
ts, ref = timed(3); arc = bytes(ref.archive_view)
print(f"one piece: min {min(ts):.2f} ms")
for parts, first in ((3, None), (4, None), (6, None), (8, None), (12, None), (16, None), (24, None), (8, 0.04), (12, 0.03)):
    os.environ["S3G_PARTS"] = str(parts)
    if first: os.environ["S3G_FIRST"] = str(first)
    else: os.environ.pop("S3G_FIRST", None)
    timed(2); ts, r = timed()
    print(f"parts={parts} first={first}: min {min(ts):.2f} median {sorted(ts)[len(ts) // 2]:.2f} ms same_archive={bytes(r.archive_view) == arc}")
os.environ["S3G_PARTS"] = "8"; os.environ.pop("S3G_FIRST", None); os.environ["S3G_TIMING"] = "1"
timed(1)
