import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
bed = synth.bed(2, 10_000_000)
pinned = torch.empty(bed.nbytes, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = bed
hv = pinned.numpy()
ctx = s3.Context(0)
def timed(n=6):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = ctx.compress_bed(hv, 9, lazy=True); ts.append((time.perf_counter() - t0) * 1e3)
    return ts, r
os.environ["S3G_CHAIN"] = "0"
timed(2); ts, ref = timed()
print(f"by chromosome: min {min(ts):.2f} median {sorted(ts)[len(ts) // 2]:.2f} ms")
arc = bytes(ref.archive_view)
os.environ["S3G_CHAIN"] = "1"
for mb in [int(x) for x in sys.argv[1:]] or [16, 32, 64, 128]:
    os.environ["S3G_CHAIN_BYTES"] = str(mb << 20)
    timed(2); ts, r = timed()
    print(f"chained, units of {mb} MiB: min {min(ts):.2f} median {sorted(ts)[len(ts) // 2]:.2f} ms same_archive={bytes(r.archive_view) == arc} entry {ctx.last_host_entry}")
os.environ["S3G_CHAIN_BYTES"] = str(32 << 20); os.environ["S3G_TIMING"] = "1"
timed(1)
