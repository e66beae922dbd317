import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
bed = synth.bed(2, 10_000_000)
pinned = torch.empty(bed.nbytes, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = bed
hv = pinned.numpy()
ctx = s3.Context(0)
for i in range(4):
    r = ctx.compress_bed(hv, 9, lazy=True); torch.cuda.synchronize()
os.environ["S3G_TIMING"] = "1"
for i in range(2):
    t0 = time.perf_counter(); r = ctx.compress_bed(hv, 9, lazy=True); torch.cuda.synchronize()
    print("host call", round((time.perf_counter() - t0) * 1e3, 2), "ms; device_ms", round(r.device_ms, 2), file=sys.stderr)
