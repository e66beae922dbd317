#!/usr/bin/env python3
"""Host-side jitter of repeated device-resident calls (S3G_TIMING=1 prints the per-stage host clock)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
bed = synth.bed(2, 10_000_000)
d = torch.from_numpy(bed).cuda()
ctx = s3.Context(0)
for i in range(14):
    t0 = time.perf_counter(); r = ctx.compress_bed_device(d.data_ptr(), d.numel(), 9, want_archive=False); torch.cuda.synchronize()
    print("call", i, round((time.perf_counter() - t0) * 1e3, 2), "ms wall; device_ms", round(r.device_ms, 2), file=sys.stderr)
