#!/usr/bin/env python3
"""Wall-clock breakdown of the host-to-host call (s3g_compress_bed) on the GPU box."""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
lines = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
bed = synth.bed(2, lines)
pinned = torch.empty(bed.nbytes, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = bed
hv = pinned.numpy()
ctx = s3.Context(0)
d = pinned.cuda()
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,pstate,clocks_event_reasons.active,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader", "-lms", "50"], stdout=subprocess.PIPE, text=True)
T0 = time.perf_counter()
def rd():
    for ln in p.stdout: rows.append((round((time.perf_counter() - T0) * 1e3), ln.strip()))
threading.Thread(target=rd, daemon=True).start()
time.sleep(0.5)
def stamp(): return round((time.perf_counter() - T0) * 1e3)
for i in range(8):
    t0 = time.perf_counter(); r = ctx.compress_bed_device(d.data_ptr(), d.numel(), 9, want_archive=False); torch.cuda.synchronize()
    print(stamp(), "device-resident call", i, round((time.perf_counter() - t0) * 1e3, 2), "ms wall; device_ms", round(r.device_ms, 2))
for i in range(12):
    t0 = time.perf_counter(); d2 = pinned.cuda(); torch.cuda.synchronize(); t1 = time.perf_counter()
    r = ctx.compress_bed(hv, 9, lazy=True); torch.cuda.synchronize()
    print(stamp(), "plain H2D", round((t1 - t0) * 1e3, 2), "host call", i, round((time.perf_counter() - t1) * 1e3, 2), "ms wall; device_ms", round(r.device_ms, 2))
p.terminate()
for t, r in rows: print(t, r)
