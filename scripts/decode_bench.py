#!/usr/bin/env python3
"""Times the decoder path (s3g_decompress_archive, decode.cu) on the archive of a synthetic BED: per-kernel table
(CUDA events around every launch), whole-call device time, host-to-host time.  One JSON line.
usage: scripts/decode_bench.py [--cfg 2] [--lines 10000000] [--steps 3]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import starch3_b200 as s3
from starch3_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--lines", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
bed = synth.bed(a.cfg, a.lines, seed=42)
with s3.Context(0) as ctx:
    res = ctx.compress_bed(bed, 9, lazy=True)
    arc = bytes(res.archive_view)
    del res
    got, info = ctx.decompress_archive(arc)                      # warm-up (buffers grow) + check
    ok = len(got) == bed.nbytes and np.array_equal(np.frombuffer(got, dtype=np.uint8), bed)
    ctx.profile(True)
    ctx.decompress_archive(arc, want_bed=False)
    table = ctx.profile_report()
    ctx.profile(False)
    dev, wall = [], []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        _, inf = ctx.decompress_archive(arc, want_bed=False)
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(inf["device_ms"])
    print(json.dumps({"what": "decoder path: archive -> BED on the GPU", "cfg": a.cfg, "lines": a.lines, "roundtrip_equal": bool(ok),
                      "archive_mb": len(arc) / 1e6, "bed_mb": bed.nbytes / 1e6, "bzip2_blocks": info["n_blocks"], "streams": info["n_streams"],
                      "device_ms": sum(dev) / len(dev), "call_ms_bed_left_on_device": sum(wall) / len(wall),
                      "bed_MBps_device": bed.nbytes / 1e6 / (sum(dev) / len(dev) / 1e3),
                      "kernels_ms": {k: round(v[1], 3) for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])}}))
