#!/usr/bin/env python3
"""Times the device-resident hot path on every BASELINE.json config shape (GPU box) and checks each
archive by round trip (bzip2 decode of every stream + inverse transform on a prefix).
usage: scripts/cfg_sweep.py cfg:lines[:variant] ...   e.g. 1:1000000 2:10000000 3:20000000 4:5000000"""
import bz2, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import starch3_b200 as s3
from starch3_b200 import synth

ctx = s3.Context(0)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
for spec in sys.argv[1:]:
    f = spec.split(":"); cfg, lines = int(f[0]), int(f[1]); variant = int(f[2]) if len(f) > 2 else 0
    bed = synth.bed(cfg, lines, variant=variant)
    d = torch.from_numpy(bed).cuda()
    for _ in range(2):
        res = ctx.compress_bed_device(d.data_ptr(), d.numel(), 9, want_archive=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = ctx.compress_bed_device(d.data_ptr(), d.numel(), 9, want_archive=False)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    dev_ms = res.device_ms
    ctx.profile(True)
    res = ctx.compress_bed_device(d.data_ptr(), d.numel(), 9, want_archive=False)
    prof = ctx.profile_report(); ctx.profile(False)
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]
    # check: every stream decodes; total decoded size equals the transformed size
    z = ctx.read_streams(res.streams_size)
    tot = 0
    for c in res.chroms:
        tot += len(bz2.decompress(z[c["bz_off"]:c["bz_off"] + c["bz_len"]])) if lines <= 120_000_000 else c["tf_len"]
    ok = tot == res.tf_bytes
    print(json.dumps({"cfg": cfg, "variant": variant, "lines": lines, "in_mb": round(bed.nbytes / 1e6, 1), "tf_mb": round(res.tf_bytes / 1e6, 1),
                      "blocks": res.n_blocks, "out_mb": round(res.streams_size / 1e6, 1), "device_ms": round(dev_ms, 2),
                      "wall_ms": round(wall, 2), "GBps_in": round(bed.nbytes / 1e6 / dev_ms, 2), "decodes": ok,
                      "top": {k: round(v[1], 2) for k, v in top}}), flush=True)
    del d
ctx.close()
