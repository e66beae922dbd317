#!/usr/bin/env python3
"""cfg3 with a constant gap (every line the same delta: the transformed stream is "5\n5\n5\n...", its tail block periodic):
device-resident time of the whole path and of k_fallback_exact, archive checked against the reference-libbz2 oracle.
usage: scripts/periodic_probe.py [lines]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import starch3_b200 as s3
from starch3_b200 import synth
from oracle import oracle as O
lines = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
bed = synth.bed(3, lines, variant=1)
d = torch.from_numpy(bed).cuda()
with s3.Context(0) as ctx:
    for _ in range(2):
        r = ctx.compress_bed_device(d.data_ptr(), d.numel(), 9, want_archive=False)
    ctx.profile(True)
    r = ctx.compress_bed_device(d.data_ptr(), d.numel(), 9, want_archive=True)
    t = ctx.profile_report()
    ctx.profile(False)
    ok = r.archive == O.archive_mt(bed, 9, "")
    top = sorted(t.items(), key=lambda kv: -kv[1][1])[:8]
    print(json.dumps({"what": "cfg3, constant gap (periodic tail block)", "lines": lines, "input_mb": bed.nbytes / 1e6, "device_ms": r.device_ms,
                      "archive_equals_oracle": bool(ok), "blocks": r.n_blocks, "k_fallback_exact_ms": t.get("k_fallback_exact", (0, 0, 0))[1],
                      "top_kernels_ms": {k: round(v[1], 3) for k, v in top}}))
