import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
cfg = int(sys.argv[1]); lines = int(sys.argv[2])
bed = synth.bed(cfg, lines)
pinned = torch.empty(bed.nbytes, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = bed
os.environ["S3G_CHAIN"] = "1"; os.environ["S3G_PARTS"] = "2"
ctx = s3.Context(0)
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    r = ctx.compress_bed(pinned.numpy(), 9, lazy=True)
print(r.n_blocks, ctx.last_host_entry)
