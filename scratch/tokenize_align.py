"""Device time of the range tokenizer (s3g_shard_tokenize) against where the range starts in the buffer and whether it has a
halo line.  usage: scratch/tokenize_align.py [cfg] [lines]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starch3_b200 as s3
from starch3_b200 import synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lines = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
bed = synth.bed(cfg, lines)
b = bed.tobytes()
ctx = s3.Context(0)
d = torch.from_numpy(bed.copy()).cuda()
# line starts to cut at
first_nl = b.index(b"\n") + 1
second = b.index(b"\n", first_nl) + 1
def run(off, halo, label):
    n = len(b) - off
    ts = []
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        sm = ctx.shard_tokenize(d.data_ptr() + off, n, halo)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{label}: start offset {off} (mod 16 = {off % 16}), halo {halo}: {min(ts):.3f} ms, lines {sm['n_lines']}")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
run(0, 0, "whole buffer")
run(first_nl, 0, "from the second line, no halo")
run(0, first_nl, "whole buffer, first line as halo")
run(first_nl, second - first_nl, "from the second line, it as halo")
