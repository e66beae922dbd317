"""host-to-host call from pageable memory against the number of copier threads.  usage: pageable_sweep.py cfg lines"""
import os, sys, time, subprocess
cfg, lines = sys.argv[1], sys.argv[2]
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np
import starch3_b200 as s3
from starch3_b200 import synth
bed = synth.bed(%s, %s)
ctx = s3.Context(0)
ts = []
for i in range(6):
    t0 = time.perf_counter(); r = ctx.compress_bed(bed, 9, lazy=True); ts.append((time.perf_counter() - t0) * 1e3)
print("threads", os.environ.get("S3G_STAGE_THREADS"), "min %%.2f median %%.2f ms" %% (min(ts[2:]), sorted(ts[2:])[2]), "entry", ctx.last_host_entry)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), cfg, lines)
for t in (2, 4, 6, 8, 12, 16):
    env = dict(os.environ, S3G_STAGE_THREADS=str(t))
    subprocess.run([sys.executable, "-c", code], env=env)
