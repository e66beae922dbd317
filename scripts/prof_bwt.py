#!/usr/bin/env python3
"""Runs the block sort alone on a few transformed cfg2 blocks (for ncu: few launches, one batch).
usage: scripts/prof_bwt.py [n_blocks] [cfg]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starch3_b200 as s3
from starch3_b200 import synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = s3.Context(0)
tf, chroms, _ = ctx.transform(synth.bed(cfg, 40000 * nb))
blocks = [tf[i * 899981:(i + 1) * 899981] for i in range(nb) if len(tf) >= (i + 1) * 899981]
for _ in range(2):
    out = ctx.bwt(blocks)
print("blocks", len(blocks), "stats", ctx.sort_stats)
ctx.close()
