#!/usr/bin/env python3
"""Generates tests/golden/transform_*.{bed,tf}: inputs and the transformed stream the
REFERENCE starch3 binary itself dumps on stderr ("Content [..]", starch3api.hpp:395).

Run in the build container (needs oracle/_ref/starch3_ref, i.e. /root/reference):
    python tests/golden/make_transform_golden.py
Only single-chromosome, newline-terminated inputs are used: across chromosome
changes the reference loses lines, so those cases are pinned by
the worked examples of SURVEY.md section 8(c) instead (tests/test_oracle.py).
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from starch3_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "starch3_ref")


def ref_transform(bed: bytes) -> bytes:
    p = subprocess.run([REF], input=bed, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.stdout == bytes([0xca, 0x5c, 0xad, 0x1a]), p.stdout
    m = re.search(rb"Content \[(.*?)\]\n--- starch3::Starch::reset_transformation_state", p.stderr, re.S)
    assert m, p.stderr[-400:]
    return m.group(1)


def main():
    cases = {
        "cfg1": synth.bed(1, 3000, seed=42).tobytes(),
        "cfg3": synth.bed(3, 3000, seed=42).tobytes(),
        "cfg3const": synth.bed(3, 1500, seed=42, variant=1).tobytes(),
        "cfg4": synth.bed(4, 1500, seed=42).tobytes(),
        "overlap": b"chr1\t100\t200\tid1\t5\t+\nchr1\t150\t250\tid2\t7\t-\nchr1\t300\t400\tid3\t1\t+\nchr1\t400\t450\n",
        "zerolen": b"chr1\t0\t0\nchr1\t0\t10\nchr1\t20\t30\nchr1\t30\t40\nchr1\t35\t50\n",
    }
    for name, bed in cases.items():
        tf = ref_transform(bed)
        open(os.path.join(HERE, f"transform_{name}.bed"), "wb").write(bed)
        open(os.path.join(HERE, f"transform_{name}.tf"), "wb").write(tf)
        print(name, len(bed), "->", len(tf))


if __name__ == "__main__":
    main()
