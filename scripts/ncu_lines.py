#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel from an ncu report (needs -lineinfo and --import-source on).
usage: scripts/ncu_lines.py report.ncu-rep [top-n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = fn = None; hdr = None; lines = {}
for row in rows:
    if not row: continue
    if row[0] == "File Path": cur_file = row[1]; continue
    if row[0] == "Function Name": fn = row[1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or not row[0].isdigit(): continue
    si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    try: s = int(row[si]); i = int(row[ii])
    except ValueError: continue
    key = (fn.split("(")[0], cur_file.split("/")[-1], int(row[0]), row[1].strip())
    a = lines.setdefault(key, [0, 0]); a[0] += s; a[1] += i
ts = sum(a[0] for a in lines.values()) or 1; ti = sum(a[1] for a in lines.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for k, a in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[1]}:{k[2]:4d} smp {100 * a[0] / ts:5.1f}% ins {100 * a[1] / ti:5.1f}%  {k[3][:120]}")
