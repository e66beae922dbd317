#!/usr/bin/env python3
"""Top CUDA source lines by warp-stall samples, per kernel, from an ncu report captured with
--set full --import-source on.   usage: scripts/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn, hdr, sec = None, None, []
def _i(x):
    try: return int(x or 0)
    except ValueError: return 0
def flush():
    if not sec: return
    i_s = hdr.index("# Samples"); i_ie = hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(_i(r[i_s]) for r in sec)
    print(f"== {fn[:100]}  samples={tot}")
    for r in sorted(sec, key=lambda r: -_i(r[i_s]))[:top]:
        s = _i(r[i_s])
        st = sorted(((_i(r[i]), hdr[i][6:]) for i in stalls), reverse=True)[:3]
        print(f"{r[0]:>5} {100.0*s/max(tot,1):5.1f}% inst={r[i_ie]:>10} {' '.join(f'{n}:{100*v//max(s,1)}' for v,n in st if v)} | {r[1].strip()[:110]}")
for r in rows:
    if not r: continue
    if r[0] == "Function Name": flush(); fn = r[1]; sec = []; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "File Path" or hdr is None: continue
    if r[0] != "" and len(r) >= len(hdr): sec.append(r)
flush()
