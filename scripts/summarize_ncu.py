#!/usr/bin/env python3
"""Turns the scratch ncu outputs of scripts/gpu_profile.sh (gpurun_out/) into the tracked summaries under
profiles/: the launch-list share table, the per-kernel table of the `--set full` capture, and
profiles/ncu_traffic.json (dram bytes per launch, read by bench.py for roofline.traffic).
usage: scripts/summarize_ncu.py <tag> <lines-of-full-capture>"""
import csv, io, json, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
lines = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")

MODES = {"0": "MODE_INIT", "1": "MODE_MM", "2": "MODE_KV", "3": "MODE_KVX"}

def short(name):
    """ncu's demangled name -> the name the library's per-kernel report (and bench.py) uses"""
    name = name.replace("s3g::", "")
    i = name.find("(")
    name = name[:i] if i > 0 else name
    if name.startswith("void "):
        name = name[5:]
    m = re.match(r"(k_hist|k_scatter)<(\d)>$", name)
    if m:
        return f"{m.group(1)}<{MODES[m.group(2)]}>"
    m = re.match(r"k_sweep<(?:\(bool\))?([01]), (?:\(bool\))?([01])>$", name)
    if m:
        return {"00": "k_sweep_first", "10": "k_sweep_ordered", "11": "k_sweep_masks"}.get(m.group(1) + m.group(2), name)
    m = re.match(r"k_mtf_small<(?:\(int\))?(\d+), (?:\(bool\))?([01])>$", name)
    if m:
        return "k_mtf_list_small" if m.group(2) == "1" else "k_mtf_list_big"
    m = re.match(r"k_huff<(?:\(int\))?(\d+)>$", name)
    if m:
        return f"k_huff<{m.group(1)}>"
    m = re.match(r"(k_bound_\w+)<([01])>$", name)
    if m:
        return f"{m.group(1)}<{'true' if m.group(2) == '1' else 'false'}>"
    return name

# ---- launch list -------------------------------------------------------------------------------
f = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(f):
    rows = [l for l in open(f, errors="replace") if l.startswith('"')]
    r = list(csv.reader(rows))
    h = r[0]; ki = h.index("Kernel Name"); mi = h.index("Metric Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
    tot = collections.OrderedDict()
    for row in r[1:]:
        if row[mi] != "gpu__time_duration.sum": continue
        v = float(row[vi].replace(",", "")); u = row[ui]
        us = v / 1000.0 if u in ("ns", "nsecond") else v * 1000.0 if u in ("ms", "msecond") else v
        k = short(row[ki]); a = tot.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
    total = sum(a[1] for a in tot.values())
    with open(os.path.join(P, f"{tag}_launch_list_summary.md"), "w") as o:
        o.write(f"# {tag} -- ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
                "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --lines 10000000 --steps 1 --warmup 1 --no-cpu-baseline`\n"
                "(cfg2, 10 M lines = the bench workload; warm-up, timed and e2e steps all captured; cold-cache, serialised: compare SHARES).\n\n"
                "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            o.write(f"| {k} | {a[0]} | {a[1]:.1f} | {100 * a[1] / total:.1f}% |\n")
        o.write(f"\ntotal {total:.0f} us over {sum(a[0] for a in tot.values())} launches\n")
    print("launch list:", len(tot), "kernels", total, "us")

# ---- full capture ------------------------------------------------------------------------------
f = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(f):
    txt = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    h, units = r[0], r[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "launch__grid_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    ix = {w: h.index(w) for w in want if w in h}
    ki = h.index("Kernel Name")
    def tobytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    def tous(v, u):
        v = float(v.replace(",", ""))
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(u, 1)
    traffic = collections.OrderedDict()
    with open(os.path.join(P, f"{tag}_ncu_full_summary.md"), "w") as o:
        o.write(f"# {tag} -- `ncu --set full --clock-control none` of the dominant kernels\n\n"
                f"Command: `python bench.py --lines {lines} --steps 1 --warmup 1 --no-cpu-baseline` (cfg2), one launch per row, in launch order.\n"
                "dram = dram__bytes_read.sum + dram__bytes_write.sum; the other columns are % of peak sustained.\n"
                "smem wf = shared-memory wavefronts (one per SM per cycle at best), of which `conflict` are bank-conflict replays; issue % = issue slots used.\n\n"
                "| kernel | us | dram MB | dram GB/s | L2 MB | dram % | L2 % | SM % | issue % | warps active % | smem wf M | conflict M | smem pipe busy % | regs | grid |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for row in r[2:]:
            g = lambda w: row[ix[w]] if w in ix else "0"
            us = tous(g("gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]])
            dr = tobytes(g("dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]]) + tobytes(g("dram__bytes_write.sum"), units[ix["dram__bytes_write.sum"]])
            l2 = tobytes(g("lts__t_bytes.sum"), units[ix["lts__t_bytes.sum"]]) if "lts__t_bytes.sum" in ix else 0
            k = short(row[ki])
            traffic.setdefault(k, []).append(dr)
            o.write(f"| {k} | {us:.1f} | {dr / 1e6:.1f} | {dr / us / 1e3:.0f} | {l2 / 1e6:.1f} | {float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                    f"{float(g('lts__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {float(g('sm__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                    f"{float(g('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
                    f"{float(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | {float(g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum').replace(',', '')) / 1e6:.0f} | "
                    f"{float(g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum').replace(',', '')) / 1e6:.0f} | "
                    f"{100 * float(g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum').replace(',', '')) / 148 / (us * 1965.0):.0f} | "
                    f"{g('launch__registers_per_thread')} | {g('launch__grid_size')} |\n")
    json.dump({"tag": tag, "lines": lines, "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over captured launches), bytes",
               "kernels": {k: sum(v) / len(v) for k, v in traffic.items()}}, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    print("full capture:", {k: len(v) for k, v in traffic.items()})
