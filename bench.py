#!/usr/bin/env python3
"""bench.py -- throughput of the starch3 compression hot path (BED -> starch transform ->
per-chromosome bzip2) on B200, in input BED MB/s, with the roofline and the CPU baselines beside it.

    python bench.py --gpus N --steps K --warmup W                    # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path, same input
    python bench.py --cfg 3 [--lines L]                              # another BASELINE.json config

A "step" is one pass of the whole hot path over ONE synthetic input: by default the configuration
BASELINE.json's metric is quoted on, configs[1] (hg38-shaped BED6, 10 M elements over 24 chromosomes, bzip2 at
900k blocks).  One process per GPU.  With N > 1 the ranks compress ONE input together (strong scaling): rank r
tokenises and transforms its newline-aligned byte range, the transformed pieces are exchanged over NVLink
(all-gather), every rank compresses its share of the bzip2 blocks, and rank 0 joins the block bit strings into
the archive -- the same bytes as the 1-GPU archive (checked against the reference-libbz2 oracle after the
timed loops; no value is printed otherwise).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "input BED MB/s compressed (byte-identical archive)"
UNIT = "MB/s"
WORKLOADS = {
    1: ("cfg1: synthetic sorted BED3, single chromosome (chr1), bzip2 900k blocks", 1_000_000),
    2: ("cfg2: synthetic hg38-shaped BED6, 24 chromosomes, ids and scores, bzip2 900k blocks", 10_000_000),
    3: ("cfg3: dense DNase-footprint-like BED3, short uniform-length intervals, one chromosome, bzip2 900k blocks", 100_000_000),
    4: ("cfg4: sparse wide-interval BED6, high-entropy names and scores, one chromosome, bzip2 900k blocks", 20_000_000),
    5: ("cfg5: whole-genome mix of cfg1-4 shapes across 24 chromosomes, bzip2 900k blocks", 400_000_000),
}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append(ln.strip())

    def mark(self):
        """rows read so far: brackets the timed region (nvidia-smi needs ~0.1 s before its first row, so it is started early)"""
        return len(self.rows)

    def stop(self, i0=0, i1=None):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[max(0, i0 - 1):(i1 + 1 if i1 is not None else None)]      # the rows of the timed region and its two neighbours
        for r in rows or self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# the reference arm / CPU baselines: the reference's own algorithm on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_compress(bed_np, threads):
    """Restated transform (oracle) + the reference's vendored libbz2 when oracle/_ref is present, chromosomes and
    the blocks of a chromosome spread over `threads` host threads (BASELINE.md plan B2; block-level threads keep
    all cores busy on the single-chromosome configs too, and the bytes are the serial reference's:
    oracle.archive_mt).  Returns (seconds, kind, archive bytes)."""
    from oracle import oracle as O
    kind = "reference" if O.have_ref() else "port"
    t0 = time.perf_counter()
    arc = O.archive_mt(bed_np, 9, "", threads=threads)
    return time.perf_counter() - t0, kind, arc


def cpu_baselines(cfg, lines, threads):
    """B0 (the reference binary as shipped, cfg1 shape only: it performs the transform but never compresses,
    SURVEY.md F2) and B1 (the completed CPU path on one core) on bounded samples."""
    from oracle import oracle as O
    from starch3_b200 import synth
    out = {}
    if os.path.exists(O.REF_BINARY):
        s = synth.bed(1, 60_000, seed=42)
        t0 = time.perf_counter()
        p = subprocess.run([O.REF_BINARY], input=s.tobytes(), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        if p.returncode == 0:
            out["B0_reference_binary"] = {"value": s.nbytes / 1e6 / dt, "unit": UNIT, "cores": 2, "kind": "reference",
                                          "sample": f"60000 lines of cfg1 ({s.nbytes / 1e6:.2f} MB), {dt:.1f} s; oracle/_ref/starch3_ref "
                                                    "as shipped: transform only, no compression, stderr to /dev/null"}
    s = synth.bed(cfg, min(lines, 1_000_000), seed=42)
    dt, kind, _ = cpu_compress(s, 1)
    out["B1_one_core"] = {"value": s.nbytes / 1e6 / dt, "unit": UNIT, "cores": 1, "kind": kind,
                          "sample": f"{min(lines, 1_000_000)} lines of cfg{cfg} ({s.nbytes / 1e6:.1f} MB BED), one pass, {dt:.1f} s"}
    return out


def run_reference(args, rank, world):
    from starch3_b200 import synth
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 32)
    workload, _ = WORKLOADS[args.cfg]
    bed = synth.bed(args.cfg, args.lines, seed=42)
    # the whole K/W run has to end within a few minutes: a step is the whole input unless that is too slow,
    # then a bounded sample of it (stated in `sample`)
    sample_lines = args.lines
    if args.ref_lines and args.ref_lines < args.lines:
        sample_lines = args.ref_lines
        bed = synth.bed(args.cfg, sample_lines, seed=42)
    times = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        dt, kind, _ = cpu_compress(bed, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = bed.nbytes / 1e6 / (ms / 1000.0)
    sample = f"{sample_lines} lines of cfg{args.cfg} ({bed.nbytes / 1e6:.1f} MB BED) per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "lines_per_step": sample_lines, "input_mb_per_step": bed.nbytes / 1e6,
                   "same_input_as_b200_arm": sample_lines == args.lines,
                   "note": "CPU: restated transform + reference libbz2 1.0.6 (level 9, workFactor 30), chromosomes and blocks over host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# this repository's arm
# -------------------------------------------------------------------------------------------------
def stage_table(res, nbytes, ms_scale, peak):
    """SURVEY.md section 8(d) bytes per stage over the CUDA-event time of the stage (s3g_result.stage_ms)."""
    b_in, b_tf, b_blk, m, b_out = nbytes, res.tf_bytes, res.rle_bytes, res.mtf_symbols, res.streams_size
    alg = {"tokenise+transform": b_in + b_tf, "rle1+cut+crc": b_tf + b_blk, "blocksort": 2 * b_blk, "mtf": b_blk + 2 * m,
           "huffman": 10 * m + b_out, "assemble": None}
    out = {}
    for nm, ms in res.stage_ms.items():
        ms *= ms_scale
        a = alg.get(nm)
        out[nm] = {"ms": round(ms, 3), "algorithmic_mb": round(a / 1e6, 1) if a else None,
                   "GBps": round(a / ms / 1e6, 1) if a and ms > 0 else None,
                   "frac": round(a / ms / 1e6 / peak, 4) if a and ms > 0 else None}
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import starch3_b200 as s3
    from starch3_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import bench_multigpu
        return bench_multigpu.bench(args, rank, world, local_rank, METRIC, UNIT, WORKLOADS, measured_peak_hbm, ClockSampler, cpu_compress)

    workload, _ = WORKLOADS[args.cfg]
    bed = synth.bed(args.cfg, args.lines, seed=42)
    nbytes = int(bed.nbytes)
    pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = bed
    d_bed = pinned.cuda(non_blocking=False)
    ctx = s3.Context(local_rank)
    stream = torch.cuda.Stream()                      # the launching stream: library kernels and the timing events share it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def sync():
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)                # started before the warm-up: its rows are bracketed around the timed regions below
    # ---- device-resident: input already in HBM, output left in HBM ----
    # warm-up; the last warm-up step is timed kernel by kernel (two events around every launch) to get the
    # per-kernel table and to find the dominant kernel
    for i in range(args.warmup):
        if i == args.warmup - 1:
            ctx.profile(True)
        res = ctx.compress_bed_device(d_bed.data_ptr(), nbytes, 9, want_archive=False)
    table = ctx.profile_report() if args.warmup else {}
    if not table:
        ctx.profile(True)
        res = ctx.compress_bed_device(d_bed.data_ptr(), nbytes, 9, want_archive=False)
        table = ctx.profile_report()
    top_name = max(table.items(), key=lambda kv: kv[1][1])[0]
    # timed region: K steps; only the dominant kernel keeps its events (its duration is measured live, here)
    ctx.profile_filter(top_name)
    sync()
    launches0 = ctx.launch_count
    clk0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    stage_sum = {}
    for _ in range(args.steps):
        res = ctx.compress_bed_device(d_bed.data_ptr(), nbytes, 9, want_archive=False)
        for k, v in res.stage_ms.items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    e1.record(stream)
    sync()
    dev_ms = e0.elapsed_time(e1)
    prof = ctx.profile_report()
    ctx.profile(False)
    ctx.profile_filter(None)
    lib_ms = res.device_ms
    launches = ctx.launch_count - launches0
    res.stage_ms = {k: v / args.steps for k, v in stage_sum.items()}
    # the same K steps once more without any per-kernel events
    sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        ctx.compress_bed_device(d_bed.data_ptr(), nbytes, 9, want_archive=False)
    f1.record(stream)
    sync()
    noprof_ms = f0.elapsed_time(f1) / args.steps

    # ---- end to end: host buffer in, archive in host memory out, through the C ABI ----
    def e2e(host_view):
        for _ in range(min(args.warmup, 2)):
            r = ctx.compress_bed(host_view, 9, lazy=True)
        sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = ctx.compress_bed(host_view, 9, lazy=True)     # archive left in the library's pinned buffer
        sync()
        return (time.perf_counter() - t0) * 1000.0 / args.steps, r

    e2e_ms, r2 = e2e(pinned.numpy())
    clocks = sampler.stop(clk0, sampler.mark())               # sampled over the timed regions: device-resident steps and host-to-host calls
    he = ctx.last_host_entry
    host_entry = ("one piece" if he == 0 else "%d ranges by chromosome, two worker contexts" % he if he > 0 else
                  "%d ranges chained at bzip2-block granularity" % -he) + " (DESIGN.md section 5b)"
    archive = bytes(r2.archive_view)                          # what the parity check below compares
    archive_bytes = int(r2.archive_size)
    e2e_pageable_ms, _ = e2e(bed)                             # the same call fed from ordinary (pageable) memory, as the CLI does

    ms_per_step = dev_ms / args.steps
    value = nbytes / 1e6 / (ms_per_step / 1000.0)
    peak, peak_src = measured_peak_hbm()
    # ---- parity: the archive the timed call produced against the reference-libbz2 oracle (outside the timed region) ----
    parity = None
    cpu_dt = None
    if not args.no_parity:
        cpu_dt, okind, expect = cpu_compress(bed, min(os.cpu_count() or 1, 32))
        parity = archive == expect
        if not parity:
            raise SystemExit(f"bench.py: the archive differs from the {okind} oracle's ({len(archive)} vs {len(expect)} bytes): no value reported")
    # ---- roofline: SURVEY.md section 8(d), A = B_in + 2 B_tf + 4 B_blk + 12 M + B_out over first-to-last-kernel time ----
    b_tf, b_blk, m_sym, b_out = res.tf_bytes, res.rle_bytes, res.mtf_symbols, res.streams_size
    a_total = nbytes + 2 * b_tf + 4 * b_blk + 12 * m_sym + b_out
    achieved = a_total / (ms_per_step / 1000.0) / 1e9
    tot_kernel_ms = sum(v[1] for v in table.values()) or 1.0       # one step, every kernel (last warm-up step)
    top_n, top_ms, top_bytes = prof[top_name]                        # the timed region, dominant kernel only
    k_bytes = top_bytes / top_n if top_n else 0.0
    k_ach = k_bytes / (top_ms / top_n / 1000.0) / 1e9 if top_n and top_ms else 0.0
    stage_scale = ms_per_step / (sum(res.stage_ms.values()) or 1.0)  # stage marks cover first to last kernel of the call
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "parity_checked": parity,
        "config": {"workload": workload, "lines": args.lines, "input_mb": nbytes / 1e6,
                   "transformed_mb": b_tf / 1e6, "rle_mb": b_blk / 1e6, "mtf_symbols_m": m_sym / 1e6, "bzip2_blocks": res.n_blocks,
                   "compressed_mb": b_out / 1e6, "l2": "input (%.0f MB) larger than the 126 MB L2" % (nbytes / 1e6),
                   "parallelism": "1 GPU"},
        "e2e": {"value": nbytes / 1e6 / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": archive_bytes,
                "ms_per_step": e2e_ms, "host_buffer": "pinned (caller-provided); archive left in the library's pinned buffer",
                "upload_overlap": host_entry},
        "e2e_pageable": {"value": nbytes / 1e6 / (e2e_pageable_ms / 1000.0), "unit": UNIT, "ms_per_step": e2e_pageable_ms,
                         "host_buffer": "pageable (malloc), as the CLI client feeds it"},
        "gpu_launches": int(launches),
        "ms_per_step_without_any_events": noprof_ms, "library_first_to_last_kernel_ms": lib_ms,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": a_total,
                     "formula": "A = B_in + 2 B_tf + 4 B_blk + 12 M + B_out (SURVEY.md 8(d)) with the measured B_blk and M; frac = A / t_device / peak",
                     "kernel": top_name, "kernel_frac": k_ach / peak, "kernel_achieved": k_ach,
                     "kernel_share_of_step": table[top_name][1] / tot_kernel_ms, "kernel_launches": top_n,
                     "kernel_avg_launch_ms": top_ms / top_n if top_n else None, "kernel_algorithmic_bytes_per_launch": k_bytes,
                     "traffic": ncu_traffic(top_name, args.cfg, args.lines),
                     "note": "kernel_* = the dominant kernel, timed live with CUDA events inside the timed region; its algorithmic bytes per launch are "
                             "accounted inside the library (S3G_BYTES, DESIGN.md section 4); traffic = ncu dram bytes per launch from profiles/ncu_traffic.json"},
        "stages": stage_table(res, nbytes, stage_scale, peak),
        "kernels_note": "one step (the last warm-up step) with events around every launch",
        "kernels": {k: {"launches": v[0], "ms": round(v[1], 3),
                        "GBps": round(v[2] / (v[1] / 1000.0) / 1e9, 1) if v[1] > 0 and v[2] > 0 else None,
                        "frac": round(v[2] / (v[1] / 1000.0) / 1e9 / peak, 4) if v[1] > 0 and v[2] > 0 else None}
                    for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = min(cores, 32)
        sample_lines = min(args.lines, args.ref_lines or args.lines)
        sample = bed if sample_lines == args.lines else synth.bed(args.cfg, sample_lines, seed=42)
        if sample_lines == args.lines and cpu_dt is not None:
            dt, kind = cpu_dt, okind               # the parity check above just timed exactly this
        else:
            dt, kind, _ = cpu_compress(sample, threads)
        line["cpu_baseline"] = {"value": sample.nbytes / 1e6 / dt, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"{sample_lines} lines of cfg{args.cfg} ({sample.nbytes / 1e6:.1f} MB BED), one pass, {dt:.1f} s (B2: all host threads)"}
        line["cpu_baseline"].update(cpu_baselines(args.cfg, args.lines, threads))
    print(json.dumps(line), flush=True)
    ctx.close()


def ncu_traffic(kernel, cfg, lines):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture
    (profiles/ncu_traffic.json, written from one `ncu --set full` run of this same workload), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
        if int(d.get("lines", -1)) != int(lines) or int(d.get("cfg", 2)) != int(cfg):
            return None
        return d["kernels"].get(kernel)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cfg", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (default: 2, the one the metric is quoted on)")
    ap.add_argument("--lines", type=int, default=0, help="lines of the input (default: the config's BASELINE size; cfg5: 40 M)")
    ap.add_argument("--ref-lines", type=int, default=0, help="bound the CPU arm to a sample of this many lines (default: the same input)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the archive comparison with the oracle (never for a reported number)")
    args = ap.parse_args()
    if not args.lines:
        args.lines = WORKLOADS[args.cfg][1] if args.cfg != 5 else 40_000_000
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
