#!/usr/bin/env python3
"""bench.py -- throughput of the starch3 compression hot path (BED -> starch transform ->
per-chromosome bzip2) on B200, in input BED MB/s, with the kernel roofline and the CPU
baseline beside it.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

A "step" is one pass of the whole hot path over one batch of synthetic input: the
configuration BASELINE.json's metric is quoted on, configs[1] (hg38-shaped BED6, 10 M elements
over 24 chromosomes, bzip2 at 900k blocks).  One process per GPU; with N > 1 every rank
compresses its own 10 M-line input (independent chromosomes/blocks, no data-path collective:
"weak" scaling) and the value is the total MB of all ranks over the max-over-ranks device time.
"""
import argparse
import ctypes
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
WORKLOAD = "cfg2: synthetic hg38-shaped BED6, 24 chromosomes, ids and scores, bzip2 900k blocks"


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

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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
# the reference arm / CPU baseline: the reference's own algorithm on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_compress(bed_np, threads):
    """Restated transform (oracle) + the reference's vendored libbz2 when oracle/_ref is present,
    one bzip2 stream per chromosome, whole chromosomes spread over `threads` host threads
    (BASELINE.md plan B2).  Returns (seconds, kind, compressed bytes)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    kind = "reference" if O.have_ref() else "port"
    comp = O.ref_bz_compress if O.have_ref() else O.bz_compress
    t0 = time.perf_counter()
    tf, chroms, _ = O.transform(bed_np)
    tfv = memoryview(tf)
    streams = [tfv[c["tf_off"]:c["tf_off"] + c["tf_len"]] for c in chroms]
    order = sorted(range(len(streams)), key=lambda i: -len(streams[i]))
    out = [None] * len(streams)

    def work(i):
        out[i] = comp(np.frombuffer(streams[i], dtype=np.uint8), 9)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(work, order))
    dt = time.perf_counter() - t0
    return dt, kind, sum(len(z) for z in out)


def run_reference(args, rank, world):
    from starch3_b200 import synth
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 24)            # one stream per chromosome: 24 is all the parallelism the reference path has
    sample_lines = args.ref_lines
    bed = synth.bed(2, sample_lines, seed=42)
    times = []
    for i in range(args.warmup + args.steps):
        dt, kind, zbytes = cpu_compress(bed, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = bed.nbytes / 1e6 / (ms / 1000.0)
    sample = f"{sample_lines} lines of cfg2 ({bed.nbytes / 1e6:.1f} MB BED) per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "lines_per_step": sample_lines, "input_mb_per_step": bed.nbytes / 1e6,
                   "note": "CPU: restated transform + reference libbz2 1.0.6 (level 9, workFactor 30), chromosomes over host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# this repository's arm
# -------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import starch3_b200 as s3
    from starch3_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    bed = synth.bed(2, args.lines, seed=42 + rank)              # each rank: its own genome
    nbytes = int(bed.nbytes)
    pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = bed
    d_bed = pinned.cuda(non_blocking=False)
    ctx = s3.Context(local_rank)
    stream = torch.cuda.Stream()                      # the launching stream: library kernels and the timing events share it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        res = ctx.compress_bed_device(d_bed.data_ptr(), nbytes, 9, want_archive=False)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    prof = ctx.profile_report()
    ctx.profile(False)
    ctx.profile_filter(None)
    lib_ms = res.device_ms
    launches = ctx.launch_count - launches0
    # the same K steps once more without any events
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        res = ctx.compress_bed_device(d_bed.data_ptr(), nbytes, 9, want_archive=False)
    f1.record(stream)
    barrier()
    noprof_ms = f0.elapsed_time(f1) / args.steps
    clocks = sampler.stop() if sampler else None

    # ---- end to end: host buffer in, archive in host memory out, through the C ABI ----
    host_view = pinned.numpy()
    for _ in range(min(args.warmup, 2)):
        r2 = ctx.compress_bed(host_view, 9, lazy=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r2 = ctx.compress_bed(host_view, 9, lazy=True)     # archive left in the library's pinned buffer
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    archive_bytes = int(r2.archive_size)

    t = torch.tensor([dev_ms, e2e_s * 1000.0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(nbytes)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    total_bytes = float(tot[0])

    if rank == 0:
        ms_per_step = dev_ms_max / args.steps
        value = total_bytes / 1e6 / (ms_per_step / 1000.0)
        e2e_value = total_bytes / 1e6 / (e2e_ms_max / args.steps / 1000.0)
        peak, peak_src = measured_peak_hbm()
        # dominant kernel: largest share of the per-kernel CUDA-event time in the timed region; its
        # algorithmic bytes are accounted per launch inside the library (S3G_BYTES)
        tot_kernel_ms = sum(v[1] for v in table.values()) or 1.0       # one step, every kernel (last warm-up step)
        top_n, top_ms, top_bytes = prof[top_name]                        # the timed region, dominant kernel only
        n_blocks, tf_bytes = res.n_blocks, res.tf_bytes
        bytes_per_launch = top_bytes / top_n if top_n else 0.0
        achieved = bytes_per_launch / (top_ms / top_n / 1000.0) / 1e9 if top_n and top_ms else 0.0
        traffic = ncu_traffic(top_name, args.lines)
        # whole-pipeline figure of SURVEY.md section 8(d): A = B_in + 2 B_tf + 4 B_blk + 12 M + B_out
        b_out = res.streams_size
        m_sym = 0.67 * tf_bytes
        a_total = nbytes + 2 * tf_bytes + 4 * tf_bytes + 12 * m_sym + b_out
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "lines_per_gpu": args.lines, "input_mb_per_gpu": nbytes / 1e6,
                       "transformed_mb_per_gpu": tf_bytes / 1e6, "bzip2_blocks_per_gpu": n_blocks,
                       "compressed_mb_per_gpu": b_out / 1e6, "l2": "input (%.0f MB) larger than the 126 MB L2" % (nbytes / 1e6),
                       "parallelism": f"{world} independent rank(s), chromosomes/blocks per rank, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": archive_bytes,
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(launches),
            "ms_per_step_without_any_events": noprof_ms, "library_first_to_last_kernel_ms": lib_ms,
            "roofline": {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_share_of_step": table[top_name][1] / tot_kernel_ms, "launches": top_n,
                         "avg_launch_ms": top_ms / top_n if top_n else None, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "note": "algorithmic bytes per launch are accounted inside the library (S3G_BYTES, DESIGN.md section 4); traffic = ncu dram bytes per launch from profiles/ncu_traffic.json",
                         "pipeline": {"algorithmic_bytes_per_step": a_total, "achieved": a_total / (ms_per_step / 1000.0) / 1e9,
                                      "frac": a_total / (ms_per_step / 1000.0) / 1e9 / peak}},
            "kernels_note": "one step (the last warm-up step) with events around every launch",
            "kernels": {k: {"launches": v[0], "ms": round(v[1], 3),
                            "GBps": round(v[2] / (v[1] / 1000.0) / 1e9, 1) if v[1] > 0 and v[2] > 0 else None,
                            "frac": round(v[2] / (v[1] / 1000.0) / 1e9 / peak, 4) if v[1] > 0 and v[2] > 0 else None}
                        for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            threads = min(cores, 24)
            sample = synth.bed(2, args.ref_lines, seed=42)
            dt, kind, _ = cpu_compress(sample, threads)
            line["cpu_baseline"] = {"value": sample.nbytes / 1e6 / dt, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{args.ref_lines} lines of cfg2 ({sample.nbytes / 1e6:.1f} MB BED), one pass, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def ncu_traffic(kernel, lines):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture
    (profiles/ncu_traffic.json, written from one `ncu --set full` run of this same workload), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
        if int(d.get("lines", -1)) != int(lines):
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
    ap.add_argument("--lines", type=int, default=10_000_000, help="lines per GPU (cfg2: 10 M)")
    ap.add_argument("--ref-lines", type=int, default=6_000_000, help="bounded CPU sample, lines of cfg2 (about 20 core-seconds)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
