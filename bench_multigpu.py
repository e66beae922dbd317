"""bench.py --gpus N (N > 1): ONE input compressed by N ranks together (strong scaling); see bench.py.
Kept outside the product package: it calls the CPU checker (through bench.py's cpu_compress) for the parity check."""
import numpy as np

from starch3_b200.multigpu import Phases, build_header, compress_sharded, plan_ranges


def bench(args, rank, world, local_rank, METRIC, UNIT, WORKLOADS, measured_peak_hbm, ClockSampler, cpu_compress):
    import json
    import os
    import time
    import torch
    import torch.distributed as dist
    import starch3_b200 as s3
    from starch3_b200 import synth

    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    workload, _ = WORKLOADS[args.cfg]
    bed = synth.bed(args.cfg, args.lines, seed=42)               # every rank can read the one input; it uploads its range only
    nbytes = int(bed.nbytes)
    cut, halo = plan_ranges(bed, world)
    lo, hi = cut[rank] - halo[rank], cut[rank + 1]
    pinned = torch.empty(max(hi - lo, 16), dtype=torch.uint8).pin_memory()
    pinned.numpy()[:hi - lo] = bed[lo:hi]
    d_range = pinned.to(device)
    ctx = s3.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ph = Phases(ctx, device)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return compress_sharded(ph, dist, rank, world, d_range, hi - lo, halo[rank], bed, lo, 9, device, torch)

    sampler = ClockSampler(local_rank) if rank == 0 else None      # started before the warm-up: its rows are bracketed around the timed region
    for _ in range(max(args.warmup, 1)):
        out = step_resident()
    barrier()
    launches0 = ctx.launch_count
    clk0 = sampler.mark() if sampler else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_sum = {}
    e0.record(stream)
    for _ in range(args.steps):
        out = step_resident()
        for k, v in ctx.stage_times().items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(clk0, sampler.mark()) if sampler else None

    # ---- end to end: every rank uploads its range from pinned host memory, rank 0 ends with the archive in host memory ----
    arc_host = None

    def step_e2e():
        nonlocal arc_host
        d = pinned.to(device, non_blocking=True)
        o = compress_sharded(ph, dist, rank, world, d, hi - lo, halo[rank], bed, lo, 9, device, torch)
        if rank == 0:
            hdr = build_header(o["streams"], o["blocks_of"], o["stream_off"], o["stream_len"], 9, "")
            need = len(hdr) + o["total"]
            if arc_host is None or arc_host.numel() < need:
                arc_host = torch.empty(need + need // 8, dtype=torch.uint8).pin_memory()
            arc_host[:len(hdr)] = torch.frombuffer(bytearray(hdr), dtype=torch.uint8)
            arc_host[len(hdr):need].copy_(o["payload"][:o["total"]], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return need
        return 0

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arc_n = step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    barrier()

    t = torch.tensor([dev_ms, e2e_ms, float(launches)], dtype=torch.float64, device=device)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    st = torch.tensor([stage_sum.get(k, 0.0) / args.steps for k in s3.api.STAGE_NAMES], dtype=torch.float64, device=device)
    st_max = st.clone(); dist.all_reduce(st_max, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_per_step = float(tmax[0]) / args.steps
        e2e_step = float(tmax[1]) / args.steps
        value = nbytes / 1e6 / (ms_per_step / 1000.0)
        peak, peak_src = measured_peak_hbm()
        parity = None
        cpu_dt = None
        if not args.no_parity:
            cpu_dt, okind, expect = cpu_compress(bed, min(os.cpu_count() or 1, 32))
            got = bytes(arc_host[:arc_n].numpy())
            parity = got == expect
            if not parity:
                raise SystemExit(f"bench.py: the {world}-GPU archive differs from the {okind} oracle's ({len(got)} vs {len(expect)} bytes): no value reported")
        a_total = nbytes + 2 * out["tf_total"] + 4 * out["rle_bytes"] + 12 * out["mtf_symbols"] + out["total"]
        achieved = a_total / (ms_per_step / 1000.0) / 1e9
        stages = {k: round(float(st_max[i]), 3) for i, k in enumerate(s3.api.STAGE_NAMES)}
        stages["exchange+host"] = round(ms_per_step - sum(stages.values()), 3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "parity_checked": parity,
            "config": {"workload": workload, "lines": args.lines, "input_mb": nbytes / 1e6, "transformed_mb": out["tf_total"] / 1e6,
                       "rle_mb": out["rle_bytes"] / 1e6, "mtf_symbols_m": out["mtf_symbols"] / 1e6, "bzip2_blocks": out["n_blocks"],
                       "compressed_mb": out["total"] / 1e6, "l2": "input (%.0f MB) larger than the 126 MB L2" % (nbytes / 1e6),
                       "parallelism": f"ONE input over {world} ranks: byte ranges for tokenise + transform, all-gather of the transformed bytes over "
                                      "NVLink (NCCL), the block plan on every rank, contiguous shares of the bzip2 blocks, byte strings gathered on rank 0"},
            "e2e": {"value": nbytes / 1e6 / (e2e_step / 1000.0), "unit": UNIT, "h2d_bytes_per_step": nbytes + int(sum(halo)),
                    "d2h_bytes_per_step": int(arc_n), "ms_per_step": e2e_step,
                    "host_buffer": "pinned; every rank uploads its own range, rank 0 ends with the whole archive in pinned host memory"},
            "gpu_launches": int(float(tsum[2])),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "peak_source": peak_src + f" x {world} GPUs", "algorithmic_bytes_per_step": a_total,
                         "formula": "A = B_in + 2 B_tf + 4 B_blk + 12 M + B_out (SURVEY.md 8(d)); frac = A / t_device / (N x peak)", "traffic": None},
            "stages_max_over_ranks_ms": stages,
            "clocks": clocks,
        }
        if cpu_dt is not None:
            line["cpu_baseline"] = {"value": nbytes / 1e6 / cpu_dt, "unit": UNIT, "cores": min(os.cpu_count() or 1, 32), "kind": okind,
                                    "sample": f"{args.lines} lines of cfg{args.cfg} ({nbytes / 1e6:.1f} MB BED), one pass, {cpu_dt:.1f} s (all host threads)"}
        print(json.dumps(line), flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
