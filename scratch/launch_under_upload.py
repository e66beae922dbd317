"""Round trip of a tiny launch + synchronise (and of a launch + 8-byte read-back) while a large host-to-device copy runs on
another stream: idle, one big copy, chunked copies, chunked copies with gaps."""
import time, torch
dev = torch.device("cuda:0")
big = torch.empty(2 << 30, dtype=torch.uint8).pin_memory()
dbig = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
y = torch.zeros(64, device=dev)
h = torch.zeros(64).pin_memory()
side = torch.cuda.Stream()
main = torch.cuda.current_stream()
def loop(n, readback):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        y.add_(1)
        if readback: h.copy_(y, non_blocking=True)
        main.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    ts.sort()
    return ts[len(ts) // 2], ts[int(len(ts) * 0.9)], sum(ts) / len(ts)
def burst(n):
    # n launches back to back, one synchronise
    t0 = time.perf_counter()
    for _ in range(n): y.add_(1)
    main.synchronize()
    return (time.perf_counter() - t0) * 1e6 / n
def upload(chunk, gap_us):
    with torch.cuda.stream(side):
        if chunk == 0:
            dbig.copy_(big, non_blocking=True)
        else:
            for o in range(0, big.numel(), chunk):
                dbig[o:o + chunk].copy_(big[o:o + chunk], non_blocking=True)
                if gap_us: torch.cuda._sleep(int(gap_us * 1900))
for _ in range(3): loop(50, True)
torch.cuda.synchronize()
print("idle: launch+sync median/p90/mean us", loop(300, False), " with read-back", loop(300, True), " burst of 20 per launch", burst(20))
for name, chunk, gap in (("one 2 GiB copy", 0, 0), ("16 MiB chunks", 16 << 20, 0), ("2 MiB chunks", 2 << 20, 0), ("2 MiB chunks + 10 us gaps", 2 << 20, 10), ("1 MiB chunks + 10 us gaps", 1 << 20, 10)):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); upload(chunk, gap); t_issue = time.perf_counter() - t0
    a = loop(150, False); b = loop(150, True); c = burst(20)
    done_early = side.query()
    side.synchronize(); t_all = time.perf_counter() - t0
    print(f"{name}: issue {t_issue * 1e3:.1f} ms, copy done after {t_all * 1e3:.1f} ms (finished before the loops ended: {done_early}); launch+sync {a}; with read-back {b}; burst {c:.1f}")
