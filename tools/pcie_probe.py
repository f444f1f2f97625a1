"""Raw PCIe throughput of the box with pinned host memory: H2D alone, D2H alone, both at once (two streams)."""
import json, time, torch
n = 84 * 10 ** 6  # doubles: 672 MB, the size of one bench step
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_in.normal_()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda"); d_b = torch.randn(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return dt
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    run(a, b, 2); dt = run(a, b)
    print(json.dumps({"mode": name, "ms": round(dt * 1e3, 2), "GBs_per_direction": round(8 * n / dt / 1e9, 1)}))
# chunked: does chunk size matter?
for chunk_mb in (2, 8, 32, 128):
    c = chunk_mb * (1 << 20) // 8
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for off in range(0, n, c):
        e = min(n, off + c)
        with torch.cuda.stream(s1): d_a[off:e].copy_(h_in[off:e], non_blocking=True)
        with torch.cuda.stream(s2): h_out[off:e].copy_(d_b[off:e], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"mode": "both, chunks of %d MB" % chunk_mb, "ms": round(dt * 1e3, 2), "GBs_per_direction": round(8 * n / dt / 1e9, 1)}))
