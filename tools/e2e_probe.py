"""Host C ABI (bsls_proj_multi_simplex on pinned buffers): time per call for the three bench arrays."""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, bsls_b200
NB = 10 ** 6
rng = np.random.RandomState(1)
out = {}
for K in (4, 16, 64):
    host = torch.empty(NB * K, dtype=torch.float64).pin_memory()
    src = rng.randn(NB * K)
    blocks = np.arange(0, NB * K, K, dtype=np.int32)
    ts = []
    for it in range(4):
        host.numpy()[:] = src
        t0 = time.perf_counter()
        bsls_b200.proj_multi_simplex_c(host.numpy(), blocks)
        ts.append(time.perf_counter() - t0)
    out[K] = round(1e3 * min(ts[1:]), 3)
out["sum_ms"] = round(sum(out.values()), 3)
out["ideal_ms_at_47GBs_per_direction"] = round(sum(8 * NB * K for K in (4, 16, 64)) / 47e9 * 1e3, 2)
out["chunk_mb"] = os.environ.get("BSLS_PIPE_MB", "8")
print(json.dumps(out))
