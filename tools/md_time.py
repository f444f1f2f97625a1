import json, sys
sys.path.insert(0, ".")
import numpy as np, torch, bsls_b200
from bsls_b200.sparse import default_workspace
for K in (16, 20, 5):
    nb = 2 * 10 ** 7 // K; n = nb * K
    starts = torch.arange(0, n, K, dtype=torch.int64, device="cuda")
    plan = bsls_b200.BlockPlan(starts, n)
    ws = default_workspace(torch.device("cuda", 0))
    x = torch.rand(n, dtype=torch.float64, device="cuda"); g = torch.randn(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    ws.md_update(plan, y, x, g, 0.01); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); ws.md_update(plan, y, x, g, 0.01); b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    print(json.dumps({"op": "md_update", "K": K, "n": n, "ms": round(ms, 4), "GBs": round(24 * n / ms / 1e6), "frac": round(24 * n / ms / 1e6 / 6552, 3)}))
