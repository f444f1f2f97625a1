import json, sys, os
sys.path.insert(0, ".")
import numpy as np, torch, bsls_b200
from bsls_b200 import c_extensions as cx
for K in (16, 5, 64):
    nb = 10 ** 8 // K; n = nb * K
    starts = torch.arange(0, n, K, dtype=torch.int64, device="cuda")
    plan = bsls_b200.BlockPlan(starts, n)
    x = torch.rand(n, dtype=torch.float64, device="cuda"); z = torch.empty(n - nb, dtype=torch.float64, device="cuda")
    for name, fn in (("x2z", lambda: cx.x2z_c(x, z, plan)), ("z2x", lambda: cx.z2x_c(x, z, plan))):
        fn(); torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
        print(json.dumps({"op": name, "K": K, "ms": round(ms, 4), "GBs": round(16 * n / ms / 1e6), "frac": round(16 * n / ms / 1e6 / 6552, 3)}))
