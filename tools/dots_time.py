import json, sys
sys.path.insert(0, ".")
import numpy as np, torch, bsls_b200
from bsls_b200.sparse import default_workspace
n = 2 * 10 ** 7
ws = default_workspace(torch.device("cuda", 0))
x, y, g = (torch.randn(n, dtype=torch.float64, device="cuda") for _ in range(3))
for pairs, nm in (([(x, y), (y, y), (g, y)], "3 pairs / 3 arrays"), ([(x, y)], "1 pair")):
    ws.dots(pairs); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
    for a, b in ev:
        a.record(); ws.dots(pairs); b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    arrays = 3 if len(pairs) == 3 else 2
    print(json.dumps({"op": "dots " + nm, "ms": round(ms, 4), "GBs": round(8 * n * arrays / ms / 1e6), "frac": round(8 * n * arrays / ms / 1e6 / 6552, 3)}))
