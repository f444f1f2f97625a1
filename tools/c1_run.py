"""C1 (1,000 OD blocks x 5 routes, 2,000 links) BATCH.solve_BB once; for ncu launch lists of a latency-bound solve."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bsls_b200
from bsls_b200.generate import SyntheticProblem
sp = SyntheticProblem.config("C1", noise=0.1, implicit_ones=False)
parts = sp.solver_parts()
for _ in range(2):
    t0 = time.perf_counter()
    sol = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=2000)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(json.dumps({"iterations": sol["iterations"], "ms": round(dt * 1e3, 3), "us_per_iteration": round(dt * 1e6 / sol["iterations"], 1)}))
