import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsls_b200
from bsls_b200.generate import SyntheticProblem
for (nb, K, m, L) in [(100000, 20, 50000, 10), (10000, 20, 5000, 10), (1000, 20, 500, 10)]:
    sp = SyntheticProblem(nb, K, m, L)
    x = sp.x_init.clone(); g = torch.empty_like(x)
    f = sp.problem.obj(x, g)
    # torch check
    rows = sp.problem.t_idx.long(); cols = torch.arange(sp.n, device="cuda").repeat_interleave(L)
    r = torch.zeros(m, dtype=torch.float64, device="cuda").index_add_(0, rows, x[cols]) - sp.b
    gref = r[rows].reshape(sp.n, L).sum(1)
    print(nb, "f", f, float(0.5 * r.dot(r)), "g err", float((g - gref).abs().max()), "b nan", bool(torch.isnan(sp.b).any()))
    step_size, proj, line_search, obj = sp.solver_parts()
    sol = bsls_b200.BATCH.solve_BB(obj, proj, line_search, sp.x_init, max_iter=12, prog_tol=0.0)
    print(" native", [p[1] for p in sol["progress"]], sol["backtracks"])
    sol = bsls_b200.BATCH.solve_BB(lambda x, g=None: obj(x, g), lambda x: proj(x), lambda *a: line_search(*a), sp.x_init, max_iter=12, prog_tol=0.0)
    print(" generic", [p[1] for p in sol["progress"]])
