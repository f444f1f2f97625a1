"""Randomised differential test of the three device-resident BATCH loops (single CTA, cluster of 8 CTAs, multi-kernel
loop) against each other and against the oracle's restatement of python/BATCH.py: random small problems -- 1..300 OD
blocks of 1..64 routes (uniform or ragged), 1..700 links, 1..6 links per route, 0/1 or real-valued matrices, simplex or
l1-ball (lasso) feasible sets, BB and projected gradient with line search.  The first objective values of the trace must
agree to 1e-9 and the final objective of a converged run to 1e-6 (north_star).

    python tools/solver_fuzz.py [seconds] [seed0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sps, torch
import bsls_b200
from oracle import solvers_np as S

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def problem(rng):
    nb = int(rng.choice([1, 2, 7, 8, 9, 31, 33, 100, 300]))
    kind = rng.randint(3)
    if kind == 0:
        sizes = np.full(nb, int(rng.choice([1, 2, 3, 5, 8, 9, 16, 17, 32, 33, 64])))
    elif kind == 1:
        sizes = rng.randint(1, int(rng.choice([4, 9, 20, 64])) + 1, size=nb)
    else:
        sizes = rng.randint(1, 6, size=nb)
        sizes[rng.randint(nb)] = int(rng.choice([40, 64]))
    n = int(sizes.sum())
    while n > 4000:       # keep every loop kind applicable (shared-memory limits of the one-launch solvers)
        sizes = sizes[:len(sizes) // 2]
        n = int(sizes.sum())
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
    m = int(rng.choice([1, 5, 31, 32, 33, 100, 255, 256, 257, 700]))
    L = int(min(m, rng.randint(1, 7)))
    base = np.sort(rng.randint(0, m - L + 1, size=(n, L)), axis=1) + np.arange(L)
    vals = np.ones(n * L) if rng.randint(2) else 0.5 + rng.rand(n * L)
    A = sps.csr_matrix((vals, (base.reshape(-1), np.repeat(np.arange(n), L))), shape=(m, n))
    x_true = np.concatenate([rng.dirichlet(np.ones(k)) for k in sizes])
    b = A.dot(x_true) + 0.1 * rng.randn(m) * rng.randint(2)
    x0 = np.concatenate([np.full(k, 1.0 / k) for k in sizes])
    return A, b, starts, x0


t0 = time.time()
cases, seed, failures = 0, seed0, 0
while time.time() - t0 < budget:
    rng = np.random.RandomState(seed)
    A, b, starts, x0 = problem(rng)
    lasso = bool(rng.randint(2))
    method = "bb" if rng.randint(3) else "pg"
    tag = (seed, A.shape, len(starts), lasso, method)
    rp = S.get_solver_parts(A, b, starts, 0.1, lasso=lasso)
    parts = bsls_b200.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True, lasso=lasso)
    if method == "bb":
        ref = S.solve_BB(rp[3], rp[1], rp[2], x0, max_iter=200)
        run = lambda: bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=200)
    else:
        ref = S.solve(rp[3], rp[1], rp[0], x0, rp[2], max_iter=60)
        run = lambda: bsls_b200.BATCH.solve(parts[3], parts[1], parts[0], dev(x0), parts[2], max_iter=60)
    sols = {}
    for kind, env in (("tiny", {"BSLS_TINY_CLUSTER": "0"}), ("cluster", {"BSLS_TINY_CLUSTER": "1"}), ("loop", {"BSLS_NO_TINY": "1"})):
        for k in ("BSLS_TINY_CLUSTER", "BSLS_NO_TINY"):
            os.environ.pop(k, None)
        os.environ.update(env)
        sols[kind] = run()
    for k in ("BSLS_TINY_CLUSTER", "BSLS_NO_TINY"):
        os.environ.pop(k, None)
    scale = max(1.0, abs(ref["f"]))
    ref_trace = np.array([p[1] for p in ref["progress"]])
    bad = []
    for kind, sol in sols.items():
        tr = np.array([p[1] for p in sol["progress"]])
        k = min(5, len(tr), len(ref_trace))
        x = sol["x"].cpu().numpy()
        # a run cut off by max_iter is still on BB's chaotic path, where last-bit differences (the loops take the exact
        # quadratic line search, the reference re-evaluates the objective) have grown to 1e-3: 5 % (+ 1e-4 of the starting objective) there, 1e-6 when converged
        cut = "max_iter" in str(ref["stop"]) or "max_iter" in str(sol["stop"])
        tol = 5e-2 * abs(ref["f"]) + 1e-4 * abs(ref_trace[0]) if cut else 1e-6 * scale + 1e-10
        if not (abs(sol["f"] - ref["f"]) <= tol):
            bad.append((kind, "f", sol["f"], sol["iterations"], sol["stop"]))
        if not np.allclose(tr[:k], ref_trace[:k], rtol=1e-9, atol=1e-12 * scale):
            bad.append((kind, "trace", list(tr[:k])))
        if not (np.isfinite(x).all() and x.min() >= 0.0):
            bad.append((kind, "x"))
    if bad:
        failures += 1
        print("MISMATCH", tag, "oracle:", ref["f"], ref["iterations"], ref["stop"], list(ref_trace[:5]), bad, flush=True)
    cases += 1
    seed += 1
print("%s %d cases in %.0f s (seeds %d..%d), %d mismatches" % ("ok" if failures == 0 else "FAILED", cases, time.time() - t0, seed0, seed - 1, failures))
sys.exit(1 if failures else 0)
