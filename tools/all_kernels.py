"""Runs every kernel of libbsls_b200 once (after one warm-up call) on inputs far larger than L2, so that
ncu can list duration and DRAM traffic per kernel:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off ...

Prints, for every call, the algorithmic bytes (SURVEY 8d formulas) so that the table in profiles/ can put
achieved GB/s = algorithmic bytes / duration next to the measured DRAM traffic."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bsls_b200
import bench
from bsls_b200 import _lib
from bsls_b200.generate import SyntheticProblem
from bsls_b200.sparse import axpby, default_workspace

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
L = _lib.lib()
calls = []


def run(name, bytes_, fn, setup=None):
    args = setup() if setup else ()
    fn(*args)                      # warm-up (not profiled)
    args = setup() if setup else ()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn(*args)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    calls.append({"call": name, "algorithmic_bytes": int(bytes_)})


# ---- projection / PAVA, uniform layouts, 10^8 variables -------------------------------------------------
for K in (4, 16, 64):
    nb = 10 ** 8 // K
    n = nb * K
    plan = bsls_b200.BlockPlan(torch.arange(0, n, K, dtype=torch.int64, device=dev), n)
    mk = lambda n=n: (torch.randn(n, dtype=torch.float64, device=dev, generator=gen),)
    run("proj_multi_simplex K=%d n=1e8" % K, 16 * n + 4 * nb, lambda y, plan=plan: bsls_b200.proj_multi_simplex_c(y, plan), mk)
    if K == 16:
        run("proj_multi_ball K=16 n=1e8", 16 * n + 4 * nb, lambda y, plan=plan: bsls_b200.proj_multi_ball_c(y, plan), mk)
    ramp = 50.0 * torch.log1p(torch.arange(K, dtype=torch.float64, device=dev))
    mkp = lambda nb=nb, K=K, ramp=ramp: ((torch.randint(-50, 50, (nb, K), device=dev, generator=gen).to(torch.float64) + ramp).reshape(-1),)
    run("isotonic_regression_multi K=%d n=1e8" % K, 16 * n + 4 * nb, lambda y, plan=plan: bsls_b200.isotonic_regression_multi_c(y, plan, None, 1), mkp)
    del plan
torch.cuda.empty_cache()

# ---- ragged layout C3 (power law 2..4096, 10^7 variables) ---------------------------------------------------
sizes = bench.power_law_sizes(10 ** 7)
n, nb = int(sizes.sum()), len(sizes)
starts = torch.as_tensor(np.concatenate(([0], np.cumsum(sizes)[:-1]))).to(dev)
plan = bsls_b200.BlockPlan(starts, n)
pos = torch.arange(n, device=dev) - torch.repeat_interleave(starts, torch.as_tensor(sizes).to(dev))
ramp = 50.0 * torch.log1p(pos.to(torch.float64))
run("proj_multi_simplex C3 ragged n=1e7", 16 * n + 4 * nb, lambda y: bsls_b200.proj_multi_simplex_c(y, plan),
    lambda: (torch.randn(n, dtype=torch.float64, device=dev, generator=gen),))
run("isotonic_regression_multi C3 ragged n=1e7", 16 * n + 4 * nb, lambda y: bsls_b200.isotonic_regression_multi_c(y, plan, None, 1),
    lambda: (torch.randint(-50, 50, (n,), device=dev, generator=gen).to(torch.float64) + ramp,))
del plan, pos, ramp
torch.cuda.empty_cache()

# ---- sparse least squares: C5 / 8 (what one of eight GPUs holds): n = 2e7, nnz = 1.6e8, m = 1e6 -----------------
sp = SyntheticProblem(1250000, 16, 1000000, 8, noise=0.1)
P = sp.problem.set_panels()
prob, ws = sp.problem, sp.problem.ws
nn, m, nnz = sp.n, sp.m, sp.nnz
x = sp.x_init.clone()
g = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
run("residual r = A x - b (index-only, %d panels)" % P, 12 * nnz + 8 * (m + 1) + 8 * nn + 16 * m,
    lambda: _lib.check(L.bsls_dev_lsq_residual_f64(prob.handle, x.data_ptr(), st)))
run("gradient g = A^T r (index-only)", 12 * nnz + 8 * (nn + 1) + 8 * m + 8 * nn,
    lambda: _lib.check(L.bsls_dev_lsq_gradient_f64(prob.handle, g.data_ptr(), st)))
parts = sp.solver_parts()
run("BATCH.solve_BB, 3 iterations (update+projection, SpMV pair with BB epilogue)", 3 * (24 * nnz + 72 * nn + 32 * m + 4 * sp.nb),
    lambda: bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=4))
y = torch.randn(nn, dtype=torch.float64, device=dev, generator=gen)
run("axpby n=2e7", 24 * nn, lambda: axpby(g, 1.0, x, -0.5, y))
run("dots (3 pairs) n=2e7", 32 * nn, lambda: ws.dots([(x, y), (y, y), (g, y)]))
run("md_update (exp + block normalise) n=2e7", 24 * nn + 4 * sp.nb, lambda: ws.md_update(sp.plan, g, sp.x_true, y, 0.01))
z = bsls_b200.bsls_utils.x2z(sp.x_true, block_starts=sp.starts)
run("x2z n=2e7", 16 * nn, lambda: bsls_b200.bsls_utils.x2z(sp.x_true, block_starts=sp.starts))
run("z2x n=2e7", 16 * nn, lambda: bsls_b200.bsls_utils.z2x(z, block_starts=sp.starts, n=nn))
N = bsls_b200.bsls_utils.NOperator(sp.starts, nn, _plan=sp.plan)
run("N z n=2e7", 16 * nn, lambda: N.dot(z, g))
run("N^T v n=2e7", 16 * nn, lambda: N.T.dot(y))
print(json.dumps(calls))
