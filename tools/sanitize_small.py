"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): a few tiles each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse as sps
import bsls_b200
from oracle import cpu
rng = np.random.RandomState(3)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()
chk = cpu.port()
# uniform projections (sort + select kernels, dense fallback), ball mode
for K in (4, 16, 20, 64, 128, 300):
    nb = 700
    for kind in ("normal", "dense"):
        y = rng.randn(nb * K) if kind == "normal" else (rng.dirichlet(np.ones(K), size=nb).reshape(-1) + 1e-3 * rng.randn(nb * K))
        starts = np.arange(0, nb * K, K)
        for ball in (False, True):
            want = y.copy(); (chk.proj_multi_ball if ball else chk.proj_multi_simplex)(want, starts)
            got = dev(y); (bsls_b200.proj_multi_ball_c if ball else bsls_b200.proj_multi_simplex_c)(got, torch.as_tensor(starts).cuda())
            assert np.array_equal(got.cpu().numpy(), want), (K, kind, ball)
# ragged projection + PAVA
sizes = np.concatenate([rng.randint(1, 40, size=900), rng.randint(33, 600, size=40), [700, 2000]]); rng.shuffle(sizes)
starts = np.concatenate(([0], np.cumsum(sizes)[:-1])); n = int(sizes.sum())
for kind in ("normal", "dense"):
    y = rng.randn(n) if kind == "normal" else np.concatenate([rng.dirichlet(np.ones(k)) for k in sizes]) + 1e-4 * rng.randn(n)
    want = y.copy(); chk.proj_multi_simplex(want, starts)
    got = dev(y); bsls_b200.proj_multi_simplex_c(got, torch.as_tensor(starts).cuda())
    assert np.array_equal(got.cpu().numpy(), want), kind
yp = rng.randint(-50, 50, size=n).astype(np.float64)
want = yp.copy(); wref = chk.pava_multi(want, starts)
got = dev(yp); w = torch.ones(n, dtype=torch.int32, device="cuda")
bsls_b200.isotonic_regression_multi_c(got, torch.as_tensor(starts).cuda(), w, 1)
assert np.array_equal(got.cpu().numpy(), want)
for K in (5, 16, 64, 100):
    nb = 600; yp = rng.randint(-50, 50, size=nb * K).astype(np.float64); st = np.arange(0, nb * K, K)
    want = yp.copy(); chk.pava_multi(want, st)
    got = dev(yp); bsls_b200.isotonic_regression_multi_c(got, torch.as_tensor(st).cuda(), None, 1, clip01=False)
    assert np.array_equal(got.cpu().numpy(), want), K
# sparse least squares + solvers
nb, K, m, L = 300, 16, 200, 8
nn = nb * K
base = np.sort(rng.randint(0, m - L + 1, size=(nn, L)), axis=1) + np.arange(L)
A = sps.csr_matrix((np.ones(nn * L), (base.reshape(-1), np.repeat(np.arange(nn), L))), shape=(m, nn))
b = A.dot(rng.dirichlet(np.ones(K), size=nb).reshape(-1)) + 0.1 * rng.randn(m)
st = np.arange(0, nn, K)
prob = bsls_b200.LsqProblem(A, b, implicit_ones=True); prob.set_panels(panel_cols=1000)
parts = bsls_b200.algorithm_utils.get_solver_parts(prob, st, 0.1)
x0 = dev(np.ones(nn) / K)
print("BB", bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], x0, max_iter=30)["f"])
print("MD", bsls_b200.BATCH.solve_MD(parts[3], st, parts[0], x0, max_iter=10)["f"])
print("LBFGS", bsls_b200.BATCH.solve_LBFGS(parts[3], parts[1], parts[2], x0, max_iter=12)["f"])
z = bsls_b200.bsls_utils.x2z(x0, block_starts=st); bsls_b200.bsls_utils.z2x(z, block_starts=st, n=nn)
torch.cuda.synchronize(); print("sanitize_small ok")
