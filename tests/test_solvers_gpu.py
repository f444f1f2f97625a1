"""GPU parity tests of the sparse least-squares / solver surface against (i) the golden
fixtures produced by the reference's own modules and (ii) the pinned CPU oracle.

Tolerances (BASELINE.json north_star): objective values and vectors within 1e-6 relative in
fp64; the SpMV in STREAM mode keeps the reference's left-to-right row sums and is compared
bit for bit."""
import os
from collections import deque

import numpy as np
import pytest
import scipy.sparse as sps

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TAGS = ("c1mini", "k16", "noisy", "noisy20")
SHAPES = {"c1mini": (60, 5), "k16": (20, 16), "noisy": (80, 5), "noisy20": (30, 20)}


@pytest.fixture(scope="module")
def B():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import bsls_b200
    return bsls_b200


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "solvers.npz"))


def problem(gold, tag):
    m, n = gold[tag + "_shape"]
    A = sps.csr_matrix((gold[tag + "_val"], gold[tag + "_idx"], gold[tag + "_ptr"]), shape=(m, n))
    return A, gold[tag + "_b"], gold[tag + "_starts"], gold[tag + "_xinit"]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def random_problem(rng, nb, K, m, L, noise=0.1):
    n = nb * K
    rows = np.concatenate([rng.choice(m, L, replace=False) for _ in range(n)])
    cols = np.repeat(np.arange(n), L)
    A = sps.csr_matrix((np.ones(n * L), (rows, cols)), shape=(m, n))
    x_true = rng.dirichlet(np.ones(K), size=nb).reshape(-1)
    b = A.dot(x_true) + noise * rng.randn(m)
    return A, b, np.arange(0, n, K, dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# objective / gradient (a9)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("modes", [(1, 1), (0, 0), (8, 4), (32, 16)])
@pytest.mark.parametrize("implicit", [False, True])
def test_objective_matches_reference(B, gold, tag, modes, implicit):
    A, b, starts, x0 = problem(gold, tag)
    prob = B.LsqProblem(A, b, implicit_ones=implicit)
    prob.set_modes(*modes)
    x = dev(x0)
    g = torch.zeros_like(x)
    f = prob.obj(x, g)
    assert f == pytest.approx(float(gold[tag + "_f0"]), rel=1e-13)
    if modes == (1, 1):
        # left-to-right row sums, no FMA: identical to scipy's csr_matvec
        assert np.array_equal(host(g), gold[tag + "_g0"])
        assert np.array_equal(host(prob.residual()), A.dot(x0) - b)
    else:
        np.testing.assert_allclose(host(g), gold[tag + "_g0"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("shape", [(2000, 5, 700, 6), (500, 20, 300, 10), (37, 3, 11, 2), (1, 2, 1, 1), (300, 64, 5000, 12)])
def test_objective_matches_oracle_random(B, shape):
    from oracle import solvers_np as S
    rng = np.random.RandomState(4242 + shape[0])
    A, b, starts = random_problem(rng, *shape)
    A.data[:] = rng.rand(A.nnz) + 0.5          # general values, not only an incidence matrix
    x0 = rng.rand(A.shape[1])
    _, _, _, obj = S.get_solver_parts(A, b, starts, 0.1)
    g_ref = np.zeros_like(x0)
    f_ref = obj(x0, g_ref)
    for modes in ((1, 1), (0, 0), (32, 32), (4, 4)):
        prob = B.LsqProblem(A, b)
        prob.set_modes(*modes)
        g = torch.zeros(A.shape[1], dtype=torch.float64, device="cuda")
        f = prob.obj(dev(x0), g)
        assert f == pytest.approx(f_ref, rel=1e-12)
        np.testing.assert_allclose(host(g), g_ref, rtol=1e-11, atol=1e-12)
        if modes == (1, 1):
            assert np.array_equal(host(g), g_ref)
        out = host(prob.matvec(dev(x0)))
        np.testing.assert_allclose(out, A.dot(x0), rtol=1e-12, atol=1e-13)
        w = rng.randn(A.shape[0])
        np.testing.assert_allclose(host(prob.rmatvec(dev(w))), A.T.dot(w), rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("panel_cols", [1, 7, 64, 1000])
def test_column_panelled_product(B, panel_cols):
    """bsls_lsq_set_panels: A x formed panel by panel equals the plain product."""
    rng = np.random.RandomState(21)
    A, b, starts = random_problem(rng, 60, 5, 40, 4)
    A.data[:] = rng.rand(A.nnz) + 0.5
    x0 = rng.rand(A.shape[1])
    r = A.dot(x0) - b
    for implicit in (False, True):
        Ause = A.copy()
        if implicit:
            Ause.data[:] = 1.0
            r = Ause.dot(x0) - b
        prob = B.LsqProblem(Ause, b, implicit_ones=implicit)
        P = prob.set_panels(panel_cols)
        assert P == -(-A.shape[1] // panel_cols)
        g = torch.zeros(A.shape[1], dtype=torch.float64, device="cuda")
        f = prob.obj(dev(x0), g)
        assert f == pytest.approx(.5 * r.dot(r), rel=1e-12)
        np.testing.assert_allclose(host(prob.residual()), r, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(host(g), Ause.T.dot(r), rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(host(prob.matvec(dev(x0))), Ause.dot(x0), rtol=1e-12, atol=1e-13)


def test_empty_rows_and_columns(B):
    rng = np.random.RandomState(5)
    A = sps.random(50, 120, density=0.02, random_state=rng, format="csr")
    b = rng.randn(50)
    x0 = rng.rand(120)
    r = A.dot(x0) - b
    for modes in ((1, 1), (8, 8)):
        prob = B.LsqProblem(A, b)
        prob.set_modes(*modes)
        g = torch.zeros(120, dtype=torch.float64, device="cuda")
        f = prob.obj(dev(x0), g)
        assert f == pytest.approx(.5 * r.dot(r), rel=1e-13)
        np.testing.assert_allclose(host(g), A.T.dot(r), rtol=1e-12, atol=1e-14)


# ---------------------------------------------------------------------------------------------
# vector kernels
# ---------------------------------------------------------------------------------------------
def test_vector_kernels(B):
    from bsls_b200.sparse import axpby, default_workspace
    rng = np.random.RandomState(11)
    for n in (1, 7, 1000, 300001):
        x, y, z = rng.randn(n), rng.randn(n), rng.randn(n)
        out = axpby(torch.empty(n, dtype=torch.float64, device="cuda"), 1.0, dev(x), -0.37, dev(y))
        assert np.array_equal(host(out), x + (-0.37) * y)
        t = 0.8 * 0.8
        xn = dev(y)
        axpby(xn, 1.0 - t, dev(x), t, xn)          # in place, as the line search does
        assert np.array_equal(host(xn), (1.0 - t) * x + t * y)
        ws = default_workspace("cuda")
        d = ws.dots([(dev(x), dev(y)), (dev(y), dev(y)), (dev(x), dev(z))], want_max=True)
        np.testing.assert_allclose(d[:3], [x.dot(y), y.dot(y), x.dot(z)], rtol=1e-11, atol=1e-11)
        assert d[3] == np.max(np.abs(x - y))
        # deterministic: same bits on every run
        assert d == ws.dots([(dev(x), dev(y)), (dev(y), dev(y)), (dev(x), dev(z))], want_max=True)


def test_lbfgs_helper_matches_oracle(B):
    from oracle import solvers_np as S
    from bsls_b200.BATCH import LBFGS_helper
    rng = np.random.RandomState(3)
    n, m = 500, 7
    ys = [rng.randn(n) for _ in range(m)]
    ss = [rng.randn(n) for _ in range(m)]
    rho = [1.0 / y.dot(s) for y, s in zip(ys, ss)]
    g = rng.randn(n)
    d_ref = np.zeros(n)
    S.LBFGS_helper(deque(ys), deque(ss), deque(rho), g, d_ref, np.zeros(50))
    d = torch.zeros(n, dtype=torch.float64, device="cuda")
    alpha = torch.zeros(104, dtype=torch.float64, device="cuda")
    LBFGS_helper(deque(dev(y) for y in ys), deque(dev(s) for s in ss), deque(rho), dev(g), d, alpha)
    np.testing.assert_allclose(host(d), d_ref, rtol=1e-9, atol=1e-10)


# ---------------------------------------------------------------------------------------------
# change of variables (a7) and the N operator (a10)
# ---------------------------------------------------------------------------------------------
def test_x2z_z2x_golden(B, golden_dir):
    d = np.load(os.path.join(golden_dir, "x2z.npz"))
    for k in range(int(d["count"])):
        x, starts, z, xback = d["x%d" % k], d["starts%d" % k], d["z%d" % k], d["xback%d" % k]
        zd = torch.empty(len(z), dtype=torch.float64, device="cuda")
        B.x2z_c(dev(x), zd, starts)
        assert np.array_equal(host(zd), z)
        xd = torch.empty(len(x), dtype=torch.float64, device="cuda")
        B.z2x_c(xd, dev(z), starts)
        assert np.array_equal(host(xd), xback)
        assert np.array_equal(host(B.bsls_utils.x2z(dev(x), block_starts=starts)), z)
    with pytest.raises(AssertionError):
        B.x2z_c(dev(np.ones(4)), torch.empty(2, dtype=torch.float64, device="cuda"), np.array([1, 2]))
    with pytest.raises(AssertionError):
        B.x2z_c(dev(np.ones(4)), torch.empty(2, dtype=torch.float64, device="cuda"), np.array([0, 2, 2]))

@pytest.mark.parametrize("K", [2, 3, 5, 16, 20, 64, 65, 100])
def test_x2z_z2x_uniform_against_oracle(B, K):
    """x2z / z2x on uniform layouts (tiled kernel for K <= 64, thread-per-block beyond) bit for bit against the oracle
    (c_extensions.pyx:195-248): running sums left to right, adjacent differences, last entry 1 - z_last."""
    from oracle import cpu
    port = cpu.port()
    rng = np.random.RandomState(100 + K)
    nb = 1000 + 131 * (K % 7)          # not a multiple of the tile height
    x = rng.dirichlet(np.ones(K), size=nb).reshape(-1) + 1e-3 * rng.randn(nb * K)
    starts = np.arange(0, nb * K, K)
    z = port.x2z(x.copy(), np.zeros(nb * (K - 1)), starts)
    zd = torch.empty(nb * (K - 1), dtype=torch.float64, device="cuda")
    B.x2z_c(dev(x), zd, starts)
    assert np.array_equal(host(zd), z)
    xb = port.z2x(np.zeros(nb * K), z.copy(), starts)
    xd = torch.empty(nb * K, dtype=torch.float64, device="cuda")
    B.z2x_c(xd, dev(z), starts)
    assert np.array_equal(host(xd), xb)


def test_x2z_matches_oracle_ragged(B):
    from oracle import cpu
    rng = np.random.RandomState(8)
    sizes = rng.randint(1, 40, size=3000)
    sizes[::97] = 300
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    n = int(sizes.sum())
    x = rng.rand(n)
    z_ref = np.empty(n - len(sizes))
    cpu.port().x2z(x, z_ref, starts)
    z = B.bsls_utils.x2z(dev(x), block_sizes=sizes)
    assert np.array_equal(host(z), z_ref)
    x_ref = np.empty(n)
    cpu.port().z2x(x_ref, z_ref, starts)
    assert np.array_equal(host(B.bsls_utils.z2x(dev(z_ref), block_sizes=sizes)), x_ref)


@pytest.mark.parametrize("sizes", [[5] * 40, [2, 3, 7, 2, 30, 4, 64, 2], [1, 4, 1, 1, 6], [16] * 1000])
def test_N_operator(B, sizes):
    from oracle import solvers_np as S
    sizes = np.asarray(sizes)
    n = int(sizes.sum())
    m = 9
    rng = np.random.RandomState(2)
    A = sps.random(m, n, density=0.3, random_state=rng, format="csr")
    Nref, x0ref, *_ = S.z_space_closures(A, rng.randn(m), sizes)
    N = B.bsls_utils.block_sizes_to_N(sizes)
    assert N.shape == Nref.shape
    z = rng.randn(Nref.shape[1])
    v = rng.randn(n)
    assert np.array_equal(host(N.dot(dev(z))), Nref.dot(z))
    assert np.array_equal(host(N.T.dot(dev(v))), Nref.T.dot(v))
    assert np.array_equal(host(B.bsls_utils.particular_x0(sizes)), x0ref)


# ---------------------------------------------------------------------------------------------
# BATCH solvers (a13, a14)
# ---------------------------------------------------------------------------------------------
def run_batch(B, name, parts, starts, x0, native):
    step_size, proj, line_search, obj = parts
    if not native:  # strip the handles: forces the generic (closure-driven) loop
        obj_ = lambda x, g=None: obj(x, g)
        proj_ = lambda x: proj(x)
        ls_ = lambda *a: line_search(*a)
        ss_ = lambda i: step_size(i)
    else:
        obj_, proj_, ls_, ss_ = obj, proj, line_search, step_size
    if name == "bb":
        return B.BATCH.solve_BB(obj_, proj_, ls_, x0, max_iter=300)
    if name == "pg":
        return B.BATCH.solve(obj_, proj_, ss_, x0, ls_, max_iter=100)
    if name == "md":
        return B.BATCH.solve_MD(obj_, starts, ss_, x0, max_iter=100)
    return B.BATCH.solve_LBFGS(obj_, proj_, ls_, x0, max_iter=150)


@pytest.fixture(params=["tiny", "cluster", "loop"])
def native_loop(request, monkeypatch):
    """The native BATCH solvers have three device-resident loops: one CTA with everything in shared memory for problems
    that fit, the same on a cluster of 8 CTAs that exchange their slices through distributed shared memory
    (solver_tiny.cuh), and the multi-kernel loop (lsq.cu).  The small test problems qualify for the first two
    (BSLS_TINY_CLUSTER picks); BSLS_NO_TINY=1 sends them through the third."""
    monkeypatch.delenv("BSLS_NO_TINY", raising=False)
    monkeypatch.delenv("BSLS_TINY_CLUSTER", raising=False)
    if request.param == "loop":
        monkeypatch.setenv("BSLS_NO_TINY", "1")
    else:
        monkeypatch.setenv("BSLS_TINY_CLUSTER", "1" if request.param == "cluster" else "0")
    return request.param


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name", ["bb", "pg", "md", "lbfgs"])
@pytest.mark.parametrize("native", [True, False])
def test_batch_solvers_match_reference(B, gold, tag, name, native, native_loop):
    if native_loop != "tiny" and (not native or name == "lbfgs"):
        pytest.skip("the generic / L-BFGS loops do not depend on the native loop kind")
    A, b, starts, x0 = problem(gold, tag)
    parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True)
    parts[3].problem.set_modes(1, 1)
    sol = run_batch(B, name, parts, starts, dev(x0), native)
    key = "%s_%s_" % (tag, name)
    f_ref = float(gold[key + "f"])
    trace = np.array([p[1] for p in sol["progress"]])
    ref_trace = gold[key + "ftrace"]
    if native and name != "lbfgs":
        assert "kernel_launches" in sol and sol["kernel_launches"] > 0
    # objective of the full solve: 1e-6 relative (north_star); a noiseless problem converges to f = 0
    assert sol["f"] == pytest.approx(f_ref, rel=1e-6, abs=1e-10)
    # the first iterations follow the reference's trajectory closely
    k = min(6, len(trace), len(ref_trace))
    np.testing.assert_allclose(trace[:k], ref_trace[:k], rtol=1e-9, atol=1e-12)
    if name in ("pg", "md"):
        # no chaotic step rule: whole trajectory, iteration count, stop reason and solution agree
        assert sol["iterations"] == int(gold[key + "iters"])
        assert sol["stop"].split("=")[0] == str(gold[key + "stop"]).split("=")[0]
        np.testing.assert_allclose(trace, ref_trace, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(host(sol["x"]), gold[key + "x"], atol=1e-7)
    xs = host(sol["x"])
    K = SHAPES[tag][1]
    np.testing.assert_allclose(xs.reshape(-1, K).sum(1), 1.0, atol=1e-9)
    assert xs.min() >= 0.0


def test_batch_solvers_accept_host_vectors(B, gold):
    """The reference's solvers take and return NumPy vectors: host input (ndarray, pinned or pageable CPU tensor) is staged
    to the device, the result comes back the same way and equals the device-vector call."""
    A, b, starts, x0 = problem(gold, "noisy")
    parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True)
    ref = B.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=300)
    for make in (lambda: x0.copy(), lambda: torch.from_numpy(x0.copy()), lambda: torch.from_numpy(x0.copy()).pin_memory()):
        xin = make()
        keep = xin.copy() if isinstance(xin, np.ndarray) else xin.clone()
        sol = B.BATCH.solve_BB(parts[3], parts[1], parts[2], xin, max_iter=300)
        out = sol["x"]
        assert isinstance(out, np.ndarray) if isinstance(xin, np.ndarray) else (torch.is_tensor(out) and not out.is_cuda)
        assert np.array_equal(np.asarray(out), host(ref["x"])) and sol["f"] == ref["f"]
        assert np.array_equal(np.asarray(xin), np.asarray(keep))            # the caller's start vector is not touched
    for name in ("pg", "md", "lbfgs"):
        a = run_batch(B, name, parts, starts, dev(x0), True)
        h = run_batch(B, name, parts, starts, x0.copy(), True)
        assert isinstance(h["x"], np.ndarray) and np.array_equal(h["x"], host(a["x"]))


def test_batch_solvers_float32_inputs(B, gold):
    """north_star's fp32 bar: projected vectors and objectives within 1e-4 relative.  The solvers take float32 problems
    and start vectors (NumPy, CPU tensor, CUDA tensor), widen them on the way in -- the loop is bound by 8-byte gathers
    that cost a 32-byte sector each whatever the element size -- and return x as float32.  Checked against the float64
    ORACLE run on the float32-rounded data, and against the float64 run on the original data at 1e-4."""
    from oracle import solvers_np as S
    A, b, starts, x0 = problem(gold, "noisy")
    A32, b32, x32 = A.astype(np.float32), b.astype(np.float32), x0.astype(np.float32)
    parts64 = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True)
    ref64 = B.BATCH.solve_BB(parts64[3], parts64[1], parts64[2], dev(x0), max_iter=300)
    parts = B.algorithm_utils.get_solver_parts((A32, b32), starts, 0.1, is_sparse=True)
    op = S.get_solver_parts(A32.astype(np.float64), b32.astype(np.float64), starts, 0.1)
    want = S.solve_BB(op[3], op[1], op[2], x32.astype(np.float64), max_iter=300)
    for make in (lambda: x32.copy(), lambda: torch.from_numpy(x32.copy()), lambda: torch.from_numpy(x32.copy()).cuda()):
        xin = make()
        sol = B.BATCH.solve_BB(parts[3], parts[1], parts[2], xin, max_iter=300)
        out = sol["x"]
        if isinstance(xin, np.ndarray):
            assert isinstance(out, np.ndarray) and out.dtype == np.float32
        else:
            assert torch.is_tensor(out) and out.dtype == torch.float32 and out.is_cuda == xin.is_cuda
        got = host(out) if torch.is_tensor(out) else out
        assert sol["f"] == pytest.approx(want["f"], rel=1e-6, abs=1e-10)
        np.testing.assert_allclose(got, want["x"], rtol=1e-4, atol=1e-6)
        assert sol["f"] == pytest.approx(ref64["f"], rel=1e-4, abs=1e-8)
        np.testing.assert_allclose(got, host(ref64["x"]), rtol=1e-4, atol=1e-4)
    for name in ("pg", "md"):
        h = run_batch(B, name, parts, starts, x32.copy(), True)
        a = run_batch(B, name, parts64, starts, dev(x0), True)
        assert h["x"].dtype == np.float32
        np.testing.assert_allclose(h["x"], host(a["x"]), rtol=1e-4, atol=1e-4)


def test_small_qp_known_answer(B, gold):
    """tests/fast/test_BATCH.py of the reference: 2-variable QP, solution [.25, .75], f_min 1.875."""
    Q, c, x_true, f_min, min_eig = B.bsls_utils.generate_small_qp()
    step_size, proj, line_search, obj = B.algorithm_utils.get_solver_parts((Q, c), np.array([0]), min_eig)
    x0 = dev([.5, .5])
    sol = B.BATCH.solve_BB(obj, proj, line_search, x0)
    np.testing.assert_allclose(host(sol["x"]), x_true, atol=1e-3)
    assert sol["stop"].split("=")[0] == str(gold["qp_bb_stop"]).split("=")[0]
    np.testing.assert_allclose(host(sol["x"]), gold["qp_bb_x"], atol=1e-9)
    sol = B.BATCH.solve_BB(obj, proj, line_search, x0, f_min=f_min)
    assert sol["stop"].split("=")[0] == str(gold["qp_bb_fmin_stop"]).split("=")[0]
    sol = B.BATCH.solve(obj, proj, step_size, x0, line_search)
    np.testing.assert_allclose(host(sol["x"]), x_true, atol=1e-3)
    sol = B.BATCH.solve_LBFGS(obj, proj, line_search, x0)
    np.testing.assert_allclose(host(sol["x"]), x_true, atol=1e-3)
    sol = B.BATCH.solve_MD(obj, np.array([0]), step_size, x0)
    np.testing.assert_allclose(host(sol["x"]), x_true, atol=1e-2)


@pytest.mark.parametrize("seed", range(3))
def test_batch_random_least_squares_vs_oracle(B, seed):
    """The reference's own solver test shape (tests/fast/test_BATCH.py: random_least_squares(10, 7)
    generalised): GPU and oracle reach the same objective from the same start."""
    from oracle import solvers_np as S
    rng = np.random.RandomState(100 + seed)
    A, b, starts = random_problem(rng, 150, 7, 260, 5, noise=0.2)
    x0 = np.ones(A.shape[1]) / 7
    ref = S.solve_BB(*[S.get_solver_parts(A, b, starts, 0.1)[k] for k in (3, 1, 2)], x0, max_iter=400)
    parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True)
    sol = B.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=400)
    assert sol["f"] == pytest.approx(ref["f"], rel=1e-6)
    # lasso (l1-ball) feasible set
    ref = S.solve_BB(*[S.get_solver_parts(A, b, starts, 0.1, lasso=True)[k] for k in (3, 1, 2)], x0, max_iter=400)
    parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True, lasso=True)
    sol = B.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=400)
    assert sol["f"] == pytest.approx(ref["f"], rel=1e-6)


def test_batch_in_z_and_scaled(B):
    """get_solver_parts(in_z=True) (PAVA + clip) and f-scaled projections (algorithm_utils.py:219-265)."""
    from oracle import solvers_np as S
    rng = np.random.RandomState(77)
    nb, K = 40, 6
    A, b, starts = random_problem(rng, nb, K, 90, 4, noise=0.2)
    # problem posed in z: Az = A N, bz = b - A x0  (ls_to_ls_in_z, bsls_utils.py:251-264)
    sizes = np.full(nb, K)
    N, x0p, *_ = S.z_space_closures(A, b, sizes)
    Az = sps.csr_matrix(A.dot(N))
    bz = b - A.dot(x0p)
    z0 = np.concatenate([np.cumsum(np.ones(K) / K)[:-1] for _ in range(nb)])
    ref_parts = S.get_solver_parts(Az, bz, starts, 0.1, in_z=True)
    ref = S.solve_BB(ref_parts[3], ref_parts[1], ref_parts[2], z0, max_iter=300)
    for native in (True, False):
        parts = B.algorithm_utils.get_solver_parts((Az, bz), starts, 0.1, is_sparse=True, in_z=True)
        sol = run_batch(B, "bb", parts, starts, dev(z0), native)
        assert sol["f"] == pytest.approx(ref["f"], rel=1e-6)
    # f-scaled simplex: blocks sum to f_k
    fk = rng.rand(nb) + 0.5
    parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True, f=fk)
    y = rng.randn(nb * K)
    yd = dev(y)
    parts[1](yd)
    from oracle import cpu
    want = y.copy()
    for k in range(nb):
        want[k * K:(k + 1) * K] /= fk[k]
    cpu.port().proj_multi_simplex(want, starts)
    for k in range(nb):
        want[k * K:(k + 1) * K] *= fk[k]
    np.testing.assert_allclose(host(yd), want, rtol=1e-14, atol=1e-15)


# ---------------------------------------------------------------------------------------------
# functional drivers in z (a10-a12, a15, a16, a19) and mirror descent (a17)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", TAGS)
def test_z_space_drivers_match_reference(B, gold, tag):
    A, b, starts, xinit = problem(gold, tag)
    nb, K = SHAPES[tag]
    sizes = np.full(nb, K, dtype=np.int64)
    z0, target, f, nabla_f, proj, prob, N = B.main.z_space_parts(A, b, host(B.bsls_utils.particular_x0(sizes)), None, sizes)
    # the fixtures start from x2z(x_init), not from x0
    z_init = B.bsls_utils.x2z(dev(xinit), block_sizes=sizes)
    np.testing.assert_allclose(host(z_init), gold[tag + "_z0"], rtol=0, atol=0)
    log = lambda it, state, dur: 0.0
    opts = {"max_iter": 200, "opt_tol": 1e-30, "verbose": 0}
    # BB in z: the library's device-resident loop (bsls_zbb_run_f64) and the generic closure-driven loop
    zbb = B.BB.solve(z_init.clone(), f, nabla_f, B.solvers.stopping, proj=proj, log=log, options=opts)
    assert B.BB.solve.last["kernel_launches"] > 0 and B.BB.solve.last["iterations"] <= 200
    assert f(zbb) == pytest.approx(float(gold[tag + "_zbb_f"]), rel=1e-6, abs=1e-10)
    zbb_g = B.BB.solve(z_init.clone(), f, nabla_f, B.solvers.stopping, proj=proj, log=log, options=dict(opts, generic_loop=True))
    assert f(zbb_g) == pytest.approx(float(gold[tag + "_zbb_f"]), rel=1e-6, abs=1e-10)
    # states are recorded every `record_every` iterations plus the first and the last, as the reference's log does
    seen = []
    B.BB.solve(z_init.clone(), f, nabla_f, B.solvers.stopping, proj=proj, log=lambda it, state, dur: seen.append((it, state.clone())) or 0.0,
               options={"max_iter": 35, "opt_tol": 1e-30, "verbose": 0}, record_every=10)
    assert [it for it, _ in seen] == [0, 10, 20, 30, 35]
    seen_g = []
    B.BB.solve(z_init.clone(), f, nabla_f, B.solvers.stopping, proj=proj, log=lambda it, state, dur: seen_g.append((it, state.clone())) or 0.0,
               options={"max_iter": 35, "opt_tol": 1e-30, "verbose": 0, "generic_loop": True}, record_every=10)
    assert [it for it, _ in seen_g] == [0, 10, 20, 30, 35]
    for (_, a), (_, b_) in zip(seen[:3], seen_g[:3]):      # the first iterations follow the same trajectory
        np.testing.assert_allclose(host(a), host(b_), rtol=1e-9, atol=1e-12)
    ones = torch.ones_like(z_init)
    zl = B.LBFGS.solve(z_init + ones, f, nabla_f, B.solvers.stopping, proj=proj, log=log,
                       options={"max_iter": 40, "opt_tol": 1e-30, "verbose": 0})
    assert f(zl) == pytest.approx(float(gold[tag + "_zlbfgs_f"]), rel=1e-6, abs=1e-10)
    lsv = B.bsls_utils.lsv_operator(prob, N)
    assert lsv == pytest.approx(float(gold[tag + "_lsv"]), rel=1e-9)
    gd = B.gradient_descent.GradientDescent(z0=z_init.clone(), f=f, nabla_f=nabla_f, proj=proj, method="DORE",
                                            options={"max_iter": 150, "opt_tol": 1e-30, "verbose": 0}, A=prob, N=N, target=target)
    iters, times, states = gd.run()
    assert f(states[-1]) == pytest.approx(float(gold[tag + "_zdore_f"]), rel=1e-6, abs=1e-10)
    np.testing.assert_allclose(host(states[-1]), gold[tag + "_zdore_z"], atol=1e-6)


@pytest.mark.parametrize("tag", TAGS)
def test_mirror_descent_least_squares(B, gold, tag):
    A, b, starts, xinit = problem(gold, tag)
    nb, K = SHAPES[tag]
    Lf = float(gold[tag + "_Lf"])
    x = B.mirror_descent.least_squares(A, b, [K] * nb, iters=60, tolerance=1e-9, Lf=Lf)
    np.testing.assert_allclose(host(x), gold[tag + "_md_ls_x"], rtol=1e-9, atol=1e-12)
    # the library's own Lanczos estimate of sigma_max(A) agrees with ARPACK's
    assert B.bsls_utils.largest_singular_value(A) == pytest.approx(Lf, rel=1e-9)


def test_mirror_descent_ragged_blocks(B):
    from oracle import solvers_np as S
    rng = np.random.RandomState(31)
    sizes = [2, 5, 3, 17, 40, 2, 2, 9] * 20
    n = sum(sizes)
    A = sps.random(60, n, density=0.05, random_state=rng, format="csr")
    A.data[:] = 1.0
    b = rng.rand(60)
    Lf = float(sps.linalg.svds(A, 1, return_singular_vectors=False)[0])
    ref = S.md_least_squares(A, b, sizes, iters=40, Lf=Lf)
    x = B.mirror_descent.least_squares(A, b, sizes, iters=40, Lf=Lf)
    np.testing.assert_allclose(host(x), ref, rtol=1e-9, atol=1e-13)


def test_block_isotonic_regression_drop_in(B):
    from oracle import cpu
    rng = np.random.RandomState(9)
    sizes = rng.randint(2, 30, size=200)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    y = rng.randn(int(sizes.sum()))
    want = y.copy()
    cpu.port().pava_multi(want, starts)
    yd = dev(y)
    B.block_isotonic_regression.block_isotonic_regression_2(yd, starts)
    assert np.array_equal(host(yd), want)


def test_block_isotonic_regression_new_array_and_clip(B):
    """block_isotonic_regression (block_isotonic_regression.py:8-16): a NEW vector, z-blocks of block_sizes - 1 entries
    regressed and box-clipped to [0, 1]; blocks of one route have no z entries and are dropped.  Checked against the
    reference's own formulation (scikit-learn per block, 1e-8 as in tests/fast/test_isotonic_regression.py) and,
    bit for bit, against the oracle's PAVA + clip."""
    from oracle import cpu
    from sklearn.isotonic import IsotonicRegression
    ir = IsotonicRegression()
    rng = np.random.RandomState(11)
    block_sizes = rng.randint(1, 25, size=300)                # includes single-route blocks (empty z-blocks)
    zsizes = block_sizes - 1
    blocks_start = np.concatenate(([0], np.cumsum(zsizes)[:-1]))
    blocks_end = np.cumsum(zsizes)
    x = 0.5 + 0.6 * rng.randn(int(zsizes.sum()))              # values on both sides of [0, 1]
    box = lambda y: np.maximum(np.minimum(y, 1), 0)
    ref = np.concatenate([ir.fit_transform(np.arange(k), x[s:e]) if k > 1 else box(x[s:e])
                          for k, s, e in zip(zsizes, blocks_start, blocks_end) if k > 0])
    ref = box(ref)
    xd = dev(x)
    out = B.block_isotonic_regression.block_isotonic_regression(xd, None, block_sizes, blocks_start, blocks_end)
    assert out.data_ptr() != xd.data_ptr() and np.array_equal(host(xd), x)          # input untouched
    assert np.abs(host(out) - ref).max() < 1e-8
    want = x.copy()
    cpu.port().pava_multi(want, blocks_start[zsizes > 0])
    cpu.port().clip01(want)
    assert np.array_equal(host(out), want)
    # torch tensors for the layout arguments work too
    out2 = B.block_isotonic_regression.block_isotonic_regression(xd, None, torch.as_tensor(block_sizes), torch.as_tensor(blocks_start),
                                                                 torch.as_tensor(blocks_end))
    assert torch.equal(out, out2)


def test_f64_only_entry_points_reject_f32(B):
    """x2z / z2x / N / N^T / block_scale have fp64 kernels only: an fp32 buffer must raise the reference's dtype error
    instead of being read as 8-byte elements."""
    starts = np.arange(0, 40, 4)
    x64 = torch.rand(40, dtype=torch.float64, device="cuda")
    z64 = torch.empty(30, dtype=torch.float64, device="cuda")
    x32, z32 = x64.float(), z64.float()
    cx = B.c_extensions
    for call in (lambda: cx.x2z_c(x32, z64, starts), lambda: cx.x2z_c(x64, z32, starts), lambda: cx.z2x_c(x32, z64, starts),
                 lambda: cx.z2x_c(x64, z32, starts), lambda: cx.n_dot(x32, z64, starts), lambda: cx.n_dot(x64, z32, starts),
                 lambda: cx.nt_dot(z32, x64, starts), lambda: cx.nt_dot(z64, x32, starts),
                 lambda: cx.block_scale(x32, starts, torch.ones(10, dtype=torch.float64, device="cuda"))):
        with pytest.raises(ValueError, match="dtype mismatch"):
            call()
    cx.x2z_c(x64, z64, starts)   # the fp64 call still works


# ---------------------------------------------------------------------------------------------
# BASELINE configs as parity cases: C1 at full size, C4 / C5 shapes reduced in size
# ---------------------------------------------------------------------------------------------
def config_problem(nb, K, m, L, seed, noise=0.1):
    rng = np.random.RandomState(seed)
    n = nb * K
    base = np.sort(rng.randint(0, m - L + 1, size=(n, L)), axis=1) + np.arange(L)   # L distinct links per route
    A = sps.csr_matrix((np.ones(n * L), (base.reshape(-1), np.repeat(np.arange(n), L))), shape=(m, n))
    x_true = rng.dirichlet(np.ones(K), size=nb).reshape(-1)
    b = A.dot(x_true) + noise * rng.randn(m)
    return A, b, np.arange(0, n, K, dtype=np.int64)


def test_config1_bb_full_size(B):
    """BASELINE config 1: 1,000 OD blocks x 5 routes, 2,000 links, BB with proj_simplex (the reference path)."""
    from oracle import solvers_np as S
    A, b, starts = config_problem(1000, 5, 2000, 10, 237423433 + 1)
    x0 = np.ones(5000) / 5
    ref_parts = S.get_solver_parts(A, b, starts, 0.1)
    ref = S.solve_BB(ref_parts[3], ref_parts[1], ref_parts[2], x0, max_iter=2000)
    for implicit in (False, True):
        parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True, implicit_ones=implicit)
        sol = B.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=2000)
        assert sol["f"] == pytest.approx(ref["f"], rel=1e-6)
        k = min(8, len(sol["progress"]), len(ref["progress"]))
        np.testing.assert_allclose([p[1] for p in sol["progress"][:k]], [p[1] for p in ref["progress"][:k]], rtol=1e-9)
        xs = host(sol["x"])
        np.testing.assert_allclose(xs.reshape(-1, 5).sum(1), 1.0, atol=1e-9)


@pytest.mark.parametrize("nb,K,m,L,ragged", [(1000, 6, 2500, 6, False), (700, 9, 1500, 5, True), (40, 33, 300, 4, True)])
def test_cluster_solver_shapes(B, monkeypatch, nb, K, m, L, ragged):
    """The cluster loop (solver_cluster.cuh) on shapes around its limits: a problem whose vectors do not fit ONE CTA's
    shared memory but fit a cluster's shares (the environment switch is not consulted there), ragged blocks whose
    column shares start in the middle of a 32-column group, and blocks longer than a group (halo of a share).  Checked
    against the oracle's solve and, to 1e-9 on the objective trace, against the multi-kernel loop."""
    from oracle import solvers_np as S
    rng = np.random.RandomState(nb + K)
    if ragged:
        sizes = rng.randint(max(2, K - 7), K + 1, size=nb)
        sizes[rng.randint(nb)] = min(64, K + 20)
    else:
        sizes = np.full(nb, K)
    n = int(sizes.sum())
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
    base = np.sort(rng.randint(0, m - L + 1, size=(n, L)), axis=1) + np.arange(L)
    A = sps.csr_matrix((np.ones(n * L), (base.reshape(-1), np.repeat(np.arange(n), L))), shape=(m, n))
    x_true = np.concatenate([rng.dirichlet(np.ones(k)) for k in sizes])
    b = A.dot(x_true) + 0.1 * rng.randn(m)
    x0 = np.concatenate([np.full(k, 1.0 / k) for k in sizes])
    ref_parts = S.get_solver_parts(A, b, starts, 0.1)
    ref = S.solve_BB(ref_parts[3], ref_parts[1], ref_parts[2], x0, max_iter=400)
    parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True)
    monkeypatch.delenv("BSLS_NO_TINY", raising=False)
    monkeypatch.setenv("BSLS_TINY_CLUSTER", "1")
    sol = B.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=400)
    monkeypatch.setenv("BSLS_NO_TINY", "1")
    loop = B.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=400)
    assert sol["kernel_launches"] == 1 and loop["kernel_launches"] > 1      # one launch of the cluster; the multi-kernel loop
    assert sol["f"] == pytest.approx(ref["f"], rel=1e-6, abs=1e-10)
    assert loop["f"] == pytest.approx(ref["f"], rel=1e-6, abs=1e-10)
    k = min(8, len(sol["progress"]), len(ref["progress"]), len(loop["progress"]))
    np.testing.assert_allclose([p[1] for p in sol["progress"][:k]], [p[1] for p in ref["progress"][:k]], rtol=1e-9)
    np.testing.assert_allclose([p[1] for p in sol["progress"][:k]], [p[1] for p in loop["progress"][:k]], rtol=1e-9)
    xs = host(sol["x"])
    assert xs.min() >= 0.0
    np.testing.assert_allclose(np.add.reduceat(xs, starts), 1.0, atol=1e-9)


def test_config4_shape_md_and_lbfgs(B):
    """BASELINE config 4 shape (K = 20, L = 10; 1/10 of the blocks and links): mirror descent and L-BFGS."""
    from oracle import solvers_np as S
    nb, K, m, L = 10000, 20, 5000, 10
    A, b, starts = config_problem(nb, K, m, L, 237423433 + 4)
    Lf = float(sps.linalg.svds(A, 1, return_singular_vectors=False)[0])
    ref = S.md_least_squares(A, b, [K] * nb, iters=15, Lf=Lf)
    x = B.mirror_descent.least_squares(A, b, [K] * nb, iters=15, Lf=Lf)
    np.testing.assert_allclose(host(x), ref, rtol=1e-9, atol=1e-14)
    assert B.bsls_utils.largest_singular_value(A) == pytest.approx(Lf, rel=1e-8)
    x0 = np.ones(nb * K) / K
    ref_parts = S.get_solver_parts(A, b, starts, 0.1)
    ref = S.solve_LBFGS(ref_parts[3], ref_parts[1], ref_parts[2], x0, max_iter=12)
    parts = B.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True)
    sol = B.BATCH.solve_LBFGS(parts[3], parts[1], parts[2], dev(x0), max_iter=12)
    np.testing.assert_allclose([p[1] for p in sol["progress"]], [p[1] for p in ref["progress"]], rtol=1e-7)
    ref = S.solve_MD(ref_parts[3], starts, ref_parts[0], x0, max_iter=10)
    sol = B.BATCH.solve_MD(parts[3], starts, parts[0], dev(x0), max_iter=10)
    np.testing.assert_allclose([p[1] for p in sol["progress"]], [p[1] for p in ref["progress"]], rtol=1e-9)
    np.testing.assert_allclose(host(sol["x"]), ref["x"], rtol=1e-9, atol=1e-14)


def test_config5_shape_bb_with_panels(B):
    """BASELINE config 5 shape (K = 16, L = 8; 1/500 of the blocks): the BB solve with the column-panelled
    product of A (several panels) reaches the objective of the oracle's BATCH.solve_BB."""
    from oracle import solvers_np as S
    nb, K, m, L = 20000, 16, 2000, 8
    A, b, starts = config_problem(nb, K, m, L, 237423433 + 5)
    x0 = np.ones(nb * K) / K
    ref_parts = S.get_solver_parts(A, b, starts, 0.1)
    ref = S.solve_BB(ref_parts[3], ref_parts[1], ref_parts[2], x0, max_iter=200)
    prob = B.LsqProblem(A, b, implicit_ones=True)
    assert prob.set_panels(panel_cols=50000) == 7
    parts = B.algorithm_utils.get_solver_parts(prob, starts, 0.1)
    sol = B.BATCH.solve_BB(parts[3], parts[1], parts[2], dev(x0), max_iter=200)
    assert sol["f"] == pytest.approx(ref["f"], rel=1e-6)
    k = min(6, len(sol["progress"]), len(ref["progress"]))
    np.testing.assert_allclose([p[1] for p in sol["progress"][:k]], [p[1] for p in ref["progress"][:k]], rtol=1e-9)


def test_synthetic_generator_matches_its_own_matrix(B):
    """generate.SyntheticProblem builds CSR(A) and CSR(A^T) on the device: the two sides describe the
    same matrix, columns have L distinct links, b = A x_true."""
    from bsls_b200.generate import SyntheticProblem
    sp = SyntheticProblem(300, 16, 500, 8, implicit_ones=True)
    a_ptr, a_idx = host_i(sp.problem.a_ptr), host_i(sp.problem.a_idx)
    t_ptr, t_idx = host_i(sp.problem.t_ptr), host_i(sp.problem.t_idx)
    n, m = sp.n, sp.m
    A1 = sps.csr_matrix((np.ones(len(a_idx)), a_idx, a_ptr), shape=(m, n))
    A2 = sps.csr_matrix((np.ones(len(t_idx)), t_idx, t_ptr), shape=(n, m)).T.tocsr()
    assert (A1 != A2).nnz == 0
    assert np.all(np.diff(t_ptr) == 8) and A1.max() == 1.0
    np.testing.assert_allclose(host(sp.b), A1.dot(host(sp.x_true)), rtol=1e-12)
    xs = host(sp.x_true).reshape(-1, 16)
    np.testing.assert_allclose(xs.sum(1), 1.0, atol=1e-12)


def host_i(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("n,m1,m2", [(100, 5, 10), (100, 20, 10), (1000, 50, 100)])
def test_solve_in_z_on_generate_data(B, n, m1, m2):
    """The reference's end-to-end tests (tests/fast/test_main.py:31-47, tests/slow/test_main.py:32-36):
    generate_data -> solve_in_z with BB drives 0.5 |A x - b|^2 to (numerically) zero."""
    np.random.seed(237423433)
    data = B.bsls_utils.generate_data(n=n, m1=m1, m2=m2, scale=False)
    A, b, sizes = sps.csr_matrix(data['A']), data['b'], data['block_sizes']
    x0 = host(B.bsls_utils.particular_x0(sizes))
    iters, times, states = B.main.solve_in_z(A, b, x0, None, sizes, 'BB',
                                             options={'max_iter': 30000, 'opt_tol': 1e-30, 'verbose': 0})
    z = states[-1]
    N = B.bsls_utils.block_sizes_to_N(sizes)
    x = host(N.dot(z)) + x0
    assert np.all(x >= -1e-12)
    np.testing.assert_allclose(x.reshape(-1)[np.cumsum(sizes) - 1] + 0, x[np.cumsum(sizes) - 1])
    r = A.dot(x) - b
    assert 0.5 * r.dot(r) < 1e-12
    for s, e in zip(data['block_starts'], np.append(data['block_starts'][1:], n)):
        assert abs(x[s:e].sum() - 1.0) < 1e-9
