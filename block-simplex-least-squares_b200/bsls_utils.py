"""Hot-path helpers of the reference's ``python/bsls_utils.py`` on the GPU: the x <-> z change
of variables, the bidiagonal N as an operator, the particular solution x0, the largest
singular value of A N (for DORE's scaling) and the small synthetic problems of the
reference's tests.  IO / plotting / analysis helpers of that file are out of scope."""
import numpy as np
import torch

from . import c_extensions as cx
from .plan import BlockPlan
from .sparse import LsqProblem, axpby, default_workspace

__all__ = ["generate_data", "x2z", "z2x", "block_sizes_to_N", "block_starts_to_N", "block_starts_to_x0", "particular_x0", "lsv_operator",
           "largest_singular_value", "generate_small_qp", "random_least_squares", "block_starts_to_block_sizes"]


def _host(a):
    return np.asarray(a.cpu() if torch.is_tensor(a) else a)


def block_starts_to_block_sizes(block_starts, n):
    """bsls_utils.py:111-118"""
    block_starts = _host(block_starts)
    assert False not in ((block_starts[1:] - block_starts[:-1]) > 0)
    assert block_starts[0] == 0 and block_starts[-1] < n
    return np.append(block_starts[1:], [n]) - block_starts


def _starts_from_sizes(block_sizes):
    sizes = _host(block_sizes).astype(np.int64)
    return np.concatenate(([0], np.cumsum(sizes)[:-1])), int(sizes.sum())


def x2z(x, block_sizes=None, block_starts=None, lasso=False):
    """Convert x (original splits) to z (eliminated equality constraint): per-block running sums
    without the last entry (bsls_utils.py:267-287).  Returns a new device vector."""
    assert block_sizes is not None or block_starts is not None
    assert not lasso, "lasso z-space keeps all entries; not on the hot path"
    n = x.shape[0]
    if block_starts is None:
        block_starts, total = _starts_from_sizes(block_sizes)
        assert total == n
    starts = _host(block_starts)
    z = torch.empty(n - len(starts), dtype=torch.float64, device=x.device)
    return cx.x2z_c(x, z, starts)


def z2x(z, block_sizes=None, block_starts=None, n=None):
    """Inverse of :func:`x2z` (c_extensions.pyx:223-248): new device vector x."""
    if block_starts is None:
        block_starts, n = _starts_from_sizes(block_sizes)
    starts = _host(block_starts)
    x = torch.empty(int(n), dtype=torch.float64, device=z.device)
    return cx.z2x_c(x, z, starts)


class NOperator:
    """The matrix N of ``x = x0 + N z`` (bsls_utils.py:139-162) as an operator on device
    vectors: ``N.dot(z)``, ``N.T.dot(v)``, ``N.shape``.  Never materialised."""

    def __init__(self, block_starts, n, device=None, transposed=False, _plan=None):
        self._starts = _host(block_starts).astype(np.int64)
        self._n = int(n)
        self._plan = _plan if _plan is not None else BlockPlan(self._starts, self._n, device)
        self._transposed = transposed
        nz = self._n - len(self._starts)
        self.shape = (nz, self._n) if transposed else (self._n, nz)

    @property
    def T(self):
        return NOperator(self._starts, self._n, transposed=not self._transposed, _plan=self._plan)

    @property
    def plan(self):
        return self._plan

    def dot(self, v, out=None):
        if self._transposed:
            if out is None:
                out = torch.empty(self.shape[0], dtype=torch.float64, device=v.device)
            return cx.nt_dot(out, v, self._plan)
        if out is None:
            out = torch.empty(self.shape[0], dtype=torch.float64, device=v.device)
        return cx.n_dot(out, v, self._plan)

    def tocsr(self):
        return self


def block_sizes_to_N(block_sizes, device=None):
    """bsls_utils.py:139-162"""
    starts, n = _starts_from_sizes(block_sizes)
    return NOperator(starts, n, device)


def block_starts_to_N(block_starts, n, lasso=False, device=None):
    """bsls_utils.py:165-188 (simplex case)"""
    assert not lasso
    return NOperator(block_starts, n, device)


def block_starts_to_x0(block_starts, n, f=None, device=None):
    """x0 with f_k (default 1) at the last entry of every block (bsls_utils.py:121-129)."""
    starts = _host(block_starts)
    if f is None:
        f = np.ones(starts.shape[0])
    x0 = np.zeros(n)
    x0[starts[1:] - 1] = np.asarray(f)[:-1]
    x0[n - 1] = np.asarray(f)[-1]
    return torch.as_tensor(x0).to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))


def particular_x0(block_sizes, device=None):
    """bsls_utils.py:327-328: e_{K-1} in every block."""
    starts, n = _starts_from_sizes(block_sizes)
    return block_starts_to_x0(starts, n, device=device)


# ---------------------------------------------------------------------------------------------
# largest singular values (the reference calls ARPACK: bsls_utils.py:334-369, mirror_descent.py:19-24)
# ---------------------------------------------------------------------------------------------
def _lanczos_largest(apply_M, n, device, max_steps=80, tol=1e-13):
    """Largest eigenvalue of the symmetric PSD operator ``apply_M`` by the Lanczos iteration;
    all vector work is library kernels, the tiny tridiagonal eigenproblem is solved on the host."""
    ws = default_workspace(device)
    gen = np.random.RandomState(0)
    v = torch.as_tensor(gen.rand(n) + 0.5).to(device)
    nv = ws.norm(v)
    axpby(v, 1.0 / nv, v, 0.0, v)
    v_prev = torch.zeros_like(v)
    alphas, betas = [], []
    beta = 0.0
    last = None
    for k in range(max_steps):
        w = apply_M(v)
        a = ws.dot(w, v)
        axpby(w, 1.0, w, -a, v)
        if k > 0:
            axpby(w, 1.0, w, -beta, v_prev)
        alphas.append(a)
        T = np.diag(alphas)
        if betas:
            T += np.diag(betas, 1) + np.diag(betas, -1)
        top = float(np.linalg.eigvalsh(T)[-1])
        if last is not None and abs(top - last) <= tol * abs(top):
            return top
        last = top
        beta = ws.norm(w)
        if beta <= 1e-300:
            return top
        betas.append(beta)
        v_prev, v = v, axpby(w, 1.0 / beta, w, 0.0, w)
    return last


def largest_singular_value(A, device=None):
    """sigma_max(A) (what ``svds(A, 1)`` returns in mirror_descent.py:19-24)."""
    problem = A if isinstance(A, LsqProblem) else LsqProblem(A, np.zeros(A.shape[0]), device=device)
    tmp = torch.empty(problem.m, dtype=torch.float64, device=problem.device)

    def apply_M(v):
        problem.matvec(v, tmp)
        return problem.rmatvec(tmp)
    return float(np.sqrt(_lanczos_largest(apply_M, problem.n, problem.device)))


def lsv_operator(A, N):
    """Largest singular value of A N without forming it (bsls_utils.py:334-369)."""
    problem = A if isinstance(A, LsqProblem) else LsqProblem(A, np.zeros(A.shape[0]))
    x = torch.empty(problem.n, dtype=torch.float64, device=problem.device)
    r = torch.empty(problem.m, dtype=torch.float64, device=problem.device)
    gx = torch.empty(problem.n, dtype=torch.float64, device=problem.device)
    NT = N.T

    def apply_M(v):
        N.dot(v, x)
        problem.matvec(x, r)
        problem.rmatvec(r, gx)
        return NT.dot(gx)
    return float(np.sqrt(_lanczos_largest(apply_M, N.shape[1], problem.device)))


# ---------------------------------------------------------------------------------------------
# the reference's small synthetic problems (host-side data generation, as there)
# ---------------------------------------------------------------------------------------------
def generate_small_qp():
    """bsls_utils.py:510-517"""
    Q = 2 * np.array([[2, .5], [.5, 1]])
    c = np.array([1.0, 1.0])
    x_true = np.array([.25, .75])
    w, v = np.linalg.eig(Q)
    f_min = 1.875
    min_eig = w[-1]
    return Q, c, x_true, f_min, min_eig


def random_least_squares(m, n, block_starts, sparsity=0.0, in_z=False, lasso=False, truncated=False,
                         distribution='normal'):
    """Dense random least squares with x_true on the block simplices (bsls_utils.py:520-569;
    the 'normal', 'truncated' and 'exponential' designs).  Host arrays, as in the reference."""
    assert sparsity < 1.0
    block_starts = _host(block_starts)
    A = np.random.randn(m, n)
    if distribution == 'truncated':
        A = abs(A)
    if distribution == 'exponential':
        A = np.random.exponential(size=(m, n))
    x_true = abs(np.random.randn(n, 1))
    if int(sparsity * n) > 0:
        zeros = np.random.choice(n, int(sparsity * n), replace=False)
        for i in zeros:
            x_true[i] = 0.0
    block_ends = np.append(block_starts[1:], [n])
    for s, e in zip(block_starts, block_ends):
        x_true[s:e] = x_true[s:e] / np.sum(x_true[s:e])
    if lasso:
        for start, end in zip(block_starts, block_ends):
            if np.random.uniform() > 0.7:
                alpha = np.random.uniform(0.5, 1)
                x_true[start:end] = x_true[start:end] * alpha
    b = A.dot(x_true)
    x_true = x_true.flatten()
    Q = A.T.dot(A)
    c = -A.T.dot(b).flatten()
    w, v = np.linalg.eig(Q)
    g = Q.dot(x_true) + c
    f_min = .5 * x_true.T.dot(g + c)
    min_eig = w[-1]
    return {'Q': Q, 'c': c, 'x_true': x_true, 'f_min': f_min, 'min_eig': min_eig, 'A': A, 'b': b}


def generate_data(fname=None, n=100, m1=5, m2=10, A_sparse=0.5, alpha=1.0, tolerance=1e-10, permute=False, scale=True,
                  in_z=False, distribution='uniform'):
    """The reference's synthetic traffic data (bsls_utils.py:590-655), host arrays as there: A (m1 x n, 0/1),
    U (m2 x n block indicator), x_true ~ Dirichlet(alpha) per block (scaled by f when ``scale``), b = A x."""
    import scipy.linalg as ssla
    if distribution == 'uniform':
        A = (np.random.random((m1, n)) > A_sparse).astype(float)
    elif distribution == 'affine':
        tmp = 2 * (1 - A_sparse)
        line = (1 - tmp) + tmp * np.arange(n) / (n - 1)
        lines = []
        for i in range(m1):
            j = np.random.randint(n)
            lines.append(np.append(line[j:], line[:j]))
        A = (np.random.random((m1, n)) > np.array(lines)).astype(float)
    elif distribution == 'aggregated':
        num_zeros = int(n * A_sparse)
        line = np.array([0.1] * num_zeros + [.9] * (n - num_zeros))
        lines = []
        for i in range(m1):
            j = np.random.randint(n)
            lines.append(np.append(line[j:], line[:j]))
        A = (np.random.random((m1, n)) > np.array(lines)).astype(float)
    else:
        raise ValueError(distribution)
    block_sizes = (np.random.multinomial(n - m2, np.ones(m2) / m2) + np.ones(m2)).astype(int)
    assert sum(block_sizes) == n, 'all-zero row present!'
    block_starts = np.append([0], np.cumsum(block_sizes[:-1])).astype(int)
    x = np.concatenate([np.random.dirichlet(alpha * np.ones(bs)) for bs in block_sizes])
    U = ssla.block_diag(*[np.ones(bs) for bs in block_sizes])
    if scale:
        f = np.floor(np.random.random(m2) * 1000)
        x = U.T.dot(f) * x
    else:
        f = np.ones(len(U))
    b = A.dot(x)
    assert np.linalg.norm(U.dot(x) - f) < tolerance, "Ux!=f"
    assert np.linalg.norm(A.dot(x) - b) < tolerance, "Ax!=b"
    if permute:
        reorder = np.random.permutation(n)
        A = A[:, reorder]
        U = U[:, reorder]
        x = x[reorder]
    data = {'A': A, 'b': b, 'x_true': x, 'U': U, 'f': f, 'block_starts': block_starts, 'block_sizes': block_sizes}
    if fname:
        import scipy.io
        scipy.io.savemat(fname, data, oned_as='column')
    return data
