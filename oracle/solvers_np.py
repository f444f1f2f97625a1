"""CPU restatement of the reference's solver drivers -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product (block-simplex-least-squares_b200/) never does.

NumPy + scipy.sparse CSR products (the un-vendored third-party arithmetic the reference itself
uses: scipy is "also needed", README.md:13) + the C oracle for the projections.  Each function
follows the reference statement for statement:

    sparse_least_squares_obj   python/algorithm_utils.py:88-94
    line_search_np             python/algorithm_utils.py:113-137
    stopping                   python/algorithm_utils.py:158-172
    solve / solve_BB / solve_LBFGS / LBFGS_helper / solve_MD    python/BATCH.py:7-250
    bb_solve                   python/BB.py:7-45
    solvers_stopping           python/solvers.py:40-63
    weak_wolfe_ls, lbfgs_solve python/LBFGS.py:9-123
    dore_solve                 python/DORE.py:6-90
    md_least_squares           python/mirror_descent.py:7-53
    z_space_closures           python/main.py:47-65

Pinned by tests/test_oracle_solvers.py against tests/golden/solvers.npz, which was produced by
the reference's own modules (tests/golden/make_golden.py).
"""
from collections import deque

import numpy as np
import numpy.linalg as la
import scipy.sparse as sps

from . import cpu


# ---------------------------------------------------------------------------------------------
# parts (get_solver_parts, python/algorithm_utils.py:182-271, sparse x-space branch)
# ---------------------------------------------------------------------------------------------
def sparse_least_squares_obj(x, A_sparse_T, A_sparse, b, g):
    tmp = A_sparse.dot(x) - b
    np.copyto(g, A_sparse_T.dot(tmp))
    return .5 * tmp.T.dot(tmp)


def line_search_np(x, f, g, x_new, f_new, g_new, obj):
    t = 1.0
    suffDec = 1e-4
    progTol = 1e-12
    upper_line = f + suffDec * g.dot(x_new - x)
    while f_new > upper_line:
        t *= .8
        step = np.linalg.norm(x_new - x, np.inf)
        if step < progTol:
            t = 0.0
            f_new = f
            np.copyto(g_new, g)
            np.copyto(x_new, x)
            break
        np.copyto(x_new, (1.0 - t) * x + t * x_new)
        f_new = obj(x_new, g_new)
        upper_line = f + suffDec * g.dot(x_new - x)
    return f_new


def stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min=None):
    flag = False
    stop = 'continue'
    if i == max_iter:
        stop = 'max_iter'
        flag = True
    if f_min is not None and f - f_min < opt_tol:
        stop = 'f-f_min = {} < opt_tol'.format(f - f_min)
        flag = True
    if abs(f_old - f) < prog_tol:
        stop = '|f_old-f| = {} < prog_tol'.format(abs(f_old - f))
        flag = True
    return flag, stop


def normalization(x, block_starts, block_ends):
    for start, end in zip(block_starts, block_ends):
        np.copyto(x[start:end], x[start:end] / np.sum(x[start:end]))


def get_solver_parts(A, b, block_starts, min_eig, lasso=False, in_z=False, threads=0, project=None):
    """``threads`` > 1 (bench.py's reference arm only): the two CSR products run on that many host threads over row
    ranges (same arithmetic per row as scipy's csr_matvec) and ``project`` replaces the projection (the compiled
    reference over block ranges).  The reference itself has no threading; see bench.py."""
    A_sparse = sps.csr_matrix(A)
    A_sparse_T = sps.csr_matrix(A.T)
    block_starts = np.asarray(block_starts)

    if threads and threads > 1:
        port = cpu.port()
        Ap, Ai, Av = A_sparse.indptr.astype(np.int64), A_sparse.indices.astype(np.int32), A_sparse.data
        Tp, Ti, Tv = A_sparse_T.indptr.astype(np.int64), A_sparse_T.indices.astype(np.int32), A_sparse_T.data
        tmp_m = np.empty(A_sparse.shape[0])

        def obj(x, g=None):
            port.csr_matvec_mt(Ap, Ai, Av, x, tmp_m, threads)
            np.subtract(tmp_m, b, tmp_m)
            port.csr_matvec_mt(Tp, Ti, Tv, tmp_m, g, threads)
            return .5 * tmp_m.dot(tmp_m)
    else:
        def obj(x, g=None):
            return sparse_least_squares_obj(x, A_sparse_T, A_sparse, b, g)

    def step_size(i):
        return 1.0 / (min_eig * i + 1.0)

    chk = cpu.port()
    if in_z:
        tmp = block_starts - np.arange(len(block_starts))

        def proj(x):
            chk.pava_multi(x, tmp)
            np.maximum(0., x, x)
            np.minimum(1., x, x)
    elif lasso:
        def proj(x):
            chk.proj_multi_ball(x, block_starts)
    else:
        def proj(x):
            chk.proj_multi_simplex(x, block_starts)
    if project is not None:
        proj = project

    def line_search(x, f, g, x_new, f_new, g_new, i):
        return line_search_np(x, f, g, x_new, f_new, g_new, obj)

    return step_size, proj, line_search, obj


# ---------------------------------------------------------------------------------------------
# BATCH.py
# ---------------------------------------------------------------------------------------------
def solve(obj, proj, step_size, x_init, line_search=None, f_min=None, opt_tol=1e-6, max_iter=2000, prog_tol=1e-12):
    n = x_init.shape[0]
    x = np.copy(x_init)
    g = np.zeros(n)
    g_new = np.zeros(n)
    x_new = np.zeros(n)
    f_old = np.inf
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        t = step_size(i)
        np.add(x, -t * g, x_new)
        proj(x_new)
        f_new = obj(x_new, g_new)
        if line_search is not None:
            f_new = line_search(x, f, g, x_new, f_new, g_new, i)
        f_old = f
        f = f_new
        np.copyto(x, x_new)
        np.copyto(g, g_new)
        i += 1
        progress.append([0.0, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}


def solve_BB(obj, proj, line_search, x_init, f_min=None, opt_tol=1e-6, max_iter=2000, prog_tol=1e-12):
    n = x_init.shape[0]
    x = np.copy(x_init)
    g = np.zeros(n)
    delta_x = np.zeros(n)
    delta_g = np.zeros(n)
    g_new = np.zeros(n)
    x_new = np.zeros(n)
    f_old = np.inf
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        if i == 1:
            np.add(x, -g, x_new)
        else:
            t = delta_x.T.dot(delta_g) / delta_g.T.dot(delta_g)
            np.add(x, -t * g, x_new)
        proj(x_new)
        f_new = obj(x_new, g_new)
        f_new = line_search(x, f, g, x_new, f_new, g_new, i)
        f_old = f
        f = f_new
        np.add(x_new, -x, delta_x)
        np.add(g_new, -g, delta_g)
        np.copyto(x, x_new)
        np.copyto(g, g_new)
        i += 1
        progress.append([0.0, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}


def LBFGS_helper(q_delta_g, q_delta_x, q_rho, g, d, alpha):
    m = len(q_delta_g)
    np.copyto(d, g)
    for j in range(1, m + 1):
        alpha[-j] = q_rho[-j] * q_delta_x[-j].T.dot(d)
        d -= alpha[-j] * q_delta_g[-j]
    t = q_delta_x[-1].T.dot(q_delta_g[-1]) / q_delta_g[-1].T.dot(q_delta_g[-1])
    d *= t
    for j in range(m):
        beta = q_rho[j] * q_delta_g[j].T.dot(d)
        d += q_delta_x[j] * (alpha[-m + j] - beta)
    d *= -1.0


def solve_LBFGS(obj, proj, line_search, x_init, f_min=None, opt_tol=1e-6, max_iter=1000, prog_tol=1e-12, corrections=50):
    # the deques receive the SAME two arrays every iteration (BATCH.py:154-156), as in the reference
    q_delta_g = deque()
    q_delta_x = deque()
    q_rho = deque()
    n = x_init.shape[0]
    x = np.copy(x_init)
    g = np.zeros(n)
    d = np.zeros(n)
    alpha = np.zeros(corrections)
    delta_x = np.zeros(n)
    delta_g = np.zeros(n)
    g_new = np.zeros(n)
    x_new = np.zeros(n)
    f_old = np.inf
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        if i == 1:
            np.add(x, -g, x_new)
        else:
            q_delta_g.append(delta_g)
            q_delta_x.append(delta_x)
            q_rho.append(1 / delta_g.T.dot(delta_x))
            if i > corrections + 1:
                q_delta_g.popleft()
                q_delta_x.popleft()
                q_rho.popleft()
            if i <= 5:
                d = -(delta_x.T.dot(delta_g) / delta_g.T.dot(delta_g)) * g
            else:
                LBFGS_helper(q_delta_g, q_delta_x, q_rho, g, d, alpha)
            np.add(x, d, x_new)
        proj(x_new)
        f_new = obj(x_new, g_new)
        f_new = line_search(x, f, g, x_new, f_new, g_new, i)
        f_old = f
        f = f_new
        np.add(x_new, -x, delta_x)
        np.add(g_new, -g, delta_g)
        np.copyto(x, x_new)
        np.copyto(g, g_new)
        i += 1
        progress.append([0.0, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}


def solve_MD(obj, block_starts, step_size, x_init, line_search=None, f_min=None, opt_tol=1e-6, max_iter=1000, prog_tol=0.0):
    n = x_init.shape[0]
    block_ends = np.append(block_starts[1:], [n])
    x = np.copy(x_init)
    g = np.zeros(n)
    g_new = np.zeros(n)
    x_new = np.zeros(n)
    f_old = np.inf
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        t = step_size(i)
        np.copyto(x_new, x * np.exp(-t * g))
        normalization(x_new, block_starts, block_ends)
        f_new = obj(x_new, g_new)
        f_old = f
        f = f_new
        np.copyto(x, x_new)
        np.copyto(g, g_new)
        i += 1
        progress.append([0.0, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}


# ---------------------------------------------------------------------------------------------
# functional drivers (BB.py, LBFGS.py, DORE.py, solvers.py, mirror_descent.py)
# ---------------------------------------------------------------------------------------------
def solvers_stopping(g, fx, i, t, d=None, delta_g=None, options=None, TOLER=1e-6):
    if options and 'max_iter' in options:
        if i >= options['max_iter']:
            return True
    if options and 'opt_tol' in options:
        TOLER = options['opt_tol']
    norm2_nabla_f = np.square(la.norm(g))
    thresh = TOLER * (1 + abs(fx))
    if norm2_nabla_f <= thresh:
        return True
    if d is not None and la.norm(t * d) <= 1e-12:
        return True
    if delta_g is not None and la.norm(delta_g) == 0:
        return True
    return False


def bb_solve(x0, f, nabla_f, stopping_fn=solvers_stopping, proj=None, options=None):
    i, stop = 0, False
    x = x0
    x_prev = x + 1
    g_prev = nabla_f(x_prev)
    while not stop:
        i += 1
        g = nabla_f(x)
        delta_g = g - g_prev
        if sum(delta_g) == 0:
            break
        delta_x = x - x_prev
        t = delta_x.dot(delta_g) / delta_g.dot(delta_g)
        x_next = x - t * g
        x_prev, x = x, x_next
        g_prev = g
        if proj:
            x = proj(x)
        fx = f(x)
        stop = stopping_fn(g, fx, i, t, delta_g=delta_g, options=options)
    return x


def weak_wolfe_ls(x, d, f, nabla_f, proj=lambda x: x, c1=1e-3, c2=0.9):
    alpha, beta = 0, float('inf')
    t = 1
    stop = False
    proj_x = proj(x)
    nabla_fx = nabla_f(proj_x)
    while not stop:
        proj_xtd = proj(x + t * d)
        if f(proj_xtd) >= f(proj_x) + c1 * t * d.dot(nabla_fx):
            beta = t
            t = 0.5 * (alpha + beta)
        elif d.dot(nabla_f(proj_xtd)) < c2 * d.dot(nabla_fx):
            alpha = t
            t = 2 * alpha if beta == float('inf') else 0.5 * (alpha + beta)
        else:
            stop = True
        if np.abs(alpha - beta) <= 1e-14:
            stop = True
        if la.norm(t * d) <= 1e-8:
            stop = True
    return t


def lbfgs_solve(x0, f, nabla_f, stopping_fn=solvers_stopping, m=50, proj=None, options=None):
    def search_dir(g_new, y_new, s_new, rho, y, s, m=10):
        q = g_new
        alpha = [0] * m
        for i in range(m - 1, -1, -1):
            alpha[i] = rho[i] * (s[i].dot(q))
            q = q - alpha[i] * y[i]
        H = y_new.dot(s_new) / (y_new.dot(y_new))
        r = H * q
        for i in range(0, m):
            beta = rho[i] * y[i].dot(r)
            r = r + s[i] * (alpha[i] - beta)
        return -r

    i, stop = 0, False
    x = x0
    n = x.shape[0]
    y, s = [np.zeros((n))] * m, [np.zeros((n))] * m
    g_new = nabla_f(x)
    y_new, s_new = g_new, np.ones((n))
    rho, rho_new = [0] * m, 1 / (y_new.dot(s_new))
    while not stop:
        i += 1
        d = search_dir(g_new, y_new, s_new, rho, y, s, m=m)
        y.pop(0)
        y.append(y_new)
        s.pop(0)
        s.append(s_new)
        rho.pop(0)
        rho.append(rho_new)
        t = weak_wolfe_ls(x, d, f, nabla_f, proj=proj)
        s_new = t * d
        x_next = x + s_new
        if proj:
            x_next = proj(x_next)
        g = g_new
        g_new = nabla_f(x_next)
        y_new = g_new - g
        if y_new.dot(s_new) == 0:
            break
        rho_new = 1 / (y_new.dot(s_new))
        x = x_next
        fx = f(x)
        stop = stopping_fn(g_new, fx, i, t, d=d, options=options)
    return x


def dore_solve(x0, linop, linop_T, target, proj=None, options=None, i=10000, eps=10 ** -16):
    if options and 'max_iter' in options:
        i = options['max_iter']
    if options and 'opt_tol' in options:
        eps = options['opt_tol']
    b = -np.array(target)
    x = np.array(x0)
    x_prev = x
    Ax = 0
    Ax_prev = 0
    Ax_prev_prev = 0
    for iter_ in range(i):
        Ax_prev_prev = Ax_prev
        Ax_prev = Ax
        Ax = linop(x)
        err = b - Ax
        norm_change = ((la.norm(x - x_prev) ** 2))
        if iter_ > 0 and (norm_change <= eps):
            break
        x_new = x + linop_T(err)
        x_new = proj(x_new)
        Ax = linop(x_new)
        err = b - Ax
        if iter_ > 2:
            delta_Ax = Ax - Ax_prev
            dp = delta_Ax.dot(delta_Ax)
            if dp > 0:
                a1 = delta_Ax.dot(err) / dp
                Ax_1 = (1 + a1) * Ax - a1 * Ax_prev
                x_1 = x_new + a1 * (x_new - x)
                err_1 = b - Ax_1
                delta_Ax = Ax_1 - Ax_prev_prev
                dp = delta_Ax.dot(delta_Ax)
                if dp > 0:
                    a2 = delta_Ax.dot(err_1) / dp
                    x_2 = x_1 + a2 * (x_1 - x_prev)
                    x_2 = proj(x_2)
                    Ax_2 = linop(x_2)
                    err_2 = b - Ax_2
                    if err_2.dot(err_2) / err.dot(err) < 1:
                        x_select = x_2
                        Ax = Ax_2
                    else:
                        x_select = x_new
                else:
                    x_select = x_new
            else:
                x_select = x_new
        else:
            x_select = x_new
        x_prev = x
        x = x_select
    return x


def md_least_squares(A, b, blocks, iters=1000, tolerance=1e-9, Lf=None):
    """mirror_descent.py:7-53 with the Lipschitz constant passed in (the reference draws it from
    ARPACK with a random start vector)."""
    n_vector = np.concatenate([[bs] * bs for bs in blocks]).astype(float)
    x = np.divide(1.0, n_vector)
    if Lf is None:
        Lf = sps.linalg.svds(A, 1, return_singular_vectors=False)[0]

    def t_(k):
        return np.sqrt(2 * np.log(n_vector)) / (np.sqrt(k) * Lf)

    for _iter in range(1, iters + 1):
        x_prev = x
        up = A.T.dot(A.dot(x) - b)
        up *= t_(_iter)
        x = x * np.exp(-up)
        beginning = 0
        for block in blocks:
            x_section = x[beginning:block + beginning]
            x[beginning:block + beginning] = x_section / np.sum(x_section)
            beginning += block
        if np.linalg.norm(x - x_prev, np.inf) < tolerance:
            break
    return x


def z_space_closures(A, b, block_sizes):
    """main.py:47-65: N (bsls_utils.py:139-162), x0 (:327-328), f, nabla_f, proj in z."""
    block_sizes = np.asarray(block_sizes, dtype=np.int64)
    n = int(block_sizes.sum())
    nz = n - len(block_sizes)
    rows, cols, vals = [], [], []
    sr = sc = 0
    for K in block_sizes:
        for j in range(K - 1):
            rows += [sr + j, sr + j + 1]
            cols += [sc + j, sc + j]
            vals += [1.0, -1.0]
        sr += K
        sc += K - 1
    N = sps.csr_matrix((vals, (rows, cols)), shape=(n, nz))
    x0 = np.zeros(n)
    x0[np.cumsum(block_sizes) - 1] = 1.0
    target = A.dot(x0) - b
    AT = A.T.tocsr()
    NT = N.T.tocsr()
    f = lambda z: 0.5 * la.norm(A.dot(N.dot(z)) + target) ** 2
    nabla_f = lambda z: NT.dot(AT.dot(A.dot(N.dot(z)) + target))
    zstarts = np.concatenate(([0], np.cumsum(block_sizes - 1)))[:-1]
    chk = cpu.port()

    def proj(v):
        chk.pava_multi(v, zstarts)
        return np.maximum(np.minimum(v, 1.), 0.)

    return N, x0, target, f, nabla_f, proj, zstarts
