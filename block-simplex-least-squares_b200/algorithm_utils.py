"""Solver parts on the GPU -- drop-in for the reference's ``python/algorithm_utils.py``.

``get_solver_parts`` hands back the same four closures (``step_size, proj, line_search, obj``)
with the same calling conventions (in-place on their vector arguments); vectors are float64
CUDA tensors and every operation is a kernel of libbsls_b200:

    obj          bsls_lsq_obj_f64            r = A x - b, g = A^T r, f = 0.5 <r, r>
    proj         bsls_dev_proj_multi_*_f64 / bsls_dev_isotonic_regression_multi_f64 (+ clip)
    line_search  axpby + objective + dot kernels

The closures carry their handles (``obj.problem``, ``proj.plan``, ``proj.mode`` ...) so that
``BATCH.solve*`` can hand the whole loop to the library (bsls_batch_solve_f64).
"""
import numpy as np
import torch

from . import c_extensions as cx
from .plan import BlockPlan, plan_for
from .sparse import LsqProblem, axpby, copy_, default_workspace

__all__ = ["get_solver_parts", "sparse_least_squares_obj", "quad_obj_np", "decreasing_step_size", "line_search_np",
           "line_search_exact_quad_obj",
           "stopping", "normalization", "proj_simplex", "proj_multi_simplex"]


# ---------------------------------------------------------------------------------------------
# projections (reference: python/algorithm_utils.py:63-76 are NumPy re-statements of the C
# routines; here both names run the CUDA kernels)
# ---------------------------------------------------------------------------------------------
def proj_simplex(y, start, end):
    """projects subvector of y in range(start, end) (algorithm_utils.py:63-69)"""
    return cx.proj_simplex_c(y, start, end)


def proj_multi_simplex(y, blocks):
    """algorithm_utils.py:72-76"""
    return cx.proj_multi_simplex_c(y, blocks)


# ---------------------------------------------------------------------------------------------
# objectives
# ---------------------------------------------------------------------------------------------
_PROBLEMS = {}


def _problem_for(A_sparse, b):
    if isinstance(A_sparse, LsqProblem):
        return A_sparse
    key = (id(A_sparse), id(b))
    hit = _PROBLEMS.get(key)
    if hit is not None and hit[0] is A_sparse and hit[1] is b:
        return hit[2]
    prob = LsqProblem(A_sparse, b)
    _PROBLEMS[key] = (A_sparse, b, prob)
    while len(_PROBLEMS) > 8:
        _PROBLEMS.pop(next(iter(_PROBLEMS)))
    return prob


def sparse_least_squares_obj(x, A_sparse_T, A_sparse, b, g):
    """Sparse least-squares objective and gradient (algorithm_utils.py:88-94): ``g`` is
    overwritten with A^T (A x - b) and f = 0.5 |A x - b|^2 is returned.  ``A_sparse`` is a scipy
    matrix (uploaded once and cached) or an :class:`LsqProblem`; ``A_sparse_T`` is accepted for
    signature compatibility (the transpose is kept inside the problem)."""
    return _problem_for(A_sparse, b).obj(x, g)


class _DenseQuadratic:
    """0.5 x'Qx + c'x for the reference's small dense QPs (algorithm_utils.py:79-85), evaluated
    with the sparse kernels on a CSR copy of Q:  g = Q x + c  is the residual of (A, b) = (Q, -c)."""

    def __init__(self, Q, c, device=None):
        import scipy.sparse as sps
        Q = np.asarray(Q, dtype=np.float64)
        self.c_host = np.asarray(c, dtype=np.float64).reshape(-1)
        self.problem = LsqProblem(sps.csr_matrix(Q), -self.c_host, device=device)
        self.c = torch.as_tensor(self.c_host).to(self.problem.device)
        self.n = Q.shape[0]

    def obj(self, x, g):
        self.problem.value(x)               # residual = Q x + c
        copy_(g, self.problem.residual())
        d = self.problem.ws.dots([(x, g), (x, self.c)])
        return .5 * (d[0] + d[1])           # .5 * x.T.dot(g + c)


def quad_obj_np(x, Q, c, g=None):
    """algorithm_utils.py:79-85 on device vectors (``Q`` may be a prepared _DenseQuadratic)."""
    quad = Q if isinstance(Q, _DenseQuadratic) else _DenseQuadratic(Q, c, x.device)
    if g is None:
        g = torch.zeros_like(x)
    return quad.obj(x, g)


def decreasing_step_size(i, t0, alpha):
    """step size of the form t = t0 / (1 + t0*alpha*t) (algorithm_utils.py:97-101)"""
    return t0 / (alpha * i + t0)


def line_search_np(x, f, g, x_new, f_new, g_new, obj):
    """Backtracking line search (algorithm_utils.py:113-137); updates x_new and g_new in place."""
    ws = default_workspace(x.device)
    t = 1.0
    suffDec = 1e-4
    progTol = 1e-12
    tmp = torch.empty_like(x)

    def g_dot_step():
        axpby(tmp, 1.0, x_new, -1.0, x)
        return ws.dot(g, tmp)

    upper_line = f + suffDec * g_dot_step()
    while f_new > upper_line:
        t *= .8
        step = ws.max_abs_diff(x_new, x)
        if step < progTol:
            t = 0.0
            f_new = f
            copy_(g_new, g)
            copy_(x_new, x)
            break
        axpby(x_new, 1.0 - t, x, t, x_new)
        f_new = obj(x_new, g_new)
        upper_line = f + suffDec * g_dot_step()
    return f_new


def line_search_exact_quad_obj(x, f, g, x_new, f_new, g_new, Q, c):
    """Exact line search for a quadratic objective 0.5 x'Qx + c'x (algorithm_utils.py:140-155): minimises along
    d = x_new - x, overwrites x_new with x + t d and g_new with the gradient there, returns the new objective.
    ``Q`` may be a prepared :class:`_DenseQuadratic` (then ``c`` is ignored)."""
    quad = Q if isinstance(Q, _DenseQuadratic) else _DenseQuadratic(Q, c, x.device)
    ws = quad.problem.ws
    progTol = 1e-8
    d = axpby(torch.empty_like(x), 1.0, x_new, -1.0, x)
    # Check whether step has become too small
    if ws.max_abs_diff(x_new, x) < progTol:
        copy_(g_new, g)
        copy_(x_new, x)
        return f
    tmp = quad.problem.matvec(d)                                   # Q.dot(d)
    xt, dc, dt = ws.dots([(x, tmp), (d, quad.c), (d, tmp)])
    t = -(xt + dc) / dt
    axpby(x_new, 1.0, x, t, d)                                     # x + t*d
    return quad_obj_np(x_new, quad, None, g_new)


def stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min=None):
    """Simple stopping (algorithm_utils.py:158-172) -- host scalars only."""
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


def normalization(x, block_starts, block_ends=None):
    """Divide every block of x by its sum, in place (algorithm_utils.py:175-179).  ``block_ends``
    is implied by the starts (the reference passes np.append(block_starts[1:], [n]))."""
    plan = plan_for(block_starts, x.shape[0], x.device)
    default_workspace(x.device).md_update(plan, x, x, x, 0.0)  # x * exp(-0 * x) = x, then normalise


def _z_starts(block_starts, lasso):
    tmp = np.array(block_starts.cpu() if torch.is_tensor(block_starts) else block_starts, dtype=np.int64).copy()
    if not lasso:
        tmp -= np.arange(len(tmp))
    return tmp


def get_solver_parts(data, block_starts, min_eig, in_z=False, is_sparse=False, lasso=False, f=None, device=None,
                     implicit_ones=False):
    """Returns the step_size, proj, line_search, and obj functions for the least squares problem
    (algorithm_utils.py:182-271).

    data: (Q, c) if not sparse, (A, b) if sparse -- host arrays / scipy matrices, uploaded once --
          or an :class:`LsqProblem` already on the device
    block_starts: first indices of each block (in x)
    min_eig: minimum eigenvalue of Q = A.T.dot(A)
    in_z: the variable is z (projection = isotonic regression + clip to [0,1])
    lasso: feasible set is the l1-ball instead of the simplex
    f: per-block totals; x is divided by f_k before and multiplied after the projection
    """
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if isinstance(data, LsqProblem):
        problem = data
        n = problem.n

        def obj(x, g=None):
            return problem.obj(x, g)
        obj.problem = problem
    elif is_sparse:
        A, b = data
        problem = LsqProblem(A, b, device=device, implicit_ones=implicit_ones)
        n = problem.n

        def obj(x, g=None):
            return problem.obj(x, g)
        obj.problem = problem
    else:
        Q, c = data
        quad = _DenseQuadratic(Q, c, device)
        n = quad.n

        def obj(x, g=None):
            return quad_obj_np(x, quad, None, g)
        obj.problem = None

    def step_size(i):
        return decreasing_step_size(i, 1.0, min_eig)
    step_size.min_eig = min_eig

    starts_host = np.asarray(block_starts.cpu() if torch.is_tensor(block_starts) else block_starts, dtype=np.int64)
    if in_z:
        zstarts = _z_starts(starts_host, lasso)
        plan = BlockPlan(zstarts, n, device)
        mode = 2
    else:
        plan = BlockPlan(starts_host, n, device)
        mode = 1 if lasso else 0
    scale = None
    if f is not None:
        scale = torch.as_tensor(np.asarray(f, dtype=np.float64)).to(device)
        assert scale.shape[0] == plan.numblocks

    def proj(x):
        if scale is not None:
            cx.block_scale(x, plan, scale, divide=True)
        if mode == 2:
            # block_isotonic_regression_2 + np.maximum(0.,x,x) + np.minimum(1.,x,x)  (:219-224)
            cx.isotonic_regression_multi_c(x, plan, None, 1, clip01=True)
        elif mode == 1:
            cx.proj_multi_ball_c(x, plan)
        else:
            cx.proj_multi_simplex_c(x, plan)
        if scale is not None:
            cx.block_scale(x, plan, scale, divide=False)
    proj.plan = plan
    proj.mode = mode
    proj.scaled = scale is not None

    def line_search(x, f, g, x_new, f_new, g_new, i):
        return line_search_np(x, f, g, x_new, f_new, g_new, obj)
    line_search.obj = obj

    return step_size, proj, line_search, obj
