"""Exponentiated-gradient (mirror descent) least squares on block simplices -- drop-in for
``python/mirror_descent.py``.  One iteration = the SpMV pair plus ONE fused kernel that
exponentiates, normalises every block and measures max |x - x_prev|."""
import numpy as np
import torch

from .plan import BlockPlan
from .sparse import LsqProblem

__all__ = ["least_squares"]


def least_squares(A, b, blocks, iters=1000, tolerance=1e-9, Lf=None, device=None):
    """mirror_descent.py:7-53.  ``blocks`` holds the block SIZES.  ``A`` is a scipy matrix or an
    :class:`LsqProblem`; ``Lf`` (largest singular value of A) is computed on the GPU with a
    Lanczos iteration when not given (the reference calls ARPACK ``svds``, whose random start
    vector makes it reproducible only to solver tolerance -- pass the same ``Lf`` for parity)."""
    problem = A if isinstance(A, LsqProblem) else LsqProblem(A, b, device=device)
    sizes = np.asarray(blocks, dtype=np.int64)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    n = int(sizes.sum())
    assert n == problem.n
    plan = BlockPlan(starts, n, problem.device)
    # x = 1 / block size
    x = torch.as_tensor(np.repeat(1.0 / sizes.astype(np.float64), sizes)).to(problem.device)
    if Lf is None:
        from .bsls_utils import largest_singular_value
        Lf = largest_singular_value(problem)
    g = torch.empty_like(x)
    x_new = torch.empty_like(x)
    from . import _lib
    L = _lib.lib()
    for _iter in range(1, iters + 1):
        with torch.cuda.device(problem.device):
            st = torch.cuda.current_stream(problem.device).cuda_stream
            _lib.check(L.bsls_dev_lsq_residual_f64(problem.handle, x.data_ptr(), st))
            _lib.check(L.bsls_dev_lsq_gradient_f64(problem.handle, g.data_ptr(), st))
        # t_k = sqrt(2 ln K_block) / (sqrt(k) Lf);  x <- normalise(x exp(-t_k g))
        problem.ws.md_update(plan, x_new, x, g, np.sqrt(_iter) * Lf, per_block_log=True)
        change = problem.ws.scalars()[10]
        x, x_new = x_new, x
        if change < tolerance:
            break
    return x
