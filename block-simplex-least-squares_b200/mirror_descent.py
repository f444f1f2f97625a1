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
    # the whole loop runs inside the library (bsls_md_least_squares_f64): SpMV pair + ONE fused exponentiate / normalise /
    # max-change kernel per iteration, step t_k = sqrt(2 ln K_block) / (sqrt(k) Lf) and the stop test on the device
    from . import _lib
    import ctypes
    res = _lib.BatchResult()
    with torch.cuda.device(problem.device):
        st = torch.cuda.current_stream(problem.device).cuda_stream
        _lib.check(_lib.lib().bsls_md_least_squares_f64(problem.handle, plan.handle, x.data_ptr(), int(iters), float(tolerance), float(Lf),
                                                        ctypes.byref(res), st), "md_least_squares")
    least_squares.last = {"iterations": res.iterations, "device_ms": res.device_ms, "change": res.stop_value,
                          "kernel_launches": res.kernel_launches}
    return x
