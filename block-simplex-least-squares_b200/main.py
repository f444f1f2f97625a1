"""``solve_in_z`` of the reference's ``python/main.py`` (:41-79) -- the z-space hot loop that the
reference's CLI drives -- on the GPU.  The CLI itself (argparse, .mat loading, plots) is out of
scope; see SURVEY.md section 2, row 18."""
import numpy as np
import torch

from . import c_extensions as cx
from .bsls_utils import x2z, particular_x0, NOperator
from .gradient_descent import GradientDescent
from .plan import BlockPlan
from .sparse import LsqProblem, axpby

__all__ = ["solve_in_z", "z_space_parts"]


def z_space_parts(A, b, x0, N, block_sizes, device=None):
    """The closures of main.py:47-65: z0, target, f, nabla_f, proj and the problem handle.

    f(z) = 0.5 |A N z + target|^2,  nabla_f(z) = N^T A^T (A N z + target),  target = A x0 - b,
    proj = isotonic regression of every z-block, clipped to [0, 1]."""
    sizes = np.asarray(block_sizes, dtype=np.int64)
    if isinstance(A, LsqProblem):
        problem = A.with_b(torch.zeros(A.m, dtype=torch.float64, device=A.device))   # private handle: the caller's b stays
    else:
        problem = LsqProblem(A, np.zeros(A.shape[0]), device=device)
    dev = problem.device
    if not torch.is_tensor(x0):
        x0 = torch.as_tensor(np.asarray(x0, dtype=np.float64).reshape(-1)).to(dev)
    if N is None or not isinstance(N, NOperator):
        from .bsls_utils import block_sizes_to_N
        N = block_sizes_to_N(sizes, dev)
    NT = N.T
    z0 = x2z(x0, block_sizes=sizes)
    bdev = torch.as_tensor(np.asarray(b, dtype=np.float64).reshape(-1)).to(dev) if not torch.is_tensor(b) else b
    # target = A x0 - b ; the library subtracts its `b`, so it is handed -target
    problem.set_b(bdev)
    problem.value(x0)
    target = problem.residual().clone()
    problem.set_b(axpby(torch.empty_like(target), 0.0, target, -1.0, target))
    xbuf = torch.empty(problem.n, dtype=torch.float64, device=dev)
    gx = torch.empty(problem.n, dtype=torch.float64, device=dev)

    def f(z):
        N.dot(z, xbuf)
        return problem.value(xbuf)

    def nabla_f(z):
        N.dot(z, xbuf)
        problem.obj(xbuf, gx)
        return NT.dot(gx)

    cum_blocks = np.concatenate(([0], np.cumsum(sizes - 1)))
    zplan = BlockPlan(cum_blocks[:-1], int(cum_blocks[-1]), dev)

    def proj(z):
        cx.isotonic_regression_multi_c(z, zplan, None, 1, clip01=True)
        return z

    return z0, target, f, nabla_f, proj, problem, N


def solve_in_z(A, b, x0, N, block_sizes, method, options=None):
    """main.py:41-79: returns (iters, times, states); states[-1] is the final z."""
    sizes = np.asarray(block_sizes, dtype=np.int64)
    n = A.n if isinstance(A, LsqProblem) else A.shape[1]
    if block_sizes is not None and len(sizes) == n:
        raise SystemExit('Trivial example: nblocks == nroutes, exiting solver')
    z0, target, f, nabla_f, proj, problem, N = z_space_parts(A, b, x0, N, sizes)
    if method == 'DORE':
        gd = GradientDescent(z0=z0, f=f, nabla_f=nabla_f, proj=proj, method=method, options=options, A=problem, N=N,
                             target=target)
    else:
        gd = GradientDescent(z0=z0, f=f, nabla_f=nabla_f, proj=proj, method=method, options=options)
    iters, times, states = gd.run()
    return iters, times, states
