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

__all__ = ["solve_in_z", "z_space_parts", "LS_postprocess", "main"]


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

    # the handles behind the closures: lets BB.solve run the whole loop inside the library (bsls_zbb_run_f64)
    zspace = {"problem": problem, "N": N, "zplan": zplan}
    f.zspace = nabla_f.zspace = proj.zspace = zspace
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


def LS_postprocess(states, x0, A, b, x_true, scaling=None, block_sizes=None, output=None, N=None, is_x=False):
    """Objective-error and route-flow-error metrics over the recorded iterates (main.py:81-136) -- the step immediately
    AFTER the hot loop.  ``states`` are device vectors (z, or x when ``is_x``), ``A`` an
    :class:`LsqProblem` (or anything LsqProblem accepts) with right-hand side ``b``.  Every iterate costs one
    ``x = N z + x0`` kernel, one product ``A x - b`` and one fused pass for the four flow metrics; nothing is
    materialised as a (variables x iterates) matrix as the reference does.  Returns (x_last, error, output)."""
    if x_true is None:
        return [], [], output
    if output is None:
        output = {}
    problem = A if isinstance(A, LsqProblem) else LsqProblem(A, b)
    if isinstance(A, LsqProblem) and b is not None:
        problem = A.with_b(b if torch.is_tensor(b) else torch.as_tensor(np.asarray(b, dtype=np.float64)).to(A.device))
    dev = problem.device
    as_dev = lambda v: v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(-1)).to(dev)
    x_true, x0 = as_dev(x_true), as_dev(x0)
    scaling = None if scaling is None else as_dev(scaling)
    d = len(states)
    ws = problem.ws
    # Convert back to x (from z) if necessary
    x_hat = torch.empty(problem.n, dtype=torch.float64, device=dev)
    error = np.zeros(d)
    dist, wrong, per_flow = np.zeros(d), np.zeros(d, dtype=np.int64), np.zeros(d)
    x_last = None
    for k, st in enumerate(states):
        if not is_x and N is not None and N.shape[1] > 0:
            N.dot(st, x_hat)
            axpby(x_hat, 1.0, x_hat, 1.0, x0)              # N z + x0
        else:
            x_hat.copy_(st)
        error[k] = problem.value(x_hat)                    # 0.5 ||A x - b||^2
        sabs, sref, cnt, _, smax = ws.flow_metrics(scaling, x_true, x_hat, 1e-3)
        dist[k], wrong[k], per_flow[k] = smax, int(cnt), sabs / sref
        if k == d - 1:
            x_last = x_hat.clone()
    output['AA'] = (problem.m, problem.n)
    output['x_hat'] = (problem.n, d)
    output['blocks'] = None if block_sizes is None else np.asarray(block_sizes).shape
    output['0.5norm(Ax-b)^2'] = error
    output['0.5norm(Ax_init-b)^2'] = problem.value(x0)
    output['0.5norm(Ax*-b)^2'] = problem.value(x_true)
    output['max|f * (x-x_true)|'] = dist                       # most incorrect entry (route flow)
    output['incorrect x entries'] = wrong                      # entries with x_true - x > 1e-3
    output['percent flow allocated incorrectly'] = per_flow
    sabs0 = ws.flow_metrics(scaling, x_true, x0, 1e-3)
    neg = ws.flow_metrics(scaling, x0, x_true, 1e-3)
    output['max|f * (x_init-x_true)|'] = max(sabs0[4], neg[4])  # max over entries of s |x_true - x0|
    return x_last, error, output


def main(args=None, plot=False):
    """main.py:146-176 without the command line and the plots: ``args`` carries ``file``, ``method``, ``eq``, ``init``,
    ``noise`` (an argparse.Namespace as the reference's tests build it, tests/fast/test_main.py:20-29).  File ->
    BSLSMatrices (device) -> solve_in_z -> LS_postprocess; returns (iters, times, states, output)."""
    from .bsls_matrices import BSLSMatrices
    config = {'full': True, 'L': True, 'OD': True, 'CP': True, 'LP': True, 'eq': args.eq, 'init': args.init}
    bm = BSLSMatrices(fname=args.file, **config)
    bm.degree_reduced_form()
    AA, bb, N, block_sizes, x_split, nz, scaling, rsort_index, x0 = bm.get_LS()
    output = bm.info
    if getattr(args, "noise", None):
        bb = bb + torch.as_tensor(np.random.normal(scale=np.abs(bb.cpu().numpy()) * args.noise)).to(bb.device)
    problem = bm.problem()
    options = getattr(args, "options", None)
    iters, times, states = solve_in_z(problem, bb, x0, N, block_sizes, args.method, options=options)
    x_last, error, output = LS_postprocess(states, x0, problem, bb, x_split, scaling=scaling, block_sizes=block_sizes, N=N,
                                           output=output)
    return iters, times, states, output
