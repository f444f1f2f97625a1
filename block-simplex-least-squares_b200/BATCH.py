"""Batch solvers on the GPU -- drop-in for the reference's ``python/BATCH.py``.

Same signatures, same return dictionaries (``f, x, stop, iterations, progress``).  Vectors are
float64 CUDA tensors.  Two execution paths:

* NATIVE: when ``obj`` / ``proj`` / ``line_search`` are the closures of
  :func:`algorithm_utils.get_solver_parts` for a sparse problem, the whole loop runs inside
  libbsls_b200 (``bsls_batch_solve_f64``): per iteration one update kernel, one projection
  kernel, the SpMV pair with the step / line-search dot products fused into its epilogue, and
  one 128-byte read-back of scalars.  No interpreter in the loop.
* GENERIC: arbitrary closures; the loop below is the reference's, statement for statement, with
  every NumPy vector expression replaced by a library kernel.
"""
import ctypes
import time
from collections import deque

import numpy as np
import torch

from . import _lib
from .algorithm_utils import stopping, normalization
from .sparse import axpby, copy_, default_workspace

__all__ = ["solve", "solve_BB", "solve_LBFGS", "LBFGS_helper", "solve_MD"]

_STOP = {0: 'continue', 1: 'max_iter'}


def _stage_in(x_init, obj):
    """The reference's solvers take and return NumPy vectors.  Host input (ndarray, or a CPU tensor -- pinned memory makes
    the copy asynchronous) is copied to the device of the problem; the result then goes back the same way.

    float32 input (host or device) is accepted as well: the path is bound by 8-byte gathers that cost one 32-byte sector
    each whatever the element size, so single precision would buy no time; it is widened on the way in, the loop runs in
    float64 and ``x`` comes back as float32 (north_star's fp32 bar, 1e-4, is met with room: tests/test_solvers_gpu.py)."""
    if torch.is_tensor(x_init) and x_init.is_cuda:
        if x_init.dtype == torch.float32:
            staged = x_init.to(torch.float64)
            staged._bsls_private = True
            return staged, ("cuda", True)
        return x_init, None
    problem = getattr(obj, "problem", None)
    device = problem.device if problem is not None else torch.device("cuda", torch.cuda.current_device())
    f32 = (x_init.dtype == torch.float32) if torch.is_tensor(x_init) else (np.asarray(x_init).dtype == np.float32)
    if f32:
        host = x_init if torch.is_tensor(x_init) else torch.from_numpy(np.ascontiguousarray(x_init))
        assert host.dim() == 1, "x_init: vector expected"
        staged = host.to(device, non_blocking=True).to(torch.float64)
    else:
        host = x_init if torch.is_tensor(x_init) else torch.from_numpy(np.ascontiguousarray(x_init, dtype=np.float64))
        assert host.dtype == torch.float64 and host.dim() == 1, "x_init: float64 (or float32) vector expected"
        staged = host.to(device, non_blocking=True)
    staged._bsls_private = True        # a fresh device copy: the native loop may work in it instead of cloning again
    return staged, ("tensor" if torch.is_tensor(x_init) else "numpy", f32)


_PINNED_OUT = {}


def _stage_out(sol, kind, like=None):
    """Result back to the host.  NumPy in -> a new NumPy array out (as the reference).  A pinned CPU tensor in -> a pinned
    CPU tensor out; pinning memory costs far more than the copy, so ONE pinned result buffer per vector length is kept
    and reused by the next call of that length (copy it if two results must be alive at once)."""
    if kind is None:
        return sol
    where, f32 = kind
    x = sol['x'].to(torch.float32) if f32 else sol['x']
    if where == "cuda":
        sol['x'] = x
    elif where == "numpy":
        sol['x'] = x.cpu().numpy()
    elif like is not None and like.is_pinned():
        key = (x.shape[0], x.dtype)
        out = _PINNED_OUT.get(key)
        if out is None:
            out = _PINNED_OUT[key] = torch.empty(x.shape, dtype=x.dtype).pin_memory()
        out.copy_(x, non_blocking=False)
        sol['x'] = out
    else:
        sol['x'] = x.cpu()
    return sol


def _host_api(fn):
    """Wraps a solver so that host vectors are accepted for ``x_init`` (see _stage_in)."""
    def wrapped(*args, **kw):
        names = fn.__code__.co_varnames[:fn.__code__.co_argcount]
        k = names.index("x_init")
        x_init = args[k] if k < len(args) else kw["x_init"]
        obj = args[0] if args else kw["obj"]
        xd, kind = _stage_in(x_init, obj)
        if k < len(args):
            args = args[:k] + (xd,) + args[k + 1:]
        else:
            kw["x_init"] = xd
        return _stage_out(fn(*args, **kw), kind, x_init if torch.is_tensor(x_init) else None)
    wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
    return wrapped


def _native_parts(obj, proj, line_search=None, need_proj=True):
    problem = getattr(obj, "problem", None)
    if problem is None:
        return None
    if need_proj:
        plan = getattr(proj, "plan", None)
        if plan is None or getattr(proj, "scaled", False) or plan.n != problem.n:
            return None
    if line_search is not None and getattr(line_search, "obj", None) is not obj:
        return None
    return problem


def _solve_native(problem, plan, method, proj_mode, x_init, use_line_search, f_min, opt_tol, max_iter, prog_tol, min_eig=0.0,
                  corrections=0):
    L = _lib.lib()
    x = x_init if getattr(x_init, "_bsls_private", False) else x_init.clone()
    opts = _lib.BatchOpts(method=method, proj_mode=proj_mode, use_line_search=int(bool(use_line_search)),
                          has_f_min=int(f_min is not None), f_min=0.0 if f_min is None else float(f_min),
                          opt_tol=float(opt_tol), prog_tol=float(prog_tol), min_eig=float(min_eig), max_iter=int(max_iter),
                          corrections=int(corrections))
    res = _lib.BatchResult()
    cap = max(2, int(max_iter) + 1)
    pf = (ctypes.c_double * cap)()
    pt = (ctypes.c_double * cap)()
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(L.bsls_batch_solve_f64(problem.handle, plan.handle, x.data_ptr(), ctypes.byref(opts), ctypes.byref(res),
                                          pf, pt, cap, stream), "batch_solve")
    if res.stop_code == 2:
        stop = 'f-f_min = {} < opt_tol'.format(res.stop_value)
    elif res.stop_code == 3:
        stop = '|f_old-f| = {} < prog_tol'.format(res.stop_value)
    else:
        stop = _STOP.get(res.stop_code, 'continue')
    npts = min(cap, res.iterations)
    progress = [[pt[k], pf[k]] for k in range(npts)]
    return {'f': res.f, 'x': x, 'stop': stop, 'iterations': res.iterations, 'progress': progress,
            'obj_evals': res.obj_evals, 'backtracks': res.backtracks, 'kernel_launches': res.kernel_launches,
            'device_ms': res.device_ms}


@_host_api
def solve(obj, proj, step_size, x_init, line_search=None, f_min=None, opt_tol=1e-6,
          max_iter=2000, prog_tol=1e-12):
    """Projected batch gradient descent with line search (BATCH.py:7-52)
    obj: f = obj(x, g), g overwritten with the gradient at x
    proj: proj(x), in place
    step_size: step_size(i)
    """
    problem = _native_parts(obj, proj, line_search)
    if problem is not None and hasattr(step_size, "min_eig"):
        return _solve_native(problem, proj.plan, 0, proj.mode, x_init, line_search is not None, f_min, opt_tol, max_iter,
                             prog_tol, step_size.min_eig)
    x = x_init.clone()
    g = torch.zeros_like(x)
    g_new = torch.zeros_like(x)
    x_new = torch.zeros_like(x)
    f_old = float('inf')
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    start_time = time.time()
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        t = step_size(i)
        axpby(x_new, 1.0, x, -t, g)
        proj(x_new)
        f_new = obj(x_new, g_new)
        if line_search is not None:
            f_new = line_search(x, f, g, x_new, f_new, g_new, i)
        f_old = f
        f = f_new
        copy_(x, x_new)
        copy_(g, g_new)
        i += 1
        progress.append([time.time() - start_time, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}


@_host_api
def solve_BB(obj, proj, line_search, x_init, f_min=None, opt_tol=1e-6,
             max_iter=2000, prog_tol=1e-12):
    """Projected batch gradient descent with Barzilai-Borwein step (BATCH.py:55-106)"""
    problem = _native_parts(obj, proj, line_search)
    if problem is not None:
        return _solve_native(problem, proj.plan, 1, proj.mode, x_init, True, f_min, opt_tol, max_iter, prog_tol)
    ws = default_workspace(x_init.device)
    x = x_init.clone()
    g = torch.zeros_like(x)
    delta_x = torch.zeros_like(x)
    delta_g = torch.zeros_like(x)
    g_new = torch.zeros_like(x)
    x_new = torch.zeros_like(x)
    f_old = float('inf')
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    start_time = time.time()
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        if i == 1:
            axpby(x_new, 1.0, x, -1.0, g)
        else:
            sxy, syy = ws.dots([(delta_x, delta_g), (delta_g, delta_g)])
            t = sxy / syy
            axpby(x_new, 1.0, x, -t, g)
        proj(x_new)
        f_new = obj(x_new, g_new)
        f_new = line_search(x, f, g, x_new, f_new, g_new, i)
        f_old = f
        f = f_new
        axpby(delta_x, 1.0, x_new, -1.0, x)
        axpby(delta_g, 1.0, g_new, -1.0, g)
        copy_(x, x_new)
        copy_(g, g_new)
        i += 1
        progress.append([time.time() - start_time, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}


@_host_api
def solve_LBFGS(obj, proj, line_search, x_init, f_min=None, opt_tol=1e-6,
                max_iter=1000, prog_tol=1e-12, corrections=50):
    """Projected L-BFGS (BATCH.py:110-193).  As in the reference, the history deques hold
    REFERENCES to the two difference buffers, which the loop overwrites in place every
    iteration -- so every stored pair aliases the latest (delta_x, delta_g) while the stored
    curvatures ``rho`` stay distinct.  That is what the reference computes, and what its
    results (and tests) are pinned to, so it is kept.

    NATIVE path (closures of get_solver_parts, at most 64 corrections): the loop runs inside the library.  Because all
    pairs are the same two vectors, the two-loop recursion reduces to scalar recurrences on <s,g>, <y,g>, <s,y>, <y,y>
    (taken on the device) and d = cg g + cy delta_g + cs delta_x: one vector pass instead of 2 x corrections."""
    problem = _native_parts(obj, proj, line_search)
    if problem is not None and corrections <= 64:
        return _solve_native(problem, proj.plan, 5, proj.mode, x_init, True, f_min, opt_tol, max_iter, prog_tol,
                             corrections=corrections)
    ws = default_workspace(x_init.device)
    q_delta_g = deque()
    q_delta_x = deque()
    q_rho = deque()
    n = x_init.shape[0]
    x = x_init.clone()
    g = torch.zeros_like(x)
    d = torch.zeros_like(x)
    alpha = torch.zeros(2 * corrections + 4, dtype=torch.float64, device=x.device)  # device scalars of the two-loop recursion
    delta_x = torch.zeros_like(x)
    delta_g = torch.zeros_like(x)
    g_new = torch.zeros_like(x)
    x_new = torch.zeros_like(x)
    f_old = float('inf')
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    start_time = time.time()
    sxy = syy = 0.0
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        if i == 1:
            axpby(x_new, 1.0, x, -1.0, g)
        else:
            sxy, syy = ws.dots([(delta_x, delta_g), (delta_g, delta_g)])
            q_delta_g.append(delta_g)
            q_delta_x.append(delta_x)
            q_rho.append(1 / sxy)
            if i > corrections + 1:
                q_delta_g.popleft()
                q_delta_x.popleft()
                q_rho.popleft()
            if i <= 5:
                # d more Barzilai-Borwein steps
                axpby(d, 0.0, g, -(sxy / syy), g)
            else:
                LBFGS_helper(q_delta_g, q_delta_x, q_rho, g, d, alpha, ws=ws, bb=(sxy, syy))
            axpby(x_new, 1.0, x, 1.0, d)
        proj(x_new)
        f_new = obj(x_new, g_new)
        f_new = line_search(x, f, g, x_new, f_new, g_new, i)
        f_old = f
        f = f_new
        axpby(delta_x, 1.0, x_new, -1.0, x)
        axpby(delta_g, 1.0, g_new, -1.0, g)
        copy_(x, x_new)
        copy_(g, g_new)
        i += 1
        progress.append([time.time() - start_time, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}


def LBFGS_helper(q_delta_g, q_delta_x, q_rho, g, d, alpha, ws=None, bb=None):
    """Two-loop recursion (BATCH.py:196-214), chained on the device: every step is ONE kernel
    that applies the previous correction to ``d`` and forms the next inner product in the same
    pass; the scalars alpha_j / beta_j never visit the host.

    ``alpha`` is a float64 device tensor with at least 2*m+1 entries (scratch for the inner
    products): alpha[j] = <s_j, d> of the first loop, alpha[m + j] = <y_j, d> of the second."""
    if ws is None:
        ws = default_workspace(g.device)
    m = len(q_delta_g)
    base = alpha.data_ptr()
    slot = lambda k: base + 8 * k
    copy_(d, g)
    # first loop: alpha_j = rho_j <s_j, d> ; d -= alpha_j y_j     (j = m-1 .. 0)
    ws.axpy_dot(d, 1.0, None, None, None, q_delta_x[m - 1], slot(m - 1))      # <s_{m-1}, d>
    for j in range(m - 1, -1, -1):
        nxt = q_delta_x[j - 1] if j > 0 else None
        ws.axpy_dot(d, -q_rho[j], slot(j), None, q_delta_g[j], nxt, slot(j - 1) if j > 0 else None)
    if bb is None:
        sxy, syy = ws.dots([(q_delta_x[-1], q_delta_g[-1]), (q_delta_g[-1], q_delta_g[-1])])
    else:
        sxy, syy = bb
    t = sxy / syy
    # d *= t, and <y_0, d> for the second loop
    ws.axpy_dot(d, t, None, None, None, q_delta_g[0], slot(m))
    # second loop: beta_j = rho_j <y_j, d> ; d += s_j (alpha_j - beta_j)    (j = 0 .. m-1)
    for j in range(m):
        nxt = q_delta_g[j + 1] if j + 1 < m else None
        ws.axpy_dot(d, q_rho[j], slot(j), slot(m + j), q_delta_x[j], nxt, slot(m + j + 1) if j + 1 < m else None)
    ws.axpy_dot(d, -1.0, None, None, None, None, None)  # d *= -1.0


@_host_api
def solve_MD(obj, block_starts, step_size, x_init, line_search=None, f_min=None, opt_tol=1e-6,
             max_iter=1000, prog_tol=0.0):
    """mirror descent algorithm (BATCH.py:217-250)"""
    from .plan import plan_for
    problem = getattr(obj, "problem", None)
    plan = plan_for(block_starts, x_init.shape[0], x_init.device)
    if problem is not None and hasattr(step_size, "min_eig"):
        return _solve_native(problem, plan, 2, 0, x_init, False, f_min, opt_tol, max_iter, prog_tol, step_size.min_eig)
    ws = default_workspace(x_init.device)
    x = x_init.clone()
    g = torch.zeros_like(x)
    g_new = torch.zeros_like(x)
    x_new = torch.zeros_like(x)
    f_old = float('inf')
    i = 1
    f = obj(x, g)
    progress = [[0.0, f]]
    start_time = time.time()
    while True:
        flag, stop = stopping(i, max_iter, f, f_old, opt_tol, prog_tol, f_min)
        if flag is True:
            break
        t = step_size(i)
        ws.md_update(plan, x_new, x, g, t)  # x * exp(-t*g), then normalize
        f_new = obj(x_new, g_new)
        f_old = f
        f = f_new
        copy_(x, x_new)
        copy_(g, g_new)
        i += 1
        progress.append([time.time() - start_time, f])
    return {'f': f, 'x': x, 'stop': stop, 'iterations': i, 'progress': progress}
