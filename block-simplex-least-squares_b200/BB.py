"""Barzilai-Borwein projected gradient -- drop-in for ``python/BB.py`` on device vectors.

``f(x) -> float``, ``nabla_f(x) -> new device vector``, ``proj(x) -> device vector`` (may work
in place), ``stopping`` as in :mod:`solvers`, ``log(iteration, x, duration) -> start time``.
"""
import time

import numpy as np
import torch

from .sparse import axpby, default_workspace

__all__ = ["solve"]


def _solve_native(x0, f, nabla_f, zs, record_every, log, options):
    """The loop of BB.solve inside the library for the closures of main.z_space_parts (bsls_zbb_run_f64): segments of
    ``record_every`` iterations, a state recorded after each as the reference's log callback does (BB.py:39-43)."""
    import ctypes
    from . import _lib
    problem, N, zplan = zs["problem"], zs["N"], zs["zplan"]
    L = _lib.lib()
    start = log(0, x0, 0)
    max_iter = int(options['max_iter']) if options and 'max_iter' in options else 2 ** 31 - 1
    opt_tol = float(options['opt_tol']) if options and 'opt_tol' in options else 1e-6
    z = x0.clone()
    z_prev = axpby(torch.empty_like(z), 1.0, z, 1.0, torch.ones_like(z))       # x + 1
    g_prev = nabla_f(z_prev).clone()
    res = _lib.BatchResult()
    i = 0
    solve.last = {"device_ms": 0.0, "kernel_launches": 0}
    while True:
        i_end = min(max_iter, (i // record_every + 1) * record_every)
        with torch.cuda.device(z.device):
            st = torch.cuda.current_stream(z.device).cuda_stream
            _lib.check(L.bsls_zbb_run_f64(problem.handle, N.plan.handle, zplan.handle, z.data_ptr(), z_prev.data_ptr(), g_prev.data_ptr(),
                                          i, i_end, max_iter, opt_tol, ctypes.byref(res), st), "zbb_run")
        solve.last["device_ms"] += res.device_ms
        solve.last["kernel_launches"] += res.kernel_launches
        i = res.iterations
        if res.stop_code == 5:
            print('Exiting... no change in gradient')
            break
        if i % record_every == 0 and i > 0:      # BB.py:39-40 runs before the loop condition is looked at again
            start = log(i, z.clone(), time.time() - start)
        if res.stop_code != 0 or i >= max_iter:
            break
    solve.last.update(iterations=i, stop_code=res.stop_code, f=res.f)
    log(i, z, time.time() - start)
    return z


def solve(x0, f, nabla_f, stopping, record_every=500, proj=None, log=None, options=None):
    """BB.py:7-45.  With the closures of main.z_space_parts and solvers.stopping the loop runs inside the library."""
    if log is None:
        log = lambda it, state, dur: time.time()
    zs = getattr(f, "zspace", None)
    from . import solvers as _solvers
    if (zs is not None and getattr(nabla_f, "zspace", None) is zs and getattr(proj, "zspace", None) is zs
            and stopping is _solvers.stopping and not (options and options.get("generic_loop"))):
        return _solve_native(x0, f, nabla_f, zs, record_every, log, options)
    start = log(0, x0, 0)
    ws = default_workspace(x0.device)
    ones = torch.ones_like(x0)

    i, stop = 0, False
    x = x0
    x_prev = axpby(torch.empty_like(x), 1.0, x, 1.0, ones)      # x + 1
    g_prev = nabla_f(x_prev)
    delta_g = torch.empty_like(x)
    delta_x = torch.empty_like(x)

    while not stop:
        i += 1
        g = nabla_f(x)
        axpby(delta_g, 1.0, g, -1.0, g_prev)
        axpby(delta_x, 1.0, x, -1.0, x_prev)
        sxy, syy, total = ws.dots([(delta_x, delta_g), (delta_g, delta_g), (delta_g, ones)])
        if total == 0:                                           # sum(delta_g) == 0
            print('Exiting... no change in gradient')
            break
        t = sxy / syy                                            # BB step
        if np.abs(t) <= 1e-10 or np.abs(t) > 1e10:
            print('BB update is having some trouble, implement fix! t=%8.5e' % t)
        x_next = axpby(torch.empty_like(x), 1.0, x, -t, g)       # next position

        x_prev, x = x, x_next
        g_prev = g

        if proj:
            x = proj(x)
        fx = f(x)
        stop = stopping(g, fx, i, t, delta_g=delta_g, options=options)

        if i % record_every == 0:
            start = log(i, x, time.time() - start)

    log(i, x, time.time() - start)
    return x
