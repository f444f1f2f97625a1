"""Projected L-BFGS with a weak-Wolfe bisection line search -- drop-in for ``python/LBFGS.py``
on device vectors (closures as in :mod:`BB`)."""
import math
import time

import numpy as np
import torch

from .sparse import axpby, default_workspace

__all__ = ["weak_wolfe_ls", "solve"]


def weak_wolfe_ls(x, d, f, nabla_f, proj=lambda x: x, c1=1e-3, c2=0.9):
    """Weak Wolfe line search by bisection (LBFGS.py:9-53).  ``proj`` may work in place: it is
    always handed a fresh vector."""
    ws = default_workspace(x.device)
    alpha, beta = 0, float('inf')
    t = 1
    stop = False
    proj_x = proj(x.clone())
    nabla_fx = nabla_f(proj_x)
    f_px = f(proj_x)
    d_dot_g, d_dot_d = ws.dots([(d, nabla_fx), (d, d)])
    norm_d = np.sqrt(d_dot_d)
    while not stop:
        proj_xtd = proj(axpby(torch.empty_like(x), 1.0, x, t, d))
        # armijo condition violated
        if f(proj_xtd) >= f_px + c1 * t * d_dot_g:
            beta = t
            t = 0.5 * (alpha + beta)
        # curvature condition violated
        elif ws.dot(d, nabla_f(proj_xtd)) < c2 * d_dot_g:
            alpha = t
            t = 2 * alpha if beta == float('inf') else 0.5 * (alpha + beta)
        else:
            stop = True
        if np.abs(alpha - beta) <= 1e-14:
            stop = True
        if abs(t) * norm_d <= 1e-8:          # la.norm(t*d)
            stop = True
    return t


def solve(x0, f, nabla_f, stopping, m=50, record_every=500, proj=None, log=None, options=None):
    """Limited memory BFGS (LBFGS.py:56-123)"""
    if log is None:
        log = lambda it, state, dur: time.time()
    ws = default_workspace(x0.device)

    def search_dir(g_new, y_new, s_new, rho, y, s, m=10):
        # two-loop recursion (LBFGS.py:60-71); pairs with rho == 0 are the zero-filled
        # history slots of the first iterations and contribute nothing
        q = g_new.clone()
        alpha = [0] * m
        for i in range(m - 1, -1, -1):
            if rho[i] == 0:
                continue
            alpha[i] = rho[i] * ws.dot(s[i], q)
            axpby(q, 1.0, q, -alpha[i], y[i])
        yy, ys = ws.dots([(y_new, y_new), (y_new, s_new)])
        H = ys / yy
        r = axpby(q, 0.0, q, H, q)
        for i in range(0, m):
            if rho[i] == 0:
                continue
            beta = rho[i] * ws.dot(y[i], r)
            axpby(r, 1.0, r, alpha[i] - beta, s[i])
        return axpby(r, 0.0, r, -1.0, r)

    start = log(0, x0, 0)

    i, stop = 0, False
    x = x0
    zero = torch.zeros_like(x)
    y, s = [zero] * m, [zero] * m
    g_new = nabla_f(x)
    y_new, s_new = g_new, torch.ones_like(x)

    rho, rho_new = [0] * m, 1 / ws.dot(y_new, s_new)
    while not stop:
        i += 1
        d = search_dir(g_new, y_new, s_new, rho, y, s, m=m)
        y.pop(0)
        y.append(y_new)
        s.pop(0)
        s.append(s_new)
        rho.pop(0)
        rho.append(rho_new)

        t = weak_wolfe_ls(x, d, f, nabla_f, proj=proj if proj else (lambda v: v))
        s_new = axpby(torch.empty_like(x), 0.0, d, t, d)          # t * d
        x_next = axpby(torch.empty_like(x), 1.0, x, 1.0, s_new)
        if proj:
            x_next = proj(x_next)
        g = g_new
        g_new = nabla_f(x_next)
        y_new = axpby(torch.empty_like(x), 1.0, g_new, -1.0, g)
        ys = ws.dot(y_new, s_new)
        if ys == 0:
            print("iter=%d, f=%8.5e" % (i, f(x_next)))
            print("Exiting... no change in gradient")
            break
        rho_new = 1 / ys

        x = x_next
        fx = f(x)
        if math.isnan(fx):
            raise ArithmeticError("objective function evaluates to NaN")
        stop = stopping(g_new, fx, i, t, d=d, options=options)

        if i % record_every == 0:
            start = log(i, x, time.time() - start)

    log(i, x, time.time() - start)
    return x
