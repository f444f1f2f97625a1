"""DORE (double over-relaxation) accelerated projected Landweber iteration -- drop-in for
``python/DORE.py`` on device vectors.  ``linop(x) -> A x``, ``linop_T(r) -> A^T r`` return new
device vectors; the caller pre-scales the operator so that |A| < 1
(python/gradient_descent.py:57-61)."""
import logging
import time

import numpy as np
import torch

from .sparse import axpby, default_workspace

__all__ = ["solve"]


def solve(x0, linop, linop_T, target, record_every=5, proj=None, log=None, options=None, i=10000, eps=10 ** -16):
    """Solves DORE accelerated least squares via projection (DORE.py:6-90)"""
    if log is None:
        log = lambda it, state, dur: time.time()
    start = log(0, x0, 0)
    if options and 'max_iter' in options:
        i = options['max_iter']
    if options and 'opt_tol' in options:
        eps = options['opt_tol']
    ws = default_workspace(x0.device)
    new = lambda ref: torch.empty_like(ref)

    b = axpby(new(target), 0.0, target, -1.0, target)            # -target
    x = x0.clone()
    x_prev = x
    Ax = None
    Ax_prev = None
    Ax_prev_prev = None
    iter_ = 0
    for iter_ in range(i):
        Ax_prev_prev = Ax_prev
        Ax_prev = Ax
        Ax = linop(x)
        err = axpby(new(b), 1.0, b, -1.0, Ax)
        diff = axpby(new(x), 1.0, x, -1.0, x_prev)
        norm_change = ws.dot(diff, diff)                          # la.norm(x - x_prev)**2

        if iter_ > 0 and (norm_change <= eps):
            break
        x_new = axpby(new(x), 1.0, x, 1.0, linop_T(err))

        x_new = proj(x_new)
        Ax = linop(x_new)
        err = axpby(new(b), 1.0, b, -1.0, Ax)

        x_select = x_new
        if iter_ > 2:
            delta_Ax = axpby(new(Ax), 1.0, Ax, -1.0, Ax_prev)
            dp, de = ws.dots([(delta_Ax, delta_Ax), (delta_Ax, err)])
            if dp > 0:
                a1 = de / dp
                Ax_1 = axpby(new(Ax), 1 + a1, Ax, -a1, Ax_prev)
                dx = axpby(new(x), 1.0, x_new, -1.0, x)
                x_1 = axpby(new(x), 1.0, x_new, a1, dx)
                err_1 = axpby(new(b), 1.0, b, -1.0, Ax_1)

                delta_Ax = axpby(new(Ax), 1.0, Ax_1, -1.0, Ax_prev_prev)
                dp, de = ws.dots([(delta_Ax, delta_Ax), (delta_Ax, err_1)])
                if dp > 0:
                    a2 = de / dp
                    dx = axpby(new(x), 1.0, x_1, -1.0, x_prev)
                    x_2 = axpby(new(x), 1.0, x_1, a2, dx)
                    x_2 = proj(x_2)

                    Ax_2 = linop(x_2)
                    err_2 = axpby(new(b), 1.0, b, -1.0, Ax_2)
                    e2, e0 = ws.dots([(err_2, err_2), (err, err)])
                    if e2 / e0 < 1:
                        x_select = x_2
                        Ax = Ax_2

        x_prev = x
        x = x_select

        if iter_ % record_every == 0:
            start = log(iter_, x, time.time() - start)
        if options and 'verbose' in options and options['verbose'] >= 1 and iter_ % 100 == 0:
            logging.debug("iter=%d: %e %e %e" % (iter_, ws.dot(err, err), norm_change, ws.norm(x)))

    log(iter_, x, time.time() - start)
    return x
