"""Barzilai-Borwein projected gradient -- drop-in for ``python/BB.py`` on device vectors.

``f(x) -> float``, ``nabla_f(x) -> new device vector``, ``proj(x) -> device vector`` (may work
in place), ``stopping`` as in :mod:`solvers`, ``log(iteration, x, duration) -> start time``.
"""
import time

import numpy as np
import torch

from .sparse import axpby, default_workspace

__all__ = ["solve"]


def solve(x0, f, nabla_f, stopping, record_every=500, proj=None, log=None, options=None):
    """BB.py:7-45"""
    if log is None:
        log = lambda it, state, dur: time.time()
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
