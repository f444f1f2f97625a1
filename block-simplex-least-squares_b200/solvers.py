"""Stopping rule and the least-squares front door -- drop-in for ``python/solvers.py``
(the ``qp`` / ``qp2`` stubs of the reference, which only ``pass``, are not carried over)."""
import logging

import numpy as np

from .sparse import default_workspace

__all__ = ["stopping", "least_squares"]


def least_squares(x, linop, linop_transpose, target, proj=None, diagnostics=None, options=None, log=None):
    """solvers.py:35-37 -> DORE.solve"""
    from . import DORE
    if log is None:
        log = lambda it, state, dur: 0.0
    return DORE.solve(x, linop, linop_transpose, target, proj=proj, log=log, options=options)


def stopping(g, fx, i, t, d=None, delta_g=None, options=None, TOLER=1e-6):
    """Stopping condition (solvers.py:40-63).  ``g``, ``d``, ``delta_g`` are device vectors; the
    three norms come from ONE pass (a fused multi-dot kernel)."""
    if options and 'max_iter' in options:
        if i >= options['max_iter']:
            return True
    if options and 'opt_tol' in options:
        TOLER = options['opt_tol']
    ws = default_workspace(g.device)
    pairs = [(g, g)]
    if d is not None:
        pairs.append((d, d))
    if delta_g is not None:
        pairs.append((delta_g, delta_g))
    dots = ws.dots(pairs)
    norm2_nabla_f = dots[0]          # np.square(la.norm(g))
    thresh = TOLER * (1 + abs(fx))
    if options and 'verbose' in options and options['verbose'] >= 1 and i % 100 == 0:
        logging.debug("iter=%d: %e %e %e %f" % (i, t, norm2_nabla_f, thresh, fx))
    if norm2_nabla_f <= thresh:
        logging.info("iter=%d: %e %e %e %f" % (i, t, norm2_nabla_f, thresh, fx))
        logging.warning('Exiting... norm(grad) too small')
        return True
    k = 1
    if d is not None:
        if abs(t) * np.sqrt(dots[k]) <= 1e-12:   # la.norm(t*d)
            logging.info("iter=%d: %e %e %e %f" % (i, t, norm2_nabla_f, thresh, fx))
            logging.warning('Exiting... step too small')
            return True
        k += 1
    if delta_g is not None and np.sqrt(dots[k]) == 0:
        logging.info("iter=%d: %e %e %e %f" % (i, t, norm2_nabla_f, thresh, fx))
        logging.warning('Exiting... no change in gradient')
        return True
    return False
