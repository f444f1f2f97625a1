"""Solver dispatcher -- drop-in for ``python/gradient_descent.py`` (class ``GradientDescent``: same constructor
arguments, same ``run()`` result ``(iters, times, states)``, same ``log`` callback contract).

The three methods are entries of a table; each one binds the device-side pieces its solver needs:

* ``'BB'``     -> :func:`BB.solve` (native z-space loop when the closures come from ``main.z_space_parts``)
* ``'LBFGS'``  -> :func:`LBFGS.solve`, started from ``z0 + 1`` as the reference does (gradient_descent.py:44)
* ``'DORE'``   -> :func:`DORE.solve` on the operator ``z -> (alpha / sigma_max) A N z`` with ``sigma_max`` from the Lanczos
  estimate of :func:`bsls_utils.lsv_operator` (gradient_descent.py:51-66)
"""
import logging
import time

import numpy as np
import torch

from . import BB, LBFGS, DORE, solvers
from .bsls_utils import lsv_operator
from .sparse import LsqProblem, axpby

__author__ = 'cathywu (reference); GPU mirror'

_DEFAULT_OPTIONS = {'max_iter': 300000, 'verbose': 1, 'opt_tol': 1e-30, 'suff_dec': 0.003, 'corrections': 500}
_DORE_ALPHA = 0.99          # safety factor on 1 / sigma_max (gradient_descent.py:52)
_DORE_RECORD_EVERY = 100


def _scaled(t, c, out=None):
    """c * t through the library's axpby (no torch arithmetic on the path)."""
    out = torch.empty_like(t) if out is None else out
    return axpby(out, 0.0, t, c, t)


class GradientDescent:
    """gradient_descent.py:13-69.  ``A`` (DORE only) is an :class:`LsqProblem`, ``N`` an
    :class:`bsls_utils.NOperator`, vectors are device tensors."""

    def __init__(self, z0=None, f=None, nabla_f=None, proj=None, method='BB', options=None, A=None, N=None, target=None):
        self.z0, self.f, self.nabla_f, self.proj = z0, f, nabla_f, proj
        self.method, self.A, self.N, self.target = method, A, N, target
        self.options = dict(_DEFAULT_OPTIONS) if options is None else options
        self.iters, self.times, self.states = [], [], []

    def log(self, iter_, state, duration):
        """The solvers' progress callback: records one (iteration, state, seconds) triple and restarts their clock."""
        self.iters.append(iter_)
        self.times.append(duration)
        self.states.append(state)
        return time.time()

    # ---- one entry per method ----------------------------------------------------------------------------
    def _bb(self):
        BB.solve(self.z0, self.f, self.nabla_f, solvers.stopping, log=self.log, proj=self.proj, options=self.options)

    def _lbfgs(self):
        start = axpby(torch.empty_like(self.z0), 1.0, self.z0, 1.0, torch.ones_like(self.z0))   # z0 + 1
        LBFGS.solve(start, self.f, self.nabla_f, solvers.stopping, log=self.log, proj=self.proj, options=self.options)
        logging.debug("Took %s time" % str(np.sum(self.times)))

    def _dore(self):
        A, N = self.A, self.N
        sigma = lsv_operator(A, N)
        logging.info("Largest singular value: %s" % sigma)
        c = _DORE_ALPHA / sigma
        NT = N.T
        x = torch.empty(A.n, dtype=torch.float64, device=A.device)

        def forward(z):          # c * A (N z)
            N.dot(z, x)
            r = A.matvec(x)
            return _scaled(r, c, out=r)

        def adjoint(r):          # N^T (c * A^T r)
            g = A.rmatvec(r)
            return NT.dot(_scaled(g, c, out=g))

        DORE.solve(self.z0, forward, adjoint, _scaled(self.target, c), proj=self.proj, log=self.log, options=self.options,
                   record_every=_DORE_RECORD_EVERY)

    _METHODS = {'BB': _bb, 'LBFGS': _lbfgs, 'DORE': _dore}

    def run(self):
        logging.debug('Starting %s solver...' % self.method)
        entry = self._METHODS.get(self.method)
        if entry is not None:       # an unknown name runs nothing and returns the empty logs, as in the reference
            entry(self)
        logging.debug('Stopping %s solver...' % self.method)
        return self.iters, self.times, self.states
