"""Solver dispatcher -- drop-in for ``python/gradient_descent.py``."""
import logging
import time

import numpy as np

from . import BB, LBFGS, DORE, solvers
from .bsls_utils import lsv_operator
from .sparse import LsqProblem, axpby

__author__ = 'cathywu (reference); GPU mirror'


class GradientDescent:
    """gradient_descent.py:13-69.  ``A`` (DORE only) is an :class:`LsqProblem`, ``N`` an
    :class:`bsls_utils.NOperator`, vectors are device tensors."""

    def __init__(self, z0=None, f=None, nabla_f=None, proj=None, method='BB', options=None, A=None, N=None, target=None):
        self.z0 = z0
        self.f = f
        self.nabla_f = nabla_f
        self.proj = proj
        self.method = method
        self.A = A
        self.N = N
        self.target = target
        if options is None:
            self.options = {'max_iter': 300000, 'verbose': 1, 'opt_tol': 1e-30, 'suff_dec': 0.003, 'corrections': 500}
        else:
            self.options = options
        self.iters, self.times, self.states = [], [], []

        def log(iter_, state, duration):
            self.iters.append(iter_)
            self.times.append(duration)
            self.states.append(state)
            start = time.time()
            return start
        self.log = log

    def run(self):
        logging.debug('Starting %s solver...' % self.method)
        if self.method == 'LBFGS':
            import torch
            z1 = axpby(torch.empty_like(self.z0), 1.0, self.z0, 1.0, torch.ones_like(self.z0))   # z0 + 1
            LBFGS.solve(z1, self.f, self.nabla_f, solvers.stopping, log=self.log, proj=self.proj, options=self.options)
            logging.debug("Took %s time" % str(np.sum(self.times)))
        elif self.method == 'BB':
            BB.solve(self.z0, self.f, self.nabla_f, solvers.stopping, log=self.log, proj=self.proj, options=self.options)
        elif self.method == 'DORE':
            import torch
            alpha = 0.99
            lsv = lsv_operator(self.A, self.N)
            logging.info("Largest singular value: %s" % lsv)
            scale = alpha / lsv
            A, N = self.A, self.N
            NT = N.T
            x = torch.empty(A.n, dtype=torch.float64, device=A.device)
            target_dore = axpby(torch.empty_like(self.target), 0.0, self.target, scale, self.target)

            def linop(z):          # A_dore.dot(N.dot(z))
                N.dot(z, x)
                out = A.matvec(x)
                return axpby(out, 0.0, out, scale, out)

            def linop_T(r):        # N.T.dot(A_dore.T.dot(r))
                gx = A.rmatvec(r)
                axpby(gx, 0.0, gx, scale, gx)
                return NT.dot(gx)

            DORE.solve(self.z0, linop, linop_T, target_dore, proj=self.proj, log=self.log, options=self.options,
                       record_every=100)
        logging.debug('Stopping %s solver...' % self.method)
        return self.iters, self.times, self.states
