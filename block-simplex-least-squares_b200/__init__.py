"""bsls_b200 -- B200-native (sm_100a) inner loop of block-simplex least squares.

Import as ``bsls_b200`` (see bsls_b200.py at the repository root).  The public surface
mirrors the reference's operator API for the hot path:

    bsls_b200.c_extensions    proj_simplex_c, proj_multi_simplex_c, proj_multi_ball_c, isotonic_regression_*_c,
                              x2z_c, z2x_c          (reference: python/c_extensions/c_extensions.pyx)
    bsls_b200.isotonic_regression.block_isotonic_regression
    bsls_b200.algorithm_utils get_solver_parts, sparse_least_squares_obj, line_search_np, stopping, normalization
    bsls_b200.BATCH           solve, solve_BB, solve_LBFGS, solve_MD
    bsls_b200.BB / LBFGS / DORE / mirror_descent / solvers / gradient_descent / bsls_utils / main.solve_in_z
    bsls_b200.bsls_matrices   BSLSMatrices (problem construction on the device); main.LS_postprocess, main.main

All compute runs in libbsls_b200.so (hand-written CUDA behind a C ABI, include/bsls_b200.h).
"""
from . import _lib
from . import c_extensions
from .c_extensions import (proj_simplex_c, proj_multi_simplex_c, proj_multi_ball_c, isotonic_regression_c,
                           isotonic_regression_multi_c, isotonic_regression_c_2, isotonic_regression_multi_c_2,
                           isotonic_regression_c_3, isotonic_regression_multi_c_3)
from .c_extensions import x2z_c, z2x_c
from .plan import BlockPlan, plan_for
from .sparse import LsqProblem, Communicator, Workspace
from . import algorithm_utils, BATCH, BB, LBFGS, DORE, mirror_descent, solvers, gradient_descent, bsls_utils, main, bsls_matrices
from .isotonic_regression import block_isotonic_regression

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
