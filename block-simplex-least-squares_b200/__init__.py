"""bsls_b200 -- B200-native (sm_100a) inner loop of block-simplex least squares.

Import as ``bsls_b200`` (see bsls_b200.py at the repository root).  The public surface
mirrors the reference's operator API for the hot path:

    bsls_b200.c_extensions    proj_simplex_c, proj_multi_simplex_c, proj_multi_ball_c, ...
                              (reference: python/c_extensions/c_extensions.pyx)

All compute runs in libbsls_b200.so (hand-written CUDA behind a C ABI, include/bsls_b200.h).
"""
from . import _lib
from . import c_extensions
from .c_extensions import (proj_simplex_c, proj_multi_simplex_c, proj_multi_ball_c, isotonic_regression_c,
                           isotonic_regression_multi_c, isotonic_regression_c_2, isotonic_regression_multi_c_2,
                           isotonic_regression_c_3, isotonic_regression_multi_c_3)
from .plan import BlockPlan, plan_for

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
