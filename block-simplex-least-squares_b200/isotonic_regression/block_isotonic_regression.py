"""Per-block isotonic regression -- drop-in for
``python/isotonic_regression/block_isotonic_regression.py``, which loops over the blocks in
Python and calls scikit-learn's IsotonicRegression on each.  Here all blocks are regressed by
one launch of the segmented PAVA kernel of libbsls_b200."""
import numpy as np
import torch

from ..c_extensions import isotonic_regression_multi_c

__all__ = ["block_isotonic_regression", "block_isotonic_regression_2"]


def block_isotonic_regression_2(x, blocks_start):
    """In place, no clipping (block_isotonic_regression.py:19-24)."""
    isotonic_regression_multi_c(x, blocks_start, None, 1)


def block_isotonic_regression(x, ir, block_sizes, blocks_start, blocks_end):
    """Returns a NEW vector: the regression of the z-blocks x[s:e] (sizes block_sizes - 1),
    clipped to [0, 1]; empty z-blocks are dropped (block_isotonic_regression.py:8-16).  ``ir`` (the
    scikit-learn regressor of the reference) is ignored."""
    sizes = np.asarray(block_sizes.cpu() if torch.is_tensor(block_sizes) else block_sizes) - 1
    s = np.asarray(blocks_start.cpu() if torch.is_tensor(blocks_start) else blocks_start)
    e = np.asarray(blocks_end.cpu() if torch.is_tensor(blocks_end) else blocks_end)
    keep = sizes > 0
    contiguous = np.all(s[1:] == e[:-1]) and np.all(e - s == sizes)
    assert contiguous, "z-blocks must tile the vector (as bsls_utils / main.py build them)"
    out = x[int(s[0]):int(e[-1])].clone()
    starts = (s[keep] - s[0]).astype(np.int64)
    if len(starts):
        isotonic_regression_multi_c(out, starts, None, 1, clip01=True)
    return out
