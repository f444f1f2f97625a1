"""Drop-in for the reference's Cython module ``python/c_extensions/c_extensions.pyx``.

Same function names, positional order, in-place semantics and ``AssertionError`` behaviour.
Two kinds of buffers are accepted:

* ``torch.Tensor`` on a CUDA device (fp64, or fp32 as an extension): the device entry points
  of libbsls_b200 run asynchronously on the current stream, nothing leaves HBM;
* ``numpy.ndarray`` (C-contiguous float64, exactly what the reference takes): the host
  entry points are called, which copy to the GPU, run the same kernels and copy back.

Every computation happens in the CUDA library; there is no NumPy/CPU implementation here.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .plan import BlockPlan, plan_for

__all__ = ["proj_simplex_c", "proj_multi_simplex_c", "proj_multi_ball_c"]


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_dev_vector(y, name="y"):
    assert torch.is_tensor(y) and y.is_cuda, "%s must be a CUDA tensor" % name
    assert y.dim() == 1, "%s must be 1-D" % name
    # the reference silently works on a private copy of non-contiguous input
    # (c_extensions.pyx:27 np.ascontiguousarray) and drops the result; we refuse instead
    if not y.is_contiguous():
        raise ValueError("%s must be contiguous (in-place operation)" % name)
    if y.dtype not in (torch.float64, torch.float32):
        raise ValueError("Buffer dtype mismatch: expected float64 (or float32), got %s" % y.dtype)


def _check_host_vector(y, name="y"):
    if y.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got %s" % y.dtype)
    assert y.ndim == 1
    if not y.flags.c_contiguous:
        raise ValueError("%s must be C-contiguous (in-place operation)" % name)


def _host_blocks(blocks):
    b = np.asarray(blocks)
    # c_extensions.pyx:33-34
    assert False not in ((b[1:] - b[:-1]) > 0)
    return np.ascontiguousarray(b, dtype=np.int32)


def _project_multi(y, blocks, ball):
    L = _lib.lib()
    if isinstance(y, np.ndarray):
        _check_host_vector(y)
        b = _host_blocks(blocks)
        n = y.shape[0]
        assert b[0] >= 0 and b[-1] < n
        fn = L.bsls_proj_multi_ball if ball else L.bsls_proj_multi_simplex
        _lib.check(fn(y.ctypes.data, b.ctypes.data, int(b.shape[0]), int(n)), fn.__name__)
        return None
    _check_dev_vector(y)
    n = y.shape[0]
    plan = plan_for(blocks, n, y.device)
    if y.dtype == torch.float64:
        fn = L.bsls_dev_proj_multi_ball_f64 if ball else L.bsls_dev_proj_multi_simplex_f64
    else:
        fn = L.bsls_dev_proj_multi_ball_f32 if ball else L.bsls_dev_proj_multi_simplex_f32
    with torch.cuda.device(y.device):
        _lib.check(fn(plan.handle, y.data_ptr(), _stream(y)), fn.__name__)
    return None


def proj_simplex_c(y, start, end):
    """Project ``y[start:end]`` on the unit simplex, in place
    (reference: c_extensions.pyx:22-28 -> proj_simplex.h:17-34)."""
    n = y.shape[0]
    assert start >= 0 and start < n and end > 0 and end <= n
    if start >= end:
        return
    if isinstance(y, np.ndarray):
        _check_host_vector(y)
        L = _lib.lib()
        _lib.check(L.bsls_proj_simplex(y.ctypes.data, int(start), int(end)), "bsls_proj_simplex")
        return
    _check_dev_vector(y)
    # a single block [start, end) of the sub-vector y[:end]
    sub = y[:end]
    _project_multi(sub, BlockPlan(np.array([start]), end, y.device), ball=False)


def proj_multi_simplex_c(y, blocks):
    """Project every block of ``y`` on the unit simplex, in place
    (reference: c_extensions.pyx:31-39 -> proj_simplex.h:37-47).  ``blocks`` holds start
    offsets; it may also be a prebuilt :class:`BlockPlan`."""
    return _project_multi(y, blocks, ball=False)


def proj_multi_ball_c(y, blocks):
    """Clip negatives and project the blocks whose sum exceeds one ("lasso" feasible set)
    (reference: c_extensions.pyx:42-50 -> proj_simplex.h:50-74)."""
    return _project_multi(y, blocks, ball=True)
