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

__all__ = ["proj_simplex_c", "proj_multi_simplex_c", "proj_multi_ball_c",
           "isotonic_regression_c", "isotonic_regression_multi_c",
           "isotonic_regression_c_2", "isotonic_regression_multi_c_2",
           "isotonic_regression_c_3", "isotonic_regression_multi_c_3", "x2z_c", "z2x_c", "block_scale", "n_dot", "nt_dot"]


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


def _check_f64_vectors(*pairs):
    """x2z / z2x / N / N^T / block_scale exist in fp64 only: an fp32 buffer would be read as 8-byte
    elements.  Same error as the reference's Cython layer for a wrong dtype."""
    for t, name in pairs:
        _check_dev_vector(t, name)
        if t.dtype != torch.float64:
            raise ValueError("Buffer dtype mismatch, expected 'double' but got %s for %s" % (t.dtype, name))
    dev = pairs[0][0].device
    for t, name in pairs[1:]:
        assert t.device == dev, "%s is on %s, expected %s" % (name, t.device, dev)


def _check_host_vector(y, name="y"):
    if y.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got %s" % y.dtype)
    assert y.ndim == 1
    if not y.flags.c_contiguous:
        raise ValueError("%s must be C-contiguous (in-place operation)" % name)


def _host_blocks(blocks):
    """Block starts for the host C ABI.  The reference's asserts (c_extensions.pyx:33-34: strictly increasing,
    in range) are checked by the library itself in one vectorised pass over the int32 array and come back as
    AssertionError through ``_lib.check``; a second NumPy pass here cost about a millisecond per million blocks
    before the first byte moved."""
    b = np.asarray(blocks)
    assert b.ndim == 1 and b.shape[0] > 0
    if b.dtype != np.int32:
        if b.dtype.kind not in "iu":
            raise ValueError("Buffer dtype mismatch: block starts must be integers, got %s" % b.dtype)
        assert int(b.max()) < 2 ** 31 and int(b.min()) >= 0  # the C layer indexes with int (proj_simplex.h:37)
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


# ---------------------------------------------------------------------------------------------
# isotonic regression (reference: c_extensions.pyx:52-138 -> isotonic_regression.h)
# ---------------------------------------------------------------------------------------------
def _host_weight(weight, n):
    """The reference updates ``weight`` in place only when it is already a C-contiguous int32
    array (np.ascontiguousarray otherwise copies, c_extensions.pyx:72); same here."""
    if weight is None:
        return None, None
    w = np.ascontiguousarray(weight, dtype=np.int32)
    assert w.shape[0] == n
    return w, w.ctypes.data


def _pava_multi(y, blocks, weight, update, clip01=False, variant=1):
    L = _lib.lib()
    if isinstance(y, np.ndarray):
        _check_host_vector(y)
        b = _host_blocks(blocks)
        n = y.shape[0]
        assert b[0] >= 0 and b[-1] < n
        if variant == 2:
            _lib.check(L.bsls_isotonic_regression_multi_2(y.ctypes.data, b.ctypes.data, int(b.shape[0]), int(n)),
                       "bsls_isotonic_regression_multi_2")
        else:
            w, wptr = _host_weight(weight, n)
            fn = L.bsls_isotonic_regression_multi if variant == 1 else L.bsls_isotonic_regression_multi_3
            _lib.check(fn(y.ctypes.data, b.ctypes.data, int(b.shape[0]), int(n), wptr, int(update)), fn.__name__)
        if clip01:
            np.clip(y, 0.0, 1.0, out=y)
        return None
    _check_dev_vector(y)
    n = y.shape[0]
    plan = plan_for(blocks, n, y.device)
    wptr = None
    if weight is not None and variant != 2:
        assert torch.is_tensor(weight) and weight.is_cuda and weight.dtype == torch.int32 and weight.is_contiguous(), \
            "weight must be a contiguous int32 CUDA tensor"
        assert weight.shape[0] == n
        wptr = weight.data_ptr()
    if variant != 1:
        # variants 2 / 3: the reference's routines as written (values and weights bit-identical), fp64 only
        if y.dtype != torch.float64:
            raise ValueError("Buffer dtype mismatch, expected 'double' but got %s" % y.dtype)
        with torch.cuda.device(y.device):
            if variant == 2:
                _lib.check(L.bsls_dev_isotonic_regression_multi_2_f64(plan.handle, y.data_ptr(), _stream(y)), "isotonic_regression_multi_2")
            else:
                if wptr is None:  # weight=None: ones in, result dropped (c_extensions.pyx:118-119)
                    weight = torch.ones(n, dtype=torch.int32, device=y.device)
                    wptr = weight.data_ptr()
                _lib.check(L.bsls_dev_isotonic_regression_multi_3_f64(plan.handle, y.data_ptr(), wptr, int(update), _stream(y)),
                           "isotonic_regression_multi_3")
        assert not clip01, "the fused clamp exists for variant 1 only"
        return None
    fn = L.bsls_dev_isotonic_regression_multi_f64 if y.dtype == torch.float64 else L.bsls_dev_isotonic_regression_multi_f32
    with torch.cuda.device(y.device):
        _lib.check(fn(plan.handle, y.data_ptr(), wptr, int(update), int(bool(clip01)), _stream(y)), fn.__name__)
    return None


def _pava_single(y, start, end, weight, update, variant):
    n = y.shape[0]
    assert start >= 0 and start < n and end > 0 and end <= n
    if start >= end:
        return
    if isinstance(y, np.ndarray):
        _check_host_vector(y)
        L = _lib.lib()
        if variant == 2:
            _lib.check(L.bsls_isotonic_regression_2(y.ctypes.data, int(start), int(end)), "bsls_isotonic_regression_2")
            return
        w, wptr = _host_weight(weight, n)
        fn = L.bsls_isotonic_regression if variant == 1 else L.bsls_isotonic_regression_3
        _lib.check(fn(y.ctypes.data, int(start), int(end), wptr, int(update)), fn.__name__)
        return
    _check_dev_vector(y)
    sub_w = None if weight is None else weight[:end]
    _pava_multi(y[:end], BlockPlan(np.array([start]), end, y.device), sub_w, update, variant=variant)


def isotonic_regression_c(y, start, end, weight=None, update=1):
    """Non-decreasing least-squares fit of ``y[start:end]``, in place (c_extensions.pyx:63-73).
    ``weight`` (int32, in/out) holds the pool size at every pool head."""
    return _pava_single(y, start, end, weight, update, 1)


def isotonic_regression_multi_c(y, blocks, weight=None, update=1, clip01=False):
    """Isotonic regression of every block, in place (c_extensions.pyx:76-89).  ``clip01`` is
    an extension that fuses the [0,1] clamp of the z-space projection (python/main.py:65)."""
    return _pava_multi(y, blocks, weight, update, clip01, 1)


def isotonic_regression_c_2(y, start, end):
    """Variant 2 of the reference (c_extensions.pyx:92-98): weight-free sweeps, bit-identical values."""
    return _pava_single(y, start, end, None, 1, 2)


def isotonic_regression_multi_c_2(y, blocks):
    return _pava_multi(y, blocks, None, 1, False, 2)


def isotonic_regression_c_3(y, start, end, weight=None, update=1):
    """Variant 3 of the reference (c_extensions.pyx:112-122): one pass with back-tracking; values and the
    weight array (pool sizes at the heads, tail markers ``w[k-1]``) are the reference's, bit for bit."""
    return _pava_single(y, start, end, weight, update, 3)


def isotonic_regression_multi_c_3(y, blocks, weight=None, update=1):
    return _pava_multi(y, blocks, weight, update, False, 3)


# ---------------------------------------------------------------------------------------------
# x <-> z (reference: c_extensions.pyx:195-248) and the bidiagonal N (bsls_utils.py:139-162)
# ---------------------------------------------------------------------------------------------
def _host_x2z_blocks(blocks, n):
    b = np.asarray(blocks.cpu() if torch.is_tensor(blocks) else blocks)
    # c_extensions.pyx:202-203
    assert False not in ((b[1:] - b[:-1]) > 0)
    assert b[0] == 0 and b[-1] < n
    return b


def _z_plan(x, blocks):
    n = x.shape[0]
    if isinstance(blocks, BlockPlan):
        assert blocks.first == 0 and blocks.n == n
        return blocks
    _host_x2z_blocks(blocks, n)
    return plan_for(blocks, n, x.device)


def x2z_c(x, z, blocks):
    """z <- per-block running sums of x without every block's last entry; returns z
    (c_extensions.pyx:195-220).  Device tensors, or NumPy arrays (copied through the GPU)."""
    if isinstance(x, np.ndarray):
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
        zd = torch.empty(z.shape[0], dtype=torch.float64, device=xd.device)
        x2z_c(xd, zd, np.asarray(blocks))
        z[:] = zd.cpu().numpy()
        return z
    _check_f64_vectors((x, "x"), (z, "z"))
    plan = _z_plan(x, blocks)
    assert z.shape[0] == x.shape[0] - plan.numblocks, "z must have n - numblocks entries"
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().bsls_dev_x2z_f64(plan.handle, x.data_ptr(), z.data_ptr(), _stream(x)), "x2z")
    return z


def z2x_c(x, z, blocks):
    """x <- adjacent differences of z, last entry of a block = 1 - z_last; returns x
    (c_extensions.pyx:223-248)."""
    if isinstance(x, np.ndarray):
        zd = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64)).cuda()
        xd = torch.empty(x.shape[0], dtype=torch.float64, device=zd.device)
        z2x_c(xd, zd, np.asarray(blocks))
        x[:] = xd.cpu().numpy()
        return x
    _check_f64_vectors((x, "x"), (z, "z"))
    plan = _z_plan(x, blocks)
    assert z.shape[0] == x.shape[0] - plan.numblocks, "z must have n - numblocks entries"
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().bsls_dev_z2x_f64(plan.handle, x.data_ptr(), z.data_ptr(), _stream(x)), "z2x")
    return x


def n_dot(x, z, blocks, add_x0=False):
    """x <- N z (+ x0): the change of variables of python/bsls_utils.py:139-162,327-328 as an
    operator (x_l = z_l - z_{l-1} inside a block)."""
    _check_f64_vectors((x, "x"), (z, "z"))
    plan = _z_plan(x, blocks)
    assert z.shape[0] == x.shape[0] - plan.numblocks
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().bsls_dev_nz_f64(plan.handle, x.data_ptr(), z.data_ptr(), int(bool(add_x0)), _stream(x)), "nz")
    return x


def nt_dot(zg, v, blocks):
    """zg <- N^T v ((N^T v)_l = v_l - v_{l+1})."""
    _check_f64_vectors((v, "v"), (zg, "zg"))
    plan = _z_plan(v, blocks)
    assert zg.shape[0] == v.shape[0] - plan.numblocks
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().bsls_dev_ntv_f64(plan.handle, zg.data_ptr(), v.data_ptr(), _stream(v)), "ntv")
    return zg


def block_scale(y, blocks, f, divide=False):
    """y[block k] *= f[k] (or /= f[k]), in place: the per-block (de)normalisation that
    get_solver_parts wraps around the projection when ``f`` is given (algorithm_utils.py:232-265)."""
    _check_f64_vectors((y, "y"))
    plan = plan_for(blocks, y.shape[0], y.device)
    assert torch.is_tensor(f) and f.is_cuda and f.dtype == torch.float64 and f.is_contiguous() and f.shape[0] == plan.numblocks
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().bsls_dev_block_scale_f64(plan.handle, y.data_ptr(), f.data_ptr(), int(bool(divide)), _stream(y)),
                   "block_scale")
    return None
