"""ctypes door onto libbsls_b200.so -- the only compute backend of this package.

There is deliberately no fallback: if the shared library has not been built, or a call
fails (no sm_100 device, CUDA error), a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsls_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC = 0, 1, 2, 3, 4

_lib = None


class BslsError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BslsError(
                "libbsls_b200.so is missing (%s). Build it with `python block-simplex-least-squares_b200/build.py` "
                "or `python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        c_int, c_void_p, c_i64 = ctypes.c_int, ctypes.c_void_p, ctypes.c_int64
        L.bsls_last_error.restype = ctypes.c_char_p
        L.bsls_version.restype = ctypes.c_char_p
        L.bsls_device_ok.restype = c_int
        for name in ("bsls_proj_multi_simplex", "bsls_proj_multi_ball"):
            getattr(L, name).argtypes = [c_void_p, c_void_p, c_int, c_int]
        L.bsls_proj_simplex.argtypes = [c_void_p, c_int, c_int]
        L.bsls_isotonic_regression.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int]
        L.bsls_isotonic_regression_3.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int]
        L.bsls_isotonic_regression_2.argtypes = [c_void_p, c_int, c_int]
        L.bsls_isotonic_regression_multi.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]
        L.bsls_isotonic_regression_multi_3.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]
        L.bsls_isotonic_regression_multi_2.argtypes = [c_void_p, c_void_p, c_int, c_int]
        for name in ("bsls_dev_isotonic_regression_multi_f64", "bsls_dev_isotonic_regression_multi_f32"):
            getattr(L, name).argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
        L.bsls_host_alloc.argtypes = [ctypes.POINTER(c_void_p), c_i64]
        L.bsls_host_free.argtypes = [c_void_p]
        L.bsls_plan_create.argtypes = [c_void_p, c_int, c_int, c_void_p, ctypes.POINTER(c_void_p)]
        L.bsls_plan_destroy.argtypes = [c_void_p]
        L.bsls_plan_info.argtypes = [c_void_p, ctypes.POINTER(c_i64 * 8)]
        for name in ("bsls_dev_proj_multi_simplex_f64", "bsls_dev_proj_multi_ball_f64",
                     "bsls_dev_proj_multi_simplex_f32", "bsls_dev_proj_multi_ball_f32"):
            getattr(L, name).argtypes = [c_void_p, c_void_p, c_void_p]
        _lib = L
    return _lib


def last_error():
    return lib().bsls_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    """Status -> exception.  BSLS_ERR_ARG mirrors the reference's Python `assert`s
    (python/c_extensions/c_extensions.pyx:24,33-34), so it raises AssertionError."""
    if rc == OK:
        return
    msg = "%s: %s" % (what, last_error()) if what else last_error()
    if rc == ERR_ARG:
        raise AssertionError(msg)
    raise BslsError("libbsls_b200 status %d: %s" % (rc, msg))
