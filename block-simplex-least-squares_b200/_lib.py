"""ctypes door onto libbsls_b200.so -- the only compute backend of this package.

There is deliberately no fallback: if the shared library has not been built, or a call
fails (no sm_100 device, CUDA error), a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsls_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC, ERR_UNSUPPORTED = 0, 1, 2, 3, 4, 5

_lib = None


class BatchOpts(ctypes.Structure):  # bsls_batch_opts (include/bsls_b200.h)
    _fields_ = [("method", ctypes.c_int), ("proj_mode", ctypes.c_int), ("use_line_search", ctypes.c_int),
                ("has_f_min", ctypes.c_int), ("f_min", ctypes.c_double), ("opt_tol", ctypes.c_double),
                ("prog_tol", ctypes.c_double), ("min_eig", ctypes.c_double), ("max_iter", ctypes.c_int),
                ("corrections", ctypes.c_int)]


class BatchResult(ctypes.Structure):  # bsls_batch_result
    _fields_ = [("f", ctypes.c_double), ("iterations", ctypes.c_int), ("stop_code", ctypes.c_int),
                ("stop_value", ctypes.c_double), ("obj_evals", ctypes.c_int), ("backtracks", ctypes.c_int),
                ("kernel_launches", ctypes.c_int), ("device_ms", ctypes.c_double)]


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
        L.bsls_dev_isotonic_regression_multi_2_f64.argtypes = [c_void_p, c_void_p, c_void_p]
        L.bsls_dev_isotonic_regression_multi_3_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p]
        L.bsls_host_alloc.argtypes = [ctypes.POINTER(c_void_p), c_i64]
        L.bsls_host_free.argtypes = [c_void_p]
        L.bsls_plan_create.argtypes = [c_void_p, c_int, c_int, c_void_p, ctypes.POINTER(c_void_p)]
        L.bsls_plan_destroy.argtypes = [c_void_p]
        L.bsls_plan_info.argtypes = [c_void_p, ctypes.POINTER(c_i64 * 8)]
        for name in ("bsls_dev_proj_multi_simplex_f64", "bsls_dev_proj_multi_ball_f64",
                     "bsls_dev_proj_multi_simplex_f32", "bsls_dev_proj_multi_ball_f32"):
            getattr(L, name).argtypes = [c_void_p, c_void_p, c_void_p]
        c_dbl, c_char_p = ctypes.c_double, ctypes.c_char_p
        PP = ctypes.POINTER(c_void_p)
        for name in ("bsls_dev_x2z_f64", "bsls_dev_z2x_f64", "bsls_dev_ntv_f64"):
            getattr(L, name).argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
        L.bsls_dev_nz_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p]
        L.bsls_dev_block_scale_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p]
        L.bsls_comm_unique_id.argtypes = [c_char_p, c_void_p]
        L.bsls_comm_create.argtypes = [c_char_p, c_int, c_int, c_void_p, PP]
        L.bsls_comm_destroy.argtypes = [c_void_p]
        L.bsls_comm_allreduce_sum_f64.argtypes = [c_void_p, c_void_p, c_i64, c_void_p]
        L.bsls_comm_p2p_alloc.argtypes = [c_void_p, c_i64, c_void_p]
        L.bsls_comm_p2p_open.argtypes = [c_void_p, c_void_p]
        L.bsls_comm_p2p_ready.argtypes = [c_void_p]
        L.bsls_comm_p2p_disable.argtypes = [c_void_p]
        L.bsls_ws_create.argtypes = [PP]
        L.bsls_ws_destroy.argtypes = [c_void_p]
        L.bsls_ws_set_comm.argtypes = [c_void_p, c_void_p]
        L.bsls_ws_scalar_ptr.argtypes = [c_void_p]
        L.bsls_ws_scalar_ptr.restype = c_void_p
        L.bsls_ws_scalars.argtypes = [c_void_p, ctypes.POINTER(c_dbl * 16), c_void_p]
        L.bsls_lsq_create.argtypes = [c_i64, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, PP]
        L.bsls_lsq_destroy.argtypes = [c_void_p]
        L.bsls_lsq_set_comm.argtypes = [c_void_p, c_void_p]
        L.bsls_lsq_ws.argtypes = [c_void_p]
        L.bsls_lsq_ws.restype = c_void_p
        L.bsls_lsq_set_b.argtypes = [c_void_p, c_void_p]
        L.bsls_lsq_set_modes.argtypes = [c_void_p, c_int, c_int]
        L.bsls_lsq_set_panels.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        L.bsls_lsq_obj_f64.argtypes = [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_dbl), c_void_p]
        L.bsls_dev_lsq_residual_f64.argtypes = [c_void_p, c_void_p, c_void_p]
        L.bsls_dev_lsq_gradient_f64.argtypes = [c_void_p, c_void_p, c_void_p]
        L.bsls_dev_lsq_gradient_bb_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.bsls_dev_lsq_matvec_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
        L.bsls_dev_lsq_rmatvec_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
        L.bsls_lsq_scalars.argtypes = [c_void_p, ctypes.POINTER(c_dbl * 16), c_void_p]
        L.bsls_lsq_residual_ptr.argtypes = [c_void_p]
        L.bsls_lsq_residual_ptr.restype = c_void_p
        L.bsls_lsq_scalar_ptr.argtypes = [c_void_p]
        L.bsls_lsq_scalar_ptr.restype = c_void_p
        L.bsls_dev_axpby_f64.argtypes = [c_void_p, c_dbl, c_void_p, c_dbl, c_void_p, c_i64, c_void_p]
        L.bsls_ws_dots_f64.argtypes = [c_void_p, c_int, ctypes.POINTER(c_void_p * 4), ctypes.POINTER(c_void_p * 4), c_i64, c_int,
                                       ctypes.POINTER(c_dbl * 5), c_void_p]
        L.bsls_ws_flow_metrics_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_dbl, ctypes.POINTER(c_dbl * 5), c_void_p]
        L.bsls_dev_axpy_dot_f64.argtypes = [c_void_p, c_void_p, c_dbl, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64,
                                            c_void_p]
        L.bsls_dev_md_update_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_dbl, c_int, c_void_p]
        L.bsls_batch_solve_f64.argtypes = [c_void_p, c_void_p, c_void_p, ctypes.POINTER(BatchOpts), ctypes.POINTER(BatchResult),
                                           c_void_p, c_void_p, c_int, c_void_p]
        L.bsls_md_least_squares_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_dbl, c_dbl, ctypes.POINTER(BatchResult), c_void_p]
        L.bsls_zbb_run_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_dbl,
                                       ctypes.POINTER(BatchResult), c_void_p]
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
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise BslsError("libbsls_b200 status %d: %s" % (rc, msg))
