"""ctypes doors onto the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

``port``  = oracle/liboracle.so, the C restatement (oracle/bsls_oracle.c).
``ref``   = oracle/_ref/libbsls_ref.so, the reference's own unmodified C++ headers
            (python/c_extensions/proj_simplex.h, isotonic_regression.h) behind the
            C ABI of oracle/ref_shim.cpp; None when it has not been built.

All functions work in place on C-contiguous float64 / int32 NumPy arrays and mirror the
argument meaning of the reference's Cython layer (python/c_extensions/c_extensions.pyx).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_D = ctypes.POINTER(ctypes.c_double)
_I = ctypes.POINTER(ctypes.c_int32)
_L = ctypes.POINTER(ctypes.c_int64)
_i64 = ctypes.c_int64
_int = ctypes.c_int


def build(force=False):
    """Compile liboracle.so (always) and _ref/libbsls_ref.so (when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bsls_oracle.c")
    stale = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    ref_missing = not os.path.exists(os.path.join(_HERE, "_ref", "libbsls_ref.so")) and os.path.exists(
        "/root/reference/python/c_extensions/proj_simplex.h")
    if stale or ref_missing:
        if stale and os.path.exists(so):
            os.remove(so)
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def _dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_D)


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_I)


def _lp(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(_L)


def _starts32(starts):
    return np.ascontiguousarray(starts, dtype=np.int32)


class Port:
    """The C restatement."""

    kind = "port"

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        L = self.lib
        L.orc_proj_simplex.argtypes = [_D, _i64, _i64]
        L.orc_proj_multi_simplex.argtypes = [_D, _I, _i64, _i64]
        L.orc_proj_multi_ball.argtypes = [_D, _I, _i64, _i64]
        L.orc_proj_multi_simplex_mt.argtypes = [_D, _I, _i64, _i64, _int]
        L.orc_max_threads.restype = _int
        L.orc_pava.argtypes = [_D, _i64, _i64, _I, _int]
        L.orc_pava2.argtypes = [_D, _i64, _i64]
        L.orc_pava3.argtypes = [_D, _i64, _i64, _I, _int]
        L.orc_pava_multi.argtypes = [_D, _I, _i64, _i64, _I, _int]
        L.orc_pava_multi2.argtypes = [_D, _I, _i64, _i64]
        L.orc_pava_multi3.argtypes = [_D, _I, _i64, _i64, _I, _int]
        L.orc_pava_multi_mt.argtypes = [_D, _I, _i64, _i64, _I, _int, _int]
        L.orc_clip01.argtypes = [_D, _i64]
        L.orc_x2z.argtypes = [_D, _D, _I, _i64, _i64]
        L.orc_z2x.argtypes = [_D, _D, _I, _i64, _i64]
        L.orc_csr_matvec.argtypes = [_i64, _L, _I, _D, _D, _D]
        L.orc_lsq_obj.argtypes = [_i64, _i64, _L, _I, _D, _L, _I, _D, _D, _D, _D, _D]
        L.orc_lsq_obj.restype = ctypes.c_double

    # -- projection ---------------------------------------------------------------
    def proj_simplex(self, y, start, end):
        self.lib.orc_proj_simplex(_dp(y), start, end)

    def proj_multi_simplex(self, y, starts, threads=0):
        s = _starts32(starts)
        if threads:
            self.lib.orc_proj_multi_simplex_mt(_dp(y), _ip(s), len(s), len(y), threads)
        else:
            self.lib.orc_proj_multi_simplex(_dp(y), _ip(s), len(s), len(y))

    def proj_multi_ball(self, y, starts):
        s = _starts32(starts)
        self.lib.orc_proj_multi_ball(_dp(y), _ip(s), len(s), len(y))

    def max_threads(self):
        return int(self.lib.orc_max_threads())

    # -- isotonic regression ---------------------------------------------------------
    def pava(self, y, start, end, weight=None, update=1, variant=1):
        if variant == 2:
            self.lib.orc_pava2(_dp(y), start, end)
            return None
        w = np.ones(len(y), dtype=np.int32) if weight is None else weight
        (self.lib.orc_pava if variant == 1 else self.lib.orc_pava3)(_dp(y), start, end, _ip(w), update)
        return w

    def pava_multi(self, y, starts, weight=None, update=1, variant=1, threads=0):
        s = _starts32(starts)
        if variant == 2:
            self.lib.orc_pava_multi2(_dp(y), _ip(s), len(s), len(y))
            return None
        w = np.ones(len(y), dtype=np.int32) if weight is None else weight
        if threads and variant == 1:
            self.lib.orc_pava_multi_mt(_dp(y), _ip(s), len(s), len(y), _ip(w), update, threads)
        else:
            (self.lib.orc_pava_multi if variant == 1 else self.lib.orc_pava_multi3)(
                _dp(y), _ip(s), len(s), len(y), _ip(w), update)
        return w

    def clip01(self, y):
        self.lib.orc_clip01(_dp(y), len(y))

    # -- change of variables ------------------------------------------------------------
    def x2z(self, x, z, starts):
        s = _starts32(starts)
        self.lib.orc_x2z(_dp(x), _dp(z), _ip(s), len(s), len(x))
        return z

    def z2x(self, x, z, starts):
        s = _starts32(starts)
        self.lib.orc_z2x(_dp(x), _dp(z), _ip(s), len(s), len(x))
        return x

    # -- sparse objective --------------------------------------------------------------------
    def csr_matvec(self, ptr, idx, val, v, rows):
        out = np.empty(rows)
        self.lib.orc_csr_matvec(rows, _lp(ptr), _ip(idx), _dp(val), _dp(v), _dp(out))
        return out

    def csr_matvec_mt(self, ptr, idx, val, v, out, threads):
        """out = M v with the rows cut into ``threads`` contiguous ranges, one host thread each (the C routine runs
        without the GIL); every row is still summed left to right, so the result equals csr_matvec's bit for bit."""
        import threading
        rows = len(ptr) - 1
        if threads <= 1:
            self.lib.orc_csr_matvec(rows, _lp(ptr), _ip(idx), _dp(val), _dp(v), _dp(out))
            return out
        nnz_cuts = np.linspace(0, int(ptr[-1]), threads + 1)
        cuts = np.searchsorted(ptr, nnz_cuts, side="left").astype(np.int64)
        cuts[0], cuts[-1] = 0, rows

        def work(k):
            lo, hi = int(cuts[k]), int(cuts[k + 1])
            if hi > lo:
                self.lib.orc_csr_matvec(hi - lo, _lp(ptr[lo:hi + 1]), _ip(idx), _dp(val), _dp(v), _dp(out[lo:hi]))
        ths = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return out

    def lsq_obj(self, A_csr, AT_csr, x, b, g):
        """f = 0.5 |Ax-b|^2 and g <- A^T(Ax-b); A_csr/AT_csr are (ptr int64, idx int32, val f64)."""
        m = len(A_csr[0]) - 1
        n = len(AT_csr[0]) - 1
        res = np.empty(m)
        f = self.lib.orc_lsq_obj(m, n, _lp(A_csr[0]), _ip(A_csr[1]), _dp(A_csr[2]),
                                 _lp(AT_csr[0]), _ip(AT_csr[1]), _dp(AT_csr[2]),
                                 _dp(x), _dp(b), _dp(res), _dp(g))
        return float(f), res


class Ref:
    """The reference's own headers (oracle/_ref).  Block sizes are limited by the
    reference's stack VLA (proj_simplex.h:21): keep blocks below ~10^5 elements."""

    kind = "reference"

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.ref_proj_simplex.argtypes = [_D, _int, _int]
        L.ref_proj_multi_simplex.argtypes = [_D, _I, _int, _int]
        L.ref_proj_multi_ball.argtypes = [_D, _I, _int, _int]
        L.ref_isotonic_regression.argtypes = [_D, _int, _int, _I, _int]
        L.ref_isotonic_regression_multi.argtypes = [_D, _I, _int, _int, _I, _int]
        L.ref_isotonic_regression_2.argtypes = [_D, _int, _int]
        L.ref_isotonic_regression_multi_2.argtypes = [_D, _I, _int, _int]
        L.ref_isotonic_regression_3.argtypes = [_D, _int, _int, _I, _int]
        L.ref_isotonic_regression_multi_3.argtypes = [_D, _I, _int, _int, _I, _int]

    def proj_simplex(self, y, start, end):
        self.lib.ref_proj_simplex(_dp(y), start, end)

    def proj_multi_simplex(self, y, starts, threads=0):
        s = _starts32(starts)
        self.lib.ref_proj_multi_simplex(_dp(y), _ip(s), len(s), len(y))

    def proj_multi_ball(self, y, starts):
        s = _starts32(starts)
        self.lib.ref_proj_multi_ball(_dp(y), _ip(s), len(s), len(y))

    def pava(self, y, start, end, weight=None, update=1, variant=1):
        if variant == 2:
            self.lib.ref_isotonic_regression_2(_dp(y), start, end)
            return None
        w = np.ones(len(y), dtype=np.int32) if weight is None else weight
        fn = self.lib.ref_isotonic_regression if variant == 1 else self.lib.ref_isotonic_regression_3
        fn(_dp(y), start, end, _ip(w), update)
        return w

    def pava_multi(self, y, starts, weight=None, update=1, variant=1, threads=0):
        s = _starts32(starts)
        if variant == 2:
            self.lib.ref_isotonic_regression_multi_2(_dp(y), _ip(s), len(s), len(y))
            return None
        w = np.ones(len(y), dtype=np.int32) if weight is None else weight
        fn = self.lib.ref_isotonic_regression_multi if variant == 1 else self.lib.ref_isotonic_regression_multi_3
        fn(_dp(y), _ip(s), len(s), len(y), _ip(w), update)
        return w


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref():
    """The compiled reference, or None when oracle/_ref has not been built."""
    global _ref
    if _ref is None:
        build()
        p = os.path.join(_HERE, "_ref", "libbsls_ref.so")
        if os.path.exists(p):
            _ref = Ref(p)
    return _ref


def pools_from_weights(w, starts, n):
    """Canonical pool structure: walk heads i += w[i] inside every block
    and return the list of head indices.  Interior entries of w are stale and ignored."""
    starts = np.asarray(starts)
    ends = np.append(starts[1:], n)
    heads = []
    for s, e in zip(starts, ends):
        i = int(s)
        while i < e:
            heads.append(i)
            i += int(w[i])
    return np.array(heads, dtype=np.int64)
