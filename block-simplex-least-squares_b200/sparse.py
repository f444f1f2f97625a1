"""Sparse least-squares problem ``0.5 |A x - b|^2`` resident on one GPU (one shard of it
when the OD blocks are split over ranks), and the vector kernels the solver drivers use.

The reference keeps ``A`` and ``A.T`` as two scipy CSR matrices and evaluates the objective
with two ``csr_matvec`` calls (python/algorithm_utils.py:88-94,199-200).  Here the same pair
lives in HBM (int64 row pointers, int32 indices, float64 values -- or no values at all for a
0/1 incidence matrix) behind a ``bsls_lsq`` handle of libbsls_b200; every product, dot
product and update is a CUDA kernel of that library.  Nothing in this module computes on the
CPU, and nothing computes with torch operators either: torch only owns the memory.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

_F64 = torch.float64


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _dev_tensor(a, dtype, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device).contiguous()


def _check_vec(v, n, name):
    assert torch.is_tensor(v) and v.is_cuda and v.dtype == _F64 and v.dim() == 1, "%s: float64 CUDA vector expected" % name
    assert v.is_contiguous(), "%s must be contiguous" % name
    assert v.shape[0] == n, "%s has %d entries, expected %d" % (name, v.shape[0], n)


def nccl_library_path():
    """The NCCL that torch itself uses (nvidia-nccl wheel), else whatever the loader finds."""
    try:
        import nvidia.nccl
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            p = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:
        pass
    return None


class Communicator:
    """NCCL communicator of libbsls_b200 spanning the ranks of the default torch.distributed
    group (one process per GPU).  torch.distributed only carries the 128-byte NCCL id."""

    def __init__(self, device=None):
        import torch.distributed as dist
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        L = _lib.lib()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        path = nccl_library_path()
        cpath = path.encode() if path else None
        ident = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(L.bsls_comm_unique_id(cpath, ident), "comm_unique_id")
        box = [bytes(ident.raw)]
        dist.broadcast_object_list(box, src=0)
        ident = ctypes.create_string_buffer(box[0], 128)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.bsls_comm_create(cpath, self.world, self.rank, ident, ctypes.byref(self._handle)), "comm_create")

    @property
    def handle(self):
        return self._handle

    def enable_p2p(self, m):
        """Peer-memory exchange for link vectors of ``m`` entries (bsls_comm_p2p_alloc / _open): every rank allocates its
        region, torch.distributed carries the 64-byte CUDA IPC handles, every rank maps the regions of the others.  The
        sharded solver loop then reduces A x over NVLink loads / stores (csrc/p2p.cuh) instead of ncclAllReduce.
        Collective: every rank of the group must call it.  Returns True when the exchange is in place; on any failure
        (no peer access, more than 8 ranks, BSLS_P2P=0) the NCCL path stays and False is returned."""
        import torch.distributed as dist
        L = _lib.lib()
        if getattr(self, "_p2p_m", None) is not None:
            return self._p2p_m == int(m) and bool(L.bsls_comm_p2p_ready(self._handle))
        self._p2p_m = int(m)
        ok = os.environ.get("BSLS_P2P", "1") != "0" and 2 <= self.world <= 8
        handle = ctypes.create_string_buffer(64)
        if ok:
            with torch.cuda.device(self.device):
                ok = L.bsls_comm_p2p_alloc(self._handle, int(m), handle) == _lib.OK
        box = [None] * self.world
        dist.all_gather_object(box, bytes(handle.raw) if ok else None)
        if any(h is None for h in box):
            L.bsls_comm_p2p_disable(self._handle)
            return False
        blob = ctypes.create_string_buffer(b"".join(box), 64 * self.world)
        with torch.cuda.device(self.device):
            opened = L.bsls_comm_p2p_open(self._handle, blob) == _lib.OK
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(opened))
        if not all(flags):
            L.bsls_comm_p2p_disable(self._handle)
        return all(flags) and bool(L.bsls_comm_p2p_ready(self._handle))

    def allreduce_sum_(self, t):
        _check_vec(t, t.shape[0], "buffer")
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().bsls_comm_allreduce_sum_f64(self._handle, t.data_ptr(), t.shape[0], _stream(t.device)))
        return t

    def close(self):
        if self._handle is not None and self._handle.value:
            _lib.lib().bsls_comm_destroy(self._handle)
            self._handle = None


class Workspace:
    """Reduction workspace (``bsls_ws``): deterministic dot products / maxima on device vectors."""

    def __init__(self, device=None, handle=None, owner=None):
        L = _lib.lib()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._owner = owner  # keeps a parent LsqProblem alive when the handle is borrowed
        self._own = handle is None
        if handle is None:
            self._handle = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(L.bsls_ws_create(ctypes.byref(self._handle)), "ws_create")
        else:
            self._handle = ctypes.c_void_p(handle)
        self._scal_ptr = L.bsls_ws_scalar_ptr(self._handle)

    @property
    def handle(self):
        return self._handle

    def set_comm(self, comm):
        self._comm = comm
        _lib.check(_lib.lib().bsls_ws_set_comm(self._handle, comm.handle if comm is not None else None))

    def scalar_ptr(self, slot):
        """Device address of scalar slot ``slot`` (0..15) of this workspace."""
        return self._scal_ptr + 8 * int(slot)

    def scalars(self):
        out = (ctypes.c_double * 16)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_ws_scalars(self._handle, ctypes.byref(out), _stream(self.device)), "ws_scalars")
        return list(out)

    def dots(self, pairs, want_max=False):
        """[<x_k, y_k> for (x_k, y_k) in pairs] (at most four, one pass, summed over ranks);
        with ``want_max`` also max |x_0 - y_0| as the last entry.  Blocks."""
        assert 1 <= len(pairs) <= 4
        n = pairs[0][0].shape[0]
        xs, ys = (ctypes.c_void_p * 4)(), (ctypes.c_void_p * 4)()
        for k, (x, y) in enumerate(pairs):
            _check_vec(x, n, "x%d" % k)
            _check_vec(y, n, "y%d" % k)
            xs[k], ys[k] = x.data_ptr(), y.data_ptr()
        out = (ctypes.c_double * 5)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_ws_dots_f64(self._handle, len(pairs), ctypes.byref(xs), ctypes.byref(ys), n,
                                                   int(bool(want_max)), ctypes.byref(out), _stream(self.device)), "ws_dots")
        # NumPy scalars, as the reference's x.dot(y) returns: 0/0 and 1/0 give nan / inf, not exceptions
        res = [np.float64(out[k]) for k in range(len(pairs))]
        if want_max:
            res.append(np.float64(out[4]))
        return res

    def dot(self, x, y):
        return self.dots([(x, y)])[0]

    def norm(self, x):
        return float(np.sqrt(self.dot(x, x)))

    def max_abs_diff(self, x, y):
        return self.dots([(x, y)], want_max=True)[1]

    def flow_metrics(self, scaling, x_true, x_hat, thresh=1e-3):
        """[sum |s (xt - xh)|, sum s xt, #{xt - xh > thresh}, |xt - xh|^2, max s (xt - xh)] in one pass
        (the per-iterate reductions of LS_postprocess, python/main.py:112-134)."""
        n = x_true.shape[0]
        _check_vec(x_true, n, "x_true")
        _check_vec(x_hat, n, "x_hat")
        if scaling is not None:
            _check_vec(scaling, n, "scaling")
        out = (ctypes.c_double * 5)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_ws_flow_metrics_f64(self._handle, None if scaling is None else scaling.data_ptr(), x_true.data_ptr(),
                                                           x_hat.data_ptr(), n, float(thresh), ctypes.byref(out), _stream(self.device)),
                       "flow_metrics")
        return [float(v) for v in out]

    def axpy_dot(self, d, scale, c0, c1, v, w, out=None):
        """d <- d + c v with c = scale * ((*c0 or 1) - (*c1 or 0)), c0/c1 DEVICE scalar addresses
        or None; v None: d <- c d.  If ``w`` is given, the DEVICE double at address ``out``
        receives <w, d_new> (summed over ranks).  Asynchronous."""
        n = d.shape[0]
        _check_vec(d, n, "d")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_dev_axpy_dot_f64(self._handle, d.data_ptr(), float(scale), c0, c1,
                                                        None if v is None else v.data_ptr(),
                                                        None if w is None else w.data_ptr(), out, n,
                                                        _stream(self.device)), "axpy_dot")

    def md_update(self, plan, x_new, x, g, step, per_block_log=False):
        """x_new = x * exp(-t g) then every block divided by its sum; returns nothing, slot 10 of
        the scalars holds max |x_new - x| (python/BATCH.py:238-241, python/mirror_descent.py:39-47)."""
        n = x.shape[0]
        for name, t in (("x_new", x_new), ("x", x), ("g", g)):
            _check_vec(t, n, name)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_dev_md_update_f64(self._handle, plan.handle, x_new.data_ptr(), x.data_ptr(), g.data_ptr(),
                                                         float(step), int(bool(per_block_log)), _stream(self.device)), "md_update")

    def __del__(self):
        try:
            if self._own and self._handle is not None and self._handle.value:
                _lib.lib().bsls_ws_destroy(self._handle)
        except Exception:
            pass
        self._handle = None


_default_ws = {}


def default_workspace(device):
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _default_ws.get(key)
    if ws is None:
        ws = _default_ws[key] = Workspace(device)
    return ws


def axpby(out, a, x, b, y):
    """out = a*x + b*y with every product rounded on its own (np.add(x, -t*g, x_new));
    ``out`` may alias ``x`` or ``y``."""
    n = out.shape[0]
    for name, t in (("out", out), ("x", x), ("y", y)):
        _check_vec(t, n, name)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().bsls_dev_axpby_f64(out.data_ptr(), float(a), x.data_ptr(), float(b), y.data_ptr(), n,
                                                 _stream(out.device)), "axpby")
    return out


def copy_(dst, src):
    """np.copyto(dst, src) for device vectors (a device-to-device copy, no kernel)."""
    _check_vec(dst, src.shape[0], "dst")
    dst.copy_(src)
    return dst


class LsqProblem:
    """``A`` (m x n) and ``b`` on the GPU.

    ``A`` may be a scipy sparse matrix (converted to CSR and its transpose to CSR on the host,
    exactly as python/algorithm_utils.py:199-200 does, then uploaded) or a tuple of device/host
    arrays ``(a_ptr, a_idx, a_val, at_ptr, at_idx, at_val, (m, n))`` when the caller built both
    sides itself (``a_val`` / ``at_val`` None = implicit ones).
    """

    def __init__(self, A, b, device=None, implicit_ones=False, comm=None):
        L = _lib.lib()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dev = self.device
        if isinstance(A, tuple):
            a_ptr, a_idx, a_val, t_ptr, t_idx, t_val, shape = A
            self.m, self.n = int(shape[0]), int(shape[1])
        else:
            import scipy.sparse as sps
            Ac = sps.csr_matrix(A)
            if not Ac.has_sorted_indices:      # never reorder the caller's matrix in place
                Ac = Ac.copy()
                Ac.sort_indices()
            At = sps.csr_matrix(Ac.T)
            At.sort_indices()
            self.m, self.n = Ac.shape
            a_ptr, a_idx, a_val = Ac.indptr, Ac.indices, Ac.data
            t_ptr, t_idx, t_val = At.indptr, At.indices, At.data
            if implicit_ones:
                assert np.all(Ac.data == 1.0), "implicit_ones needs a 0/1 matrix"
                a_val = t_val = None
        self.a_ptr = _dev_tensor(a_ptr, torch.int64, dev)
        self.a_idx = _dev_tensor(a_idx, torch.int32, dev)
        self.a_val = None if a_val is None else _dev_tensor(a_val, _F64, dev)
        self.t_ptr = _dev_tensor(t_ptr, torch.int64, dev)
        self.t_idx = _dev_tensor(t_idx, torch.int32, dev)
        self.t_val = None if t_val is None else _dev_tensor(t_val, _F64, dev)
        assert self.a_ptr.shape[0] == self.m + 1 and self.t_ptr.shape[0] == self.n + 1
        self.nnz = int(self.a_idx.shape[0])
        assert self.t_idx.shape[0] == self.nnz
        self.b = _dev_tensor(b, _F64, dev).reshape(-1)
        assert self.b.shape[0] == self.m
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.bsls_lsq_create(self.m, self.n, self.nnz, self.a_ptr.data_ptr(), self.a_idx.data_ptr(),
                                         None if self.a_val is None else self.a_val.data_ptr(),
                                         self.t_ptr.data_ptr(), self.t_idx.data_ptr(),
                                         None if self.t_val is None else self.t_val.data_ptr(),
                                         self.b.data_ptr(), ctypes.byref(self._handle)), "lsq_create")
        self.ws = Workspace(dev, handle=L.bsls_lsq_ws(self._handle), owner=self)
        self.comm = None
        if comm is not None:
            self.set_comm(comm)

    @property
    def handle(self):
        return self._handle

    def set_comm(self, comm):
        """OD blocks sharded over ranks: A x is summed over the ranks of ``comm``."""
        self.comm = comm
        _lib.check(_lib.lib().bsls_lsq_set_comm(self._handle, None if comm is None else comm.handle))
        if comm is not None and comm.world > 1:
            comm.enable_p2p(self.m)      # collective; falls back to NCCL when peer memory cannot be mapped

    def set_b(self, b):
        self.b = _dev_tensor(b, _F64, self.device).reshape(-1)
        assert self.b.shape[0] == self.m
        _lib.check(_lib.lib().bsls_lsq_set_b(self._handle, self.b.data_ptr()))

    def with_b(self, b):
        """A second problem handle over the SAME matrix arrays (no copy) with its own right-hand side, residual and
        workspace: lets a driver change ``b`` (the z-space target of main.py:48) without touching the caller's problem."""
        other = LsqProblem((self.a_ptr, self.a_idx, self.a_val, self.t_ptr, self.t_idx, self.t_val, (self.m, self.n)), b,
                           device=self.device, comm=self.comm)
        if getattr(self, "panels", 1) > 1:
            other.p_ptr, other.p_idx, other.p_val, other.panels = self.p_ptr, self.p_idx, self.p_val, self.panels
            _lib.check(_lib.lib().bsls_lsq_set_panels(other._handle, self.panels, self.p_ptr.data_ptr(), self.p_idx.data_ptr(),
                                                      None if self.p_val is None else self.p_val.data_ptr()), "lsq_set_panels")
        return other

    def set_panels(self, panel_cols=None, l2_budget_bytes=48 << 20):
        """Build the column-panelled copy of A (``bsls_lsq_set_panels``) so that the slice of x a
        panel gathers from stays L2-resident.  ``panel_cols`` columns per panel (default: as many
        as fit ``l2_budget_bytes``); a single panel removes the copy.  Set-up work, done once per
        matrix with torch's sort on the device; the per-iteration products use only library kernels."""
        L = _lib.lib()
        if panel_cols is None:
            # as few panels as fit the budget, all of the same width (a short last panel has short row pieces, which cost
            # more per entry): 160 MB of x -> 4 x 40 MB rather than 3 x 48 + 16
            budget_cols = max(1, l2_budget_bytes // 8)
            P = (self.n + budget_cols - 1) // budget_cols
            panel_cols = (self.n + P - 1) // P if P > 0 else self.n
        P = (self.n + panel_cols - 1) // panel_cols
        if P <= 1:
            _lib.check(L.bsls_lsq_set_panels(self._handle, 0, None, None, None))
            self.p_ptr = self.p_idx = self.p_val = None
            self.panels = 1
            return 1
        dev = self.device
        ptrs, idxs, vals = [], [], []
        base = 0
        for p in range(P):
            lo, hi = p * panel_cols, min(self.n, (p + 1) * panel_cols)
            n0, n1 = int(self.t_ptr[lo]), int(self.t_ptr[hi])
            rows = self.t_idx[n0:n1]
            counts = (self.t_ptr[lo + 1:hi + 1] - self.t_ptr[lo:hi])
            cols = torch.repeat_interleave(torch.arange(lo, hi, device=dev, dtype=torch.int32), counts)
            order = torch.argsort(rows.to(torch.int64), stable=True)
            idxs.append(cols[order])
            if self.t_val is not None:
                vals.append(self.t_val[n0:n1][order])
            rc = torch.bincount(rows, minlength=self.m)
            ptrs.append(base + torch.cumsum(rc, 0) - rc)          # starts of the m rows of this panel
            base += n1 - n0
            del rows, counts, cols, order, rc
        ptrs.append(torch.tensor([base], dtype=torch.int64, device=dev))
        self.p_ptr = torch.cat(ptrs).to(torch.int64).contiguous()
        self.p_idx = torch.cat(idxs).contiguous()
        self.p_val = torch.cat(vals).contiguous() if vals else None
        assert self.p_ptr.shape[0] == P * self.m + 1 and base == self.nnz
        _lib.check(L.bsls_lsq_set_panels(self._handle, P, self.p_ptr.data_ptr(), self.p_idx.data_ptr(),
                                         None if self.p_val is None else self.p_val.data_ptr()), "lsq_set_panels")
        self.panels = P
        return P

    def set_modes(self, a_mode=0, at_mode=0):
        _lib.check(_lib.lib().bsls_lsq_set_modes(self._handle, int(a_mode), int(at_mode)), "lsq_set_modes")

    # -- sparse_least_squares_obj (python/algorithm_utils.py:88-94) ----------------------------
    def obj(self, x, g):
        """g <- A^T (A x - b) in place; returns f = 0.5 |A x - b|^2 as a Python float."""
        _check_vec(x, self.n, "x")
        _check_vec(g, self.n, "g")
        f = ctypes.c_double()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_lsq_obj_f64(self._handle, x.data_ptr(), g.data_ptr(), ctypes.byref(f),
                                                   _stream(self.device)), "lsq_obj")
        return f.value

    def value(self, x):
        """f = 0.5 |A x - b|^2 only (one product)."""
        _check_vec(x, self.n, "x")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_dev_lsq_residual_f64(self._handle, x.data_ptr(), _stream(self.device)), "lsq_residual")
        return self.ws.scalars()[0]

    def residual(self):
        """The m-vector r = A x - b of the last evaluation (a view of library memory)."""
        ptr = _lib.lib().bsls_lsq_residual_ptr(self._handle)
        return _wrap_device_f64(ptr, self.m, self.device, self)

    def matvec(self, v, out=None):
        """out = A v (m entries; summed over ranks when sharded)."""
        _check_vec(v, self.n, "v")
        if out is None:
            out = torch.empty(self.m, dtype=_F64, device=self.device)
        _check_vec(out, self.m, "out")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_dev_lsq_matvec_f64(self._handle, v.data_ptr(), out.data_ptr(), _stream(self.device)),
                       "lsq_matvec")
        return out

    def rmatvec(self, w, out=None):
        """out = A^T w (n entries)."""
        _check_vec(w, self.m, "w")
        if out is None:
            out = torch.empty(self.n, dtype=_F64, device=self.device)
        _check_vec(out, self.n, "out")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bsls_dev_lsq_rmatvec_f64(self._handle, w.data_ptr(), out.data_ptr(), _stream(self.device)),
                       "lsq_rmatvec")
        return out

    def __del__(self):
        try:
            if self._handle is not None and self._handle.value:
                _lib.lib().bsls_lsq_destroy(self._handle)
        except Exception:
            pass
        self._handle = None


class _DevMem:
    """__cuda_array_interface__ shim for memory owned by the library."""

    def __init__(self, ptr, n, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def _wrap_device_f64(ptr, n, device, owner):
    with torch.cuda.device(device):
        t = torch.as_tensor(_DevMem(ptr, n, owner), device=device)
    t._bsls_owner = owner
    return t
