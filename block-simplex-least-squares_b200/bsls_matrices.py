"""Problem construction on the GPU -- drop-in for the reference's ``python/bsls_matrices.py`` (SURVEY.md section 8f,
rank 1): the step immediately BEFORE the hot loop.  Raw ``(A, b, T, d, U, f, V, g, x_true)`` as stored in a ``.mat``
file become the block-contiguous, simplex-scaled least-squares problem the solvers take:

    consolidate             stack the measurement matrices, pick the equality constraint C x = d   (:88-107)
    standard_simplex_form   scaling = C^T d; drop zero blocks; x_split = x / scaling; AA <- AA[:, nz] diag(scaling)   (:136-160)
    cleanup                 drop all-zero rows of AA and C                                         (:128-134)
    blockify                permute the columns so that every block of C is contiguous              (:109-126)
    degree_reduced_form     + N (bidiagonal, never materialised here) and the feasible start x0     (:69-86)
    reconstruct             un-order, rescale, un-zero                                             (:170-182)

File reading (``scipy.io.loadmat``) and the reference's input asserts stay on the host; every array is uploaded once and
all the transformations above run on the device: the O(nnz) work is a sort / scan / gather over the CSR triplets done
with torch's device primitives (one-off set-up work, like the synthetic generators of ``generate.py``), the result is
handed to :class:`sparse.LsqProblem` as device CSR arrays of AA and AA^T -- no NumPy / scipy preprocessing in between.
Within a block the columns keep their original order (the reference's ``np.argsort`` of the block index is not stable,
so its order inside a block is unspecified; ``reconstruct`` undoes either).
"""
import logging

import numpy as np
import torch

from .bsls_utils import NOperator, particular_x0
from .sparse import LsqProblem

__all__ = ["BSLSMatrices", "DevCSR"]

_F64 = torch.float64


class DevCSR:
    """A CSR matrix in device memory: ``ptr`` int64 (rows + 1), ``idx`` int32, ``val`` float64, ``shape``."""

    def __init__(self, ptr, idx, val, shape):
        self.ptr, self.idx, self.val, self.shape = ptr, idx, val, (int(shape[0]), int(shape[1]))

    @classmethod
    def from_host(cls, M, device):
        import scipy.sparse as sps
        M = sps.csr_matrix(M)
        if not M.has_sorted_indices:
            M = M.copy()
            M.sort_indices()
        return cls(torch.as_tensor(M.indptr.astype(np.int64)).to(device), torch.as_tensor(M.indices.astype(np.int32)).to(device),
                   torch.as_tensor(M.data.astype(np.float64)).to(device), M.shape)

    @property
    def device(self):
        return self.ptr.device

    @property
    def nnz(self):
        return int(self.idx.shape[0])

    def rows_of_entries(self):
        counts = self.ptr[1:] - self.ptr[:-1]
        return torch.repeat_interleave(torch.arange(self.shape[0], device=self.device, dtype=torch.int64), counts)

    @staticmethod
    def from_triplets(rows, cols, vals, shape):
        """Entries in any order -> CSR with ascending column ids inside every row (one device sort by row * n + col)."""
        m, n = int(shape[0]), int(shape[1])
        key = rows.to(torch.int64) * n + cols.to(torch.int64)
        key, order = torch.sort(key, stable=True)
        r = torch.div(key, n, rounding_mode="floor")
        ptr = torch.zeros(m + 1, dtype=torch.int64, device=rows.device)
        if r.numel():
            torch.cumsum(torch.bincount(r, minlength=m), 0, out=ptr[1:])
        return DevCSR(ptr, (key - r * n).to(torch.int32), vals[order].contiguous(), (m, n))

    def transpose(self):
        return DevCSR.from_triplets(self.idx.to(torch.int64), self.rows_of_entries(), self.val, (self.shape[1], self.shape[0]))

    def select_rows(self, keep_rows):
        """Rows ``keep_rows`` (ascending int64 indices), in that order."""
        counts = (self.ptr[1:] - self.ptr[:-1])[keep_rows]
        ptr = torch.zeros(keep_rows.numel() + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=ptr[1:])
        mask = torch.zeros(self.shape[0], dtype=torch.bool, device=self.device)
        mask[keep_rows] = True
        emask = mask[self.rows_of_entries()]
        return DevCSR(ptr, self.idx[emask].contiguous(), self.val[emask].contiguous(), (keep_rows.numel(), self.shape[1]))

    def map_columns(self, newcol, n_new, scale=None):
        """Column j -> newcol[j] (int64; -1 drops the column), entries multiplied by scale[j]: M[:, sel] diag(scale) and a
        column permutation in one pass; rows come out with ascending new column ids."""
        j = self.idx.to(torch.int64)
        nc = newcol[j]
        keep = nc >= 0
        vals = self.val if scale is None else self.val * scale[j]
        return DevCSR.from_triplets(self.rows_of_entries()[keep], nc[keep], vals[keep], (self.shape[0], n_new))

    def row_sums(self):
        out = torch.zeros(self.shape[0], dtype=_F64, device=self.device)
        out.index_add_(0, self.rows_of_entries(), self.val)
        return out

    def todense(self):
        out = torch.zeros(self.shape, dtype=_F64, device=self.device)
        out[self.rows_of_entries(), self.idx.to(torch.int64)] = self.val
        return out


def _vstack(X, x, Y, y):
    """stackMV (bsls_utils.py:438-454) on device matrices."""
    if X is None:
        return Y, y
    if Y is None:
        return X, x
    assert X.shape[1] == Y.shape[1]
    ptr = torch.cat((X.ptr, Y.ptr[1:] + X.ptr[-1]))
    return DevCSR(ptr, torch.cat((X.idx, Y.idx)), torch.cat((X.val, Y.val)), (X.shape[0] + Y.shape[0], X.shape[1])), torch.cat((x, y))


def _arr(x):
    return np.atleast_1d(np.squeeze(np.array(x)))


class BSLSMatrices:
    """bsls_matrices.py:15-323 with every matrix and vector resident on the GPU.  Same constructor arguments, methods and
    ``get_LS()`` tuple; ``AA`` is a :class:`DevCSR`, vectors are device tensors, ``N`` an operator
    (:class:`bsls_utils.NOperator`).  ``problem()`` wraps ``(AA, bb)`` as an :class:`LsqProblem` for the solvers."""

    def __init__(self, data=None, fname=None, full=False, L=True, OD=False, CP=False, LP=False, eq=None, init=False, thresh=1e-5,
                 noisy=False, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.eq = eq
        if data is None and fname is not None:
            import scipy.io as sio
            logging.debug('Loading %s...' % fname)
            data = sio.loadmat(fname)
        raw = self.load_raw(data, full=full, L=L, OD=OD, CP=CP, LP=LP, thresh=thresh, noisy=noisy)
        self.rA, self.b, self.rx_true, self.rT, self.d, self.rU, self.f, self.rV, self.g, self.nz, self.info = raw
        self.A, self.T, self.U, self.V = self.rA, self.rT, self.rU, self.rV
        self.x_true, self.x_split = self.rx_true, self.rx_true
        self.block_sizes, self.rsort_index, self.scaling = None, None, None
        self.N, self.x0 = None, None
        self.n_raw = int(self.rx_true.shape[0])

    # -- loading: host parsing + the reference's asserts, then one upload (bsls_matrices.py:208-274) -----------------
    def load_raw(self, data, full=False, L=True, OD=False, CP=False, LP=False, thresh=1e-5, noisy=False, info=None):
        import scipy.sparse as sps
        if info is None:
            info = {}
        sparse = lambda M: None if M is None else sps.csr_matrix(M)
        dev = self.device
        up = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(-1)).to(dev)
        A, b, nz = None, None, None
        if L and full and 'A_full' in data and 'b_full' in data:
            A, b = sparse(data['A_full']), _arr(data['b_full'])
            if len(data['A'].shape) == 1:
                A = A.T
        elif L and 'A' in data and 'b' in data:
            A, b = sparse(data['A']), _arr(data['b'])
            if len(data['A'].shape) == 1:
                A = A.T
        elif 'phi' in data and 'b' in data:
            A, b = sparse(data['phi']), _arr(data['b'])
        if 'b_full' in data:
            info['nAllLinks'] = _arr(data['b_full']).size
        if b is not None:
            info['nLinks'] = b.size
        if 'x_true' in data:
            x_true = _arr(data['x_true'])
        elif 'real_a' in data:
            x_true = _arr(data['real_a'])
        else:
            raise NotImplementedError("data holds neither x_true nor real_a")
        if A is not None:  # remove rows of zeros (unused sensors)
            counts = np.diff(sps.csr_matrix(A).indptr)
            nz = [int(i) for i in np.nonzero(counts == 0)[0]]
            keep = np.nonzero(counts > 0)[0]
            A, b = sps.csr_matrix(A)[keep, :], b[keep]
            if not noisy:
                err = np.linalg.norm(A.dot(x_true) - b)
                assert err < thresh, 'Check data input: Ax != b, norm: %s' % err
        n = x_true.shape[0]
        T = d = U = f = V = g = None
        if OD and 'T' in data and 'd' in data and data['T'] is not None and np.size(data['T']) > 0:
            T, d = sparse(data['T']), _arr(data['d'])
            assert T.shape[1] == n and np.all(np.asarray((T > 0).sum(axis=0)).ravel() <= 1)   # partial simplex incidence
            info['nOD'] = d.size
        if CP and 'U' in data and 'f' in data and data['U'] is not None and np.size(data['U']) > 0:
            U, f = sparse(data['U']), _arr(data['f'])
            assert U.shape[1] == n and np.all(np.asarray((U > 0).sum(axis=0)).ravel() == 1)   # simplex incidence
            info['nCP'] = f.size
        if LP and 'V' in data and 'g' in data and data['V'] is not None and np.size(data['V']) > 0:
            V, g = sparse(data['V']), _arr(data['g'])
            info['nLP'] = g.size
        dm = lambda M: None if M is None else DevCSR.from_host(M, dev)
        dv = lambda v: None if v is None else up(v)
        return dm(A), dv(b), up(x_true), dm(T), dv(d), dm(U), dv(f), dm(V), dv(g), nz, info

    # -- the transformations ---------------------------------------------------------------------------------
    def consolidate(self, eq=None):
        """bsls_matrices.py:88-107"""
        AA, bb = _vstack(self.A, self.b, self.V, self.g)
        if eq == 'OD':
            self.AA, self.bb = _vstack(AA, bb, self.U, self.f)
            self.C, self.d = self.T, self.d
        elif eq == 'CP':
            self.AA, self.bb = _vstack(AA, bb, self.T, self.d)
            self.C, self.d = self.U, self.f
        else:
            AA, bb = _vstack(AA, bb, self.T, self.d)
            self.AA, self.bb = _vstack(AA, bb, self.U, self.f)
            self.C, self.d = None, None

    def standard_simplex_form(self, thresh=1e-30, noisy=False):
        """AA x_true = bb -> AA' x_split = bb ;  C x_true = d -> C x_split = 1   (bsls_matrices.py:136-160)"""
        C = self.C
        n = C.shape[1]
        # scaling = C^T d: every entry of x gets the total of its block
        scaling = torch.zeros(n, dtype=_F64, device=self.device)
        scaling.index_add_(0, C.idx.to(torch.int64), C.val * self.d[C.rows_of_entries()])
        nz = torch.nonzero(scaling > thresh).reshape(-1)                 # the columns that stay
        self.nz_cols = nz
        scaling = scaling[nz].contiguous()
        self.x_split = torch.nan_to_num(self.x_true[nz] / scaling)
        newcol = torch.full((n,), -1, dtype=torch.int64, device=self.device)
        newcol[nz] = torch.arange(nz.numel(), device=self.device, dtype=torch.int64)
        full_scale = torch.zeros(n, dtype=_F64, device=self.device)
        full_scale[nz] = scaling
        self.C = C.map_columns(newcol, nz.numel())
        self.AA = self.AA.map_columns(newcol, nz.numel(), scale=full_scale)
        self.scaling = scaling

    def cleanup(self):
        """Remove zero rows (bsls_matrices.py:128-134; bsls_utils.remove_zero_rows keeps rows whose SUM is non-zero)."""
        keep = torch.nonzero(self.AA.row_sums() != 0).reshape(-1)
        self.AA, self.bb = self.AA.select_rows(keep), self.bb[keep].contiguous()
        keep = torch.nonzero(self.C.row_sums() != 0).reshape(-1)
        self.C, self.d = self.C.select_rows(keep), self.d[keep].contiguous()

    def blockify(self, noisy=False):
        """Re-arrange the columns so that C is blockwise diagonal (bsls_matrices.py:109-126)."""
        C = self.C
        pos = C.val > 0
        rows = C.rows_of_entries()
        self.block_sizes = torch.bincount(rows[pos], minlength=C.shape[0]).to(torch.int64)
        # columns grouped by the row of C they belong to (CSR order: by block, then by column id)
        sort_index = C.idx.to(torch.int64)
        n = C.shape[1]
        assert sort_index.numel() == n, "C must hold exactly one entry per column (simplex incidence)"
        newcol = torch.empty(n, dtype=torch.int64, device=self.device)
        newcol[sort_index] = torch.arange(n, device=self.device, dtype=torch.int64)
        self.AA = self.AA.map_columns(newcol, n)
        self.x_true = self.x_true[sort_index]             # as the reference: indexed in the reduced column space
        self.x_split = self.x_split[sort_index].contiguous()
        self.C = C.map_columns(newcol, n)
        self.rsort_index = newcol                                  # == argsort(sort_index): undoes the sort

    def simple_simplex_form(self, thresh=1e-5, noisy=False):
        """bsls_matrices.py:52-67"""
        self.consolidate(eq=self.eq)
        self.standard_simplex_form(thresh=thresh, noisy=noisy)
        self.cleanup()
        self.blockify(noisy=noisy)
        if self.AA is None or self.x_split is None:
            self.info['error'] = "AA,bb is empty"

    def degree_reduced_form(self, init=False):
        """bsls_matrices.py:69-86 (the equality-constrained branch: N and a feasible x0)."""
        self.simple_simplex_form()
        assert self.block_sizes is not None, "no equality constraint: nothing to eliminate (the reference falls back to lsmr)"
        sizes = self.block_sizes.cpu().numpy()
        starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
        self.N = NOperator(starts, int(sizes.sum()), self.device)
        assert not init, "init=True solves C x = 1 directly on the host in the reference; use the particular solution"
        self.x0 = particular_x0(sizes, self.device)

    @staticmethod
    def reconstruct(x_split, rsort_index=None, scaling=None, nz=None, n=None):
        """Unsort, unzero, untransform (bsls_matrices.py:170-182)."""
        x_unordered = x_split[rsort_index]
        x_rescaled = x_unordered * scaling
        x_true = torch.zeros(int(n), dtype=_F64, device=x_split.device)
        x_true[nz] = x_rescaled
        return x_true

    # -- access -------------------------------------------------------------------------------------------------
    def get_LS(self):
        """(AA, bb, N, block_sizes, x_split, nz_cols, scaling, rsort_index, x0)  (bsls_matrices.py:315-322)"""
        return (self.AA, self.bb, self.N, self.block_sizes.cpu().numpy(), self.x_split, self.nz_cols, self.scaling,
                self.rsort_index, self.x0)

    def problem(self):
        """(AA, bb) as an :class:`LsqProblem`: CSR of AA and of AA^T built on the device."""
        AT = self.AA.transpose()
        A = self.AA
        return LsqProblem((A.ptr, A.idx, A.val, AT.ptr, AT.idx, AT.val, A.shape), self.bb, device=self.device)
