"""Synthetic route-assignment problems of the BASELINE configs, generated on the device
(SURVEY.md section 8d): every route (column of A) traverses L distinct links drawn uniformly,
values 1; OD blocks of K routes; x_true ~ Dirichlet(1) per block; b = A x_true.

Data generation is set-up work, not the hot path: it uses torch's generators and sort.  The
result is handed to :class:`sparse.LsqProblem` as device CSR arrays of A and A^T, built
directly (no host round trip, no scipy), optionally for one rank's slice of the OD blocks.
"""
import numpy as np
import torch

from .plan import BlockPlan
from .sparse import LsqProblem

SEED = 237423433  # the seed of the reference's tests (tests/fast/*.py setUp)

CONFIGS = {
    # name: (nb, K, m, L)
    "C1": (1000, 5, 2000, 10),
    "C4": (100000, 20, 50000, 10),
    "C5": (10000000, 16, 1000000, 8),
}


def block_range(nb, rank, world):
    """Contiguous, block-aligned split of the OD blocks over ranks (uniform blocks: equal nnz)."""
    lo = (nb * rank) // world
    hi = (nb * (rank + 1)) // world
    return lo, hi


def route_links(n_routes, m, L, gen, device):
    """(n_routes, L) int32 link ids, distinct and ascending inside a route: a sorted draw from
    [0, m - L] plus 0..L-1."""
    assert m >= L
    base = torch.randint(0, m - L + 1, (n_routes, L), generator=gen, device=device, dtype=torch.int32)
    base, _ = torch.sort(base, dim=1)
    return base + torch.arange(L, device=device, dtype=torch.int32)


def transpose_pattern(rows_of_nnz, n_rows_out, n_cols_in_per, chunk=None):
    """CSR of the transpose of an index-only matrix whose rows all have ``n_cols_in_per``
    entries: ``rows_of_nnz`` (flat int32, row-major) holds the column ids.  Returns (ptr int64,
    idx int32) with ascending indices inside each output row."""
    nnz = rows_of_nnz.numel()
    order = torch.argsort(rows_of_nnz.to(torch.int64), stable=True)           # by output row, ties by input position
    idx = torch.div(order, n_cols_in_per, rounding_mode="floor").to(torch.int32)
    counts = torch.bincount(rows_of_nnz, minlength=n_rows_out)
    ptr = torch.zeros(n_rows_out + 1, dtype=torch.int64, device=rows_of_nnz.device)
    torch.cumsum(counts, 0, out=ptr[1:])
    assert int(ptr[-1]) == nnz
    return ptr, idx


def chunk_count(nb):
    """The blocks of a problem are generated in this many equal chunks, each from its own seed, so that the GLOBAL problem
    does not depend on how many ranks share it (a rank generates the chunks it owns)."""
    for g in (64, 8):
        if nb % g == 0:
            return g
    return 1


class SyntheticProblem:
    """One rank's slice of a synthetic problem: ``problem`` (LsqProblem over the local columns),
    ``plan`` / ``starts`` (local block layout), ``x_true``, ``x_init`` (local), ``b`` (global).
    The global problem is the same for every world size that divides ``chunk_count(nb)``."""

    def __init__(self, nb, K, m, L, device=None, seed=SEED, rank=0, world=1, comm=None, implicit_ones=True, noise=0.0):
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.nb_global, self.K, self.m, self.L = nb, K, m, L
        G = chunk_count(nb)
        assert G % world == 0 or world == 1, "world size %d must divide the %d generation chunks" % (world, G)
        if G % world:
            G = 1
        lo, hi = block_range(nb, rank, world)
        self.block_lo, self.block_hi = lo, hi
        nbl = hi - lo
        n = nbl * K
        self.nb, self.n = nbl, n
        per = nb // G                                               # blocks per chunk
        t_idx = torch.empty(n * L, dtype=torch.int32, device=device)   # CSR of A^T: row = route, L links each
        self.x_true = torch.empty(n, dtype=torch.float64, device=device)
        for c in range(lo // per, (hi + per - 1) // per):
            gen = torch.Generator(device=device).manual_seed(seed + 7919 * c)
            r0 = (c * per - lo) * K                                     # first local route of the chunk
            links = route_links(per * K, m, L, gen, device)
            t_idx[r0 * L:(r0 + per * K) * L] = links.reshape(-1)
            del links
            e = -torch.log(torch.rand(per, K, generator=gen, device=device, dtype=torch.float64))
            self.x_true[r0:r0 + per * K] = (e / e.sum(1, keepdim=True)).reshape(-1)
            del e
        t_ptr = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device=device)
        a_ptr, a_idx = transpose_pattern(t_idx, m, L)              # CSR of A: row = link
        self.starts = torch.arange(0, n, K, dtype=torch.int64, device=device)
        self.plan = BlockPlan(self.starts, n, device)
        self.x_init = torch.full((n,), 1.0 / K, dtype=torch.float64, device=device)
        ones_a = ones_t = None
        if not implicit_ones:
            ones_a = torch.ones(n * L, dtype=torch.float64, device=device)
            ones_t = ones_a
        zero_b = torch.zeros(m, dtype=torch.float64, device=device)
        self.problem = LsqProblem((a_ptr, a_idx, ones_a, t_ptr, t_idx, ones_t, (m, n)), zero_b, device=device, comm=comm)
        b = self.problem.matvec(self.x_true)                        # summed over ranks when sharded
        if noise > 0:
            gb = torch.Generator(device=device).manual_seed(seed + 1)  # same on every rank
            b = b + noise * torch.randn(m, generator=gb, device=device, dtype=torch.float64)
        self.b = b
        self.problem.set_b(self.b)
        self.nnz = n * L

    @classmethod
    def config(cls, name, **kw):
        nb, K, m, L = CONFIGS[name]
        return cls(nb, K, m, L, **kw)

    def solver_parts(self, min_eig=0.1):
        from .algorithm_utils import get_solver_parts
        step_size, proj, line_search, obj = get_solver_parts(self.problem, self.starts, min_eig)
        return step_size, proj, line_search, obj

    def bytes_bb_iteration(self, values_bytes=8):
        """Algorithmic bytes of one BB iteration without back-tracking:
        B_BB = 24 nnz + 72 n + 32 m + 4 nb for fp64 values and int32 indices."""
        per_nnz = 2 * (values_bytes + 4)
        return per_nnz * self.nnz + 72 * self.n + 32 * self.m + 4 * self.nb
