"""Column (OD-block) sharding of a block-simplex least-squares problem over ranks -- host logic.

Blocks are independent under the projections, and ``A x = sum_p A[:, cols_p] x_p``: rank p owns a
contiguous, block-aligned slice of the columns, balanced by the number of non-zeros, with its
block starts rebased to 0; ``b`` and the link vector are replicated and the partial products
are summed with one all-reduce per objective evaluation.  Pure
NumPy / scipy so that the partition can be checked on CPU ranks (gloo) without a GPU.
"""
import numpy as np


def partition_blocks(block_starts, n, col_nnz, world):
    """Cut points (in blocks) of a contiguous split into ``world`` parts with nearly equal nnz.
    Returns an int64 array of world + 1 block indices (first 0, last numblocks)."""
    starts = np.asarray(block_starts, dtype=np.int64)
    nb = len(starts)
    assert world >= 1 and nb >= world, "need at least one block per rank"
    ends = np.append(starts[1:], n)
    cum = np.concatenate(([0], np.cumsum(np.asarray(col_nnz, dtype=np.int64))))
    block_cum = cum[ends] - cum[starts[0]]                  # nnz up to the end of each block
    total = int(block_cum[-1])
    cuts = [0]
    for p in range(1, world):
        target = total * p / world
        k = int(np.searchsorted(block_cum, target, side="left")) + 1
        k = max(k, cuts[-1] + 1)
        k = min(k, nb - (world - p))
        cuts.append(k)
    cuts.append(nb)
    return np.asarray(cuts, dtype=np.int64)


def shard_problem(A, block_starts, rank, world):
    """Rank ``rank``'s slice: (A_local as scipy CSC of shape m x n_p, local block starts rebased to
    0, (col_lo, col_hi)).  ``A`` is any scipy sparse matrix whose columns are grouped by block."""
    import scipy.sparse as sps
    Ac = sps.csc_matrix(A)
    n = Ac.shape[1]
    starts = np.asarray(block_starts, dtype=np.int64)
    assert starts[0] == 0, "sharding assumes the blocks cover all columns"
    cuts = partition_blocks(starts, n, np.diff(Ac.indptr), world)
    b_lo, b_hi = int(cuts[rank]), int(cuts[rank + 1])
    col_lo = int(starts[b_lo])
    col_hi = int(starts[b_hi]) if b_hi < len(starts) else n
    local = Ac[:, col_lo:col_hi]
    local_starts = starts[b_lo:b_hi] - col_lo
    return local, local_starts, (col_lo, col_hi)
