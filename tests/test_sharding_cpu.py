"""The N > 1 path on CPU ranks (gloo, world_size 2): the block-aligned column partition, the
sum of per-rank partial link vectors, and the sharded Barzilai-Borwein loop (local x, one
all-reduce of A_p x_p per objective evaluation, one all-reduce of the dot products per step)
reach the same objective as the single-process reference loop."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sps

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_problem():
    rng = np.random.RandomState(99)
    sizes = rng.randint(2, 9, size=120)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    n, m, L = int(sizes.sum()), 70, 5
    rows = np.concatenate([rng.choice(m, L, replace=False) for _ in range(n)])
    A = sps.csc_matrix((np.ones(n * L), (rows, np.repeat(np.arange(n), L))), shape=(m, n))
    x_true = np.concatenate([rng.dirichlet(np.ones(k)) for k in sizes])
    b = A.dot(x_true) + 0.2 * rng.randn(m)
    x0 = np.concatenate([np.ones(k) / k for k in sizes])
    return A, b, starts, x0


def _allreduce(a, op=dist.ReduceOp.SUM):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
    dist.all_reduce(t, op=op)
    return t.numpy()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        import bsls_b200  # noqa: F401
        from bsls_b200.shard import shard_problem
        from oracle import cpu, solvers_np as S
        A, b, starts, x0 = _make_problem()
        Al, lstarts, (lo, hi) = shard_problem(A, starts, rank, world)
        Alr, AlT = sps.csr_matrix(Al), sps.csr_matrix(Al.T)
        # (1) partial products sum to the full product
        full = _allreduce(Alr.dot(x0[lo:hi]))
        assert np.allclose(full, A.dot(x0), rtol=1e-13, atol=1e-13)
        # (2) sharded BB: the loop of BATCH.solve_BB with the two collectives of SURVEY 8e
        chk = cpu.port()

        def obj(x, g):
            r = _allreduce(Alr.dot(x)) - b
            np.copyto(g, AlT.dot(r))
            return .5 * r.dot(r)

        def proj(x):
            chk.proj_multi_simplex(x, lstarts)

        x = x0[lo:hi].copy()
        g, g_new, x_new = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        f = obj(x, g)
        f_old, i, sxy, syy = np.inf, 1, 0.0, 0.0
        while True:
            flag, stop = S.stopping(i, 300, f, f_old, 1e-6, 1e-12, None)
            if flag:
                break
            t = 1.0 if i == 1 else sxy / syy
            np.add(x, -t * g, x_new)
            proj(x_new)
            f_new = obj(x_new, g_new)
            gd = _allreduce(np.array([g.dot(x_new - x)]))[0]
            tt = 1.0
            while f_new > f + 1e-4 * gd:
                tt *= .8
                step = _allreduce(np.array([np.abs(x_new - x).max()]), dist.ReduceOp.MAX)[0]
                if step < 1e-12:
                    f_new = f
                    np.copyto(g_new, g)
                    np.copyto(x_new, x)
                    break
                np.copyto(x_new, (1.0 - tt) * x + tt * x_new)
                f_new = obj(x_new, g_new)
                gd = _allreduce(np.array([g.dot(x_new - x)]))[0]
            dx, dg = x_new - x, g_new - g
            sxy, syy = _allreduce(np.array([dx.dot(dg), dg.dot(dg)]))
            f_old, f = f, f_new
            np.copyto(x, x_new)
            np.copyto(g, g_new)
            i += 1
        out[rank] = (f, i, lo, hi, x.copy())
    finally:
        dist.destroy_process_group()


def test_sharded_bb_two_ranks_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    from oracle import solvers_np as S
    A, b, starts, x0 = _make_problem()
    parts = S.get_solver_parts(sps.csr_matrix(A), b, starts, 0.1)
    ref = S.solve_BB(parts[3], parts[1], parts[2], x0, max_iter=300)
    f0, i0 = out[0][0], out[0][1]
    assert out[1][0] == f0 and out[1][1] == i0          # ranks stay in lock step
    assert f0 == pytest.approx(ref["f"], rel=1e-6)
    x = np.concatenate([out[r][4] for r in range(world)])
    assert out[0][3] == out[1][2] and out[1][3] == A.shape[1]
    r = A.dot(x) - b
    assert .5 * r.dot(r) == pytest.approx(f0, rel=1e-12)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_is_block_aligned_and_balanced(world):
    import sys
    from bsls_b200.shard import partition_blocks, shard_problem
    A, b, starts, x0 = _make_problem()
    cuts = partition_blocks(starts, A.shape[1], np.diff(A.indptr), world)
    assert cuts[0] == 0 and cuts[-1] == len(starts) and np.all(np.diff(cuts) > 0)
    tot = np.zeros(A.shape[0])
    covered = 0
    nnz = []
    for r in range(world):
        Al, ls, (lo, hi) = shard_problem(A, starts, r, world)
        assert ls[0] == 0 and lo == starts[cuts[r]]
        tot += Al.dot(x0[lo:hi])
        covered += hi - lo
        nnz.append(Al.nnz)
    assert covered == A.shape[1]
    np.testing.assert_allclose(tot, A.dot(x0), rtol=1e-13)
    assert max(nnz) - min(nnz) <= 2 * 8 * 5 + A.nnz // (10 * world)
