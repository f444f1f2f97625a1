"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): OD blocks sharded over ranks, the
partial link vectors summed by the library's NCCL all-reduce, native BB loop in lock step.
The sharded solve must reach the single-GPU objective (1e-6 relative)."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np, scipy.sparse as sps, torch, torch.distributed as dist
import bsls_b200
from bsls_b200.shard import shard_problem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = bsls_b200.Communicator()
rng = np.random.RandomState(5)
nb, K, m, L = 4000, 6, 900, 5
n = nb * K
rows = np.concatenate([rng.choice(m, L, replace=False) for _ in range(n)])
A = sps.csc_matrix((np.ones(n * L), (rows, np.repeat(np.arange(n), L))), shape=(m, n))
x_true = rng.dirichlet(np.ones(K), size=nb).reshape(-1)
b = A.dot(x_true) + 0.3 * rng.randn(m)
starts = np.arange(0, n, K)
x0 = np.ones(n) / K
Al, ls, (lo, hi) = shard_problem(A, starts, rank, world)
prob = bsls_b200.LsqProblem(Al, b, comm=comm)
parts = bsls_b200.algorithm_utils.get_solver_parts(prob, ls, 0.1)
xl = torch.as_tensor(x0[lo:hi]).cuda()
g = torch.empty_like(xl)
f0 = prob.obj(xl, g)
r = A.dot(x0) - b
assert abs(f0 - .5 * r.dot(r)) <= 1e-12 * abs(f0), (f0, .5 * r.dot(r))
np.testing.assert_allclose(g.cpu().numpy(), A.T.dot(r)[lo:hi], rtol=1e-11, atol=1e-12)
sol = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], xl, max_iter=300)
md = bsls_b200.BATCH.solve_MD(parts[3], ls, parts[0], xl, max_iter=40)
lb = bsls_b200.BATCH.solve_LBFGS(parts[3], parts[1], parts[2], xl, max_iter=60)
p2p = bool(bsls_b200._lib.lib().bsls_comm_p2p_ready(comm.handle))
xs = [None] * world
dist.all_gather_object(xs, sol["x"].cpu().numpy())
if rank == 0:
    x = np.concatenate(xs)
    rr = A.dot(x) - b
    print("RESULT " + json.dumps({"f": sol["f"], "f_check": float(.5 * rr.dot(rr)), "iters": sol["iterations"], "md_f": md["f"],
                                  "lbfgs_f": lb["f"], "p2p": p2p}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_bb_matches_single_gpu(tmp_path, exchange):
    """exchange = p2p: the link vector is reduced by the library's own kernel over NVLink peer memory (csrc/p2p.cuh);
    nccl: ncclAllReduce + ncclAllGather (BSLS_P2P=0)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ)
    env["BSLS_P2P"] = "1" if exchange == "p2p" else "0"
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert res["f"] == pytest.approx(res["f_check"], rel=1e-10)
    assert res["p2p"] == (exchange == "p2p")
    # single GPU, same problem
    import numpy as np
    import scipy.sparse as sps
    import bsls_b200
    rng = np.random.RandomState(5)
    nb, K, m, L = 4000, 6, 900, 5
    n = nb * K
    rows = np.concatenate([rng.choice(m, L, replace=False) for _ in range(n)])
    A = sps.csc_matrix((np.ones(n * L), (rows, np.repeat(np.arange(n), L))), shape=(m, n))
    x_true = rng.dirichlet(np.ones(K), size=nb).reshape(-1)
    b = A.dot(x_true) + 0.3 * rng.randn(m)
    starts = np.arange(0, n, K)
    parts = bsls_b200.algorithm_utils.get_solver_parts((A, b), starts, 0.1, is_sparse=True)
    x0 = torch.full((n,), 1.0 / K, dtype=torch.float64, device="cuda")
    sol = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], x0, max_iter=300)
    assert res["f"] == pytest.approx(sol["f"], rel=1e-6)
    md = bsls_b200.BATCH.solve_MD(parts[3], starts, parts[0], x0, max_iter=40)
    assert res["md_f"] == pytest.approx(md["f"], rel=1e-9)
    lb = bsls_b200.BATCH.solve_LBFGS(parts[3], parts[1], parts[2], x0, max_iter=60)
    assert res["lbfgs_f"] == pytest.approx(lb["f"], rel=1e-6)
