"""Randomised differential test of the simplex / l1-ball projection against the oracle (bit-exact): random layouts
(uniform, ragged, offsets, blocks up to 20000 entries), random inputs (Gaussian at several scales, dense supports near
the simplex, ties, sums within ulps of 1 for the ball).

    python tools/proj_fuzz.py [seconds] [seed0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bsls_b200 import c_extensions as api
from oracle import cpu

port = cpu.port()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def layout(rng):
    kind = rng.randint(5)
    if kind == 0:
        K = int(rng.choice([1, 2, 3, 4, 5, 8, 16, 20, 31, 32, 33, 64, 100, 128, 129, 300, 512, 513, 2000]))
        sizes = np.full(max(1, int(rng.randint(1, 40000) // K) + 1), K)
    elif kind == 1:
        hi = int(rng.choice([8, 40, 300, 2000, 8192]))
        total = int(rng.randint(1000, 200000))
        u = rng.rand(total // 2 + 10)
        sizes = np.clip(np.floor(1.0 * u ** (-1 / 1.3)), 1, hi).astype(np.int64)
        sizes = sizes[np.cumsum(sizes) <= total]
        if len(sizes) == 0:
            sizes = np.array([5])
    elif kind == 2:
        sizes = rng.choice([1, 8, 9, 16, 17, 31, 32, 33, 34, 511, 512, 513, 8191, 8192, 8193], size=rng.randint(1, 300))
    elif kind == 3:
        sizes = rng.randint(1, 9, size=rng.randint(1, 30000))
        for _ in range(rng.randint(0, 6)):
            sizes[rng.randint(len(sizes))] = rng.randint(33, 20000)
    else:
        sizes = rng.randint(20, 700, size=rng.randint(1, 800))
    first = int(rng.choice([0, 0, 1, 5, 33]))
    starts = first + np.concatenate(([0], np.cumsum(sizes)[:-1]))
    return first, np.asarray(sizes, dtype=np.int64), starts.astype(np.int64)


def values(rng, first, sizes):
    n = int(sizes.sum())
    kind = rng.randint(5)
    if kind == 0:
        y = rng.randn(n) * float(rng.choice([1e-3, 1.0, 1e3]))
    elif kind == 1:   # dense support: a feasible point, slightly perturbed
        y = np.concatenate([rng.dirichlet(np.ones(k)) for k in sizes]) + 1e-4 * rng.randn(n)
    elif kind == 2:   # ties
        y = rng.randint(-2, 3, size=n).astype(np.float64) * 0.5
    elif kind == 3:   # sums near 1 (ball decision)
        y = np.concatenate([rng.dirichlet(np.ones(k)) * (1.0 + rng.randint(-3, 4) * 2.220446049250313e-16) for k in sizes])
        y[rng.rand(n) < 0.1] *= -1.0
    else:
        y = rng.rand(n)
    return np.concatenate((rng.randn(first), y))


t0 = time.time()
cases = 0
seed = seed0
while time.time() - t0 < budget:
    rng = np.random.RandomState(seed)
    first, sizes, starts = layout(rng)
    y = values(rng, first, sizes)
    ball = bool(rng.randint(2))
    want = y.copy()
    (port.proj_multi_ball if ball else port.proj_multi_simplex)(want, starts)
    t = dev(y)
    (api.proj_multi_ball_c if ball else api.proj_multi_simplex_c)(t, dev(starts))
    got = t.cpu().numpy()
    assert np.array_equal(got, want), (seed, first, len(sizes), int(sizes.max()), ball)
    cases += 1
    seed += 1
print("ok %d cases in %.0f s (seeds %d..%d)" % (cases, time.time() - t0, seed0, seed - 1))
