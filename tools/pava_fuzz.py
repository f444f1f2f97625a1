"""Randomised differential test of the isotonic-regression kernels against the oracle: random layouts (uniform and
ragged, tiny to 8192-entry blocks, leading offset), random inputs (continuous, heavy ties, decreasing stretches that
cross word / tile boundaries, zeros), cold calls, calls with a weight array (all ones, and warm restarts from a
previous result), update 0 / 1, clamp.  Bit-exact comparison of values and weights.

    python tools/pava_fuzz.py [seconds] [seed0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsls_b200
from bsls_b200 import c_extensions as api
from oracle import cpu

port = cpu.port()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def layout(rng):
    kind = rng.randint(6)
    if kind == 0:   # uniform
        K = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 500, 1024, 1025, 3000]))
        nb = max(1, int(rng.randint(1, 40000) // K) + 1)
        sizes = np.full(nb, K)
    elif kind == 1:  # power law
        hi = int(rng.choice([8, 40, 300, 2000, 8192]))
        total = int(rng.randint(1000, 200000))
        u = rng.rand(total // 2 + 10)
        sizes = np.clip(np.floor(1.0 * u ** (-1 / 1.3)), 1, hi).astype(np.int64)
        sizes = sizes[np.cumsum(sizes) <= total]
        if len(sizes) == 0:
            sizes = np.array([5])
    elif kind == 2:  # around the 32 / 64 / 512 / 1024 boundaries
        sizes = rng.choice([1, 15, 16, 17, 31, 32, 33, 34, 63, 64, 65, 511, 512, 513, 1023, 1024, 1025], size=rng.randint(1, 600))
    elif kind == 3:  # mostly tiny with a few long
        sizes = rng.randint(1, 9, size=rng.randint(1, 30000))
        for _ in range(rng.randint(0, 6)):
            sizes[rng.randint(len(sizes))] = rng.randint(33, 6000)
    elif kind == 4:  # mid sizes
        sizes = rng.randint(20, 700, size=rng.randint(1, 800))
    else:            # one or two blocks
        sizes = rng.randint(1, 8193, size=rng.randint(1, 3))
    first = int(rng.choice([0, 0, 1, 5, 33]))
    starts = first + np.concatenate(([0], np.cumsum(sizes)[:-1]))
    return first, np.asarray(sizes, dtype=np.int64), starts.astype(np.int64)


def values(rng, first, sizes):
    n = int(sizes.sum())
    kind = rng.randint(6)
    if kind == 0:
        y = rng.randn(n)
    elif kind == 1:
        y = rng.randint(0, 4, size=n).astype(np.float64)
    elif kind == 2:  # long decreasing stretches
        y = -np.arange(n, dtype=np.float64) * 0.25 + 3.0 * rng.randn(n) * (rng.rand(n) < 0.02)
    elif kind == 3:
        y = np.concatenate([rng.randint(-50, 50, size=k) + 50.0 * np.log(1 + np.arange(k)) for k in sizes])
    elif kind == 4:
        y = np.where(rng.rand(n) < 0.3, 0.0, rng.randint(-2, 3, size=n)).astype(np.float64)
    else:
        y = np.cumsum(rng.randn(n) * (rng.rand(n) < 0.5))
    return np.concatenate((rng.randn(first), y))


t0 = time.time()
cases = 0
seed = seed0
while time.time() - t0 < budget:
    rng = np.random.RandomState(seed)
    first, sizes, starts = layout(rng)
    y = values(rng, first, sizes)
    n = len(y)
    update = int(rng.randint(2))
    clip = bool(rng.randint(2))
    mode = rng.randint(3)  # 0 cold, 1 ones, 2 warm restart
    tag = (seed, first, len(sizes), int(sizes.max()), update, clip, mode)
    want = y.copy()
    if mode == 0:
        port.pava_multi(want, starts, update=update)
        t = dev(y)
        api.isotonic_regression_multi_c(t, dev(starts), None, update, clip01=clip)
        if clip:
            port.clip01(want[first:])
        got = t.cpu().numpy()
        assert np.array_equal(got[first:], want[first:]) and np.array_equal(got[:first], y[:first]), tag
    else:
        w0 = np.ones(n, dtype=np.int32)
        if mode == 2:  # the state a first pass without spreading leaves, then perturbed values
            y1 = y.copy()
            w0 = port.pava_multi(y1, starts, update=0).astype(np.int32)
            y = y1 + (rng.randn(n) * (rng.rand(n) < 0.3))
            want = y.copy()
        ww = port.pava_multi(want, starts, weight=w0.copy(), update=update)
        t, tw = dev(y), dev(w0.copy())
        api.isotonic_regression_multi_c(t, dev(starts), tw, update, clip01=clip)
        if clip:
            port.clip01(want[first:])
        got, gw = t.cpu().numpy(), tw.cpu().numpy()
        assert np.array_equal(got[first:], want[first:]), tag
        assert np.array_equal(gw[first:], ww[first:]), tag
    cases += 1
    seed += 1
print("ok %d cases in %.0f s (seeds %d..%d)" % (cases, time.time() - t0, seed0, seed - 1))
