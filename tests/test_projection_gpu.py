"""GPU parity of the segmented simplex / l1-ball projection against the CPU oracle.

Bar (BASELINE.json north_star): active set bit-exact, values within 1e-6 relative in fp64.
The kernels reproduce the reference's arithmetic order, so fp64 results are compared with
exact equality (np.array_equal), which implies both.  fp32 (an extension) uses 1e-4.
All calls go through the C ABI (ctypes) of libbsls_b200.so.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SEED = 237423433


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    import bsls_b200
    return bsls_b200


@pytest.fixture(scope="module")
def oracle():
    from oracle import cpu
    return cpu.ref() or cpu.port()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_project(api, y, starts, ball=False):
    t = dev(y)
    (api.proj_multi_ball_c if ball else api.proj_multi_simplex_c)(t, dev(np.asarray(starts, dtype=np.int64)))
    torch.cuda.synchronize()
    return t.cpu().numpy()


# ------------------------------------------------------------------ reference known answers
def test_reference_known_answers_device(api):
    z = np.array([5.352, 3.23, 32.78, -1.234, 1.7, 104., 53.])
    for blocks, truth in [([0, 2, 4], [1., 0., 1., 0., 0., 1., 0.]), ([0], [0., 0., 0., 0., 0., 1., 0.]),
                          ([0, 3], [0., 0., 1., 0., 0., 1., 0.])]:
        assert (gpu_project(api, z, blocks) == np.array(truth)).all()
    for truth, start, end in zip([[5.352, 3.23, 1., 0., 1.7, 104., 53.], [0., 0., 0., 0., 0, 1., 0.], z], [2, 0, 4], [4, 7, 4]):
        t = dev(z)
        api.proj_simplex_c(t, start, end)
        assert (t.cpu().numpy() == np.array(truth)).all()
    np.random.seed(SEED)
    t = dev(np.random.rand(7))
    api.proj_simplex_c(t, 0, 7)
    truth = np.array([0., .05006376, .54108944, 0., .38841272, 0., .02043408])
    assert np.linalg.norm(t.cpu().numpy() - truth) < 1e-6
    y = np.array([0.234, 0.5, 1.3, -1.234, 1.7, -1.0, 53.])
    assert (gpu_project(api, y, [0, 2, 4], ball=True) == np.array([0.234, 0.5, 1., 0., 0., 0., 1.])).all()


def test_reference_known_answers_host_abi(api):
    """Same vectors through the HOST entry points (the reference's exact C signatures)."""
    z = np.array([5.352, 3.23, 32.78, -1.234, 1.7, 104., 53.])
    for blocks, truth in [([0, 2, 4], [1., 0., 1., 0., 0., 1., 0.]), ([0], [0., 0., 0., 0., 0., 1., 0.]),
                          ([0, 3], [0., 0., 1., 0., 0., 1., 0.])]:
        y = z.copy()
        api.proj_multi_simplex_c(y, np.array(blocks))
        assert (y == np.array(truth)).all()
    for truth, start, end in zip([[5.352, 3.23, 1., 0., 1.7, 104., 53.], [0., 0., 0., 0., 0, 1., 0.], z], [2, 0, 4], [4, 7, 4]):
        y = z.copy()
        api.proj_simplex_c(y, start, end)
        assert (y == np.array(truth)).all()
    y = np.array([0.234, 0.5, 1.3, -1.234, 1.7, -1.0, 53.])
    api.proj_multi_ball_c(y, np.array([0, 2, 4]))
    assert (y == np.array([0.234, 0.5, 1., 0., 0., 0., 1.])).all()
    for b in [np.array([-1, 2, 4]), np.array([1, 3, 7]), np.array([0, 4, 2])]:
        with pytest.raises(AssertionError):
            api.proj_multi_simplex_c(z.copy(), b)
        with pytest.raises(AssertionError):
            api.proj_multi_simplex_c(dev(z), dev(b))


# ------------------------------------------------------------------ fixtures from the reference
def test_golden_projection_fixtures(api, golden_dir):
    d = np.load(os.path.join(golden_dir, "projection.npz"))
    for i in range(int(d["count"])):
        y, starts = d["y%d" % i], d["starts%d" % i]
        assert np.array_equal(gpu_project(api, y, starts), d["simplex%d" % i]), i
        assert np.array_equal(gpu_project(api, y, starts, ball=True), d["ball%d" % i]), i
        h = y.copy()
        api.proj_multi_simplex_c(h, starts)
        assert np.array_equal(h, d["simplex%d" % i]), ("host abi", i)


# ------------------------------------------------------------------ oracle, uniform sizes
@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 19, 20, 24, 31, 32, 33, 63, 64, 100, 128, 200, 256, 511, 512])
@pytest.mark.parametrize("dist", ["normal", "uniform", "near_simplex", "ties"])
def test_uniform_blocks_match_oracle(api, oracle, K, dist):
    rng = np.random.RandomState(SEED + K)
    nb = max(3, 70000 // K) + 17     # several tiles, ragged last tile
    n = nb * K
    if dist == "normal":
        y = rng.randn(n) * rng.choice([0.01, 1.0, 100.0])
    elif dist == "uniform":
        y = rng.rand(n)
    elif dist == "near_simplex":
        y = rng.dirichlet(np.ones(K), size=nb).reshape(-1) + 1e-3 * rng.randn(n)
    else:
        y = rng.randint(-3, 4, size=n).astype(float) / 4.0
    starts = np.arange(0, n, K, dtype=np.int64)
    want = y.copy()
    oracle.proj_multi_simplex(want, starts)
    got = gpu_project(api, y, starts)
    assert np.array_equal(got > 0, want > 0)            # active set, bit-exact
    assert np.array_equal(got, want)                    # values, bit-exact
    want = y.copy()
    oracle.proj_multi_ball(want, starts)
    got = gpu_project(api, y, starts, ball=True)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("K", [5, 16, 64])
def test_offset_prefix_and_unaligned_views(api, oracle, K):
    """Entries before blocks[0] stay untouched; odd offsets defeat the 16-byte fast path."""
    rng = np.random.RandomState(SEED)
    nb = 5000
    for first in (1, 2, 7):
        n = first + nb * K
        y = rng.randn(n)
        starts = first + np.arange(0, nb * K, K, dtype=np.int64)
        want = y.copy()
        oracle.proj_multi_simplex(want, starts)
        assert np.array_equal(gpu_project(api, y, starts), want)
    # a view whose base pointer is only 8-byte aligned
    big = dev(rng.randn(nb * K + 1))
    view = big[1:]
    want = view.cpu().numpy().copy()
    oracle.proj_multi_simplex(want, np.arange(0, nb * K, K))
    api.proj_multi_simplex_c(view, dev(np.arange(0, nb * K, K, dtype=np.int64)))
    assert np.array_equal(view.cpu().numpy(), want)


@pytest.mark.parametrize("K", [4, 16, 20, 64])
def test_fp32_extension(api, K):
    rng = np.random.RandomState(SEED)
    nb = 20000
    y = rng.randn(nb * K)
    starts = np.arange(0, nb * K, K, dtype=np.int64)
    from oracle import cpu
    want = y.astype(np.float32).astype(np.float64)
    cpu.port().proj_multi_simplex(want, starts)
    t = dev(y.astype(np.float32))
    api.proj_multi_simplex_c(t, dev(starts))
    got = t.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())   # north_star fp32 tolerance
    assert np.abs(got.reshape(nb, K).sum(1) - 1).max() < 1e-5


def test_fp32_block_length_limit_is_a_distinct_error(api):
    """fp32 projection of a block longer than 8192 entries is not served (the candidate-selection margin is too wide in
    fp32): NotImplementedError -- not the AssertionError the reference raises for invalid block arrays -- and the
    buffer is left untouched; the fp64 entry point takes the same layout."""
    rng = np.random.RandomState(SEED)
    y = rng.randn(30000)
    starts = np.array([0, 100, 9000], dtype=np.int64)            # last block: 21000 entries
    t = dev(y.astype(np.float32))
    before = t.clone()
    with pytest.raises(NotImplementedError):
        api.proj_multi_simplex_c(t, dev(starts))
    assert torch.equal(t, before)
    t64 = dev(y)
    api.proj_multi_simplex_c(t64, dev(starts))
    want = y.copy()
    cpu_port().proj_multi_simplex(want, starts)
    assert np.array_equal(t64.cpu().numpy(), want)


@pytest.mark.parametrize("K", [3, 4, 16, 20, 24])
def test_ball_sum_near_one(api, K):
    """l1-ball: the decision `clipped block sums to more than 1` (proj_simplex.h:54-62) on blocks whose sum is
    within a few ulps of 1 -- the kernels add in parallel and must fall back to the reference's order there."""
    rng = np.random.RandomState(SEED + K)
    nb = 20000
    x = rng.dirichlet(np.ones(K), size=nb)                      # rows sum to 1 up to rounding
    x *= (1.0 + rng.randint(-3, 4, size=(nb, 1)) * 2.220446049250313e-16)
    x[rng.rand(nb, K) < 0.15] *= -1.0                           # some negatives (clipped away)
    x = x / np.maximum(np.where(x > 0, x, 0).sum(1, keepdims=True), 1e-300)  # clipped sums back to ~1
    x *= (1.0 + rng.randint(-2, 3, size=(nb, 1)) * 2.220446049250313e-16)
    y = x.reshape(-1).copy()
    starts = np.arange(0, nb * K, K)
    want = y.copy()
    cpu_port().proj_multi_ball(want, starts)
    got = gpu_project(api, y, starts, ball=True)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ size-independent properties
@pytest.mark.parametrize("K,nb", [(4, 10 ** 6), (16, 10 ** 6), (64, 10 ** 6)])
def test_full_size_properties(api, K, nb):
    """BASELINE config 2 at full size: feasibility, idempotence, optimality (KKT), plus an
    exact comparison with the oracle on a strided sample of blocks."""
    g = torch.Generator(device="cuda").manual_seed(SEED + K)
    y = torch.randn(nb * K, dtype=torch.float64, device="cuda", generator=g)
    starts = torch.arange(0, nb * K, K, dtype=torch.int64, device="cuda")
    x = y.clone()
    api.proj_multi_simplex_c(x, starts)
    X, Y = x.view(nb, K), y.view(nb, K)
    assert (x >= 0).all()
    assert (X.sum(1) - 1).abs().max().item() < 1e-12
    # KKT: on the support y - x is one constant per block; off the support y - shift <= 0
    shift = ((Y - X) * (X > 0)).sum(1) / (X > 0).sum(1)
    assert (((Y - X) - shift[:, None]).abs() * (X > 0)).max().item() < 1e-12
    assert ((Y - shift[:, None]) * (X == 0)).max().item() <= 1e-12
    x2 = x.clone()
    api.proj_multi_simplex_c(x2, starts)
    assert (x2 - x).abs().max().item() < 1e-15       # idempotent up to one rounding
    from oracle import cpu
    sample = torch.arange(0, nb, 997, device="cuda")
    ys = Y[sample].reshape(-1).cpu().numpy().copy()
    cpu.port().proj_multi_simplex(ys, np.arange(0, ys.size, K))
    assert np.array_equal(X[sample].reshape(-1).cpu().numpy(), ys)


# ------------------------------------------------------------------ ragged layouts
def power_law_sizes(rng, total, lo, hi, alpha=1.5):
    sizes = []
    left = total
    while left > 0:
        k = int(min(hi, max(lo, np.floor(lo * rng.random_sample() ** (-1.0 / alpha)))))
        k = min(k, left)
        sizes.append(k)
        left -= k
    return np.array(sizes, dtype=np.int64)


@pytest.mark.parametrize("lo,hi,total", [(1, 8, 30000), (2, 64, 100000), (2, 512, 200000), (2, 4096, 400000),
                                         (300, 700, 100000), (513, 8192, 300000)])
@pytest.mark.parametrize("first", [0, 3])
def test_ragged_blocks_match_oracle(api, oracle, lo, hi, total, first):
    rng = np.random.RandomState(SEED + lo + hi)
    sizes = power_law_sizes(rng, total, lo, hi) if lo < 300 else rng.randint(lo, hi + 1, size=max(2, total // hi))
    n = int(first + sizes.sum())
    starts = first + np.concatenate(([0], np.cumsum(sizes)[:-1]))
    for dist in ("normal", "near_simplex", "ties"):
        if dist == "normal":
            y = rng.randn(n)
        elif dist == "near_simplex":
            y = np.concatenate([np.zeros(first)] + [rng.dirichlet(np.ones(k)) for k in sizes]) + 1e-4 * rng.randn(n)
        else:
            y = rng.randint(-3, 4, size=n).astype(float) / 4.0
        want = y.copy()
        cpu_port().proj_multi_simplex(want, starts)
        got = gpu_project(api, y, starts)
        assert np.array_equal(got[:first], y[:first])
        assert np.array_equal(got, want), (dist, np.flatnonzero(got != want)[:5])
        want = y.copy()
        cpu_port().proj_multi_ball(want, starts)
        assert np.array_equal(gpu_project(api, y, starts, ball=True), want), dist


def cpu_port():
    from oracle import cpu
    return cpu.port()   # heap scratch: safe for blocks beyond the reference's stack VLA


@pytest.mark.parametrize("K", [513, 1000, 4096, 8192])
def test_large_uniform_blocks(api, K):
    rng = np.random.RandomState(SEED + K)
    nb = 37
    y = rng.randn(nb * K)
    starts = np.arange(0, nb * K, K, dtype=np.int64)
    want = y.copy()
    cpu_port().proj_multi_simplex(want, starts)
    assert np.array_equal(gpu_project(api, y, starts), want)


def test_config3_power_law_full_size(api):
    """BASELINE config 3 layout: power-law sizes in [2, 4096], 10^7 variables; compared with
    the multi-threaded oracle port (bit-exact) and checked for feasibility."""
    rng = np.random.RandomState(SEED + 3)
    u = rng.random_sample(3_000_000)
    sizes = np.clip(np.floor(2 * u ** (-1 / 1.5)), 2, 4096).astype(np.int64)
    csum = np.cumsum(sizes)
    cut = int(np.searchsorted(csum, 10 ** 7))
    sizes = sizes[:cut + 1]
    sizes[-1] -= int(sizes.sum() - 10 ** 7)
    if sizes[-1] <= 0:
        sizes = sizes[:-1]
    n = int(sizes.sum())
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    y = rng.randn(n)
    want = y.copy()
    cpu_port().proj_multi_simplex(want, starts, threads=8)
    got = gpu_project(api, y, starts)
    assert np.array_equal(got, want)
    sums = np.add.reduceat(got, starts)
    assert np.abs(sums - 1).max() < 1e-12 and got.min() >= 0


@pytest.mark.parametrize("K", [4, 16, 20, 64, 300])
def test_adversarial_near_ties(api, K):
    """Blocks built so that the reference's acceptance test u_k + (1 - S_k)/(k+1) > 0 lands
    exactly on, or within a few ulps of, zero at some position k: this drives the kernel's
    exact fall-back (the fast sign test cannot decide) and pins the active set bit for bit."""
    rng = np.random.RandomState(SEED + 31 * K)
    blocks = []
    for trial in range(4000):
        k = rng.randint(1, K)                      # position of the borderline element
        head = np.sort(rng.rand(k) * rng.choice([0.3, 1.0, 3.0]))[::-1]
        s = 0.0
        for v in head:
            s += v                                  # running sum, left to right
        u = (s - 1.0) / k                           # makes u*(k+1) + 1 - (s + u) ~ 0
        for _ in range(abs(int(rng.randint(-4, 5)))):
            u = np.nextafter(u, np.inf if rng.rand() < 0.5 else -np.inf)
        if u > head[-1]:
            continue
        tail = u - 1.0 - rng.rand(K - k - 1) * 5
        blk = np.concatenate([head, [u], tail])
        rng.shuffle(blk)
        blocks.append(blk)
    y = np.concatenate(blocks)
    nb = len(blocks)
    starts = np.arange(0, nb * K, K, dtype=np.int64)
    want = y.copy()
    cpu_port().proj_multi_simplex(want, starts)
    got = gpu_project(api, y, starts)
    assert np.array_equal(got > 0, want > 0)
    assert np.array_equal(got, want)
    # the same blocks behind two odd-sized blocks: a ragged layout, i.e. the tile kernel
    starts2 = np.concatenate(([0, 1], 3 + starts))
    y2 = np.concatenate((rng.randn(3), y))
    want = y2.copy()
    cpu_port().proj_multi_simplex(want, starts2)
    assert np.array_equal(gpu_project(api, y2, starts2), want)


@pytest.mark.parametrize("K", [32, 33, 64, 100, 128, 300])
def test_selection_path_mixed_scales(api, K):
    """Inputs that stress the rounding margin of the candidate selection (select_core.cuh): values
    spread over many orders of magnitude and signs, huge negative tails, blocks of equal values,
    supports of every density.  Uniform layout (thread-per-block selection + sorter fallback) and
    the same blocks in a ragged layout (thread / warp / CTA paths).  Bit-exact."""
    rng = np.random.RandomState(SEED + 977 * K)
    blocks = []
    for trial in range(1500):
        kind = trial % 6
        if kind == 0:      # wide dynamic range
            blk = rng.randn(K) * 10.0 ** rng.randint(-8, 9, size=K)
        elif kind == 1:    # a few O(1) values over a huge negative tail
            blk = -rng.rand(K) * 1e12
            blk[rng.choice(K, 3, replace=False)] = rng.rand(3) * 2
        elif kind == 2:    # all equal / two levels
            blk = np.full(K, rng.randn())
            blk[: rng.randint(0, K)] += rng.choice([0.0, 1e-16, 1e-9, 0.5])
        elif kind == 3:    # support of a prescribed size, the rest barely below the threshold
            s = rng.randint(1, K + 1)
            x = np.zeros(K)
            x[:s] = rng.dirichlet(np.ones(s))
            theta = rng.randn() * 3
            blk = x + theta
            blk[s:] -= rng.rand(K - s) * rng.choice([1e-15, 1e-12, 1e-9, 1e-3])
        elif kind == 4:    # large common offset
            blk = 1e6 + rng.rand(K)
        else:              # tiny values
            blk = rng.randn(K) * 1e-12
        rng.shuffle(blk)
        blocks.append(blk)
    y = np.concatenate(blocks)
    starts = np.arange(0, len(blocks) * K, K, dtype=np.int64)
    for ball in (False, True):
        want = y.copy()
        (cpu_port().proj_multi_ball if ball else cpu_port().proj_multi_simplex)(want, starts)
        got = gpu_project(api, y, starts, ball=ball)
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (ball, bad[:5] // K, got[bad[:3]], want[bad[:3]])
    starts2 = np.concatenate(([0, 1], 3 + starts))
    y2 = np.concatenate((rng.randn(3), y))
    want = y2.copy()
    cpu_port().proj_multi_simplex(want, starts2)
    got = gpu_project(api, y2, starts2)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, (bad[:5], got[bad[:3]], want[bad[:3]])


@pytest.mark.parametrize("n", [8193, 20000, 10 ** 5, 10 ** 6])
def test_blocks_longer_than_shared_memory(api, n):
    """The reference's own stress shape (python/experiments/test_stress_proj_simplex.py:27: ONE block of
    10^3 .. 10^6 values): blocks beyond the 8192-value shared-memory window stage only their candidates."""
    rng = np.random.RandomState(SEED + n % 1000)
    for kind in ("normal", "uniform", "sparse_support"):
        if kind == "normal":
            y = rng.randn(n)
        elif kind == "uniform":
            y = rng.rand(n)
        else:
            y = -rng.rand(n)
            idx = rng.choice(n, 200, replace=False)
            y[idx] = rng.dirichlet(np.ones(200)) + 0.3
        want = y.copy()
        cpu_port().proj_simplex(want, 0, n)
        got = torch_vec(y)
        api.proj_simplex_c(got, 0, n)
        assert np.array_equal(got.cpu().numpy(), want), kind
    # the same long block inside a ragged layout, between short blocks, and the l1-ball variant
    sizes = np.array([3, 7, n, 2, 40, 600])
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    y = rng.randn(int(sizes.sum()))
    for ball in (False, True):
        want = y.copy()
        (cpu_port().proj_multi_ball if ball else cpu_port().proj_multi_simplex)(want, starts)
        assert np.array_equal(gpu_project(api, y, starts, ball=ball), want), ball


def torch_vec(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()
