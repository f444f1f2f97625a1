"""GPU parity of the segmented isotonic regression (PAVA) against the CPU oracle.

Bar: pool structure bit-exact, values within 1e-6 relative (north_star).  The kernels replay
the reference's variant-1 sweeps with the same arithmetic, so values AND weight arrays
(including stale interior entries) are compared with exact equality.
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


def port():
    from oracle import cpu
    return cpu.port()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_pava(api, y, starts, update=1, with_weight=True, clip=False, w0=None):
    t = dev(y)
    w = None
    if with_weight:
        w = dev(np.ones(len(y), dtype=np.int32) if w0 is None else w0.astype(np.int32))
    api.isotonic_regression_multi_c(t, dev(np.asarray(starts, dtype=np.int64)), w, update, clip01=clip)
    torch.cuda.synchronize()
    return t.cpu().numpy(), (None if w is None else w.cpu().numpy())


def test_known_answers(api):
    for variant in (api.isotonic_regression_c, api.isotonic_regression_c_2, api.isotonic_regression_c_3):
        t = dev(np.array([4., 5., 1., 6., 8., 7.]))
        variant(t, 0, 6)
        assert np.allclose(t.cpu().numpy(), [10 / 3., 10 / 3., 10 / 3., 6., 7.5, 7.5], rtol=0, atol=1e-15)
        h = np.array([4., 5., 1., 6., 8., 7.])
        variant(h, 0, 6)                                  # host ABI
        assert np.allclose(h, [10 / 3., 10 / 3., 10 / 3., 6., 7.5, 7.5], rtol=0, atol=1e-15)
    y, w = gpu_pava(api, np.array([4., 5., 1., 6., 8., 7.]), [0, 2, 4])
    assert (y == np.array([4., 5., 1., 6., 7.5, 7.5])).all()
    y, w = gpu_pava(api, np.array([4., 5., 1., 6., 8., 7.]), [0])
    assert list(w) == [3, 2, 1, 1, 2, 1]
    y, w = gpu_pava(api, np.array([1., 3., 1., 2.]), [0])
    assert list(w) == [1, 2, 1, 1] and list(y) == [1., 2., 2., 2.]   # equal means stay separate pools


def test_golden_fixtures(api, golden_dir):
    d = np.load(os.path.join(golden_dir, "pava.npz"))
    for i in range(int(d["count"])):
        y, starts = d["y%d" % i], d["starts%d" % i]
        for update in (1, 0):
            gy, gw = gpu_pava(api, y, starts, update=update)
            assert np.array_equal(gy, d["v1_u%d_y%d" % (update, i)]), (i, update)
            assert np.array_equal(gw, d["v1_u%d_w%d" % (update, i)]), (i, update)
        gy, _ = gpu_pava(api, y, starts, with_weight=False)
        assert np.array_equal(gy, d["v1_u1_y%d" % i])
        h, hw = y.copy(), np.ones(len(y), dtype=np.int32)
        api.isotonic_regression_multi_c(h, starts, hw)           # host ABI, weights in/out
        assert np.array_equal(h, d["v1_u1_y%d" % i]) and np.array_equal(hw, d["v1_u1_w%d" % i])
        for name, fn in (("v2", api.isotonic_regression_multi_c_2), ("v3", api.isotonic_regression_multi_c_3)):
            t = dev(y)
            fn(t, dev(starts))
            ref = d["%s_y%d" % (name, i)] if name == "v2" else d["v3_u1_y%d" % i]
            scale = max(1.0, np.abs(ref).max())
            assert np.abs(t.cpu().numpy() - ref).max() <= 1e-8 * scale    # the reference's own tolerance


def make_input(rng, sizes, kind):
    if kind == "ref":      # tests/fast/test_isotonic_regression.py:43, ramp restarted per block
        return np.concatenate([rng.randint(-50, 50, size=(k,)) + 50. * np.log(1 + np.arange(k)) for k in sizes])
    n = int(np.sum(sizes))
    if kind == "normal":
        return rng.randn(n)
    if kind == "ints":
        return rng.randint(-2, 3, size=n).astype(float)
    if kind == "decreasing":
        return np.concatenate([np.sort(rng.randn(k))[::-1] for k in sizes])
    if kind == "zspace":   # cumulative sums of simplex points plus noise: what the solvers project
        return np.concatenate([np.cumsum(rng.dirichlet(np.ones(k + 1)))[:-1] for k in sizes]) + 0.05 * rng.randn(n)
    raise ValueError(kind)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 7, 8, 15, 16, 19, 20, 31, 32, 33, 63, 64, 65, 96, 100, 255, 256, 257, 513, 1000, 1024, 1025, 4096, 8192])
@pytest.mark.parametrize("kind", ["ref", "normal", "ints", "decreasing", "zspace"])
def test_uniform_blocks(api, K, kind):
    rng = np.random.RandomState(SEED + K)
    nb = max(3, 40000 // K) + 5
    sizes = np.full(nb, K)
    y = make_input(rng, sizes, kind)
    starts = np.arange(0, nb * K, K, dtype=np.int64)
    for update in (1, 0):
        want = y.copy()
        ww = port().pava_multi(want, starts, update=update)
        gy, gw = gpu_pava(api, y, starts, update=update)
        assert np.array_equal(gw, ww)           # pool structure (and stale entries), bit-exact
        assert np.array_equal(gy, want)         # values, bit-exact
    # cold start without a weight array (the configuration of main.py:64): the mask-driven kernels
    want = y.copy()
    port().pava_multi(want, starts)
    gy, _ = gpu_pava(api, y, starts, with_weight=False)
    assert np.array_equal(gy, want)
    port().clip01(want)
    gy, _ = gpu_pava(api, y, starts, with_weight=False, clip=True)
    assert np.array_equal(gy, want)


def power_law_sizes(rng, total, lo, hi, alpha=1.5):
    sizes, left = [], total
    while left > 0:
        k = int(min(hi, max(lo, np.floor(lo * rng.random_sample() ** (-1.0 / alpha)))))
        k = min(k, left)
        sizes.append(k)
        left -= k
    return np.array(sizes, dtype=np.int64)


@pytest.mark.parametrize("lo,hi,total", [(1, 8, 30000), (2, 64, 100000), (2, 256, 150000), (2, 4096, 400000), (200, 600, 100000)])
@pytest.mark.parametrize("first", [0, 5])
def test_ragged_blocks(api, lo, hi, total, first):
    rng = np.random.RandomState(SEED + lo * 7 + hi)
    sizes = power_law_sizes(rng, total, lo, hi) if lo < 200 else rng.randint(lo, hi + 1, size=total // hi)
    starts = first + np.concatenate(([0], np.cumsum(sizes)[:-1]))
    for kind in ("ref", "normal", "ints", "zspace"):
        y = np.concatenate((rng.randn(first), make_input(rng, sizes, kind)))
        want = y.copy()
        ww = port().pava_multi(want, starts)
        gy, gw = gpu_pava(api, y, starts)
        assert np.array_equal(gy[:first], y[:first])
        assert np.array_equal(gw, ww), kind
        assert np.array_equal(gy, want), kind
        # cold start without a weight array: row / word-per-lane kernels
        gy, _ = gpu_pava(api, y, starts, with_weight=False)
        assert np.array_equal(gy, want), kind
        port().clip01(want[first:])
        gy, _ = gpu_pava(api, y, starts, with_weight=False, clip=True)
        assert np.array_equal(gy[first:], want[first:]), kind


def test_worst_case_and_warm_start(api):
    # experiments/PAVA_worst_case.py:30-31: one sweep per element in the reference
    n = 3000
    y = np.arange(n).astype(float)
    y[-1] = -1e12
    want = y.copy()
    ww = port().pava_multi(want, np.array([0]))
    gy, gw = gpu_pava(api, y, [0])
    assert np.array_equal(gy, want) and np.array_equal(gw, ww)
    # warm start: feed the pools of a first solve back in (python/experiments/projection_comparison.py:64-105)
    rng = np.random.RandomState(SEED)
    sizes = power_law_sizes(rng, 50000, 2, 300)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    y0 = make_input(rng, sizes, "zspace")
    a = y0.copy()
    wa = port().pava_multi(a, starts, update=1)
    y1 = a + 0.01 * rng.randn(len(a))
    for s, e in zip(starts, np.append(starts[1:], len(a))):      # keep pool-constant values inside pools
        i = s
        while i < e:
            y1[i:i + wa[i]] = y1[i]
            i += wa[i]
    want = y1.copy()
    w_ref = wa.copy()
    port().pava_multi(want, starts, weight=w_ref, update=1)
    gy, gw = gpu_pava(api, y1, starts, w0=wa)
    assert np.array_equal(gy, want) and np.array_equal(gw, w_ref)


@pytest.mark.parametrize("K", [4, 16, 20, 64, 100, 1000])
def test_fp32_extension(api, K):
    """fp32 twin: values within north_star's 1e-4 of the fp64 oracle on the same (fp32-rounded) input, and isotonic."""
    rng = np.random.RandomState(SEED + K)
    nb = max(3, 200000 // K)
    sizes = np.full(nb, K)
    y = make_input(rng, sizes, "normal").astype(np.float32)
    starts = np.arange(0, nb * K, K, dtype=np.int64)
    want = y.astype(np.float64)
    port().pava_multi(want, starts)
    t = dev(y)
    api.isotonic_regression_multi_c(t, dev(starts))
    got = t.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
    assert (np.diff(got.reshape(nb, K), axis=1) >= 0).all()


def test_fp32_ragged(api):
    rng = np.random.RandomState(SEED)
    sizes = power_law_sizes(rng, 300000, 2, 4096)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    y = make_input(rng, sizes, "normal").astype(np.float32)
    want = y.astype(np.float64)
    port().pava_multi(want, starts)
    t = dev(y)
    api.isotonic_regression_multi_c(t, dev(starts))
    got = t.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("K,nb", [(15, 10 ** 6), (63, 250000)])
def test_full_size_properties(api, K, nb):
    """z-space layouts of BASELINE configs 2/5 at full size: monotone inside blocks, block
    means preserved per pool (pool value = mean of its inputs), idempotent, sample == oracle."""
    g = torch.Generator(device="cuda").manual_seed(SEED + K)
    y = torch.randn(nb * K, dtype=torch.float64, device="cuda", generator=g)
    starts = torch.arange(0, nb * K, K, dtype=torch.int64, device="cuda")
    x = y.clone()
    w = torch.ones(nb * K, dtype=torch.int32, device="cuda")
    api.isotonic_regression_multi_c(x, starts, w)
    X = x.view(nb, K)
    assert (X[:, 1:] >= X[:, :-1]).all()
    assert (X.sum(1) - y.view(nb, K).sum(1)).abs().max().item() < 1e-9      # PAVA preserves block sums
    x2 = x.clone()
    api.isotonic_regression_multi_c(x2, starts)
    assert torch.equal(x2, x)
    sample = torch.arange(0, nb, 1013, device="cuda")
    ys = y.view(nb, K)[sample].reshape(-1).cpu().numpy().copy()
    port().pava_multi(ys, np.arange(0, ys.size, K))
    assert np.array_equal(X[sample].reshape(-1).cpu().numpy(), ys)


# ---------------------------------------------------------------------------------------------
# variants 2 and 3 (isotonic_regression.h:61-82,105-155): values AND weights are the reference's bits
# ---------------------------------------------------------------------------------------------
def test_variants_golden_bit_exact(api, golden_dir):
    d = np.load(os.path.join(golden_dir, "pava.npz"))
    for i in range(int(d["count"])):
        y, starts = d["y%d" % i], d["starts%d" % i]
        t = dev(y)
        api.isotonic_regression_multi_c_2(t, dev(starts))
        assert np.array_equal(t.cpu().numpy(), d["v2_y%d" % i]), i
        h = y.copy()
        api.isotonic_regression_multi_c_2(h, starts)                      # host ABI
        assert np.array_equal(h, d["v2_y%d" % i]), i
        for update in (1, 0):
            t, w = dev(y), dev(np.ones(len(y), dtype=np.int32))
            api.isotonic_regression_multi_c_3(t, dev(starts), w, update)
            assert np.array_equal(t.cpu().numpy(), d["v3_u%d_y%d" % (update, i)]), (i, update)
            assert np.array_equal(w.cpu().numpy(), d["v3_u%d_w%d" % (update, i)]), (i, update)   # tail markers w[k-1] included
            h, hw = y.copy(), np.ones(len(y), dtype=np.int32)
            api.isotonic_regression_multi_c_3(h, starts, hw, update)      # host ABI, weights in / out
            assert np.array_equal(h, d["v3_u%d_y%d" % (update, i)]) and np.array_equal(hw, d["v3_u%d_w%d" % (update, i)])
        t = dev(y)
        api.isotonic_regression_multi_c_3(t, dev(starts))                 # weight=None: ones in, result dropped
        assert np.array_equal(t.cpu().numpy(), d["v3_u1_y%d" % i])


@pytest.mark.parametrize("kind", ["ref", "normal", "ints", "decreasing", "zspace"])
def test_variants_against_oracle_and_reference(api, kind):
    from oracle import cpu
    rng = np.random.RandomState(SEED + 77)
    sizes = np.concatenate((power_law_sizes(rng, 60000, 1, 700), [9000, 1, 2]))      # one block beyond the 8192 window
    first = 3
    starts = first + np.concatenate(([0], np.cumsum(sizes)[:-1]))
    y = np.concatenate((rng.randn(first), make_input(rng, sizes, kind)))
    checkers = [cpu.port()] + ([cpu.ref()] if cpu.ref() is not None else [])
    for chk in checkers:
        want2 = y.copy()
        chk.pava_multi(want2, starts, variant=2)
        t = dev(y)
        api.isotonic_regression_multi_c_2(t, dev(starts))
        assert np.array_equal(t.cpu().numpy(), want2), (kind, chk.kind)
        for update in (1, 0):
            want3 = y.copy()
            w3 = chk.pava_multi(want3, starts, update=update, variant=3)
            t, w = dev(y), dev(np.ones(len(y), dtype=np.int32))
            api.isotonic_regression_multi_c_3(t, dev(starts), w, update)
            assert np.array_equal(t.cpu().numpy(), want3), (kind, chk.kind, update)
            assert np.array_equal(w.cpu().numpy(), w3), (kind, chk.kind, update)
    # single-block entry points
    s, e = int(starts[5]), int(starts[9])
    for variant, fn in ((2, api.isotonic_regression_c_2), (3, api.isotonic_regression_c_3)):
        want = y.copy()
        cpu.port().pava(want, s, e, variant=variant)
        t = dev(y)
        fn(t, s, e)
        assert np.array_equal(t.cpu().numpy(), want)


# ---------------------------------------------------------------------------------------------
# blocks beyond the 8192-entry shared-memory window: served by the sequential routine, any dtype
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["uniform10000", "ragged", "single"])
def test_long_blocks(api, layout):
    rng = np.random.RandomState(SEED + 5)
    if layout == "uniform10000":
        sizes = np.full(7, 10000)
    elif layout == "ragged":
        sizes = np.concatenate((power_law_sizes(rng, 30000, 2, 3000), [20000], power_law_sizes(rng, 5000, 2, 40), [8193, 8192, 3]))
    else:
        sizes = np.array([150000])
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    for kind in ("ref", "normal", "zspace"):
        y = make_input(rng, sizes, kind)
        want = y.copy()
        ww = port().pava_multi(want, starts)
        gy, gw = gpu_pava(api, y, starts)
        assert np.array_equal(gw, ww) and np.array_equal(gy, want), (layout, kind)
        gy, _ = gpu_pava(api, y, starts, with_weight=False, clip=True)
        port().clip01(want)
        assert np.array_equal(gy, want), (layout, kind)
    # fp32 twin: no length limit either
    y = make_input(rng, sizes, "normal").astype(np.float32)
    want = y.astype(np.float64)
    port().pava_multi(want, starts)
    t = dev(y)
    api.isotonic_regression_multi_c(t, dev(starts))
    assert np.abs(t.cpu().numpy().astype(np.float64) - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
    # host ABI, single block (the reference's stress shape, test_stress_isotonic_regression.py:59)
    h = make_input(rng, sizes, "ref")
    want = h.copy()
    port().pava_multi(want, starts)
    api.isotonic_regression_multi_c(h, starts)
    assert np.array_equal(h, want)


def test_config3_full_size_with_weights(api):
    """BASELINE config 3 at full size (power-law blocks 2..4096, 10^7 values): values and the weight array
    bit-identical to the oracle (multi-threaded port), cold call and [0,1] clamp too."""
    rng = np.random.RandomState(SEED + 2)
    sizes = []
    left = 10 ** 7
    while left > 0:
        k = np.floor(2 * rng.rand(65536) ** (-1.0 / 1.5)).clip(2, 4096).astype(np.int64)
        c = np.cumsum(k)
        cut = int(np.searchsorted(c, left, side="left"))
        if cut >= len(k):
            sizes.append(k)
            left -= int(c[-1])
            continue
        take = k[:cut + 1].copy()
        take[-1] -= int(c[cut] - left)
        if take[-1] < 1:
            take = take[:-1]
        sizes.append(take)
        left = 0
    sizes = np.concatenate(sizes)
    n = int(sizes.sum())
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    pos = np.arange(n) - np.repeat(starts, sizes)
    y = rng.randint(-50, 50, size=n) + 50.0 * np.log1p(pos)            # tests/fast/test_isotonic_regression.py:43 per block
    want = y.copy()
    ww = port().pava_multi(want, starts, threads=port().max_threads())
    gy, gw = gpu_pava(api, y, starts)
    assert np.array_equal(gw, ww)
    assert np.array_equal(gy, want)
    gy, _ = gpu_pava(api, y, starts, with_weight=False)
    assert np.array_equal(gy, want)
    z = rng.randn(n) * 0.3 + 0.5
    want = z.copy()
    port().pava_multi(want, starts, threads=port().max_threads())
    port().clip01(want)
    gy, _ = gpu_pava(api, z, starts, with_weight=False, clip=True)
    assert np.array_equal(gy, want)
