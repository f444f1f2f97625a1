"""Pins the CPU oracle (oracle/) before anything is allowed to trust it.

Three anchors, strongest first:
  1. known answers written in the reference's own tests
     (tests/fast/test_proj_simplex.py:24-52,77-81; tests/fast/test_c_extensions.py:67-79;
      python/c_extensions/isotonic_regression.h:180,187,194; SURVEY section 7 tie example),
  2. fixtures produced by the reference's own Cython build (tests/golden/make_golden.py),
  3. the reference's C++ headers compiled into oracle/_ref (when present) on seeded input.
Equality is exact (bit-for-bit) everywhere the arithmetic order is the reference's own.
"""
import os

import numpy as np
import pytest

from oracle import cpu

SEED = 237423433


@pytest.fixture(scope="module")
def port():
    return cpu.port()


def checkers():
    out = [cpu.port()]
    if cpu.ref() is not None:
        out.append(cpu.ref())
    return out


# ---------------------------------------------------------------- 1. reference known answers
Z = np.array([5.352, 3.23, 32.78, -1.234, 1.7, 104., 53.])


@pytest.mark.parametrize("chk", checkers(), ids=lambda c: c.kind)
def test_single_block_known_answers(chk):
    for truth, start, end in zip([[5.352, 3.23, 1., 0., 1.7, 104., 53.], [0., 0., 0., 0., 0, 1., 0.]], [2, 0], [4, 7]):
        y = Z.copy()
        chk.proj_simplex(y, start, end)
        assert (y == np.array(truth)).all()
    np.random.seed(SEED)
    y = np.random.rand(7)
    chk.proj_simplex(y, 0, 7)
    truth = np.array([0., .05006376, .54108944, 0., .38841272, 0., .02043408])
    assert np.linalg.norm(y - truth) < 1e-6


@pytest.mark.parametrize("chk", checkers(), ids=lambda c: c.kind)
def test_multi_block_known_answers(chk):
    for blocks, truth in [([0, 2, 4], [1., 0., 1., 0., 0., 1., 0.]), ([0], [0., 0., 0., 0., 0., 1., 0.]),
                          ([0, 3], [0., 0., 1., 0., 0., 1., 0.])]:
        y = Z.copy()
        chk.proj_multi_simplex(y, np.array(blocks))
        assert (y == np.array(truth)).all()
    y = np.array([0.234, 0.5, 1.3, -1.234, 1.7, -1.0, 53.])
    chk.proj_multi_ball(y, np.array([0, 2, 4]))
    assert (y == np.array([0.234, 0.5, 1., 0., 0., 0., 1.])).all()
    y = np.array([0.234, 0.5, 1.3, -1.234, 1.7, 104., 53.])     # main.cpp / proj_simplex.h:88-90
    chk.proj_multi_ball(y, np.array([0, 2, 4]))
    assert (y == np.array([0.234, 0.5, 1., 0., 0., 1., 0.])).all()


@pytest.mark.parametrize("chk", checkers(), ids=lambda c: c.kind)
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_pava_known_answers(chk, variant):
    y = np.array([4., 5., 1., 6., 8., 7.])
    chk.pava(y, 0, 6, variant=variant)
    assert np.allclose(y, [10 / 3., 10 / 3., 10 / 3., 6., 7.5, 7.5], rtol=0, atol=1e-15)
    y = np.array([4., 5., 1., 6., 8., 7.])
    chk.pava_multi(y, np.array([0, 2, 4]), variant=variant)
    assert (y == np.array([4., 5., 1., 6., 7.5, 7.5])).all()


@pytest.mark.parametrize("chk", checkers(), ids=lambda c: c.kind)
def test_pava_pool_structure_known(chk):
    y = np.array([4., 5., 1., 6., 8., 7.])
    w = chk.pava(y, 0, 6, variant=1)
    assert list(w) == [3, 2, 1, 1, 2, 1]                      # SURVEY section 7 probe
    y = np.array([1., 3., 1., 2.])
    w1 = chk.pava(y.copy(), 0, 4, variant=1)
    w3 = chk.pava(y.copy(), 0, 4, variant=3)
    assert list(w1) == [1, 2, 1, 1] and list(w3) == [1, 2, 2, 1]
    for w in (w1, w3):
        assert list(cpu.pools_from_weights(w, [0], 4)) == [0, 1, 3]   # equal means stay separate


def test_x2z_known_answers(port):
    xs = [[.6, .1, .3], [.5, .5, .2, .8], [1., .6, .1, .3]]
    zs = [[.6, .7], [.5, .2], [.6, .7]]
    bs = [[0], [0, 2], [0, 1]]
    for x_true, z_true, b in zip(xs, zs, bs):
        z = port.x2z(np.array(x_true), np.zeros(len(z_true)), np.array(b))
        assert np.allclose(z, z_true, atol=1e-12)
        x = port.z2x(np.zeros(len(x_true)), z, np.array(b))
        assert np.allclose(x, x_true, atol=1e-12)


# ---------------------------------------------------------------- 2. fixtures from the reference
def test_projection_fixtures(port, golden_dir):
    d = np.load(os.path.join(golden_dir, "projection.npz"))
    for i in range(int(d["count"])):
        y, starts = d["y%d" % i], d["starts%d" % i]
        for chk in checkers():
            out = y.copy()
            chk.proj_multi_simplex(out, starts)
            assert np.array_equal(out, d["simplex%d" % i]), (i, chk.kind)
            out = y.copy()
            chk.proj_multi_ball(out, starts)
            assert np.array_equal(out, d["ball%d" % i]), (i, chk.kind)
        out = y.copy()
        port.proj_multi_simplex(out, starts, threads=2)
        assert np.array_equal(out, d["simplex%d" % i])


def test_pava_fixtures(port, golden_dir):
    d = np.load(os.path.join(golden_dir, "pava.npz"))
    for i in range(int(d["count"])):
        y, starts = d["y%d" % i], d["starts%d" % i]
        for chk in checkers():
            for variant, tag in ((1, "v1"), (3, "v3")):
                for update in (1, 0):
                    out = y.copy()
                    w = chk.pava_multi(out, starts, update=update, variant=variant)
                    assert np.array_equal(out, d["%s_u%d_y%d" % (tag, update, i)]), (i, tag, update, chk.kind)
                    assert np.array_equal(w, d["%s_u%d_w%d" % (tag, update, i)]), (i, tag, update, chk.kind)
            out = y.copy()
            chk.pava_multi(out, starts, variant=2)
            assert np.array_equal(out, d["v2_y%d" % i])
        # the three variants agree on values (1e-8, the reference's own tolerance) and on pools
        v1, v3, v2 = d["v1_u1_y%d" % i], d["v3_u1_y%d" % i], d["v2_y%d" % i]
        scale = max(1.0, np.abs(v1).max())
        assert np.abs(v1 - v3).max() <= 1e-8 * scale and np.abs(v1 - v2).max() <= 1e-8 * scale
        p1 = cpu.pools_from_weights(d["v1_u1_w%d" % i], starts, len(y))
        p3 = cpu.pools_from_weights(d["v3_u1_w%d" % i], starts, len(y))
        if len(np.unique(y)) == len(y):
            assert np.array_equal(p1, p3)
        else:
            # tie-heavy input: variant 3 back-tracks on ">=" and fuses equal-valued pools that
            # variant 1 (the canonical one, bound by isotonic_regression_multi_c) keeps apart,
            # e.g. [0, 2, -2] -> v1 pools {0},{1,2}; v3 pool {0,1,2}.  v3's heads are a subset.
            assert set(p3.tolist()) <= set(p1.tolist())


def test_pava_matches_sklearn(port):
    """The reference's own differential test (tests/fast/test_isotonic_regression.py:49-115)."""
    from sklearn.isotonic import IsotonicRegression
    from sklearn.utils import check_random_state
    np.random.seed(SEED)
    rs = check_random_state(0)
    for n in (10, 100):
        for _ in range(10):
            y = rs.randint(-50, 50, size=(n,)) + 50. * np.log(1 + np.arange(n))
            blocks = np.sort(np.random.choice(n, 3, replace=False))
            truth = y.copy()
            ends = np.append(blocks[1:], n)
            for s, e in zip(blocks, ends):
                truth[s:e] = IsotonicRegression().fit_transform(np.arange(s, e), y[s:e])
            for variant in (1, 2, 3):
                out = y.copy()
                port.pava_multi(out, blocks, variant=variant)
                assert np.linalg.norm(out - truth) < 1e-8


def test_x2z_fixtures(port, golden_dir):
    d = np.load(os.path.join(golden_dir, "x2z.npz"))
    for i in range(int(d["count"])):
        x, starts = d["x%d" % i], d["starts%d" % i]
        z = port.x2z(x, np.zeros_like(d["z%d" % i]), starts)
        assert np.array_equal(z, d["z%d" % i])
        xb = port.z2x(np.zeros_like(x), z, starts)
        assert np.array_equal(xb, d["xback%d" % i])


def test_sparse_objective_fixture(port, golden_dir):
    import scipy.sparse as sps
    d = np.load(os.path.join(golden_dir, "solvers.npz"))
    for tag in ("c1mini", "k16", "noisy", "noisy20"):
        m, n = d[tag + "_shape"]
        A = sps.csr_matrix((d[tag + "_val"], d[tag + "_idx"], d[tag + "_ptr"]), shape=(m, n))
        AT = A.T.tocsr()
        g = np.zeros(n)
        f, _ = port.lsq_obj((A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data),
                            (AT.indptr.astype(np.int64), AT.indices.astype(np.int32), AT.data),
                            d[tag + "_xinit"], d[tag + "_b"], g)
        assert abs(f - float(d[tag + "_f0"])) <= 1e-12 * abs(float(d[tag + "_f0"]))
        assert np.array_equal(g, d[tag + "_g0"])


# ---------------------------------------------------------------- 3. compiled reference, seeded
@pytest.mark.skipif(cpu.ref() is None, reason="oracle/_ref not built (no /root/reference on this box)")
def test_port_equals_compiled_reference_random():
    rng = np.random.RandomState(SEED)
    ref, port = cpu.ref(), cpu.port()
    for trial in range(30):
        n = int(rng.randint(5, 4000))
        cuts = np.unique(rng.randint(0, n, size=rng.randint(1, 200)))
        y = rng.randn(n) * rng.choice([0.01, 1.0, 50.0])
        a, b = y.copy(), y.copy()
        ref.proj_multi_simplex(a, cuts)
        port.proj_multi_simplex(b, cuts)
        assert np.array_equal(a, b)
        a, b = y.copy(), y.copy()
        ref.proj_multi_ball(a, cuts)
        port.proj_multi_ball(b, cuts)
        assert np.array_equal(a, b)
        yi = np.round(y * 2) / 2 if trial % 3 == 0 else y
        for variant in (1, 2, 3):
            for update in (0, 1):
                a, b = yi.copy(), yi.copy()
                wa = ref.pava_multi(a, cuts, update=update, variant=variant)
                wb = port.pava_multi(b, cuts, update=update, variant=variant)
                assert np.array_equal(a, b)
                if variant != 2:
                    assert np.array_equal(wa, wb)
