"""The steps either side of the hot loop on the GPU against fixtures produced by the
reference's own modules (tests/golden/make_golden.py: golden_pipeline): BSLSMatrices.degree_reduced_form / get_LS /
reconstruct, main.solve_in_z + LS_postprocess on bsls_utils.generate_data problems, a .mat file end to end, and
line_search_exact_quad_obj.  The invariants are those of the reference's tests/fast/test_bsls_matrices.py and
tests/fast/test_main.py."""
import argparse
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CONFIG = {'full': True, 'L': True, 'OD': True, 'CP': True, 'LP': True, 'eq': 'CP', 'init': False}
EPS = 1e-10


@pytest.fixture(scope="module")
def B():
    import __graft_entry__ as g
    g.build()
    import bsls_b200
    return bsls_b200


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "pipeline.npz"))


def host(t):
    return t.detach().cpu().numpy()


def data_of(gold, tag):
    d = {k: gold["%s_in_%s" % (tag, k)] for k in ("A", "b", "x_true", "U", "f")}
    d["block_sizes"] = gold[tag + "_in_block_sizes"]
    return d


TAGS = ["default", "sparse_x", "sparse_A", "permuted", "larger", "zeroed"]


@pytest.mark.parametrize("tag", TAGS)
def test_degree_reduced_form_matches_reference(B, gold, tag):
    data = data_of(gold, tag)
    bm = B.bsls_matrices.BSLSMatrices(data=data, **CONFIG)
    bm.degree_reduced_form()
    AA, bb, N, block_sizes, x_split, nz, scaling, rsort_index, x0 = bm.get_LS()
    n_raw = data["x_true"].size
    # layout facts: bit-exact
    assert np.array_equal(block_sizes, gold[tag + "_block_sizes"])
    assert np.array_equal(host(nz), gold[tag + "_nz"])
    assert np.array_equal(host(scaling), gold[tag + "_scaling"])
    assert np.array_equal(host(bb), gold[tag + "_bb"])
    assert np.array_equal(host(x0), gold[tag + "_x0"])
    # the order of the columns INSIDE a block is unspecified in the reference (unstable argsort): compare what does not
    # depend on it -- the matrix and the splits mapped back to the original column order
    rs, rs_ref = host(rsort_index), gold[tag + "_rsort_index"]
    dense = host(AA.todense())
    assert np.array_equal(dense[:, rs], gold[tag + "_AA"][:, rs_ref])
    assert np.array_equal(host(x_split)[rs], gold[tag + "_x_split"][rs_ref])
    rec = B.bsls_matrices.BSLSMatrices.reconstruct(x_split, rsort_index=rsort_index, scaling=scaling, nz=nz, n=n_raw)
    assert np.array_equal(host(rec), gold[tag + "_x_true_rec"])
    assert np.linalg.norm(host(rec) - data["x_true"]) < EPS                        # test_bsls_matrices.py:88-90
    # same blocks, as sets of original columns
    cum = np.concatenate(([0], np.cumsum(block_sizes)))
    inv, inv_ref = np.argsort(rs), np.argsort(rs_ref)                               # sorted position -> nz position
    for s, e in zip(cum[:-1], cum[1:]):
        assert set(inv[s:e]) == set(inv_ref[s:e])
    # the reference's own invariants (tests/fast/test_bsls_matrices.py:47-62,100-108)
    xs = host(x_split)
    C = host(bm.C.todense())
    assert np.linalg.norm(C.dot(xs) - 1.0) < EPS
    assert np.linalg.norm(dense.dot(xs) - host(bb)) < EPS
    assert np.all(xs >= 0)
    assert np.array_equal(np.nonzero(C)[1], np.arange(xs.size))
    assert abs(xs.sum() - len(block_sizes)) < EPS
    for s, e in zip(cum[:-1], cum[1:]):
        assert abs(xs[s:e].sum() - 1) < EPS
    # the operator N against its definition: x = x0 + N z reproduces x_split from z = x2z(x_split)
    z = B.bsls_utils.x2z(x_split, block_sizes=block_sizes)
    back = N.dot(z)
    back += x0
    np.testing.assert_allclose(host(back), xs, rtol=0, atol=1e-14)


def test_each_step_keeps_feasibility(B, gold):
    """tests/fast/test_bsls_matrices.py:25-52: C x = d and AA x = bb after consolidate, then C x_split = 1 and
    AA x_split = bb after every further step."""
    data = data_of(gold, "default")
    bm = B.bsls_matrices.BSLSMatrices(data=data, **CONFIG)
    bm.consolidate(eq="CP")
    xt = host(bm.x_true)
    assert np.linalg.norm(host(bm.C.todense()).dot(xt) - host(bm.d)) < EPS
    assert np.linalg.norm(host(bm.AA.todense()).dot(xt) - host(bm.bb)) < EPS
    for step in (bm.standard_simplex_form, bm.cleanup, bm.blockify):
        step()
        xs = host(bm.x_split)
        assert np.linalg.norm(host(bm.C.todense()).dot(xs) - 1.0) < EPS
        assert np.linalg.norm(host(bm.AA.todense()).dot(xs) - host(bm.bb)) < EPS
        assert np.all(xs >= 0)


@pytest.mark.parametrize("tag", ["default", "sparse_x", "zeroed", "larger"])
def test_solve_and_postprocess_match_reference(B, gold, tag):
    data = data_of(gold, tag)
    bm = B.bsls_matrices.BSLSMatrices(data=data, **CONFIG)
    bm.degree_reduced_form()
    AA, bb, N, block_sizes, x_split, nz, scaling, rsort_index, x0 = bm.get_LS()
    problem = bm.problem()
    iters, times, states = B.main.solve_in_z(problem, bb, x0, N, block_sizes, 'BB')
    x_last, error, out = B.main.LS_postprocess(states, x0, problem, bb, x_split, scaling=scaling, block_sizes=block_sizes, N=N,
                                               output={})
    assert error[-1] < 1e-16                                                       # tests/fast/test_main.py:35
    assert iters[0] == 0 and len(error) == len(states)
    # everything that does not depend on the chaotic BB trajectory agrees with the reference's LS_postprocess
    assert out['0.5norm(Ax_init-b)^2'] == pytest.approx(float(gold[tag + "_start_error"]), rel=1e-12)
    assert out['0.5norm(Ax*-b)^2'] == pytest.approx(float(gold[tag + "_opt_error"]), abs=1e-20)
    assert out['max|f * (x_init-x_true)|'] == pytest.approx(float(gold[tag + "_start_dist"]), rel=1e-12)
    assert error[0] == pytest.approx(float(gold[tag + "_error"][0]), rel=1e-12)     # state 0 is z0 itself
    assert out['max|f * (x-x_true)|'][0] == pytest.approx(float(gold[tag + "_max_f_diff"][0]), rel=1e-12)
    assert out['incorrect x entries'][0] == int(gold[tag + "_wrong"][0])
    assert out['percent flow allocated incorrectly'][0] == pytest.approx(float(gold[tag + "_per_flow"][0]), rel=1e-12)
    # the solution is feasible and, mapped back, a non-negative flow with the right block totals
    xl = host(x_last)
    assert xl.min() >= -1e-12
    rec = host(B.bsls_matrices.BSLSMatrices.reconstruct(x_last, rsort_index=rsort_index, scaling=scaling, nz=nz, n=data["x_true"].size))
    np.testing.assert_allclose(data["U"].dot(rec), data["f"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(data["A"].dot(rec)[np.abs(data["A"]).sum(1) > 0], data["b"][np.abs(data["A"]).sum(1) > 0], rtol=1e-6, atol=1e-6)


def test_mat_file_end_to_end(B, gold, tmp_path):
    """tests/fast/test_main.py: generate_data -> .mat -> main() with BB in z converges to 0.5 |Ax - b|^2 < 1e-16.  The file
    goes through BSLSMatrices on the device: no NumPy / scipy preprocessing between loadmat and the solver."""
    import scipy.io
    data = data_of(gold, "default")
    fname = str(tmp_path / "test_main.mat")
    scipy.io.savemat(fname, data, oned_as='column')
    args = argparse.Namespace(noise=0, file=fname, log='WARN', init=False, eq='CP', method='BB')
    iters, times, states, output = B.main.main(args=args)
    assert output['0.5norm(Ax-b)^2'][-1] < 1e-16
    assert output['nLinks'] == 5 and output['nCP'] == 10
    # DORE runs through the same file -> device path (the reference pins no result for it on this input; its own test,
    # tests/fast/test_main.py, uses BB only -- and the weak-Wolfe search of LBFGS.solve doubles its step without bound on
    # this badly scaled problem, in the reference as here)
    for method in ("DORE",):
        args.method = method
        args.options = {'max_iter': 200, 'verbose': 0, 'opt_tol': 1e-30}
        iters, times, states, output = B.main.main(args=args)
        assert np.all(np.isfinite(output['0.5norm(Ax-b)^2']))
        assert output['0.5norm(Ax-b)^2'][-1] <= output['0.5norm(Ax_init-b)^2'] * (1 + 1e-12)


def test_flow_metrics_kernel(B):
    rng = np.random.RandomState(3)
    for n in (1, 17, 1000, 123457):
        s, xt, xh = rng.rand(n) * 100, rng.rand(n), rng.rand(n)
        ws = B.sparse.default_workspace("cuda")
        dev = lambda a: torch.from_numpy(a).cuda()
        got = ws.flow_metrics(dev(s), dev(xt), dev(xh), 1e-3)
        d = xt - xh
        want = [np.abs(s * d).sum(), (s * xt).sum(), float((d > 1e-3).sum()), d.dot(d), max(0.0, (s * d).max())]
        np.testing.assert_allclose(got, want, rtol=1e-11)
        got1 = ws.flow_metrics(None, dev(xt), dev(xh), 1e-3)
        np.testing.assert_allclose(got1[0], np.abs(d).sum(), rtol=1e-11)


def test_line_search_exact_quad_obj(B, gold):
    """algorithm_utils.py:140-155 against the reference's own outputs (dense QPs; the last case takes the
    step-too-small branch)."""
    au = B.algorithm_utils
    for k in range(int(gold["ls_count"])):
        g = lambda name: gold["ls%d_%s" % (k, name)]
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        x, x_new, grad = dev(g("x")), dev(g("x_new_in")), dev(g("g"))
        g_new = torch.zeros_like(x)
        f_new = au.line_search_exact_quad_obj(x, float(g("f")), grad, x_new, 0.0, g_new, g("Q"), g("c"))
        assert f_new == pytest.approx(float(g("f_new")), rel=1e-10, abs=1e-12)
        np.testing.assert_allclose(host(x_new), g("x_new_out"), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(host(g_new), g("g_new_out"), rtol=1e-8, atol=1e-9)
