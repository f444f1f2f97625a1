"""Pins oracle/solvers_np.py (the CPU restatement of the reference's solver drivers) against
tests/golden/solvers.npz, which tests/golden/make_golden.py produced by running the
reference's own BATCH / BB / LBFGS / DORE / mirror_descent modules."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import solvers_np as S

TAGS = ("c1mini", "k16", "noisy", "noisy20")
SHAPES = {"c1mini": (60, 5), "k16": (20, 16), "noisy": (80, 5), "noisy20": (30, 20)}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "solvers.npz"))


def problem(gold, tag):
    m, n = gold[tag + "_shape"]
    A = sps.csr_matrix((gold[tag + "_val"], gold[tag + "_idx"], gold[tag + "_ptr"]), shape=(m, n))
    return A, gold[tag + "_b"], gold[tag + "_starts"], gold[tag + "_xinit"]


@pytest.mark.parametrize("tag", TAGS)
def test_objective_and_gradient(gold, tag):
    A, b, starts, x0 = problem(gold, tag)
    _, _, _, obj = S.get_solver_parts(A, b, starts, 0.1)
    g = np.zeros_like(x0)
    f = obj(x0, g)
    assert f == pytest.approx(float(gold[tag + "_f0"]), rel=1e-14)
    np.testing.assert_allclose(g, gold[tag + "_g0"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name", ["bb", "pg", "md", "lbfgs"])
def test_batch_solvers_replay_reference(gold, tag, name):
    A, b, starts, x0 = problem(gold, tag)
    step_size, proj, line_search, obj = S.get_solver_parts(A, b, starts, 0.1)
    if name == "bb":
        sol = S.solve_BB(obj, proj, line_search, x0, max_iter=300)
    elif name == "pg":
        sol = S.solve(obj, proj, step_size, x0, line_search, max_iter=100)
    elif name == "md":
        sol = S.solve_MD(obj, starts, step_size, x0, max_iter=100)
    else:
        sol = S.solve_LBFGS(obj, proj, line_search, x0, max_iter=150)
    key = "%s_%s_" % (tag, name)
    assert sol["iterations"] == int(gold[key + "iters"])
    assert sol["stop"].split("=")[0] == str(gold[key + "stop"]).split("=")[0]
    trace = np.array([p[1] for p in sol["progress"]])
    np.testing.assert_allclose(trace, gold[key + "ftrace"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(sol["x"], gold[key + "x"], rtol=0, atol=1e-9)
    assert sol["f"] == pytest.approx(float(gold[key + "f"]), rel=1e-9, abs=1e-14)


@pytest.mark.parametrize("tag", TAGS)
def test_z_space_drivers_replay_reference(gold, tag):
    A, b, starts, x0 = problem(gold, tag)
    nb, K = SHAPES[tag]
    sizes = np.full(nb, K, dtype=np.int64)
    N, xpart, target, f, nabla_f, proj, zstarts = S.z_space_closures(A, b, sizes)
    z0 = gold[tag + "_z0"]
    opts = {"max_iter": 200, "opt_tol": 1e-30, "verbose": 0}
    z = S.bb_solve(z0.copy(), f, nabla_f, proj=proj, options=opts)
    assert f(z) == pytest.approx(float(gold[tag + "_zbb_f"]), rel=1e-6, abs=1e-12)
    zl = S.lbfgs_solve(z0.copy() + 1, f, nabla_f, proj=proj, options={"max_iter": 40, "opt_tol": 1e-30, "verbose": 0})
    assert f(zl) == pytest.approx(float(gold[tag + "_zlbfgs_f"]), rel=1e-6, abs=1e-12)
    lsv = float(gold[tag + "_lsv"])
    A_d = A * 0.99 / lsv
    t_d = target * 0.99 / lsv
    zd = S.dore_solve(z0.copy(), lambda v: A_d.dot(N.dot(v)), lambda r: N.T.dot(A_d.T.dot(r)), t_d, proj=proj,
                      options={"max_iter": 150, "opt_tol": 1e-30, "verbose": 0})
    assert f(zd) == pytest.approx(float(gold[tag + "_zdore_f"]), rel=1e-6, abs=1e-12)
    np.testing.assert_allclose(zd, gold[tag + "_zdore_z"], atol=1e-7)


@pytest.mark.parametrize("tag", TAGS)
def test_mirror_descent_least_squares_replays_reference(gold, tag):
    A, b, starts, x0 = problem(gold, tag)
    nb, K = SHAPES[tag]
    x = S.md_least_squares(A, b, [K] * nb, iters=60, tolerance=1e-9, Lf=float(gold[tag + "_Lf"]))
    np.testing.assert_allclose(x, gold[tag + "_md_ls_x"], rtol=1e-9, atol=1e-12)
