#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference itself.

Runs ONLY in the build container (needs /root/reference).  The reference is
Python 2 + an old Cython file, so a scratch copy is made under /tmp and given the
purely mechanical py2->py3 / NumPy-2 edits listed in SURVEY.md section 8c (print
statements, xrange, ``import ipdb``, ``np.float``, ``np.int_t``); no arithmetic is
touched.  The reference's own Cython extension is then built in the scratch copy
and its public functions are called on seeded inputs.  Inputs and outputs are
stored as small .npz files which the CPU and GPU test-suites replay.

    python tests/golden/make_golden.py

Nothing under /root/reference is modified and no reference source enters the repo.
"""
import os
import re
import shutil
import subprocess
import sys

import numpy as np

REF = "/root/reference"
SCRATCH = "/tmp/bsls_ref_py3"
HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 237423433  # the seed every reference test uses (tests/fast/*.py setUp)


def patch_source(text, name):
    out = []
    for line in text.split("\n"):
        m = re.match(r"^(\s*)print (?!\()(.*)$", line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        if re.match(r"^\s*import ipdb\s*$", line):
            line = re.sub(r"import ipdb", "pass", line)
        out.append(line)
    text = "\n".join(out)
    text = text.replace("xrange", "range")
    text = text.replace("np.float)", "float)")
    if name == "algorithm_utils.py":
        text = text.replace("j=range(n+1)", "j=list(range(n+1))")
    if name == "bsls_utils.py":
        text = text.replace("block_sizes = np.random.multinomial(n-m2,np.ones(m2)/m2) + np.ones(m2)",
                            "block_sizes = (np.random.multinomial(n-m2,np.ones(m2)/m2) + np.ones(m2)).astype(int)")
    if name == "c_extensions.pyx":
        text = text.replace("np.int_t", "np.int64_t")
    return text


def prepare():
    if os.path.exists(SCRATCH):
        shutil.rmtree(SCRATCH)
    shutil.copytree(os.path.join(REF, "python"), os.path.join(SCRATCH, "python"))
    for root, _, files in os.walk(os.path.join(SCRATCH, "python")):
        for f in files:
            if f.endswith((".py", ".pyx")):
                p = os.path.join(root, f)
                with open(p) as fh:
                    src = fh.read()
                with open(p, "w") as fh:
                    fh.write(patch_source(src, f))
    ext = os.path.join(SCRATCH, "python", "c_extensions")
    with open(os.path.join(ext, "setup_py3.py"), "w") as fh:
        fh.write(
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "import numpy\n"
            "setup(ext_modules=cythonize([Extension('c_extensions', ['c_extensions.pyx'], language='c++',\n"
            "      include_dirs=[numpy.get_include()], extra_compile_args=['-O2', '-w'])],\n"
            "      compiler_directives={'language_level': 2}))\n")
    subprocess.check_call([sys.executable, "setup_py3.py", "build_ext", "--inplace"], cwd=ext,
                          stdout=subprocess.DEVNULL)
    sys.path.insert(0, os.path.join(SCRATCH, "python"))


def power_law_sizes(rng, total, lo=2, hi=64, alpha=1.5):
    sizes = []
    left = total
    while left > 0:
        k = int(min(hi, max(lo, np.floor(lo * rng.random_sample() ** (-1.0 / alpha)))))
        k = min(k, left)
        sizes.append(k)
        left -= k
    return np.array(sizes, dtype=np.int64)


def golden_projection(cx):
    """proj_simplex_c / proj_multi_simplex_c / proj_multi_ball_c on seeded inputs."""
    rng = np.random.RandomState(SEED)
    cases = {}
    idx = 0
    for nb, K in [(64, 4), (64, 5), (40, 16), (30, 20), (12, 64), (7, 33), (3, 300)]:
        for dist in ("normal", "uniform", "near_simplex"):
            n = nb * K
            if dist == "normal":
                y = rng.randn(n)
            elif dist == "uniform":
                y = rng.rand(n)
            else:
                y = rng.dirichlet(np.ones(K), size=nb).reshape(-1) + 1e-3 * rng.randn(n)
            starts = np.arange(0, n, K, dtype=np.int64)
            out = y.copy()
            cx.proj_multi_simplex_c(out, starts)
            ball = y.copy()
            cx.proj_multi_ball_c(ball, starts)
            cases["y%d" % idx] = y
            cases["starts%d" % idx] = starts
            cases["simplex%d" % idx] = out
            cases["ball%d" % idx] = ball
            idx += 1
    # ragged layouts, including a prefix that must stay untouched and size-1 blocks
    for total, hi in [(500, 16), (2000, 64), (3000, 700)]:
        sizes = power_law_sizes(rng, total, 1, hi)
        starts = np.concatenate(([0], np.cumsum(sizes)[:-1])) + 3
        n = int(total + 3)
        y = rng.randn(n)
        out = y.copy()
        cx.proj_multi_simplex_c(out, starts)
        ball = y.copy()
        cx.proj_multi_ball_c(ball, starts)
        cases["y%d" % idx] = y
        cases["starts%d" % idx] = starts
        cases["simplex%d" % idx] = out
        cases["ball%d" % idx] = ball
        idx += 1
    # ties / integer data: every sum is exact, the active set is decided by exact comparisons
    y = rng.randint(-3, 4, size=640).astype(float) / 2.0
    starts = np.arange(0, 640, 8, dtype=np.int64)
    out = y.copy()
    cx.proj_multi_simplex_c(out, starts)
    ball = y.copy()
    cx.proj_multi_ball_c(ball, starts)
    cases["y%d" % idx] = y
    cases["starts%d" % idx] = starts
    cases["simplex%d" % idx] = out
    cases["ball%d" % idx] = ball
    idx += 1
    cases["count"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "projection.npz"), **cases)
    return idx


def pava_input(rng, n, kind):
    if kind == "ref":      # tests/fast/test_isotonic_regression.py:43
        return rng.randint(-50, 50, size=(n,)) + 50. * np.log(1 + np.arange(n))
    if kind == "normal":
        return rng.randn(n)
    if kind == "ints":     # heavy ties
        return rng.randint(-2, 3, size=(n,)).astype(float)
    raise ValueError(kind)


def golden_pava(cx):
    """All three PAVA variants, values + (variant 1 and 3) the weight arrays."""
    rng = np.random.RandomState(SEED + 1)
    cases = {}
    idx = 0
    layouts = []
    for nb, K in [(50, 3), (40, 15), (25, 19), (10, 63), (4, 257)]:
        layouts.append((np.arange(0, nb * K, K, dtype=np.int64), nb * K))
    for total, hi in [(400, 12), (1500, 90), (2500, 900)]:
        sizes = power_law_sizes(rng, total, 1, hi)
        layouts.append((np.concatenate(([0], np.cumsum(sizes)[:-1])) + 2, total + 2))
    for starts, n in layouts:
        for kind in ("ref", "normal", "ints"):
            y = pava_input(rng, n, kind)
            if kind == "ref":
                # restart the log ramp inside every block as SURVEY 8d specifies
                ends = np.append(starts[1:], n)
                for s, e in zip(starts, ends):
                    y[s:e] = rng.randint(-50, 50, size=(e - s,)) + 50. * np.log(1 + np.arange(e - s))
            cases["y%d" % idx] = y
            cases["starts%d" % idx] = starts
            for variant, fn in (("v1", cx.isotonic_regression_multi_c), ("v3", cx.isotonic_regression_multi_c_3)):
                for update in (1, 0):
                    out = y.copy()
                    w = np.ones(n, dtype=np.int32)
                    fn(out, starts, w, update)
                    cases["%s_u%d_y%d" % (variant, update, idx)] = out
                    cases["%s_u%d_w%d" % (variant, update, idx)] = w
            out = y.copy()
            cx.isotonic_regression_multi_c_2(out, starts)
            cases["v2_y%d" % idx] = out
            idx += 1
    # adversarial: arange with a huge drop at the end (experiments/PAVA_worst_case.py:30-31)
    n = 200
    y = np.arange(n).astype(float)
    y[-1] = -1e12
    starts = np.array([0], dtype=np.int64)
    cases["y%d" % idx] = y
    cases["starts%d" % idx] = starts
    for variant, fn in (("v1", cx.isotonic_regression_multi_c), ("v3", cx.isotonic_regression_multi_c_3)):
        for update in (1, 0):
            out = y.copy()
            w = np.ones(n, dtype=np.int32)
            fn(out, starts, w, update)
            cases["%s_u%d_y%d" % (variant, update, idx)] = out
            cases["%s_u%d_w%d" % (variant, update, idx)] = w
    out = y.copy()
    cx.isotonic_regression_multi_c_2(out, starts)
    cases["v2_y%d" % idx] = out
    idx += 1
    cases["count"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "pava.npz"), **cases)
    return idx


def golden_x2z(cx):
    rng = np.random.RandomState(SEED + 2)
    cases = {}
    idx = 0
    for sizes in ([3], [2, 2], [1, 3], [5] * 20, list(power_law_sizes(rng, 300, 1, 40))):
        sizes = np.array(sizes, dtype=np.int64)
        n = int(sizes.sum())
        starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
        x = np.concatenate([rng.dirichlet(np.ones(k)) for k in sizes])
        z = np.zeros(n - len(sizes))
        cx.x2z_c(x, z, starts)
        xb = np.zeros(n)
        cx.z2x_c(xb, z, starts)
        cases["x%d" % idx] = x
        cases["starts%d" % idx] = starts
        cases["z%d" % idx] = z
        cases["xback%d" % idx] = xb
        idx += 1
    cases["count"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "x2z.npz"), **cases)
    return idx


def sparse_problem(rng, nb, K, m, L, noise=0.0):
    """Route/link incidence problem shaped like BASELINE config 1.
    noise > 0 perturbs b so that the optimum has a non-zero objective (a relative
    objective comparison is meaningless when f -> 0)."""
    import scipy.sparse as sps
    n = nb * K
    rows = np.concatenate([rng.choice(m, L, replace=False) for _ in range(n)])
    cols = np.repeat(np.arange(n), L)
    A = sps.csr_matrix((np.ones(n * L), (rows, cols)), shape=(m, n))
    x_true = rng.dirichlet(np.ones(K), size=nb).reshape(-1)
    b = A.dot(x_true)
    if noise > 0:
        b = b + noise * rng.randn(m)
    starts = np.arange(0, n, K, dtype=np.int64)
    return A, b, x_true, starts


def golden_solvers():
    """BATCH.solve / solve_BB / solve_LBFGS / solve_MD through get_solver_parts, and the
    functional BB / LBFGS / DORE / mirror_descent drivers, on seeded sparse problems."""
    import contextlib
    import io
    import scipy.sparse as sps
    import algorithm_utils as au
    import BATCH as batch
    import BB
    import LBFGS
    import DORE
    import solvers
    import mirror_descent
    import bsls_utils as bu

    rng = np.random.RandomState(SEED + 3)
    cases = {}
    quiet = io.StringIO()

    # --- x-space, sparse objective, proj_multi_simplex_c (the north-star path) ------------
    for tag, (nb, K, m, L, noise) in {"c1mini": (60, 5, 40, 6, 0.0), "k16": (20, 16, 50, 8, 0.0),
                                      "noisy": (80, 5, 120, 6, 0.3), "noisy20": (30, 20, 200, 10, 0.5)}.items():
        A, b, x_true, starts = sparse_problem(rng, nb, K, m, L, noise)
        n = nb * K
        x_init = np.ones(n) / K
        Acsr = A.tocsr()
        cases[tag + "_ptr"] = Acsr.indptr.astype(np.int64)
        cases[tag + "_idx"] = Acsr.indices.astype(np.int32)
        cases[tag + "_val"] = Acsr.data
        cases[tag + "_shape"] = np.array([m, n])
        cases[tag + "_b"] = b
        cases[tag + "_starts"] = starts
        cases[tag + "_xinit"] = x_init
        dense_min_eig = 0.1
        step_size, proj, line_search, obj = au.get_solver_parts((A, b), starts, dense_min_eig, is_sparse=True)
        g0 = np.zeros(n)
        cases[tag + "_f0"] = np.array(obj(x_init, g0))
        cases[tag + "_g0"] = g0
        for name, call in (
            ("bb", lambda: batch.solve_BB(obj, proj, line_search, x_init, max_iter=300)),
            ("pg", lambda: batch.solve(obj, proj, step_size, x_init, line_search, max_iter=100)),
            ("md", lambda: batch.solve_MD(obj, starts, step_size, x_init, max_iter=100)),
            ("lbfgs", lambda: batch.solve_LBFGS(obj, proj, line_search, x_init, max_iter=150)),
        ):
            with contextlib.redirect_stdout(quiet):
                sol = call()
            cases["%s_%s_x" % (tag, name)] = sol["x"]
            cases["%s_%s_f" % (tag, name)] = np.array(sol["f"])
            cases["%s_%s_iters" % (tag, name)] = np.array(sol["iterations"])
            cases["%s_%s_stop" % (tag, name)] = np.array(sol["stop"])
            cases["%s_%s_ftrace" % (tag, name)] = np.array([p[1] for p in sol["progress"]])

        # --- z-space: functional BB / LBFGS / DORE exactly as main.solve_in_z wires them ---
        block_sizes = np.full(nb, K, dtype=np.int64)
        N = bu.block_sizes_to_N(block_sizes)
        x0 = bu.particular_x0(block_sizes)
        x0 = np.asarray(x0).reshape(-1)
        z0 = bu.x2z(x_init, block_sizes)
        target = A.dot(x0) - b
        AT = A.T.tocsr()
        NT = N.T.tocsr()
        f = lambda z: 0.5 * np.linalg.norm(A.dot(N.dot(z)) + target) ** 2
        nabla_f = lambda z: NT.dot(AT.dot(A.dot(N.dot(z)) + target))
        zstarts = np.concatenate(([0], np.cumsum(block_sizes - 1)))[:-1]

        def projz(v):
            import c_extensions.c_extensions as cx
            cx.isotonic_regression_multi_c(v, zstarts)
            return np.maximum(np.minimum(v, 1.), 0.)

        log = lambda it, state, dur: 0.0
        opts = {"max_iter": 200, "opt_tol": 1e-30, "verbose": 0}
        with contextlib.redirect_stdout(quiet):
            zbb = BB.solve(z0.copy(), f, nabla_f, solvers.stopping, proj=projz, log=log, options=opts)
        cases[tag + "_z0"] = z0
        cases[tag + "_zbb_z"] = zbb
        cases[tag + "_zbb_f"] = np.array(f(zbb))
        opts2 = {"max_iter": 40, "opt_tol": 1e-30, "verbose": 0}
        with contextlib.redirect_stdout(quiet):
            zl = LBFGS.solve(z0.copy() + 1, f, nabla_f, solvers.stopping, proj=projz, log=log, options=opts2)
        cases[tag + "_zlbfgs_z"] = zl
        cases[tag + "_zlbfgs_f"] = np.array(f(zl))
        lsv = bu.lsv_operator(A, N)
        cases[tag + "_lsv"] = np.array(lsv)
        A_d = A * 0.99 / lsv
        t_d = target * 0.99 / lsv
        opts3 = {"max_iter": 150, "opt_tol": 1e-30, "verbose": 0}
        with contextlib.redirect_stdout(quiet):
            zd = DORE.solve(z0.copy(), lambda z: A_d.dot(N.dot(z)), lambda r: N.T.dot(A_d.T.dot(r)), t_d,
                            proj=projz, log=log, options=opts3, record_every=100)
        cases[tag + "_zdore_z"] = zd
        cases[tag + "_zdore_f"] = np.array(f(zd))

        # --- mirror_descent.least_squares (no caller or test in the reference: the only pin) -
        Lf = float(sps.linalg.svds(A, 1, return_singular_vectors=False)[0])
        cases[tag + "_Lf"] = np.array(Lf)
        xm = mirror_descent.least_squares(A, b, [int(K)] * nb, iters=60, tolerance=1e-9)
        cases[tag + "_md_ls_x"] = xm

    # --- the reference's dense 2-variable QP (tests/fast/test_BATCH.py) --------------------
    Q, c, x_true, f_min, min_eig = bu.generate_small_qp()
    starts = np.array([0])
    step_size, proj, line_search, obj = au.get_solver_parts((Q, c), starts, min_eig)
    sol = batch.solve_BB(obj, proj, line_search, np.array([.5, .5]))
    cases["qp_bb_x"] = sol["x"]
    cases["qp_bb_stop"] = np.array(sol["stop"])
    sol = batch.solve_BB(obj, proj, line_search, np.array([.5, .5]), f_min=f_min)
    cases["qp_bb_fmin_stop"] = np.array(sol["stop"])
    np.savez_compressed(os.path.join(HERE, "solvers.npz"), **cases)
    return len(cases)


def golden_pipeline():
    """The steps either side of the hot loop: BSLSMatrices.degree_reduced_form / get_LS /
    reconstruct (python/bsls_matrices.py:56-160), main.solve_in_z + LS_postprocess (python/main.py:41-136) on
    bsls_utils.generate_data problems (the reference's own end-to-end test input, tests/fast/test_main.py), and
    line_search_exact_quad_obj (python/algorithm_utils.py:140-155)."""
    import contextlib
    import io
    import types
    # the CLI module imports plotting and a site config that do not exist here: empty stand-ins, nothing numeric
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    cfg = types.ModuleType("config")
    cfg.ACCEPTED_LOG_LEVELS = ['CRITICAL', 'ERROR', 'WARNING', 'INFO', 'DEBUG', 'WARN']
    sys.modules["config"] = cfg
    import algorithm_utils as au
    import bsls_utils as bu
    import main as ref_main
    from bsls_matrices import BSLSMatrices

    cases = {}
    quiet = io.StringIO()
    config = {'full': True, 'L': True, 'OD': True, 'CP': True, 'LP': True, 'eq': 'CP', 'init': False}
    variants = {
        "default": dict(),
        "sparse_x": dict(alpha=0.5),
        "sparse_A": dict(A_sparse=0.05),
        "permuted": dict(permute=True),
        "larger": dict(n=400, m1=30, m2=40),
        "zeroed": dict(),          # tests/fast/test_bsls_matrices.py:64-90: a zero row of A and a zero block
    }
    for tag, kw in variants.items():
        np.random.seed(SEED)
        data = bu.generate_data(**kw)
        if tag == "zeroed":
            n = data['x_true'].size
            data['A'][0, :] = np.zeros(n)
            data['f'][0] = 0
            data['x_true'] = (data['U'].T.dot(data['f']) > 0) * data['x_true']
            data['b'] = data['A'].dot(data['x_true'])
        for k in ("A", "b", "x_true", "U", "f", "block_sizes"):
            cases["%s_in_%s" % (tag, k)] = np.asarray(data[k], dtype=np.float64)
        bm = BSLSMatrices(data=data, **config)
        bm.degree_reduced_form()
        AA, bb, N, block_sizes, x_split, nz, scaling, rsort_index, x0 = bm.get_LS()
        cases[tag + "_AA"] = np.asarray(AA.todense())
        cases[tag + "_bb"] = np.asarray(bb, dtype=np.float64)
        cases[tag + "_block_sizes"] = np.asarray(block_sizes, dtype=np.int64)
        cases[tag + "_x_split"] = np.asarray(x_split, dtype=np.float64)
        cases[tag + "_nz"] = np.asarray(nz, dtype=np.int64)
        cases[tag + "_scaling"] = np.asarray(scaling, dtype=np.float64)
        cases[tag + "_rsort_index"] = np.asarray(rsort_index, dtype=np.int64)
        cases[tag + "_x0"] = np.asarray(x0, dtype=np.float64).reshape(-1)
        cases[tag + "_x_true_rec"] = BSLSMatrices.reconstruct(x_split, rsort_index=rsort_index, scaling=scaling, nz=nz,
                                                              n=data['x_true'].size)
        with contextlib.redirect_stdout(quiet):
            iters, times, states = ref_main.solve_in_z(AA, bb, x0, N, block_sizes, 'BB')
        x_last, error, output = ref_main.LS_postprocess(states, x0, AA, bb, x_split, scaling=scaling, block_sizes=block_sizes,
                                                        N=N, output={})
        cases[tag + "_iters"] = np.asarray(iters, dtype=np.int64)
        cases[tag + "_z_last"] = np.asarray(states[-1], dtype=np.float64)
        cases[tag + "_x_last"] = np.asarray(x_last, dtype=np.float64)
        cases[tag + "_error"] = np.asarray(error, dtype=np.float64)
        cases[tag + "_start_error"] = np.array(output['0.5norm(Ax_init-b)^2'])
        cases[tag + "_opt_error"] = np.array(output['0.5norm(Ax*-b)^2'])
        cases[tag + "_max_f_diff"] = np.asarray(output['max|f * (x-x_true)|'], dtype=np.float64)
        cases[tag + "_wrong"] = np.asarray(output['incorrect x entries'], dtype=np.int64)
        cases[tag + "_per_flow"] = np.asarray(output['percent flow allocated incorrectly'], dtype=np.float64)
        cases[tag + "_start_dist"] = np.array(output['max|f * (x_init-x_true)|'])
    # line_search_exact_quad_obj on seeded dense QPs
    rng = np.random.RandomState(SEED + 11)
    for k in range(4):
        n = (3, 8, 20, 50)[k]
        M = rng.randn(n, n)
        Q = M.T.dot(M) + 0.1 * np.eye(n)
        c = rng.randn(n)
        x = rng.rand(n)
        x_new = x + (rng.randn(n) if k < 3 else 1e-10 * rng.randn(n))       # last case: step below progTol
        g = Q.dot(x) + c
        f = .5 * x.dot(Q.dot(x)) + c.dot(x)
        g_new = np.zeros(n)
        xn = x_new.copy()
        f_new = au.line_search_exact_quad_obj(x, f, g, xn, 0.0, g_new, Q, c)
        for name, v in (("Q", Q), ("c", c), ("x", x), ("x_new_in", x_new), ("g", g), ("f", np.array(f)), ("x_new_out", xn),
                        ("g_new_out", g_new), ("f_new", np.array(f_new))):
            cases["ls%d_%s" % (k, name)] = v
    cases["ls_count"] = np.array(4)
    cases["tags"] = np.array(sorted(variants))
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **cases)
    return len(cases)


def run_reference_tests():
    """Replays the reference's own unit tests for the path against the scratch build, so
    the fixtures are known to come from a reference that passes its own suite."""
    import unittest
    tdir = os.path.join(SCRATCH, "tests_fast")
    os.makedirs(tdir, exist_ok=True)
    ok = True
    for name in ("test_proj_simplex.py", "test_isotonic_regression.py", "test_BATCH.py"):
        with open(os.path.join(REF, "tests", "fast", name)) as fh:
            src = patch_source(fh.read(), name)
        src = src.replace("from python.", "from ").replace("import python.BATCH as batch", "import BATCH as batch")
        with open(os.path.join(tdir, name), "w") as fh:
            fh.write(src)
    sys.path.insert(0, tdir)
    import contextlib
    import io
    for mod in ("test_proj_simplex", "test_isotonic_regression", "test_BATCH"):
        suite = unittest.defaultTestLoader.loadTestsFromName(mod)
        with contextlib.redirect_stdout(io.StringIO()):
            res = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(suite)
        print("reference %s: ran %d, failures %d, errors %d" % (mod, res.testsRun, len(res.failures), len(res.errors)))
        ok = ok and res.wasSuccessful()
    return ok


def main():
    prepare()
    import c_extensions.c_extensions as cx
    ok = run_reference_tests()
    print("reference self-tests pass:", ok)
    print("projection cases:", golden_projection(cx))
    print("pava cases:", golden_pava(cx))
    print("x2z cases:", golden_x2z(cx))
    print("solver arrays:", golden_solvers())
    print("pipeline arrays:", golden_pipeline())


if __name__ == "__main__":
    main()
