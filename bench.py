#!/usr/bin/env python
"""bench.py -- headline benchmark of the block-simplex least-squares hot path (contract: see the task).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...       (one rank per GPU)

Headline workload = BASELINE.json configs[4]: the Barzilai-Borwein solve (BATCH.solve_BB semantics) on 10^7 OD blocks x
16 routes, 10^6 links, 8 links per route (nnz = 1.28e9), OD blocks sharded over the GPUs, the partial link vector A x
summed by one NCCL all-reduce per objective evaluation.  STRONG scaling: the same problem at every N (it fits one GPU).
One STEP = one full solve from x = 1/K to the reference's stopping rule (|f_old - f| < 1e-12).

Prints ONE JSON line.  `value` = BB iterations / s over the K timed solves (device time, max over ranks) with the
problem resident in HBM; `e2e` = the same metric through the reference-facing Python API with HOST vectors (b and x_init
uploaded from pinned memory, x downloaded, every step, on every rank); `roofline` = the dominant kernel of the solve
(the SpMV pair) against the measured HBM peak; `cpu_baseline` = the reference's path (oracle restatement of
BATCH.solve_BB + the reference's own compiled projection) on one host core, on a shard of the same problem; `parity` =
final objective against the committed value of the single-GPU solve.  Sub-objects carry the other BASELINE configs:
`c2` (projection microbench + its host-C-ABI e2e), `n1e8`, `c3`, `c3_1e8`, `c1`, `c4_*`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 237423433
SIZES = (4, 16, 64)
NB = 10 ** 6
METRIC = "bb_iterations_per_sec"
UNIT = "iter/s"
C5 = dict(nb=10 ** 7, K=16, m=10 ** 6, L=8)
C5_NNZ, C5_N, C5_M, C5_NB = 1280000000, 160000000, 1000000, 10000000
MAX_ITER = 40
WORKLOAD = ("C5: BATCH.solve_BB (max_iter 40, prog_tol 1e-12) on 10^7 OD blocks x 16 routes, 10^6 links, 8 links per route "
            "(nnz 1.28e9), fp64, noise 0.1; OD blocks sharded over the GPUs, A x summed by one NCCL all-reduce per evaluation; "
            "one step = one full solve from x = 1/K")
CONFIG = {"workload": WORKLOAD,
          "l2": "inputs larger than L2: every evaluation streams the 10.2 GB index arrays of A and A^T (126 MB L2)"}
GOLDEN_C5 = os.path.join(ROOT, "tests", "golden", "c5_bb.json")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own C++ on the host cores
# --------------------------------------------------------------------------------------------
def ref_checker():
    from oracle import cpu
    cpu.build()
    r = cpu.ref()
    return (r, "reference") if r is not None else (cpu.port(), "port")


def cpu_project_parallel(chk, y, K, threads):
    """Runs the checker's proj_multi_simplex over disjoint slices from `threads` host threads
    (blocks are independent; ctypes releases the GIL)."""
    nb = y.size // K
    cuts = np.linspace(0, nb, threads + 1).astype(np.int64)

    def work(i):
        lo, hi = int(cuts[i]) * K, int(cuts[i + 1]) * K
        if hi > lo:
            chk.proj_multi_simplex(y[lo:hi], np.arange(0, hi - lo, K, dtype=np.int32))

    if threads == 1:
        work(0)
        return
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()


def cpu_c5_scaled(scale, seed=SEED + 5):
    """Config 5 scaled down by ``scale`` for the CPU legs (10^7 / scale OD blocks x 16 routes, 10^6 / scale links, 8 links
    per route), generated on the host with the same rule as bsls_b200.generate (L distinct, ascending links per route;
    x_true ~ Dirichlet(1) per block; b = A x_true + N(0, 0.1)).  Every term of an iteration's cost (non-zeros,
    variables, links) shrinks by the same factor, so iterations/s of the full problem = measured / scale."""
    import scipy.sparse as sps
    nb, K, m, L = C5["nb"] // scale, C5["K"], C5["m"] // scale, C5["L"]
    rng = np.random.RandomState(seed)
    n = nb * K
    links = np.sort(rng.randint(0, m - L + 1, size=(n, L), dtype=np.int32), axis=1) + np.arange(L, dtype=np.int32)
    AT = sps.csr_matrix((np.ones(n * L), links.reshape(-1), np.arange(0, (n + 1) * L, L, dtype=np.int64)), shape=(n, m))
    A = sps.csr_matrix(AT.T)
    x_true = rng.dirichlet(np.ones(K), size=nb).reshape(-1)
    b = A.dot(x_true) + 0.1 * rng.randn(m)
    return A, b, np.arange(0, n, K), np.ones(n) / K


def cpu_bb(scale, threads, steps, warmup):
    """The reference's BATCH.solve_BB on a scaled-down C5, timed on the host: the oracle's statement-for-statement
    restatement of the loop (oracle/solvers_np.py), CSR products as scipy's csr_matvec, projection by the reference's
    own compiled C++ when oracle/_ref exists.  threads > 1 spreads the products and the projection over host threads
    (the reference has no threading: this is the best-effort CPU number).  Returns per-step seconds and iterations."""
    from oracle import solvers_np as S
    chk, kind = ref_checker()
    A, b, starts, x0 = cpu_c5_scaled(scale)
    K = C5["K"]
    project = (lambda x: cpu_project_parallel(chk, x, K, threads))
    parts = S.get_solver_parts(A, b, starts, 0.1, threads=threads, project=project)
    times, its, f = [], [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        sol = S.solve_BB(parts[3], parts[1], parts[2], x0, max_iter=MAX_ITER)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            its.append(sol["iterations"] - 1)
        f = sol["f"]
    return times, its, f, ("reference" if kind == "reference" else "port"), A.nnz


def cpu_scale(threads):
    """Scale factor of the CPU problem: ~10-30 s of CPU work per run (about 10 ns per non-zero and iteration on one core)."""
    return 100 if threads <= 1 else 25


def cpu_line_fields(times, its, nnz, threads, kind, scale):
    value = float(np.sum(its)) / float(np.sum(times)) / scale
    sample = ("BATCH.solve_BB on C5 scaled down by %d (%d OD blocks x 16 routes, %d links, nnz %.3g), %d iterations per solve, "
              "%d host thread(s); iterations/s of the full problem = measured / %d (every cost term is linear in the size); "
              "loop = oracle restatement of python/BATCH.py:55-106, projection = %s"
              % (scale, C5["nb"] // scale, C5["m"] // scale, nnz, int(its[0]), threads, scale,
                 "the reference's compiled C++ (oracle/_ref)" if kind == "reference" else "oracle port"))
    return value, sample


_REAL_STDOUT = None


def emit(line):
    """The result line, on the process's real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    scale = cpu_scale(threads)
    times, its, f, kind, nnz = cpu_bb(scale, threads, args.steps, args.warmup)
    value, sample = cpu_line_fields(times, its, nnz, threads, kind, scale)
    total = float(np.sum(times))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": CONFIG, "sample": sample,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "projection_kind": kind, "f_final_scaled_problem": f},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------
# secondary workloads reported beside the headline (same JSON line, own sub-objects)
# --------------------------------------------------------------------------------------------
def power_law_sizes(total, lo=2, hi=4096, alpha=1.5, seed=SEED + 2):
    """C3 layout: sizes floor(lo * u^(-1/alpha)) clipped to [lo, hi], drawn until they
    sum to `total` (last block trimmed, never below 2)."""
    rng = np.random.RandomState(seed)
    sizes = []
    left = total
    while left > 0:
        k = np.floor(lo * rng.rand(65536) ** (-1.0 / alpha)).clip(lo, hi).astype(np.int64)
        c = np.cumsum(k)
        cut = int(np.searchsorted(c, left, side="left"))
        if cut >= len(k):
            sizes.append(k)
            left -= int(c[-1])
            continue
        take = k[:cut + 1].copy()
        take[-1] -= int(c[cut] - left)
        if take[-1] < 2:                      # fold a 0/1-sized tail into the previous block
            extra = int(take[-1])
            take = take[:-1]
            take[-1] += extra
        sizes.append(take)
        left = 0
    return np.concatenate(sizes)


def bench_c3(bsls_b200, torch, dev, peak, steps=5, total=10 ** 7):
    """BASELINE config 3: power-law blocks 2..4096, 10^7 variables (and the same layout scaled to 10^8):
    projection and segmented PAVA (values only, and with the pool-size array) on fresh inputs; CUDA events."""
    sizes = power_law_sizes(total)
    n, nb = int(sizes.sum()), len(sizes)
    starts = torch.as_tensor(np.concatenate(([0], np.cumsum(sizes)[:-1]))).to(dev)
    plan = bsls_b200.BlockPlan(starts, n)
    gen = torch.Generator(device=dev).manual_seed(SEED + 3)
    pos = torch.arange(n, device=dev) - torch.repeat_interleave(starts, torch.as_tensor(sizes).to(dev))
    ramp = 50.0 * torch.log1p(pos.to(torch.float64))       # the reference's PAVA test input (test_isotonic_regression.py:43)
    out = {"workload": "C3: power-law block sizes 2..4096, %d variables in %d blocks, fp64" % (n, nb)}

    def timed(make, run, bytes_):
        bufs = [make() for _ in range(steps + 3)]
        for b in bufs[:3]:
            run(b)
        torch.cuda.synchronize()
        evs = []
        for b in bufs[3:]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(b)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        return {"avg_ms": ms, "var_per_s": n / ms * 1e3, "algorithmic_bytes": bytes_, "GBs": bytes_ / ms / 1e6,
                "frac": bytes_ / ms / 1e6 / peak}

    out["projection"] = timed(lambda: torch.randn(n, dtype=torch.float64, device=dev, generator=gen),
                              lambda y: bsls_b200.proj_multi_simplex_c(y, plan), 16 * n + 4 * nb)
    mk = lambda: torch.randint(-50, 50, (n,), device=dev, generator=gen).to(torch.float64) + ramp
    out["pava"] = timed(mk, lambda y: bsls_b200.isotonic_regression_multi_c(y, plan, None, 1), 16 * n + 4 * nb)
    w = torch.ones(n, dtype=torch.int32, device=dev)

    def run_w(y):
        w.fill_(1)
        bsls_b200.isotonic_regression_multi_c(y, plan, w, 1)
    out["pava_with_pool_sizes"] = timed(mk, run_w, 16 * n + 4 * nb + 4 * n)
    return out


def bench_1e8(bsls_b200, torch, dev, peak, reps=4):
    """north_star's target size: projection and PAVA on 10^8-variable problems (uniform blocks of 4 / 16 / 64,
    fp64, fresh inputs every launch, CUDA events)."""
    gen = torch.Generator(device=dev).manual_seed(SEED + 9)
    out = {"workload": "10^8 variables in uniform blocks of K = 4, 16, 64; fp64; one launch per measurement on a fresh input"}
    for K in SIZES:
        nb = 10 ** 8 // K
        n = nb * K
        plan = bsls_b200.BlockPlan(torch.arange(0, n, K, dtype=torch.int64, device=dev), n)
        ramp = 50.0 * torch.log1p(torch.arange(K, dtype=torch.float64, device=dev))
        for name, make, run in (
                ("projection", lambda: torch.randn(n, dtype=torch.float64, device=dev, generator=gen),
                 lambda y: bsls_b200.proj_multi_simplex_c(y, plan)),
                ("pava", lambda: (torch.randint(-50, 50, (nb, K), device=dev, generator=gen).to(torch.float64) + ramp).reshape(-1),
                 lambda y: bsls_b200.isotonic_regression_multi_c(y, plan, None, 1))):
            ms = []
            for r in range(reps + 1):
                y = make()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(y)
                e1.record()
                torch.cuda.synchronize()
                if r > 0:
                    ms.append(e0.elapsed_time(e1))
                del y
            t = float(np.mean(ms))
            bytes_ = 16 * n + 4 * nb
            out["%s_K%d" % (name, K)] = {"avg_ms": t, "var_per_s": n / t * 1e3, "algorithmic_bytes": bytes_, "GBs": bytes_ / t / 1e6,
                                         "frac": bytes_ / t / 1e6 / peak}
        del plan
        torch.cuda.empty_cache()
    # the z-space projection of config 5 (python/main.py:57-65): blocks of K - 1 = 15 running sums of a simplex point,
    # perturbed by a gradient step (here N(0, 0.05)) -- the input PAVA sees inside solve_in_z
    K = 15
    nb = 10 ** 8 // K
    n = nb * K
    plan = bsls_b200.BlockPlan(torch.arange(0, n, K, dtype=torch.int64, device=dev), n)

    def make_z():
        e = -torch.log(torch.rand(nb, K + 1, dtype=torch.float64, device=dev, generator=gen))
        z = torch.cumsum(e / e.sum(1, keepdim=True), 1)[:, :K]
        return (z + 0.05 * torch.randn(nb, K, dtype=torch.float64, device=dev, generator=gen)).reshape(-1)
    ms = []
    for r in range(reps + 1):
        y = make_z()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bsls_b200.isotonic_regression_multi_c(y, plan, None, 1, clip01=True)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            ms.append(e0.elapsed_time(e1))
        del y
    t = float(np.mean(ms))
    bytes_ = 16 * n + 4 * nb
    out["pava_zspace_K15_clip"] = {"avg_ms": t, "var_per_s": n / t * 1e3, "algorithmic_bytes": bytes_, "GBs": bytes_ / t / 1e6,
                                   "frac": bytes_ / t / 1e6 / peak,
                                   "input": "cumulative sums of Dirichlet(1) blocks + N(0, 0.05); regression + [0,1] clamp in one launch"}
    del plan
    torch.cuda.empty_cache()
    return out


def bench_c1_c4(bsls_b200, torch, dev, with_cpu):
    """BASELINE configs 1 and 4 on one GPU: the BB solve on the reference's own CPU-sized problem (C1: 1,000
    OD blocks x 5 routes, 2,000 links) and mirror descent / L-BFGS on C4 (10^5 blocks x 20 routes, 5*10^4
    links, nnz = 2*10^7).  CPU columns: the oracle's restatement of the same reference functions, one thread
    (C4 on a 1/10-size problem)."""
    import scipy.sparse as sps
    from bsls_b200.generate import SyntheticProblem
    out = {}
    # ---- C1: BATCH.solve_BB, reference defaults (max_iter 2000, prog_tol 1e-12) ------------------------
    sp = SyntheticProblem.config("C1", noise=0.1, implicit_ones=False)
    parts = sp.solver_parts()
    bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=50)
    sol = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=2000)
    its = sol["iterations"] - 1
    out["c1"] = {"workload": "C1: BATCH.solve_BB, 1,000 OD blocks x 5 routes, 2,000 links, 10 links per route (latency-bound: 1.6 MB per iteration)",
                 "iter_per_s": its / sol["device_ms"] * 1e3, "iterations": its, "objective_evaluations": sol["obj_evals"],
                 "f_final": sol["f"], "stop": sol["stop"], "us_per_evaluation": 1e3 * sol["device_ms"] / max(1, sol["obj_evals"])}
    if with_cpu:
        from oracle import solvers_np as S
        A = sps.csr_matrix((np.ones(sp.nnz), sp.problem.a_idx.cpu().numpy(), sp.problem.a_ptr.cpu().numpy()), shape=(sp.m, sp.n))
        b = sp.b.cpu().numpy()
        starts = sp.starts.cpu().numpy()
        cp = S.get_solver_parts(A, b, starts, 0.1)
        t0 = time.perf_counter()
        ref = S.solve_BB(cp[3], cp[1], cp[2], sp.x_init.cpu().numpy(), max_iter=2000)
        dt = time.perf_counter() - t0
        out["c1"]["cpu_baseline"] = {"value": (ref["iterations"] - 1) / dt, "unit": "iter/s", "cores": 1, "kind": "port",
                                     "sample": "the same problem and call, %d iterations" % (ref["iterations"] - 1), "f_final": ref["f"]}
    del sp, parts
    # ---- C4: mirror descent and L-BFGS -----------------------------------------------------------------------
    sp = SyntheticProblem.config("C4", noise=0.1)
    nb, K = sp.nb, sp.K
    Lf = bsls_b200.bsls_utils.largest_singular_value(sp.problem)
    bsls_b200.mirror_descent.least_squares(sp.problem, None, [K] * nb, iters=5, Lf=Lf)
    torch.cuda.synchronize()
    iters = 200
    bsls_b200.mirror_descent.least_squares(sp.problem, None, [K] * nb, iters=iters, tolerance=0.0, Lf=Lf)
    torch.cuda.synchronize()
    ms = bsls_b200.mirror_descent.least_squares.last["device_ms"]      # CUDA-event time of the device-resident loop
    b_md = 24 * sp.nnz + 48 * sp.n + 32 * sp.m + 4 * sp.nb
    out["c4_mirror_descent"] = {"workload": "C4: mirror_descent.least_squares, 10^5 OD blocks x 20 routes, 5*10^4 links, nnz=2e7",
                                "iter_per_s": iters / ms * 1e3, "iterations": iters, "ms_per_iteration": ms / iters, "Lf": Lf,
                                "algorithmic_bytes_per_iteration": b_md, "GBs": b_md * iters / ms / 1e6}
    parts = sp.solver_parts()
    bsls_b200.BATCH.solve_LBFGS(parts[3], parts[1], parts[2], sp.x_init, max_iter=8)
    torch.cuda.synchronize()
    sol = bsls_b200.BATCH.solve_LBFGS(parts[3], parts[1], parts[2], sp.x_init, max_iter=60)
    torch.cuda.synchronize()
    its = sol["iterations"] - 1
    out["c4_lbfgs"] = {"workload": "C4: BATCH.solve_LBFGS (corrections=50), same problem; device-resident loop, CUDA-event time of the solve",
                       "iter_per_s": its / sol["device_ms"] * 1e3, "iterations": its, "ms_per_iteration": sol["device_ms"] / max(1, its),
                       "f_final": sol["f"], "stop": sol["stop"], "objective_evaluations": sol["obj_evals"]}
    sol = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=60)
    its = sol["iterations"] - 1
    out["c4_bb"] = {"workload": "C4: BATCH.solve_BB, same problem", "iter_per_s": its / sol["device_ms"] * 1e3, "iterations": its,
                    "ms_per_iteration": sol["device_ms"] / max(1, its), "f_final": sol["f"], "stop": sol["stop"]}
    if with_cpu:
        from oracle import solvers_np as S
        rng = np.random.RandomState(SEED + 4)
        nbs, ms_, L = 10000, 5000, 10
        n = nbs * K
        base = np.sort(rng.randint(0, ms_ - L + 1, size=(n, L)), axis=1) + np.arange(L)
        A = sps.csr_matrix((np.ones(n * L), (base.reshape(-1), np.repeat(np.arange(n), L))), shape=(ms_, n))
        b = A.dot(rng.dirichlet(np.ones(K), size=nbs).reshape(-1)) + 0.1 * rng.randn(ms_)
        t0 = time.perf_counter()
        S.md_least_squares(A, b, [K] * nbs, iters=20, tolerance=0.0, Lf=10.0)
        dt = time.perf_counter() - t0
        out["c4_mirror_descent"]["cpu_baseline"] = {"value": 20 / dt, "unit": "iter/s", "cores": 1, "kind": "port",
                                                    "sample": "20 iterations on a 1/10-size C4 (nnz=2e6)", "nnz_iter_per_s": 20 * n * L / dt}
        cp = S.get_solver_parts(A, b, np.arange(0, n, K), 0.1)
        t0 = time.perf_counter()
        ref = S.solve_LBFGS(cp[3], cp[1], cp[2], np.ones(n) / K, max_iter=20)
        dt = time.perf_counter() - t0
        out["c4_lbfgs"]["cpu_baseline"] = {"value": (ref["iterations"] - 1) / dt, "unit": "iter/s", "cores": 1, "kind": "port",
                                           "sample": "%d iterations on a 1/10-size C4 (nnz=2e6)" % (ref["iterations"] - 1)}
    return out


def bench_c2(bsls_b200, torch, dev, peak, steps=10, e2e_steps=3):
    """BASELINE config 2 (round 1's headline): proj_simplex_c microbench, 10^6 uniform blocks of 4 / 16 / 64, fp64, on fresh
    N(0,1) inputs; device-resident per-K launches timed by CUDA events, and the same call through the host C ABI
    (bsls_proj_multi_simplex on pinned host buffers, copies inside the timed region)."""
    from oracle import cpu
    gen = torch.Generator(device=dev).manual_seed(SEED + 1)
    starts = {K: torch.arange(0, NB * K, K, dtype=torch.int64, device=dev) for K in SIZES}
    plans = {K: bsls_b200.BlockPlan(starts[K], NB * K) for K in SIZES}
    out = {"workload": "C2 proj_simplex_c microbench: 10^6 uniform blocks x K in {4,16,64}, fp64, N(0,1), fresh input per launch"}
    tot_ms = 0.0
    for K in SIZES:
        bufs = [torch.randn(NB * K, dtype=torch.float64, device=dev, generator=gen) for _ in range(steps + 3)]
        keep = bufs[-1][: 1024 * K].clone()
        for b in bufs[:3]:
            bsls_b200.proj_multi_simplex_c(b, plans[K])
        torch.cuda.synchronize()
        evs = []
        for b in bufs[3:]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bsls_b200.proj_multi_simplex_c(b, plans[K])
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        want = keep.cpu().numpy().copy()
        cpu.port().proj_multi_simplex(want, np.arange(0, want.size, K))
        assert np.array_equal(bufs[-1][: 1024 * K].cpu().numpy(), want), "projection differs from the oracle (K=%d)" % K
        bytes_ = 2 * 8 * NB * K + 4 * NB
        out["K=%d" % K] = {"avg_ms": ms, "algorithmic_bytes": bytes_, "GBs": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / peak,
                           "var_per_s": NB * K / ms * 1e3}
        tot_ms += ms
        del bufs
    out["var_per_s"] = NB * sum(SIZES) / tot_ms * 1e3
    # host C ABI
    rng = np.random.RandomState(SEED + 7)
    host = {K: torch.empty(NB * K, dtype=torch.float64).pin_memory() for K in SIZES}
    src = {K: rng.randn(NB * K) for K in SIZES}
    hblocks = {K: np.arange(0, NB * K, K, dtype=np.int32) for K in SIZES}
    times = []
    for it in range(e2e_steps + 1):
        for K in SIZES:
            host[K].numpy()[:] = src[K]
        t0 = time.perf_counter()
        for K in SIZES:
            bsls_b200.proj_multi_simplex_c(host[K].numpy(), hblocks[K])
        dt = time.perf_counter() - t0
        if it > 0:
            times.append(dt)
    for K in SIZES:
        want = src[K][: 512 * K].copy()
        cpu.port().proj_multi_simplex(want, np.arange(0, want.size, K))
        assert np.array_equal(host[K].numpy()[: 512 * K], want)
    t = float(np.mean(times))
    out["e2e_host_abi"] = {"var_per_s": NB * sum(SIZES) / t, "ms_per_step": 1e3 * t, "h2d_bytes": sum(8 * NB * K + 4 * NB for K in SIZES),
                           "d2h_bytes": sum(8 * NB * K for K in SIZES),
                           "api": "bsls_proj_multi_simplex(double*, const int*, int, int) on pinned host buffers via ctypes"}
    chk, kind = ref_checker()
    data = {K: src[K].copy() for K in SIZES}
    t0 = time.perf_counter()
    for K in SIZES:
        cpu_project_parallel(chk, data[K], K, 1)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": NB * sum(SIZES) / dt, "unit": "var/s", "cores": 1, "kind": kind,
                           "sample": "one pass over the same three arrays (84e6 variables), one host thread"}
    return out


def dist_parity_check(bsls_b200, torch, dist, dev, rank, world, comm):
    """Driver-visible proof that the sharded path computes what one GPU computes: a config-5-shaped problem small enough to
    be solved twice -- sharded over all ranks (NCCL all-reduce of A x, all-gather of the step scalars) and, by every rank
    on its own, whole -- must reach the same objective (1e-6 relative, north_star's bar)."""
    from bsls_b200.generate import SyntheticProblem
    nb, K, m, L = 64000, 16, 5000, 8
    whole = SyntheticProblem(nb, K, m, L, device=dev, noise=0.1)
    parts = whole.solver_parts()
    one = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], whole.x_init, max_iter=300)
    shard = SyntheticProblem(nb, K, m, L, device=dev, rank=rank, world=world, comm=comm, noise=0.1)
    lo = shard.block_lo * K
    assert torch.equal(shard.x_true, whole.x_true[lo:lo + shard.n]), "the shard is not a slice of the whole problem"
    parts = shard.solver_parts()
    many = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], shard.x_init, max_iter=300)
    rel = abs(many["f"] - one["f"]) / abs(one["f"])
    assert rel <= 1e-6, "sharded BB objective %r differs from the single-GPU objective %r" % (many["f"], one["f"])
    return {"problem": "%d OD blocks x %d routes, %d links, %d links per route" % (nb, K, m, L), "f_sharded": many["f"],
            "f_single_gpu": one["f"], "rel_err": rel, "iterations_sharded": many["iterations"] - 1,
            "iterations_single_gpu": one["iterations"] - 1, "n_gpus": world, "ok": True}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def bind_near_gpu(torch, index):
    """Run this rank's host threads on the cores NVML lists as local to its GPU, so that the pinned staging buffers of the
    e2e leg (first touched by this process) sit on the GPU's NUMA node: with 8 ranks copying at once, buffers on the
    far socket halve the PCIe rate.  Deployment placement (what `numactl` per rank does), not part of the library."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "%d cores local to the GPU (NVML cpu affinity)" % len(cpus)
    except Exception as e:   # no NVML, no affinity support: leave the placement to the OS
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bound = None if args.no_bind else bind_near_gpu(torch, local_rank)
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    import bsls_b200
    from bsls_b200.generate import SyntheticProblem
    from bsls_b200 import _lib

    steps, warmup = args.steps, args.warmup
    peak, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the problem: config 5, this rank's OD blocks ------------------------------------------------
    comm = bsls_b200.Communicator() if world > 1 else None
    sp = SyntheticProblem.config("C5", rank=rank, world=world, comm=comm, noise=0.1)
    panels = sp.problem.set_panels() if sp.n * 8 > (64 << 20) else 1
    exchange = "none (one GPU)"
    if world > 1:
        p2p = bool(_lib.lib().bsls_comm_p2p_ready(comm.handle))
        exchange = ("NVLink peer memory: reduce-scatter + all-gather of A x fused with - b and the norms in one kernel (csrc/p2p.cuh)"
                    if p2p else "NCCL: ncclAllReduce of A x + ncclAllGather of the step scalars")
    step_size, proj, line_search, obj = sp.solver_parts()
    solve = lambda x0: bsls_b200.BATCH.solve_BB(obj, proj, line_search, x0, max_iter=MAX_ITER)

    for _ in range(warmup):
        solve(sp.x_init)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()      # ncu --profile-from-start off: only the timed region is listed
    ev0.record()
    sols = [solve(sp.x_init) for _ in range(steps)]
    ev1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    its = sum(s_["iterations"] - 1 for s_ in sols)
    evals = sum(s_["obj_evals"] for s_ in sols)
    launches = sum(s_["kernel_launches"] for s_ in sols)
    value = its / (ms * 1e-3)
    sol = sols[-1]

    # ---- parity: the objective the single-GPU solve reaches (committed), and sharded == single GPU at N > 1 -----
    parity = {"f_final": sol["f"], "iterations": sol["iterations"] - 1, "stop": sol["stop"]}
    if os.path.exists(GOLDEN_C5):
        with open(GOLDEN_C5) as fh:
            gold = json.load(fh)
        parity["f_expected"] = gold["f_final"]
        parity["expected_from"] = gold["from"]
        parity["rel_err"] = abs(sol["f"] - gold["f_final"]) / abs(gold["f_final"])
        parity["ok"] = bool(parity["rel_err"] <= 1e-6)
        assert parity["ok"], "C5 objective %r differs from the committed single-GPU value %r" % (sol["f"], gold["f_final"])
    if world > 1:
        parity["sharded_vs_single_gpu"] = dist_parity_check(bsls_b200, torch, dist, dev, rank, world, comm)

    # ---- per-kernel durations of the SpMV pair: the same launches, each bracketed by its own events -----
    prob = sp.problem
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    xa, xb = sol["x"], sp.x_init
    ga, gb = torch.empty_like(xa), torch.empty_like(xa)
    tk = {"ax": [], "atr": []}
    for rep in range(3 + 5):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        _lib.check(L.bsls_dev_lsq_residual_f64(prob.handle, xa.data_ptr(), st))
        e[1].record()
        _lib.check(L.bsls_dev_lsq_gradient_bb_f64(prob.handle, ga.data_ptr(), gb.data_ptr(), xb.data_ptr(), xa.data_ptr(), st))
        e[2].record()
        torch.cuda.synchronize()
        if rep >= 3:
            tk["ax"].append(e[0].elapsed_time(e[1]))
            tk["atr"].append(e[1].elapsed_time(e[2]))
    del ga, gb
    nnz_l, n_l, m_l = sp.nnz, sp.n, sp.m
    kern = {}
    for key, name, alg, stored, launches_per in (
            ("ax", "r = A x - b: spmv_vector_kernel<EpiResidual> over %d column panels + panel_reduce_kernel" % panels,
             12 * nnz_l + 8 * n_l + 24 * m_l, 4 * nnz_l + 8 * n_l + 8 * m_l * (panels + 2) + 8 * m_l * (panels if panels > 1 else 0), panels + 1),
            ("atr", "g = A^T r + step dot products: spmv_ell_kernel<EpiGradBB, L=8>",
             12 * nnz_l + 16 * n_l + 8 * m_l + 24 * n_l, 4 * nnz_l + 32 * n_l + 8 * m_l, 1)):
        t = float(np.mean(tk[key]))
        kern[key] = {"kernel": name, "avg_ms": t, "launches": launches_per, "algorithmic_bytes": alg, "stored_bytes": stored,
                     "GBs": alg / t / 1e6, "frac": alg / t / 1e6 / peak, "GBs_stored": stored / t / 1e6,
                     "frac_stored": stored / t / 1e6 / peak, "gathers_per_s": nnz_l / t * 1e3}
    dom = max(kern.values(), key=lambda k: k["avg_ms"])
    b_bb = 24 * C5_NNZ + 72 * C5_N + 32 * C5_M + 4 * C5_NB                       # SURVEY 8d, fp64 values + int32 indices
    stored_it = 8 * C5_NNZ + 72 * C5_N + 32 * C5_M + 4 * C5_NB                   # index-only A, both sides
    roofline = {"bound": "hbm", "achieved": dom["GBs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": None,
                "kernel": dom["kernel"], "peak_source": peak_src,
                "achieved_stored": dom["GBs_stored"], "frac_stored": dom["frac_stored"],
                "note": "A is held index-only (every stored entry is 1): `achieved` credits SURVEY 8d's 12 B per non-zero, "
                        "`achieved_stored` the 4 B actually moved.  Both products are bound by the L1TEX wavefront rate of their "
                        "scattered 8-byte gathers (one per non-zero), not by HBM: see gathers_per_s and DESIGN.md",
                "per_kernel": kern,
                "whole_solve": {"algorithmic_bytes_per_evaluation": b_bb, "stored_bytes_per_evaluation": stored_it,
                                "GBs": b_bb * evals / ms / 1e6, "frac_of_n_gpus_peak": b_bb * evals / ms / 1e6 / (peak * world),
                                "GBs_stored": stored_it * evals / ms / 1e6}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as fh:
            tr = json.load(fh)
        dom_key = "ax" if dom is kern["ax"] else "atr"
        per = tr.get("spmv_c5_bytes_per_launch", {})
        # measured on the full problem of one GPU; a rank of a sharded run moves 1 / world of it
        roofline["traffic"] = per.get(dom_key) / world if per.get(dom_key) else None
        roofline["traffic_note"] = per.get(dom_key + "_note")
        roofline["traffic_source"] = tr.get("source")

    # ---- e2e: the reference-facing API with HOST vectors, on EVERY rank --------------------------------
    e2e_steps = max(1, min(steps, args.e2e_steps))
    xh = torch.full((sp.n,), 1.0 / sp.K, dtype=torch.float64).pin_memory()
    bh = sp.b.cpu().pin_memory()
    e2e_its, x_out = 0, None
    prob.set_b(bh)
    bsls_b200.BATCH.solve_BB(obj, proj, line_search, xh, max_iter=3)     # untimed: pins the result buffer once
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prob.set_b(bh)                                      # measurements arrive from the host
        res = bsls_b200.BATCH.solve_BB(obj, proj, line_search, xh, max_iter=MAX_ITER)   # x_init up, x back (pinned)
        f_host = float(res["f"])
        e2e_its += res["iterations"] - 1
        x_out = res["x"]
    torch.cuda.synchronize()
    e2e_t = max_over_ranks(time.perf_counter() - t0)
    assert not x_out.is_cuda and abs(f_host - sol["f"]) <= 1e-9 * abs(sol["f"])
    e2e = {"value": e2e_its / e2e_t, "unit": UNIT, "h2d_bytes_per_step": 8 * (C5_N + world * C5_M), "d2h_bytes_per_step": 8 * C5_N + 8 * world,
           "ms_per_step": 1e3 * e2e_t / e2e_steps, "steps": e2e_steps, "host_placement": bound or "left to the OS",
           "api": "bsls_b200.BATCH.solve_BB(obj, proj, line_search, x_init) with x_init / b in pinned host memory and x returned to "
                  "the host, on every rank (bytes are whole-job totals)"}
    del xh, bh, x_out
    if comm is not None and os.environ.get("BSLS_P2P_PROF"):   # development: the exchange kernels' wait / transfer times on stderr
        barrier()
        comm.close()

    # ---- CPU baseline and the other BASELINE configs (rank 0, device-local work only) -----------------------
    extras = {}
    del sols, sol, sp, prob, obj, proj, line_search, step_size, solve, xa, xb
    torch.cuda.empty_cache()
    if rank == 0:
        cpu_base = None
        if world == 1:
            scale = cpu_scale(1)
            times, cits, cf, kind, cnnz = cpu_bb(scale, 1, 1, 1)
            cval, csample = cpu_line_fields(times, cits, cnnz, 1, kind, scale)
            cpu_base = {"value": cval, "unit": UNIT, "cores": 1, "kind": "port", "sample": csample, "projection_kind": kind}
        if not args.skip_extras:
            extras["c2"] = bench_c2(bsls_b200, torch, dev, peak)
            extras["n1e8"] = bench_1e8(bsls_b200, torch, dev, peak)
            extras["c3"] = bench_c3(bsls_b200, torch, dev, peak)
            extras["c3_1e8"] = bench_c3(bsls_b200, torch, dev, peak, total=10 ** 8, steps=3)
            extras.update(bench_c1_c4(bsls_b200, torch, dev, with_cpu=(world == 1)))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": CONFIG,
                "solve": {"iterations_per_solve": its // steps, "objective_evaluations_per_solve": evals // steps,
                          "ms_per_iteration": ms / max(1, its), "ms_per_evaluation": ms / max(1, evals), "panels_per_gpu": panels,
                          "exchange": exchange,
                          "note": "a back-tracked line search costs no extra product: the objective is quadratic along the step "
                                  "(decide_kernel, lsq.cuh)"},
                "parity": parity, "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        for k, v in extras.items():
            if v is not None:
                line[k] = v
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank's host threads to the cores local to its GPU")
    ap.add_argument("--skip-extras", action="store_true", help="headline (BB solve on C5) only: no C1 / C2 / C3 / C4 / 10^8 sub-objects")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    # file descriptor 1 when NCCL_DEBUG is set): for the run, descriptor 1 points at stderr and the JSON line is
    # written to the real stdout at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
