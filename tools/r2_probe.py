"""Round-2 probe (development tool): SpMV variants and the device-resident BATCH loop on config-5 shapes.

    python tools/r2_probe.py [--scale 0.125] [--what spmv,bb,c1,c4]

Prints one JSON line per measurement.  Environment switches read by the library (BSLS_ELL_TEX, BSLS_BATCH_LEGACY) are
per process: run the tool once per setting.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bsls_b200  # noqa: E402
from bsls_b200.generate import SyntheticProblem, CONFIGS  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms)), float(np.min(ms))


def out(**kw):
    kw["env"] = {k: v for k, v in os.environ.items() if k.startswith("BSLS_")}
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.125)
    ap.add_argument("--what", default="spmv,bb")
    ap.add_argument("--panel-mb", default="48")
    args = ap.parse_args()
    what = args.what.split(",")
    torch.cuda.set_device(0)
    if "spmv" in what or "bb" in what:
        nb, K, m, L = CONFIGS["C5"]
        nb = int(nb * args.scale)
        sp = SyntheticProblem(nb, K, m, L, noise=0.1)
        prob = sp.problem
        x = sp.x_true.clone()
        r = torch.randn(m, dtype=torch.float64, device="cuda")
        g = torch.empty_like(x)
        rr = torch.empty(m, dtype=torch.float64, device="cuda")
        if "spmv" in what:
            ref = None
            for mode in (1, 2):
                prob.set_modes(0, mode)
                avg, best = timed(lambda: prob.rmatvec(r, g))
                if ref is None:
                    ref = g.clone()
                same = bool(torch.equal(ref, g))
                out(what="At_r", scale=args.scale, nnz=sp.nnz, mode=mode, avg_ms=avg, best_ms=best, gathers_per_s=sp.nnz / best * 1e3,
                    bit_identical_to_stream=same)
            prob.set_modes(0, 0)
            for mb in [int(v) for v in args.panel_mb.split(",")]:
                panels = prob.set_panels(l2_budget_bytes=mb << 20) if sp.n * 8 > (mb << 20) else prob.set_panels(panel_cols=sp.n)
                for pm in ("4", "8", "16", "1"):
                    os.environ["BSLS_SPMV_P"] = pm
                    if panels > 1:
                        prob.set_panels(l2_budget_bytes=mb << 20)
                    else:
                        prob.set_modes(int(pm), 0)
                    avg, best = timed(lambda: prob.matvec(x, rr))
                    out(what="A_x", scale=args.scale, nnz=sp.nnz, panel_mb=mb, panels=panels, mode=pm, avg_ms=avg, best_ms=best,
                        gathers_per_s=sp.nnz / best * 1e3)
                os.environ.pop("BSLS_SPMV_P", None)
        if "bb" in what:
            prob.set_modes(0, 0)
            panels = prob.set_panels() if sp.n * 8 > (64 << 20) else 1
            step_size, proj, line_search, obj = sp.solver_parts()
            bsls_b200.BATCH.solve_BB(obj, proj, line_search, sp.x_init, max_iter=4)
            for rep in range(2):
                t0 = time.perf_counter()
                sol = bsls_b200.BATCH.solve_BB(obj, proj, line_search, sp.x_init, max_iter=40)
                wall = time.perf_counter() - t0
                its = sol["iterations"] - 1
                out(what="bb_c5", scale=args.scale, panels=panels, iterations=its, evals=sol["obj_evals"], backtracks=sol["backtracks"],
                    device_ms=sol["device_ms"], wall_ms=1e3 * wall, ms_per_iter=sol["device_ms"] / max(1, its), f=sol["f"], stop=sol["stop"],
                    launches=sol["kernel_launches"])
        del sp, prob
        torch.cuda.empty_cache()
    if "c1" in what:
        sp = SyntheticProblem.config("C1", noise=0.1, implicit_ones=False)
        parts = sp.solver_parts()
        bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=50)
        for rep in range(4):
            os.environ["BSLS_TINY_CLUSTER"] = str(rep // 2)
            t0 = time.perf_counter()
            sol = bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=2000)
            wall = time.perf_counter() - t0
            its = sol["iterations"] - 1
            out(what="bb_c1", iterations=its, evals=sol["obj_evals"], backtracks=sol["backtracks"], device_ms=sol["device_ms"],
                wall_ms=1e3 * wall, iter_per_s=its / sol["device_ms"] * 1e3, f=sol["f"], stop=sol["stop"])
        os.environ.pop("BSLS_TINY_CLUSTER", None)
    if "c4" in what:
        sp = SyntheticProblem.config("C4", noise=0.1)
        parts = sp.solver_parts()
        for name, fn in (("bb", lambda mi: bsls_b200.BATCH.solve_BB(parts[3], parts[1], parts[2], sp.x_init, max_iter=mi)),
                         ("md", lambda mi: bsls_b200.BATCH.solve_MD(parts[3], sp.starts, parts[0], sp.x_init, max_iter=mi)),
                         ("lbfgs", lambda mi: bsls_b200.BATCH.solve_LBFGS(parts[3], parts[1], parts[2], sp.x_init, max_iter=mi)),
                         ("pg", lambda mi: bsls_b200.BATCH.solve(parts[3], parts[1], parts[0], sp.x_init, line_search=parts[2], max_iter=mi))):
            fn(5)
            t0 = time.perf_counter()
            sol = fn(100)
            wall = time.perf_counter() - t0
            its = sol["iterations"] - 1
            out(what="c4_" + name, iterations=its, evals=sol["obj_evals"], backtracks=sol["backtracks"], device_ms=sol["device_ms"],
                wall_ms=1e3 * wall, ms_per_iter=sol["device_ms"] / max(1, its), f=sol["f"], stop=sol["stop"])


if __name__ == "__main__":
    main()
