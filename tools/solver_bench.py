"""Times the sparse least-squares path on synthetic configs (development tool; bench.py carries
the judged numbers).  python tools/solver_bench.py C1 C4 C5 [--iters 30] [--explicit]
Under torchrun: OD blocks are sharded over ranks."""
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
from bsls_b200 import _lib  # noqa: E402
from bsls_b200.generate import SyntheticProblem, CONFIGS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--explicit", action="store_true", help="store fp64 values instead of implicit ones")
    ap.add_argument("--scale", type=float, default=1.0, help="scale nb (and m) of the config")
    ap.add_argument("--modes", default="0,0")
    ap.add_argument("--panel-mb", type=int, default=32, help="L2 budget of one column panel of x (0: no panels)")
    ap.add_argument("--noise", type=float, default=0.1)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = bsls_b200.Communicator()
    L = _lib.lib()
    for name in args.configs:
        nb, K, m, Lk = CONFIGS[name]
        nb, m = max(world, int(nb * args.scale)), max(Lk, int(m * args.scale))
        t0 = time.time()
        sp = SyntheticProblem(nb, K, m, Lk, rank=rank, world=world, comm=comm, implicit_ones=not args.explicit, noise=args.noise)
        panels = sp.problem.set_panels(l2_budget_bytes=args.panel_mb << 20) if args.panel_mb > 0 and sp.n * 8 > (48 << 20) else 1
        torch.cuda.synchronize()
        gen_s = time.time() - t0
        a_mode, t_mode = [int(v) for v in args.modes.split(",")]
        sp.problem.set_modes(a_mode, t_mode)
        prob = sp.problem
        x = sp.x_init.clone()
        g = torch.empty_like(x)
        st = torch.cuda.current_stream().cuda_stream
        # kernel timings
        for _ in range(3):
            prob.obj(x, g)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tr, tg = [], []
        for _ in range(5):
            ev[0].record()
            _lib.check(L.bsls_dev_lsq_residual_f64(prob.handle, x.data_ptr(), st))
            ev[1].record()
            _lib.check(L.bsls_dev_lsq_gradient_f64(prob.handle, g.data_ptr(), st))
            ev[2].record()
            torch.cuda.synchronize()
            tr.append(ev[0].elapsed_time(ev[1]))
            tg.append(ev[1].elapsed_time(ev[2]))
        vb = 8 if args.explicit else 0
        bytes_r = sp.nnz * (4 + vb) + 8 * (sp.m + 1) + 8 * sp.n + 16 * sp.m
        bytes_g = sp.nnz * (4 + vb) + 8 * (sp.n + 1) + 8 * sp.m + 8 * sp.n
        step_size, proj, line_search, obj = sp.solver_parts()
        sol = bsls_b200.BATCH.solve_BB(obj, proj, line_search, sp.x_init, max_iter=args.iters, prog_tol=0.0)
        sol = bsls_b200.BATCH.solve_BB(obj, proj, line_search, sp.x_init, max_iter=args.iters, prog_tol=0.0)
        its = sol["iterations"] - 1
        ms_it = sol["device_ms"] / max(1, sol["obj_evals"])
        if rank == 0:
            print(json.dumps({
                "config": name, "world": world, "nb": nb, "K": K, "m": m, "L": Lk, "n_local": sp.n, "nnz_local": sp.nnz,
                "explicit_values": args.explicit, "panels": panels, "gen_s": round(gen_s, 2),
                "residual_ms": float(np.median(tr)), "gradient_ms": float(np.median(tg)),
                "residual_GBs_stored": bytes_r / np.median(tr) / 1e6, "gradient_GBs_stored": bytes_g / np.median(tg) / 1e6,
                "bb_iters": its, "bb_f": sol["f"], "bb_backtracks": sol["backtracks"], "bb_evals": sol["obj_evals"],
                "bb_ms_per_eval": ms_it, "bb_iter_per_s": 1e3 * its / sol["device_ms"],
                "bb_launches": sol["kernel_launches"],
                "B_BB_GB": sp.bytes_bb_iteration() / 1e9,
                "roofline_GBs_algorithmic": sp.bytes_bb_iteration() * its / sol["device_ms"] / 1e6}))
        del sp, prob, x, g, sol
        torch.cuda.empty_cache()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
