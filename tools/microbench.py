"""Quick kernel timings on one GPU (development aid; bench.py is the contract)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import __graft_entry__ as g

pass
import bsls_b200


def make_input(n, K, kind, dtype, gen):
    if kind == "normal":        # BASELINE config 2
        return torch.randn(n, dtype=dtype, device="cuda", generator=gen)
    if kind == "uniform":       # the reference's own stress input (test_stress_proj_simplex.py:28)
        return torch.rand(n, dtype=dtype, device="cuda", generator=gen)
    if kind == "near":          # a solver iterate: a feasible point with a dense support, slightly perturbed
        e = -torch.log(torch.rand(n // K, K, dtype=dtype, device="cuda", generator=gen))
        x = e / e.sum(1, keepdim=True)
        return (x + 0.01 / K * torch.randn(n // K, K, dtype=dtype, device="cuda", generator=gen)).reshape(-1)
    raise ValueError(kind)


def time_proj(K, nb, reps=10, dtype=torch.float64, ball=False, kind="normal"):
    n = nb * K
    gen = torch.Generator(device="cuda").manual_seed(K)
    starts = torch.arange(0, n, K, dtype=torch.int64, device="cuda")
    plan = bsls_b200.BlockPlan(starts, n)
    bufs = [make_input(n, K, kind, dtype, gen) for _ in range(reps + 3)]
    fn = bsls_b200.proj_multi_ball_c if ball else bsls_b200.proj_multi_simplex_c
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for b in bufs[:3]:
        fn(b, plan)
    flush.zero_()
    torch.cuda.synchronize()
    evs = []
    for b in bufs[3:]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(b, plan)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    es = bufs[0].element_size()
    bytes_ = 2 * es * n + 4 * nb
    return {"K": K, "nb": nb, "dtype": str(dtype), "ball": ball, "ms_med": float(np.median(ms)), "ms_min": float(ms.min()),
            "gvar_s": n / np.median(ms) / 1e6, "GBs": bytes_ / np.median(ms) / 1e6}


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    print("copy GB/s", 5 * 2 * a.numel() * 8 / e0.elapsed_time(e1) / 1e6)
    del a, b
    for K, nb in [(4, 10 ** 6), (5, 10 ** 6), (16, 10 ** 6), (20, 10 ** 6), (32, 10 ** 6), (64, 10 ** 6), (128, 500000), (512, 100000),
                  (16, 6250000), (4, 25 * 10 ** 6), (64, 1562500)]:
        print(json.dumps(time_proj(K, nb)))
    print(json.dumps(time_proj(16, 10 ** 6, dtype=torch.float32)))
    print(json.dumps(time_proj(16, 10 ** 6, ball=True)))


def time_pava(K, nb, kind="ref", reps=8, with_weight=False, clip=False):
    n = nb * K
    gen = torch.Generator(device="cuda").manual_seed(K)
    starts = torch.arange(0, n, K, dtype=torch.int64, device="cuda")
    plan = bsls_b200.BlockPlan(starts, n)

    def make():
        if kind == "ref":
            ramp = 50.0 * torch.log1p(torch.arange(K, dtype=torch.float64, device="cuda"))
            r = torch.randint(-50, 50, (nb, K), device="cuda", generator=gen).to(torch.float64)
            return (r + ramp).reshape(-1)
        if kind == "normal":
            return torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
        if kind == "sorted":     # already isotonic: one sweep, nothing pools (the kernel's floor)
            return torch.sort(torch.randn(nb, K, dtype=torch.float64, device="cuda", generator=gen), dim=1)[0].reshape(-1)
        if kind == "zspace":
            e = -torch.log(torch.rand(nb, K + 1, dtype=torch.float64, device="cuda", generator=gen))
            x = e / e.sum(1, keepdim=True)
            return (torch.cumsum(x, 1)[:, :K] + 0.05 * torch.randn(nb, K, dtype=torch.float64, device="cuda", generator=gen)).reshape(-1)
        raise ValueError(kind)

    bufs = [make() for _ in range(reps + 3)]
    ws = [torch.ones(n, dtype=torch.int32, device="cuda") for _ in range(reps + 3)] if with_weight else [None] * (reps + 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for b, w in zip(bufs[:3], ws[:3]):
        bsls_b200.isotonic_regression_multi_c(b, plan, w, 1, clip01=clip)
    flush.zero_()
    torch.cuda.synchronize()
    evs = []
    for b, w in zip(bufs[3:], ws[3:]):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bsls_b200.isotonic_regression_multi_c(b, plan, w, 1, clip01=clip)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    bytes_ = 16 * n + 4 * nb + (8 * n if with_weight else 0)
    return {"pava_K": K, "nb": nb, "kind": kind, "weights": with_weight, "ms_med": float(np.median(ms)),
            "gvar_s": n / np.median(ms) / 1e6, "GBs": bytes_ / np.median(ms) / 1e6}
