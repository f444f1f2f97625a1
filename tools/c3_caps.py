"""C3 PAVA with different per-SM grid caps of the (tile, words, cta) kernels: BSLS_PAVA_CAPS=<tile>,<words>,<cta>."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsls_b200
import bench
dev = torch.device("cuda", 0)
for caps in ("0,0,0", "6,3,1", "5,4,1", "7,2,1", "4,4,1", "6,4,2", "8,3,1", "5,2,1"):
    os.environ["BSLS_PAVA_CAPS"] = caps
    r = bench.bench_c3(bsls_b200, torch, dev, 6552.0, steps=5)
    print(caps, round(r["pava"]["avg_ms"], 4), flush=True)
