"""C3 (power-law blocks) projection + PAVA a few times; for ncu launch lists."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsls_b200
import bench
dev = torch.device("cuda", 0)
sizes = bench.power_law_sizes(10 ** 7)
print("blocks", len(sizes), "max", sizes.max(), "vars in blocks >512:", int(sizes[sizes > 512].sum()), ">256:", int(sizes[sizes > 256].sum()),
      ">64:", int(sizes[sizes > 64].sum()), "<=8:", int(sizes[sizes <= 8].sum()))
print(json.dumps(bench.bench_c3(bsls_b200, torch, dev, 6552.0, steps=3)))
