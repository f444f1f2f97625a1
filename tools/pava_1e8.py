"""PAVA on 10^8 variables (north_star's target size), uniform K = 4 / 16 / 64, three input kinds."""
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
kinds = sys.argv[1:] or ["ref", "normal", "zspace", "sorted"]
for K in (4, 15, 16, 64):
    for kind in kinds:
        r = mb.time_pava(K, 10 ** 8 // K, kind, reps=5)
        r["frac_hbm_6552"] = r["GBs"] / 6552.0
        print(json.dumps(r), flush=True)
