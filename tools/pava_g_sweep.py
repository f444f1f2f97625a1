"""K = 16 / 8 / 4 / 2 with different numbers of blocks per row (BSLS_PAVA_CFG=<threads>,<G>), 10^8 values."""
import json, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K, Gs in ((16, (1, 2)), (8, (2, 4)), (4, (4, 8)), (2, (8, 16))):
    for G in Gs:
        for kind in ("ref", "normal"):
            os.environ["BSLS_PAVA_CFG"] = "64,%d" % G
            r = mb.time_pava(K, 10 ** 8 // K, kind, reps=3)
            print(json.dumps({"K": K, "G": G, "kind": kind, "ms": round(r["ms_med"], 4), "frac": round(r["GBs"] / 6552, 3)}), flush=True)
