"""Tuning sweep of the thread-per-row PAVA kernel: BSLS_PAVA_CFG=<threads>,<G> (read per call by the library)."""
import json, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K in (16, 4, 15, 32):
    for th in (32, 64, 128):
        for G in sorted({1, max(1, 16 // K), max(1, 32 // K)}):
            os.environ["BSLS_PAVA_CFG"] = "%d,%d" % (th, G)
            r = mb.time_pava(K, 10 ** 8 // K, "ref", reps=3)
            print(json.dumps({"K": K, "threads": th, "G": G, "ms": round(r["ms_med"], 4), "frac": round(r["GBs"] / 6552, 3)}), flush=True)
