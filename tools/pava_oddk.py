"""PAVA on 10^8 values for the z-space block sizes of the named configs (K - 1 = 15, 19, 4) and other odd sizes."""
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K in (15, 19, 5, 3, 7, 20, 31):
    for kind in ("ref", "zspace"):
        r = mb.time_pava(K, 10 ** 8 // K, kind, reps=3)
        print(json.dumps({"K": K, "kind": kind, "ms": round(r["ms_med"], 4), "frac": round(r["GBs"] / 6552, 3)}), flush=True)
