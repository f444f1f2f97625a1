"""A/B of the word-per-lane PAVA kernel against the warp-window / CTA kernels on uniform layouts (set BSLS_PAVA_NO_WORDS=1 for B)."""
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K in (100, 256, 512, 1000):
    for kind in ("ref", "normal"):
        r = mb.time_pava(K, 2 * 10 ** 7 // K, kind, reps=3)
        print(json.dumps({"K": K, "kind": kind, "ms": round(r["ms_med"], 4), "frac": round(r["GBs"] / 6552, 3)}), flush=True)
