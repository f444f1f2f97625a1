"""Word-per-lane PAVA kernel on uniform layouts of 100..1000 entries per block, 2*10^7 values.  (The warp-window /
CTA-per-block sweep kernels it replaced measured 0.51 / 0.50 / 1.64 / 1.12 ms on the reference generator.)"""
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K in (100, 256, 512, 1000):
    for kind in ("ref", "normal"):
        r = mb.time_pava(K, 2 * 10 ** 7 // K, kind, reps=3)
        print(json.dumps({"K": K, "kind": kind, "ms": round(r["ms_med"], 4), "frac": round(r["GBs"] / 6552, 3)}), flush=True)
