import json, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K, G in ((16, 1), (4, 4), (32, 1)):
    for th in (64, 128):
        for kind in ("ref", "zspace"):
            os.environ["BSLS_PAVA_CFG"] = "%d,%d" % (th, G)
            r = mb.time_pava(K, 10 ** 8 // K, kind, reps=3)
            print(json.dumps({"K": K, "threads": th, "kind": kind, "ms": round(r["ms_med"], 4), "frac": round(r["GBs"] / 6552, 3)}), flush=True)
