import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K in (4, 16, 20):
    for ball in (False, True):
        r = mb.time_proj(K, 10 ** 8 // K, ball=ball)
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}))
