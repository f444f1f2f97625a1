import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
for K, nb in [(15, 10 ** 6), (63, 10 ** 6), (3, 10 ** 6), (16, 6250000)]:
    for kind in ("ref", "zspace", "normal"):
        print(json.dumps(mb.time_pava(K, nb, kind)))
print(json.dumps(mb.time_pava(15, 10 ** 6, "ref", with_weight=True)))
print(json.dumps(mb.time_pava(15, 10 ** 6, "zspace", clip=True)))
