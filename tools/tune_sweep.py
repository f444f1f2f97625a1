"""Times BSLS_TUNE variants of the uniform projection kernels (development aid).
usage: tune_sweep.py K:nb[:kind] ... -- tv tv ...   (kind: normal | uniform | near)"""
import json
import os
import subprocess
import sys

CODE = r'''
import sys, json, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import microbench as mb
K, nb, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
print(json.dumps(mb.time_proj(K, nb, kind=kind)))
'''
args = sys.argv[1:]
cut = args.index("--")
cases, tvs = args[:cut], [int(t) for t in args[cut + 1:]]
for case in cases:
    parts = case.split(":")
    K, nb = int(parts[0]), int(parts[1])
    kind = parts[2] if len(parts) > 2 else "normal"
    for tv in tvs:
        env = dict(os.environ, BSLS_TUNE=str(tv))
        out = subprocess.run([sys.executable, "-c", CODE, str(K), str(nb), kind], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
        try:
            d = json.loads(line)
            print("K=%d nb=%d %s tune=%d  %.1f us  %.0f GB/s  %.0f Gvar/s" % (K, nb, kind, tv, d["ms_med"] * 1e3, d["GBs"], d["gvar_s"]))
        except Exception:
            print("K=%d tune=%d FAILED %s" % (K, tv, line))
