"""Times the BSLS_TUNE variants of the uniform projection kernel (development aid).
Each variant runs in a fresh process because the choice is read once per process."""
import json
import os
import subprocess
import sys

CODE = r'''
import sys, json, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import microbench as mb
K, nb = int(sys.argv[1]), int(sys.argv[2])
print(json.dumps(mb.time_proj(K, nb)))
'''
for K, nb in ((16, 10 ** 6), (64, 10 ** 6), (16, 6250000)):
    for tv in range(0, 6):
        env = dict(os.environ, BSLS_TUNE=str(tv))
        out = subprocess.run([sys.executable, "-c", CODE, str(K), str(nb)], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
        try:
            d = json.loads(line)
            print("K=%d nb=%d tune=%d  %.1f us  %.0f GB/s  %.0f Gvar/s" % (K, nb, tv, d["ms_med"] * 1e3, d["GBs"], d["gvar_s"]))
        except Exception:
            print("K=%d tune=%d FAILED %s" % (K, tv, line))
