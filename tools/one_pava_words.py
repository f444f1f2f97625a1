import sys, json, os
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
K = int(sys.argv[1]); kind = sys.argv[2] if len(sys.argv) > 2 else "ref"
print(json.dumps(mb.time_pava(K, 2 * 10 ** 7 // K, kind, reps=3)))
