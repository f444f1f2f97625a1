import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import microbench as mb
K, nb = int(sys.argv[1]), int(sys.argv[2]); kind = sys.argv[3] if len(sys.argv) > 3 else "ref"
print(json.dumps(mb.time_pava(K, nb, kind, reps=3)))
