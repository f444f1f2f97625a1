"""Host C ABI on the bench's three arrays: sequential calls against three concurrent calls (one host thread per
array; every calling thread owns its staging buffers and streams inside the library)."""
import json, sys, time, threading
sys.path.insert(0, ".")
import numpy as np, torch, bsls_b200
NB = 10 ** 6
rng = np.random.RandomState(1)
host, src, blocks = {}, {}, {}
for K in (4, 16, 64):
    host[K] = torch.empty(NB * K, dtype=torch.float64).pin_memory()
    src[K] = rng.randn(NB * K)
    blocks[K] = np.arange(0, NB * K, K, dtype=np.int32)
def fill():
    for K in host: host[K].numpy()[:] = src[K]
def seq():
    for K in (4, 16, 64): bsls_b200.proj_multi_simplex_c(host[K].numpy(), blocks[K])
def par():
    ts = [threading.Thread(target=bsls_b200.proj_multi_simplex_c, args=(host[K].numpy(), blocks[K])) for K in (64, 16, 4)]
    for t in ts: t.start()
    for t in ts: t.join()
out = {}
for name, fn in (("sequential", seq), ("three_threads", par), ("sequential_again", seq)):
    best = []
    for it in range(5):
        fill(); t0 = time.perf_counter(); fn(); best.append(time.perf_counter() - t0)
    out[name + "_ms"] = round(1e3 * float(np.median(best[1:])), 3)
want = src[16][:16 * 100].copy()
from oracle import cpu
cpu.port().proj_multi_simplex(want, np.arange(0, 1600, 16))
out["spot_check_ok"] = bool(np.array_equal(host[16].numpy()[:1600], want))
print(json.dumps(out))
