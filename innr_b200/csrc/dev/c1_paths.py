"""dev: C1 (10K x 128 lattice, 100 queries, top-10): time per call on the 8-query scan and on the tensor filter."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import sharded
ib.init(0)
n, d, nq = 10_000, 128, 100
shard = ib.DeviceBatch.generate("gref", 0, 0, n, d)
def gref(dim, seed):
    i = np.arange(dim, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(seed) * np.uint64(6364136223846793005) + i * np.uint64(1442695040888963407)
    return ((x >> np.uint64(33)).astype(np.float32) / np.float32(2**31) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
qs = torch.from_numpy(np.stack([gref(d, 50_000 + j) for j in range(nq)])).cuda()
sk = sharded.ShardedKnn(shard, "f32", "dot")
for min_n in (100000, 4096):
    ib.set_option("knn_tc_min_n", min_n)
    for _ in range(5): sk.knn_dev(qs, nq, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): sk.knn_dev(qs, nq, 10)
    e1.record(); torch.cuda.synchronize()
    print("knn_tc_min_n", min_n, "us per 100-query call %.1f" % (e0.elapsed_time(e1) / 50 * 1e3), ib.knn_tc_last_stats() if min_n == 4096 else "")
