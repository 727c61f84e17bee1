import sys, time, numpy as np, torch, ctypes as C
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import _lib as L, synth, sharded
ib.init(0)
ib.set_option("knn_tc", 0)
n, d = 10_000_000, 768
shard = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, 0, n, d)
sk = sharded.ShardedKnn(shard, "f32", "cosine")
for nq in (8, 16, 64):
    qs = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, nq * d).reshape(nq, d)).cuda()
    for _ in range(2): sk.knn_dev(qs, nq, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): sk.knn_dev(qs, nq, 10)
    e1.record(); torch.cuda.synchronize()
    print("QB8 path nq", nq, "ms per call", e0.elapsed_time(e1) / 5, "ms per 8 queries", e0.elapsed_time(e1) / 5 / (nq / 8))
