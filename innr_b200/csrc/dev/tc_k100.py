"""dev: tensor-filter path with k = 100 at 10M x 768 (candidate counts, overflow fallbacks, time)."""
import sys, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import synth, sharded
ib.init(0)
n, d = 10_000_000, 768
shard = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, 0, n, d)
sk = sharded.ShardedKnn(shard, "f32", "cosine")
for nq, k in ((64, 100), (1024, 100), (1024, 10)):
    qs = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, nq * d).reshape(nq, d)).cuda()
    for _ in range(2): sk.knn_dev(qs, nq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): sk.knn_dev(qs, nq, k)
    e1.record(); torch.cuda.synchronize()
    print("nq", nq, "k", k, "ms per call %.3f" % (e0.elapsed_time(e1) / 3), ib.knn_tc_last_stats())
