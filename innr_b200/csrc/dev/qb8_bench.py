"""dev: time per call of batch_knn_cosine (10M x 768, k=10) for small query batches on the bit-exact multi-query scan
(QB=8) and on the tensor-core filter path, to place the crossover (option knn_tc_min_queries)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import synth, sharded
ib.init(0)
n, d = 10_000_000, 768
shard = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, 0, n, d)
sk = sharded.ShardedKnn(shard, "f32", "cosine")
for tc in (0, 1):
    ib.set_option("knn_tc", tc)
    ib.set_option("knn_tc_min_queries", 2)
    for nq in (2, 4, 8, 16, 32, 64):
        qs = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, nq * d).reshape(nq, d)).cuda()
        for _ in range(2): sk.knn_dev(qs, nq, 10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): sk.knn_dev(qs, nq, 10)
        e1.record(); torch.cuda.synchronize()
        print("tensor filter" if tc else "QB8 scan     ", "nq", nq, "ms per call %.3f" % (e0.elapsed_time(e1) / 5))
