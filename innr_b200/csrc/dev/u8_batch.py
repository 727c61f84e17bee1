"""dev: batch_knn_u8 over 50M x 384, k = 10: time per call for 1 / 2 / 4 / 8 queries (two queries per pass)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import synth, sharded
ib.init(0)
p = ib.QuantizationParams.from_range(-1.0, 1.0)
shard = ib.U8Corpus.generate(synth.SALT_CORPUS, 0, 50_000_000, 384, p)
sk = sharded.ShardedKnn(shard, "u8")
for nq in (1, 2, 4, 8):
    q = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, nq * 384).reshape(nq, 384)).cuda()
    for _ in range(3): sk.knn_dev(q, nq, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): sk.knn_dev(q, nq, 10)
    e1.record(); torch.cuda.synchronize()
    print("u8 top-10 nq", nq, "ms per call %.3f = %.3f per query" % (e0.elapsed_time(e1) / 20, e0.elapsed_time(e1) / 20 / nq))
