"""dev: Hamming top-100 over 100M x 1024-bit codes, time per call for 1 / 2 / 4 / 8 queries."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import _lib as L, synth, sharded
ib.init(0)
shard = ib.BinaryCorpus.generate(synth.SALT_CODES, 0, 100_000_000, 1024)
sk = sharded.ShardedKnn(shard, "binary")
for nq, k in ((1, 100), (2, 100), (4, 100), (8, 100), (4, 10), (8, 10)):
    q = torch.from_numpy(synth.ghash_u64(synth.SALT_QUERY, 0, nq * 16).reshape(nq, 16).view(np.int64)).cuda()
    for _ in range(3): sk.knn_dev(q, nq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): sk.knn_dev(q, nq, k)
    e1.record(); torch.cuda.synchronize()
    print("hamming top-%d nq" % k, nq, "ms per call %.3f" % (e0.elapsed_time(e1) / 10))
