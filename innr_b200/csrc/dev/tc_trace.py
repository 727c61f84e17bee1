"""dev: phase times of the tensor-core filter path (INNR_KNN_TC_TRACE=1) for a few batch sizes."""
import sys, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import synth, sharded
ib.init(0)
n, d = 10_000_000, 768
shard = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, 0, n, d)
import os
metric = os.environ.get("METRIC", "cosine")
sk = sharded.ShardedKnn(shard, "f32", metric)
for nq in (16, 64, 1024):
    qs = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, nq * d).reshape(nq, d)).cuda()
    for _ in range(3): sk.knn_dev(qs, nq, 10)
    torch.cuda.synchronize()
    print("nq", nq, file=sys.stderr)
