"""dev: fixed cost of a fused top-k launch (small corpora: time is launch ramp + selection + ticketed finish)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import sharded, synth
ib.init(0)

def timeit(sk, q, nq, k, reps=200):
    for _ in range(10): sk.knn_dev(q, nq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): sk.knn_dev(q, nq, k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for n in (20_000, 200_000, 2_000_000, 12_500_000):
    codes = ib.BinaryCorpus.generate(synth.SALT_CODES, 0, n, 1024)
    qw = torch.from_numpy(synth.ghash_u64(synth.SALT_QUERY, 0, 16).view(np.int64)).cuda()
    sk = sharded.ShardedKnn(codes, "binary", "l2")
    for k in (10, 100):
        us = timeit(sk, qw, 1, k)
        print(f"hamming n={n:>9} k={k:>3}: {us:8.1f} us  (stream floor {n*128/6.544e12*1e6:7.1f} us)")
    del codes, sk
for n in (2_000, 20_000, 200_000, 1_250_000):
    b = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, 0, n, 768)
    q = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, 768)).cuda()
    sk = sharded.ShardedKnn(b, "f32", "cosine")
    us = timeit(sk, q, 1, 10)
    print(f"f32 cosine n={n:>9} k= 10: {us:8.1f} us  (stream floor {n*768*4/6.544e12*1e6:7.1f} us)")
    del b, sk
for n in (2_000, 20_000, 200_000, 1_250_000):
    b = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, 0, n, 768)
    q = synth.ghash_f32(synth.SALT_QUERY, 0, 768)
    for name in ("batch_dot", "batch_l2_squared"):
        fn = getattr(ib, name)
        for _ in range(3): fn(q, b)
        ms = []
        for _ in range(20):
            fn(q, b); ms.append(ib.last_kernel_ms())
        print(f"{name} n={n:>9}: kernel {float(np.median(ms))*1e3:8.1f} us  (stream floor {n*768*4/6.544e12*1e6:7.1f} us)")
    del b
