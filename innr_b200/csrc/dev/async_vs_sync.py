"""dev: host-buffer calls, synchronous against asynchronous (two in flight), per workload."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import synth, stream
ib.init(0)
def bench(name, sync, submit, reps):
    for i in range(5): sync(i)
    t0 = time.perf_counter()
    for i in range(reps): sync(i)
    ts = (time.perf_counter() - t0) / reps
    pend = None
    for i in range(5):
        t = submit(i)
        if pend: pend.wait()
        pend = t
    pend.wait(); pend = None
    t0 = time.perf_counter()
    for i in range(reps):
        t = submit(i)
        if pend: pend.wait()
        pend = t
    pend.wait()
    ta = (time.perf_counter() - t0) / reps
    print("%-8s sync %.4f ms   async %.4f ms" % (name, ts * 1e3, ta * 1e3), flush=True)
p = ib.QuantizationParams.from_range(-1.0, 1.0)
u8 = ib.U8Corpus.generate(synth.SALT_CORPUS, 0, 50_000_000, 384, p)
q8 = synth.ghash_f32(synth.SALT_QUERY, 0, 16 * 384).reshape(16, 384)
bench("u8", lambda i: ib.batch_knn_u8_many(q8[i % 16], u8, 10), lambda i: stream.submit_knn_u8(q8[i % 16], u8, 10), 60)
del u8
bc = ib.BinaryCorpus.generate(synth.SALT_CODES, 0, 100_000_000, 1024)
qw = synth.ghash_u64(synth.SALT_QUERY, 0, 16 * 16).reshape(16, 16)
bench("hamming", lambda i: ib.hamming_topk_many(qw[i % 16], bc, 100), lambda i: stream.submit_hamming_topk(qw[i % 16], bc, 100), 60)
del bc
db = ib.DeviceBatch.generate("gref", 0, 0, 10_000, 128)
qs = np.random.default_rng(0).standard_normal((100, 128)).astype(np.float32)
bench("C1", lambda i: ib.batch_knn_many("dot", qs, db, 10), lambda i: stream.submit_knn("dot", qs, db, 10), 2000)
