"""dev: where do the microseconds of a host-buffer C1 call (100 queries x 10,000 x 128, k = 10) go?"""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import _lib as L, synth
import torch
ib.init(0)
n, d, nq, k = 10_000, 128, 100, 10
db = ib.DeviceBatch.generate("gref", 0, 0, n, d) if hasattr(ib.DeviceBatch, "generate") else None
qs = np.random.default_rng(0).standard_normal((nq, d)).astype(np.float32)
qp = torch.from_numpy(qs).pin_memory().numpy()
def timeit(fn, reps=3000):
    for _ in range(200): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e6
print("batch_knn_many (pageable q)  %.1f us" % timeit(lambda: ib.batch_knn_many("cosine", qs, db, k)))
print("batch_knn_many (pinned q)    %.1f us" % timeit(lambda: ib.batch_knn_many("cosine", qp, db, k)))
idx = np.zeros((nq, k), np.uint64); sc = np.zeros((nq, k), np.float32); cnt = C.c_size_t(0)
args = (db.h, L.METRIC_COSINE, qp.ctypes.data_as(L.f32p), nq, d, k, idx.ctypes.data_as(L.u64p), sc.ctypes.data_as(L.f32p), C.byref(cnt))
print("raw L.call, preallocated     %.1f us" % timeit(lambda: L.call("innr_cuda_batch_knn", *args)))
fn = L.lib().innr_cuda_batch_knn if hasattr(L, "lib") else None
if fn is not None:
    print("raw ctypes fn                %.1f us" % timeit(lambda: fn(*args)))
print("last kernel ms", ib.last_kernel_ms())
print("no-op entry (corpus_info)    %.1f us" % timeit(lambda: L.call("innr_cuda_corpus_info", db.h, None, None, None, None, None, None)))
