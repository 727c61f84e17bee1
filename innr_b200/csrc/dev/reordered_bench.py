"""dev: batch_dimension_variance (first use, cached afterwards) and batch_knn_reordered vs batch_knn at C2 size."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import innr_b200 as ib
ib.init(0)
n, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000), 768
batch = ib.DeviceBatch.generate("ghash", 0x5EED0000, 0, n, d)
q = np.random.default_rng(0).standard_normal(d).astype(np.float32)
t0 = time.perf_counter(); var = ib.batch_dimension_variance(batch); t1 = time.perf_counter()
print("variance (first use, includes the D2H of d floats): %.1f ms wall, var[0..3] =" % ((t1 - t0) * 1e3), var[:3])
for name in ("batch_knn", "batch_knn_reordered"):
    fn = getattr(ib, name)
    for _ in range(3): r = fn(q, batch, 10)
    ms = []
    for _ in range(10):
        r = fn(q, batch, 10); ms.append(ib.last_kernel_ms())
    print(name, "kernel ms median %.3f" % float(np.median(ms)), r.indices[:4])
a, b = ib.batch_knn(q, batch, 10), ib.batch_knn_reordered(q, batch, 10)
print("same neighbours:", a.indices == b.indices)
for w in (32, 128):
    for _ in range(2): r = ib.batch_knn_adaptive(q, batch, 10, w)
    ms = []
    for _ in range(5):
        r = ib.batch_knn_adaptive(q, batch, 10, w); ms.append(ib.last_kernel_ms())
    print("batch_knn_adaptive warmup", w, "kernel ms median %.3f" % float(np.median(ms)), r.indices[:4],
          "overlap with exact top-10:", len(set(r.indices) & set(a.indices)))
