"""dev: u8 scan per call -- kernel back to back, pipelined device-resident (two streams), asynchronous host calls."""
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import synth, stream, sharded, _lib as L
ib.init(0)
p = ib.QuantizationParams.from_range(-1.0, 1.0)
u8 = ib.U8Corpus.generate(synth.SALT_CORPUS, 0, 50_000_000, 384, p)
q8 = synth.ghash_f32(synth.SALT_QUERY, 0, 16 * 384).reshape(16, 384)
dq = torch.from_numpy(q8).cuda()
sk = sharded.ShardedKnn(u8, "u8")
def dev_time(fn, reps=60):
    for i in range(5): fn(i)
    sk.drain(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    sk.drain(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("knn_dev (one stream)        %.4f ms" % dev_time(lambda i: sk.knn_dev(dq[i % 16], 1, 10)))
print("knn_dev_pipelined           %.4f ms" % dev_time(lambda i: sk.knn_dev_pipelined(dq[i % 16], 1, 10)))
hq = torch.from_numpy(q8).pin_memory()
print("pipelined, host buffers     %.4f ms" % dev_time(lambda i: sk.knn_dev_pipelined(None, 1, 10, host_queries=hq[i % 16], host_out=True)))
def wall(submit, reps=60):
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
    return (time.perf_counter() - t0) / reps * 1e3
print("async host calls            %.4f ms" % wall(lambda i: stream.submit_knn_u8(q8[i % 16], u8, 10)))
t0 = time.perf_counter()
for i in range(60): ib.batch_knn_u8_many(q8[i % 16], u8, 10)
print("sync host calls             %.4f ms" % ((time.perf_counter() - t0) / 60 * 1e3))
