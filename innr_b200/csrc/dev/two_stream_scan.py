"""dev: consecutive shard scans on ONE stream against TWO alternating streams, on a 1/8 shard of C2a / C4. The library
keeps two workspaces per device (api.cu lanes), so the two-stream form really overlaps: the head of scan i + 1 fills the
SMs under the tail of scan i. Measured: f32 0.5565 -> 0.5270 ms per scan, Hamming 0.3030 -> 0.2442 ms."""
import sys, ctypes as C, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import _lib as L, synth
ib.init(0)
for kind in ("f32", "binary"):
    if kind == "f32":
        shard = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, 0, 1_250_000, 768)
        q = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, 16 * 768).reshape(16, 768)).cuda()
        k = 10
        call = lambda i, out, st: L.call("innr_cuda_batch_knn_keys_dev", shard.h, L.METRIC_COSINE, C.c_void_p(q[i % 16].data_ptr()), 1, k, C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream))
    else:
        shard = ib.BinaryCorpus.generate(synth.SALT_CODES, 0, 12_500_000, 1024)
        import numpy as np
        q = torch.from_numpy(synth.ghash_u64(synth.SALT_QUERY, 0, 16 * 16).reshape(16, 16).view(np.int64)).cuda()
        k = 100
        call = lambda i, out, st: L.call("innr_cuda_hamming_topk_keys_dev", shard.h, C.c_void_p(q[i % 16].data_ptr()), 1, k, C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream))
    outs = [torch.empty(k, dtype=torch.int64, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for n_streams in (1, 2):
        for i in range(20):
            call(i, outs[i % 2], streams[i % n_streams])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        streams[1].wait_event(e0)
        steps = 400
        for i in range(steps):
            call(i, outs[i % 2], streams[i % n_streams])
        if n_streams == 2:
            ev = torch.cuda.Event(); ev.record(streams[1]); streams[0].wait_event(ev)
        e1.record(streams[0])
        torch.cuda.synchronize()
        print(kind, "streams", n_streams, "ms per scan %.4f" % (e0.elapsed_time(e1) / steps))
    del shard
