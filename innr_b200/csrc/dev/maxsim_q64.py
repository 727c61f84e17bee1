"""dev: MaxSim time per call at the C3 corpus for 32 / 64 / 128 query tokens (one pass per 64 tokens)."""
import sys, ctypes as C, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import _lib as L, synth
ib.init(0)
shard = ib.TokenCorpus.generate(synth.SALT_CORPUS, 0, 1_000_000, 180, 128)
out = torch.empty(1_000_000, dtype=torch.float32, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for nq in (32, 64):
    q = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, nq * 128).reshape(nq, 128)).cuda()
    for _ in range(3): L.call("innr_cuda_maxsim_dev", shard.h, C.c_void_p(q.data_ptr()), nq, 1, C.c_void_p(out.data_ptr()), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.call("innr_cuda_maxsim_dev", shard.h, C.c_void_p(q.data_ptr()), nq, 1, C.c_void_p(out.data_ptr()), s)
    e1.record(); torch.cuda.synchronize()
    print("maxsim_cosine n_q", nq, "ms per call %.3f" % (e0.elapsed_time(e1) / 10))
for nb in (2, 8):
    q = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, nb * 32 * 128).reshape(nb, 32, 128)).cuda()
    outb = torch.empty(nb * 1_000_000, dtype=torch.float32, device="cuda")
    for _ in range(2): L.call("innr_cuda_maxsim_batch_dev", shard.h, C.c_void_p(q.data_ptr()), nb, 32, 1, C.c_void_p(outb.data_ptr()), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): L.call("innr_cuda_maxsim_batch_dev", shard.h, C.c_void_p(q.data_ptr()), nb, 32, 1, C.c_void_p(outb.data_ptr()), s)
    e1.record(); torch.cuda.synchronize()
    print("maxsim_cosine batch of", nb, "queries x 32 tokens: ms per call %.3f = %.3f per query" % (e0.elapsed_time(e1) / 5, e0.elapsed_time(e1) / 5 / nb))
