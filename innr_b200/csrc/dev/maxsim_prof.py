"""dev: a short MaxSim sequence for ncu -- 3 single-query calls (32 tokens), then calls with batches of two and four
queries, at the C3 corpus (1M docs x 180 tokens x 128d). Prints CUDA-event times when not under a profiler."""
import sys, ctypes as C, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import _lib as L, synth
ib.init(0)
n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
shard = ib.TokenCorpus.generate(synth.SALT_CORPUS, 0, n_docs, 180, 128)
out = torch.empty(4 * n_docs, dtype=torch.float32, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
q = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, 4 * 32 * 128).reshape(4, 32, 128)).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
for i in range(3):
    ev[i].record()
    L.call("innr_cuda_maxsim_dev", shard.h, C.c_void_p(q.data_ptr()), 32, 1, C.c_void_p(out.data_ptr()), s)
ev[3].record()
L.call("innr_cuda_maxsim_batch_dev", shard.h, C.c_void_p(q.data_ptr()), 2, 32, 1, C.c_void_p(out.data_ptr()), s)
ev[4].record()
L.call("innr_cuda_maxsim_batch_dev", shard.h, C.c_void_p(q.data_ptr()), 2, 32, 1, C.c_void_p(out.data_ptr()), s)
ev[5].record()
L.call("innr_cuda_maxsim_batch_dev", shard.h, C.c_void_p(q.data_ptr()), 4, 32, 1, C.c_void_p(out.data_ptr()), s)
ev[6].record()
torch.cuda.synchronize()
print("single ms:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(3)])
print("batch of 2: ms", round(ev[3].elapsed_time(ev[4]), 3), round(ev[4].elapsed_time(ev[5]), 3), "| batch of 4: ms", round(ev[5].elapsed_time(ev[6]), 3))
