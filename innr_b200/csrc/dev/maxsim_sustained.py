"""dev: sustained (power-capped) rate of the single-query MaxSim launch: 100 launches back to back, CUDA-event time of the
last 80, SM clock sampled through NVML meanwhile. With INNR_MAXSIM_DEBUG timing-only variants the results are wrong by
construction; this script never looks at them."""
import sys, time, threading, ctypes as C, torch
sys.path.insert(0, ".")
import innr_b200 as ib
from innr_b200 import _lib as L, synth
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
ib.init(0)
n_docs = 1_000_000
shard = ib.TokenCorpus.generate(synth.SALT_CORPUS, 0, n_docs, 180, 128)
out = torch.empty(2 * n_docs, dtype=torch.float32, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
q = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, 2 * 32 * 128).reshape(2, 32, 128)).cuda()
pair = len(sys.argv) > 1 and sys.argv[1] == "pair"
def launch():
    if pair:
        L.call("innr_cuda_maxsim_batch_dev", shard.h, C.c_void_p(q.data_ptr()), 2, 32, 1, C.c_void_p(out.data_ptr()), s)
    else:
        L.call("innr_cuda_maxsim_dev", shard.h, C.c_void_p(q.data_ptr()), 32, 1, C.c_void_p(out.data_ptr()), s)
clocks, power, stop = [], [], False
def sample():
    while not stop:
        clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000)
        time.sleep(0.005)
for _ in range(20):
    launch()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
th = threading.Thread(target=sample); th.start()
e0.record()
for _ in range(80):
    launch()
e1.record()
torch.cuda.synchronize()
stop = True; th.join()
ms = e0.elapsed_time(e1) / 80
clocks.sort()
mhz = clocks[len(clocks) // 2]
tiles = 180 * n_docs / 128 / 148
print("ms %.3f  sm_mhz %d  cycles/tile %.0f  power_max %.0f W" % (ms, mhz, ms * 1e-3 * mhz * 1e6 / tiles, max(power)))
