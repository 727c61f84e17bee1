#!/usr/bin/env python
"""bench.py -- headline benchmark of the innr batch similarity-search hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" is one pass of the hot path over one batch of synthetic input. The default workload is the configuration
BASELINE.json's metric is quoted on: batch_knn_cosine top-10 over a 10M x 768 f32 corpus, one query per step,
row-sharded over the N GPUs (one allgather of k keys + merge per query). Prints ONE JSON line on rank 0.

  value     queries/s (docs/s for maxsim) of the whole job, inputs resident in HBM, CUDA events, max over ranks
  e2e       the same metric through the public API with HOST buffers (pinned H2D of the query, D2H of the result)
  roofline  dominant kernel: algorithmic bytes per launch / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline  the oracle (C++ restatement of innr 0.6.3, "port") timed on this box's host cores on a bounded sample

Workloads: knn_cosine_1q (default, C2a) | knn_cosine_multi (C2b, --queries Q) | batch_demo (C1) | maxsim (C3) |
hamming (C4) | u8 (C5) | knn_cosine_1q_filter (C2a through the f16 filter path, an option). `--scale f` shrinks the corpus (for quick checks; reported in config).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, metric name, unit)
    "knn_cosine_1q": ("batch_knn_cosine 10M x 768 f32, 1 query/step, k=10", "batch_knn_cosine_top10_queries_per_s", "queries/s"),
    "knn_cosine_multi": ("batch_knn_cosine 10M x 768 f32, Q queries/step, k=10", "batch_knn_cosine_top10_queries_per_s", "queries/s"),
    "knn_cosine_1q_filter": ("batch_knn_cosine 10M x 768 f32, 1 query/step, k=10, through the f16 tensor-core filter + exact rescoring (option knn_tc_min_queries=1; same bits as the scan)", "batch_knn_cosine_top10_queries_per_s", "queries/s"),
    "batch_demo": ("batch_knn_dot 10K x 128 f32 G-ref lattice, 100 queries/step, k=10", "batch_knn_dot_top10_queries_per_s", "queries/s"),
    "maxsim": ("maxsim_cosine 32 x 128 query tokens vs 1M docs x 180 tokens x 128d", "maxsim_cosine_docs_per_s", "docs/s"),
    "hamming": ("binary_hamming top-100 over 100M 1024-bit codes, 1 query/step", "hamming_top100_queries_per_s", "queries/s"),
    "u8": ("batch_knn_u8 50M x 384 u8 corpus, f32 query, k=10, 1 query/step", "batch_knn_u8_top10_queries_per_s", "queries/s"),
}


def load_traffic(workload, scale, world):
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/), only
    meaningful for the exact captured configuration (scale 1.0, 1 GPU); else null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if scale != 1.0 or world != 1 or not os.path.exists(p):
        return None
    return json.load(open(p)).get(workload, {}).get("bytes")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    """Dense 16-bit tensor peak (TFLOP/s) for a kernel timed inside a long step: the sustained cuBLAS bf16 figure of
    MEASURED_PEAKS.json (f16 and bf16 issue at the same rate), else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured sustained cuBLAS bf16 (MEASURED_PEAKS.json)"
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "measured burst cuBLAS bf16 (MEASURED_PEAKS.json)"
    return 1500.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + clocks-event (throttle) reasons sampled DURING the timed region. NVML in a thread every ~2 ms
    (a 4 ms step needs a faster sampler than `nvidia-smi -lms 100`); nvidia-smi (the B200_PROFILING.md clocks line)
    is the fallback when NVML cannot be loaded. Only samples taken between mark_begin() and mark_end() are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.stop_flag, self.t = gpu_index, [], False, None
        self.t_begin = self.t_end = None
        self.mode, self.proc, self.max_mhz = None, None, None

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self._nvml_index())], stdout=subprocess.PIPE, text=True)
            self.mode = "nvidia-smi"
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                try:
                    watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                except Exception:
                    watts = None
                self.samples.append((time.perf_counter(), mhz, tuple(n for n, bit in names if mask & bit), watts))
            except Exception:
                pass
            time.sleep(0.002)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                mhz, self.max_mhz = float(f[1]), float(f[2])
            except ValueError:
                continue
            rs = tuple(n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                       if v.lower().startswith("active"))
            try:
                watts = float(f[3])
            except ValueError:
                watts = None
            self.samples.append((time.perf_counter(), mhz, rs, watts))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampler unavailable"]}
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.t:
            self.t.join(timeout=2)
        lo, hi = self.t_begin or 0.0, self.t_end or float("inf")
        inside = [s for s in self.samples if lo <= s[0] <= hi]
        sm = [s[1] for s in inside]
        reasons = sorted({r for s in inside for r in s[2]})
        watts = [s[3] for s in inside if s[3] is not None]
        return {"power_w_max": max(watts) if watts else None,
                "sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": self.max_mhz, "samples": len(sm), "reasons": reasons, "sampler": self.mode,
                "window_ms": (hi - lo) * 1e3 if self.t_end else None}


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One config of BASELINE.json on this rank's shard. Subclasses fill: setup(), step_dev(i), step_e2e(i),
    units_per_step, kernel_bytes (algorithmic bytes of the dominant kernel per launch on THIS rank), launches."""

    def __init__(self, args, rank, world, torch):
        self.args, self.rank, self.world, self.torch = args, rank, world, torch

    @property
    def dev(self):
        return self.torch.device(f"cuda:{self.torch.cuda.current_device()}")

    def pinned(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        return t, t.numpy()

    def fetch(self, idx, score):
        """D2H of a result pair into pinned host tensors with ONE stream synchronisation."""
        key = (tuple(idx.shape), idx.dtype, score.dtype)
        if getattr(self, "_fetch_key", None) != key:
            self._fetch_key = key
            self._h_idx = self.torch.empty(idx.shape, dtype=idx.dtype).pin_memory()
            self._h_sc = self.torch.empty(score.shape, dtype=score.dtype).pin_memory()
        self._h_idx.copy_(idx, non_blocking=True)
        self._h_sc.copy_(score, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        return self._h_idx, self._h_sc


class KnnF32(Workload):
    def setup(self):
        import innr_b200 as ib
        from innr_b200 import sharded, synth
        a = self.args
        demo = a.workload == "batch_demo"
        self.n = int((10_000 if demo else 10_000_000) * a.scale)
        self.d = 128 if demo else 768
        self.k = 10
        self.metric = "dot" if demo else "cosine"
        self.nq = 100 if demo else (a.queries if a.workload == "knn_cosine_multi" else 1)
        lo, hi = sharded.shard_range(self.n, self.rank, self.world)
        self.n_local = hi - lo
        gen, salt = ("gref", 0) if demo else ("ghash", synth.SALT_CORPUS)
        self.shard = ib.DeviceBatch.generate(gen, salt, lo, self.n_local, self.d, index_base=lo)
        self.sk = sharded.ShardedKnn(self.shard, "f32", self.metric)
        n_distinct = 16
        if demo:
            from innr_b200.synth import ghash_f32  # noqa: F401
            qs = np.stack([self._gref(self.d, 50_000 + j) for j in range(self.nq)])[None].repeat(n_distinct, 0)
        else:
            qs = synth.ghash_f32(synth.SALT_QUERY, 0, n_distinct * self.nq * self.d).reshape(n_distinct, self.nq, self.d)
        self.q_host_t, self.q_host = self.pinned(qs)
        self.q_dev = self.torch.from_numpy(qs).to(self.dev)
        self.units_per_step = self.nq
        passes = (self.nq + 7) // 8 if self.nq > 1 else 1
        self.kernel_bytes = self.n_local * self.d * 4 * passes
        self.kernel_name = "pdx_scan_kernel"
        if a.workload == "knn_cosine_1q_filter":
            # the bytes the filter actually streams: the f16 unit-vector copy (the f32 corpus is only touched by the
            # ~200 rescored rows); reported against HBM like the scan
            ib.set_option("knn_tc_min_queries", 1)
            self.kernel_bytes = self.n_local * self.d * 2
            self.kernel_name = "knn_tc_filter_kernel<QRES> (4 passes + exact rescoring: whole call)"
        self.launches = passes + 1  # scan launch(es) + merge/decode
        self.h2d, self.d2h = self.nq * self.d * 4, self.nq * self.k * 12
        self.corpus_gb = self.n * self.d * 4 / 1e9

    @staticmethod
    def _gref(dim, seed):
        i = np.arange(dim, dtype=np.uint64)
        with np.errstate(over="ignore"):
            x = np.uint64(seed) * np.uint64(6364136223846793005) + i * np.uint64(1442695040888963407)
        return ((x >> np.uint64(33)).astype(np.float32) / np.float32(2**31) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)

    def step_dev(self, i):
        return self.sk.knn_dev(self.q_dev[i % self.q_dev.shape[0]], self.nq, self.k)

    def step_e2e(self, i):
        if self.world == 1:  # the C-ABI call with host buffers (pinned query; keys come back through pinned staging)
            import innr_b200 as ib
            return ib.batch_knn_many(self.metric, self.q_host[i % self.q_host.shape[0]], self.shard, self.k)
        q = self.q_host_t[i % self.q_host_t.shape[0]]
        dq = q.to(self.dev, non_blocking=True)
        idx, sc = self.sk.knn_dev(dq, self.nq, self.k)
        return self.fetch(idx, sc)

    def cpu_baseline(self, cores, budget_queries=None):
        from oracle import innr_oracle as orc
        from innr_b200 import synth
        a = self.args
        n_s = min(self.n, 10_000 if a.workload == "batch_demo" else (200_000 if budget_queries else 1_000_000))
        if getattr(self, "_cpu_sample", (None,))[0] != n_s:  # built once, reused by every step of the reference arm
            rows = (np.stack([orc.generate_embedding(self.d, i) for i in range(n_s)]) if a.workload == "batch_demo"
                    else orc.ghash_f32(synth.SALT_CORPUS, 0, n_s * self.d).reshape(n_s, self.d))
            self._cpu_sample = (n_s, orc.VerticalBatch.from_flat(rows.reshape(-1), n_s, self.d))
        ob = self._cpu_sample[1]
        nq = budget_queries or (4096 * cores if a.workload == "batch_demo" else 16 * cores)
        qs = np.ascontiguousarray(self.q_host.reshape(-1, self.d)[:1].repeat(nq, 0))
        t0 = time.perf_counter()
        orc.batch_knn_many(self.metric, qs, ob, self.k, n_threads=cores)
        dt = time.perf_counter() - t0
        return nq / dt * (n_s / self.n), f"{nq} queries over the first {n_s} of {self.n} rows x {self.d}, one query per thread; rate scaled by {n_s}/{self.n} (path is linear in N)", dt


class Hamming(Workload):
    def setup(self):
        import innr_b200 as ib
        from innr_b200 import sharded, synth
        self.n, self.dim, self.k, self.nq = int(100_000_000 * self.args.scale), 1024, 100, 1
        lo, hi = sharded.shard_range(self.n, self.rank, self.world)
        self.n_local = hi - lo
        self.shard = ib.BinaryCorpus.generate(synth.SALT_CODES, lo, self.n_local, self.dim, index_base=lo)
        self.sk = sharded.ShardedKnn(self.shard, "binary")
        qs = synth.ghash_u64(synth.SALT_QUERY, 0, 16 * 16).reshape(16, 16).view(np.int64)
        self.q_host_t, self.q_host = self.pinned(qs)
        self.q_dev = self.torch.from_numpy(qs).to(self.dev)
        self.units_per_step = 1
        self.kernel_bytes = self.n_local * 128
        self.kernel_name = "hamming_kernel"
        self.launches = 2
        self.h2d, self.d2h = 128, self.k * 16
        self.corpus_gb = self.n * 128 / 1e9

    def step_dev(self, i):
        return self.sk.knn_dev(self.q_dev[i % 16], 1, self.k)

    def step_e2e(self, i):
        if self.world == 1:
            import innr_b200 as ib
            return ib.hamming_topk_many(self.q_host[i % 16].view(np.uint64).reshape(1, -1), self.shard, self.k)
        dq = self.q_host_t[i % 16].to(self.dev, non_blocking=True)
        idx, ds = self.sk.knn_dev(dq, 1, self.k)
        return self.fetch(idx, ds)

    def cpu_baseline(self, cores, budget_queries=None):
        from oracle import innr_oracle as orc
        from innr_b200 import synth
        n_s = min(self.n, 2_000_000 if budget_queries else 10_000_000)
        if getattr(self, "_cpu_sample", (None,))[0] != n_s:
            self._cpu_sample = (n_s, orc.ghash_u64(synth.SALT_CODES, 0, n_s * 16).reshape(n_s, 16))
        codes = self._cpu_sample[1]
        nq = budget_queries or 8 * cores
        qs = np.ascontiguousarray(self.q_host.view(np.uint64)[:1].repeat(nq, 0))
        t0 = time.perf_counter()
        orc.hamming_topk_many(qs, codes, self.k, n_threads=cores)
        dt = time.perf_counter() - t0
        return nq / dt * (n_s / self.n), f"{nq} queries over the first {n_s} of {self.n} codes; rate scaled by {n_s}/{self.n}", dt


class U8(Workload):
    def setup(self):
        import innr_b200 as ib
        from innr_b200 import sharded, synth
        self.n, self.d, self.k, self.nq = int(50_000_000 * self.args.scale), 384, 10, 1
        lo, hi = sharded.shard_range(self.n, self.rank, self.world)
        self.n_local = hi - lo
        self.params = ib.QuantizationParams.from_range(-1.0, 1.0)
        self.shard = ib.U8Corpus.generate(synth.SALT_CORPUS, lo, self.n_local, self.d, self.params, index_base=lo)
        self.sk = sharded.ShardedKnn(self.shard, "u8")
        qs = synth.ghash_f32(synth.SALT_QUERY, 0, 16 * self.d).reshape(16, self.d)
        self.q_host_t, self.q_host = self.pinned(qs)
        self.q_dev = self.torch.from_numpy(qs).to(self.dev)
        self.units_per_step = 1
        self.kernel_bytes = self.n_local * self.d
        self.kernel_name = "u8_scan_kernel"
        self.launches = 2
        self.h2d, self.d2h = self.d * 4, self.k * 12
        self.corpus_gb = self.n * self.d / 1e9

    def step_dev(self, i):
        return self.sk.knn_dev(self.q_dev[i % 16], 1, self.k)

    def step_e2e(self, i):
        if self.world == 1:
            import innr_b200 as ib
            return ib.batch_knn_u8_many(self.q_host[i % 16].reshape(1, -1), self.shard, self.k)
        dq = self.q_host_t[i % 16].to(self.dev, non_blocking=True)
        idx, sc = self.sk.knn_dev(dq, 1, self.k)
        return self.fetch(idx, sc)

    def cpu_baseline(self, cores, budget_queries=None):
        from oracle import innr_oracle as orc
        from innr_b200 import synth
        n_s = min(self.n, 500_000 if budget_queries else 2_000_000)
        p = orc.QuantizationParams.from_range(-1.0, 1.0)
        if getattr(self, "_cpu_sample", (None,))[0] != n_s:
            self._cpu_sample = (n_s, orc.quantize_u8(orc.ghash_f32(synth.SALT_CORPUS, 0, n_s * self.d), p).data.reshape(n_s, self.d))
        mat = self._cpu_sample[1]
        nq = budget_queries or 32 * cores
        qs = np.ascontiguousarray(self.q_host[:1].repeat(nq, 0))
        t0 = time.perf_counter()
        orc.batch_knn_u8_many(qs, mat, p, self.k, n_threads=cores)
        dt = time.perf_counter() - t0
        return nq / dt * (n_s / self.n), f"{nq} queries over the first {n_s} of {self.n} rows x {self.d}; rate scaled by {n_s}/{self.n}", dt


class MaxSim(Workload):
    def setup(self):
        import innr_b200 as ib
        from innr_b200 import sharded, synth
        self.n, self.nt, self.dim, self.nq = int(1_000_000 * self.args.scale), 180, 128, 32
        lo, hi = sharded.shard_range(self.n, self.rank, self.world)
        self.n_local = hi - lo
        self.shard = ib.TokenCorpus.generate(synth.SALT_CORPUS, lo, self.n_local, self.nt, self.dim, index_base=lo)
        qs = synth.ghash_f32(synth.SALT_QUERY, 0, 4 * self.nq * self.dim).reshape(4, self.nq, self.dim)
        self.q_host_t, self.q_host = self.pinned(qs)
        self.q_dev = self.torch.from_numpy(qs).to(self.dev)
        self.out = self.torch.empty(self.n_local, dtype=self.torch.float32, device=self.dev)
        self.out_host_t = self.torch.empty(self.n_local, dtype=self.torch.float32).pin_memory()  # caller-owned result buffer
        self.out_host = self.out_host_t.numpy()
        self.units_per_step = self.n  # docs scored per step by the whole job
        self.kernel_bytes = self.n_local * self.nt * self.dim * 4
        self.kernel_name = "maxsim_tc_kernel"
        self.launches = 1
        self.h2d, self.d2h = self.nq * self.dim * 4, self.n_local * 4
        self.corpus_gb = self.n * self.nt * self.dim * 4 / 1e9

    def step_dev(self, i):
        from innr_b200 import _lib as L
        s = C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        L.call("innr_cuda_maxsim_dev", self.shard.h, C.c_void_p(self.q_dev[i % 4].data_ptr()), self.nq, 1,
               C.c_void_p(self.out.data_ptr()), s)
        return self.out

    def step_e2e(self, i):
        import innr_b200 as ib
        return ib.maxsim_corpus(self.q_host[i % 4], self.shard, cosine=True, out=self.out_host)

    def cpu_baseline(self, cores, budget_queries=None):
        from oracle import innr_oracle as orc
        from innr_b200 import synth
        n_s = min(self.n, (250 if budget_queries else 2000) * cores)
        reps = 1 if budget_queries else 24
        if getattr(self, "_cpu_sample", (None,))[0] != n_s:
            self._cpu_sample = (n_s, orc.ghash_f32(synth.SALT_CORPUS, 0, n_s * self.nt * self.dim).reshape(n_s * self.nt, self.dim))
        toks = self._cpu_sample[1]
        off = np.arange(0, n_s * self.nt + 1, self.nt, dtype=np.uint64)
        t0 = time.perf_counter()
        for r in range(reps):
            orc.maxsim_corpus(self.q_host[r % self.q_host.shape[0]], toks, off, cosine_flag=True, n_threads=cores)
        dt = time.perf_counter() - t0
        return n_s * reps / dt, f"{n_s} docs x {self.nt} tokens x {self.dim}d scored {reps}x, docs split over threads", dt


def make_workload(args, rank, world, torch):
    cls = {"knn_cosine_1q": KnnF32, "knn_cosine_1q_filter": KnnF32, "knn_cosine_multi": KnnF32, "batch_demo": KnnF32, "maxsim": MaxSim,
           "hamming": Hamming, "u8": U8}[args.workload]
    return cls(args, rank, world, torch)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    """The reference's own CPU implementation of the path (oracle port: the Rust crate cannot be built here) on this
    box's host cores, all threads, on a bounded sample per step."""
    if rank != 0:
        return
    desc, metric, unit = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1

    w = make_workload(args, 0, 1, None)  # no device: only the query generators and cpu_baseline() are used
    # queries only (no device): reuse the generators directly
    from innr_b200 import synth
    if isinstance(w, KnnF32):
        demo = args.workload == "batch_demo"
        w.n, w.d, w.k = int((10_000 if demo else 10_000_000) * args.scale), (128 if demo else 768), 10
        w.metric = "dot" if demo else "cosine"
        w.q_host = (np.stack([KnnF32._gref(w.d, 50_000 + j) for j in range(4)]) if demo
                    else synth.ghash_f32(synth.SALT_QUERY, 0, 4 * w.d).reshape(4, w.d))
    elif isinstance(w, Hamming):
        w.n, w.k = int(100_000_000 * args.scale), 100
        w.q_host = synth.ghash_u64(synth.SALT_QUERY, 0, 16 * 16).reshape(16, 16).view(np.int64)
    elif isinstance(w, U8):
        w.n, w.d, w.k = int(50_000_000 * args.scale), 384, 10
        w.q_host = synth.ghash_f32(synth.SALT_QUERY, 0, 16 * w.d).reshape(16, w.d)
    else:
        w.n, w.nt, w.dim, w.nq = int(1_000_000 * args.scale), 180, 128, 32
        w.q_host = synth.ghash_f32(synth.SALT_QUERY, 0, 4 * w.nq * w.dim).reshape(4, w.nq, w.dim)
    for _ in range(args.warmup):
        w.cpu_baseline(cores, budget_queries=cores)
    vals, total = [], 0.0
    sample = ""
    for _ in range(args.steps):
        v, sample, dt = w.cpu_baseline(cores, budget_queries=cores)
        vals.append(v)
        total += dt
    value = len(vals) / sum(1.0 / v for v in vals)  # harmonic mean == total units / total time
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype_of(args.workload),
            "data": "synthetic", "config": {"workload": desc, "scale": args.scale},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                             "sample": sample + " (C++ restatement of innr 0.6.3, not the Rust crate)"},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def dtype_of(workload):
    return {"hamming": "u64", "u8": "f32xu8"}.get(workload, "f32")


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="knn_cosine_1q", choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=1024, help="queries per step for knn_cosine_multi")
    ap.add_argument("--scale", type=float, default=1.0, help="corpus size multiplier (1.0 = BASELINE.json size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import innr_b200 as ib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: innr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ib.init(local_rank)

    desc, metric, unit = WORKLOADS[args.workload]
    w = make_workload(args, rank, world, torch)
    w.setup()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timed region ---------------------------------------------------------------
    for i in range(args.warmup):
        w.step_dev(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.mark_begin()
    ev[0].record()
    for i in range(args.steps):
        kev[i][0].record()
        w.step_dev(i)
        kev[i][1].record()
    ev[1].record()
    barrier()
    sampler.mark_end()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
    launches = ib.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in kev]
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel alone (N == 1: the step IS one scan launch + a 1-warp merge) ------------------
    # kernel time = CUDA events around the library's scan launch only, on the launching stream
    from innr_b200 import _lib as L
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    kern_ms = None
    if hasattr(w, "sk"):
        b = w.sk._buffers(w.nq, w.k, w.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for i in range(min(args.steps, 20)):
            q = w.q_dev[i % w.q_dev.shape[0]]
            e0.record()
            if w.sk.kind == "f32":
                L.call("innr_cuda_batch_knn_keys_dev", w.shard.h, w.sk._metric_id, C.c_void_p(q.data_ptr()), w.nq, w.k,
                       C.c_void_p(b["local"].data_ptr()), stream)
            elif w.sk.kind == "u8":
                L.call("innr_cuda_batch_knn_u8_keys_dev", w.shard.h, C.c_void_p(q.data_ptr()), w.nq, w.k,
                       C.c_void_p(b["local"].data_ptr()), stream)
            else:
                L.call("innr_cuda_hamming_topk_keys_dev", w.shard.h, C.c_void_p(q.data_ptr()), w.nq, w.k,
                       C.c_void_p(b["local"].data_ptr()), stream)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        kern_ms = sum(times) / len(times)
    else:
        kern_ms = sum(step_ms) / len(step_ms)
    barrier()

    # ---- end-to-end through the public API with host buffers ---------------------------------------------
    for i in range(min(args.warmup, 3)):
        w.step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        w.step_e2e(i)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()

    value = w.units_per_step * args.steps / (total_ms / 1e3)
    e2e_value = w.units_per_step * args.steps / e2e_s
    peak, peak_src = load_peaks()
    achieved = w.kernel_bytes / (kern_ms / 1e3) / 1e9

    roofline = {"bound": "hbm", "kernel": w.kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": load_traffic(args.workload, args.scale, world),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(w.kernel_bytes), "kernel_ms": kern_ms}
    tc = ib.knn_tc_last_stats() if args.workload == "knn_cosine_multi" else None
    if tc and tc["passes"] > 0 and tc["filter_ms"] > 0:
        # large query batches: the dominant kernel is the tcgen05 filter (csrc/knn_tc.cu); algorithmic flops per launch =
        # 2 * rows * d * queries of THIS rank, time = CUDA events around that launch inside the library
        tpeak, tsrc = load_tensor_peak()
        flops = 2.0 * w.n_local * w.d * w.nq
        tf = flops / (tc["filter_ms"] / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "knn_tc_filter_kernel (kind::f16, exact rescoring after it)",
                    "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "traffic": None,
                    "peak_source": tsrc, "algorithmic_flops_per_launch": flops, "kernel_ms": tc["filter_ms"],
                    "call_ms": tc["total_ms"], "filter_passes": tc["passes"], "rescored_pairs": tc["candidates"],
                    "exact_scan_queries": tc["exact_scan_queries"]}
    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype_of(args.workload), "data": "synthetic",
            "config": {"workload": desc, "scale": args.scale, "corpus_gb": round(w.corpus_gb, 3),
                       "sharding": f"rows/{world}", "l2": "inputs larger than L2 (corpus >> 126 MB), no flush"
                       if w.corpus_gb / world > 0.5 else "corpus is L2-resident by design of this config (latency-bound)",
                       "generator": "G-ref lattice" if args.workload == "batch_demo" else "G-hash (splitmix64)"},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, sample, _ = w.cpu_baseline(cores)
            line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                    "sample": sample + " (C++ restatement of innr 0.6.3, not the Rust crate)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
