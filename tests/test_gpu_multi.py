"""Tests that need TWO OR MORE GPUs in one box (gpurun --gpus 2): the peer-mapped exchange between devices of one process,
between processes (CUDA IPC), and the C-ABI's own sharded entries. On a 1-GPU box they skip with an explicit reason."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


needs2 = pytest.mark.skipif("_n_gpus() < 2", reason="needs >= 2 GPUs in one box (run with gpurun --gpus 2)")


@needs2
def test_exchange_between_processes_equals_nccl_and_oracle():
    n = min(_n_gpus(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mp_exchange_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mp_exchange_check ok" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@needs2
def test_exchange_between_devices_of_one_process(oracle):
    """One host thread drives two devices: each rank's kernel waits for the other's flag, which arrives as soon as the
    other device's launch (queued right behind it by the same thread) has published."""
    import ctypes as C
    import torch
    import innr_b200 as ib
    from innr_b200 import _lib as L, sharded
    world, n, d, k, nq = 2, 40_000, 64, 10, 4
    rng = np.random.default_rng(8)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    qs = rng.standard_normal((nq, d)).astype(np.float32)
    ob = oracle.VerticalBatch.from_flat(rows.reshape(-1), n, d)
    shards, exs, dqs, outs = [], [], [], []
    for r in range(world):
        ib.init(r)
        lo, hi = sharded.shard_range(n, r, world)
        shards.append(ib.DeviceBatch.from_rows_flat(rows[lo:hi].reshape(-1), hi - lo, d, index_base=lo))
        exs.append(sharded.PeerExchange(world, r))
        with torch.cuda.device(r):
            dqs.append(torch.from_numpy(qs).cuda())
            outs.append((torch.empty(nq * k, dtype=torch.int64, device=f"cuda:{r}"), torch.empty(nq * k, dtype=torch.int64, device=f"cuda:{r}"),
                         torch.empty(nq * k, dtype=torch.float32, device=f"cuda:{r}")))
    sharded.PeerExchange.connect_local(exs)
    for rep in range(4):
        for r in range(world):
            with torch.cuda.device(r):
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                L.call("innr_cuda_batch_knn_keys_dev", shards[r].h, L.METRIC_COSINE, C.c_void_p(dqs[r].data_ptr()), nq, k,
                       C.c_void_p(outs[r][0].data_ptr()), st)
                exs[r].merge_dev(outs[r][0].data_ptr(), nq, k, L.METRIC_DOT, st, idx=outs[r][1], score=outs[r][2])
        for r in range(world):
            torch.cuda.synchronize(r)
            assert exs[r].status() == 0
            idx, sc = outs[r][1].cpu().numpy().reshape(nq, k), outs[r][2].cpu().numpy().reshape(nq, k)
            for j in range(nq):
                w = oracle.batch_knn_cosine(qs[j], ob, k)
                assert idx[j].tolist() == w.indices and sc[j].tobytes() == w.scores.tobytes(), (rep, r, j)
    ib.init(0)


@needs2
def test_c_abi_sharded_entries_over_devices(oracle):
    """innr_cuda_*_sharded with one shard per DEVICE: persistent worker threads enqueue the shard scans, the lists meet in
    the root device's mailbox (peer-mapped, csrc/exchange.cu) and are merged in the same launch. Results equal the
    unsharded oracle answer, repeatedly (both mailbox parities), also with an empty shard and with requests that do
    not fit the mailbox route (k > 128 -> host merge)."""
    import torch
    import innr_b200 as ib
    from innr_b200 import sharded
    n_dev = min(torch.cuda.device_count(), 4)
    n, d, nq = 30_001, 40, 5
    rng = np.random.default_rng(14)
    rows = rng.integers(-3, 4, size=(n, d)).astype(np.float32)   # heavy ties across shard boundaries
    ob = oracle.VerticalBatch.from_flat(rows.reshape(-1), n, d)
    cuts = [n * r // n_dev for r in range(n_dev + 1)]
    if n_dev >= 3:
        cuts[1] = cuts[2]          # an empty shard on device 1... (rows move to device 0)
    shards, bsh, ush = [], [], []
    codes = rng.integers(0, 2**62, size=(n, 3), dtype=np.uint64)
    mat = rng.integers(0, 256, size=(n, 48), dtype=np.uint8)
    gp, op = ib.QuantizationParams.from_range(-1.0, 1.0), oracle.QuantizationParams.from_range(-1.0, 1.0)
    for dev, (a, b) in enumerate(zip(cuts, cuts[1:])):
        ib.init(dev)
        shards.append(ib.DeviceBatch.from_rows_flat(rows[a:b].reshape(-1), b - a, d, index_base=a))
        bsh.append(ib.BinaryCorpus.from_words(codes[a:b], b - a, 192, index_base=a))
        ush.append(ib.U8Corpus.from_rows(mat[a:b], gp, index_base=a, dimension=48))
    ib.init(0)
    for rep in range(3):
        qs = rng.integers(-3, 4, size=(nq, d)).astype(np.float32)
        for metric in ("dot", "cosine", "l2"):
            for k in (1, 10, 100, 300):
                idx, sc = sharded.batch_knn_sharded(metric, qs, shards, k)
                widx, wsc = oracle.batch_knn_many(metric, qs, ob, k, n_threads=4)
                assert sc.tobytes() == wsc.tobytes(), (rep, metric, k)
                if metric != "l2":
                    assert np.array_equal(idx, widx), (rep, metric, k)
        qc = rng.integers(0, 2**62, size=(2, 3), dtype=np.uint64)
        gi, gd = sharded.hamming_topk_sharded(qc, bsh, 100)
        wi, wd = oracle.hamming_topk_many(qc, codes, 100, n_threads=2)
        assert np.array_equal(gi, wi) and np.array_equal(gd, wd), rep
        q8 = rng.uniform(-1, 1, size=(3, 48)).astype(np.float32)
        ui, us = sharded.batch_knn_u8_sharded(q8, ush, 10)
        wi, ws = oracle.batch_knn_u8_many(q8, mat, op, 10, n_threads=2)
        assert np.array_equal(ui, wi) and us.tobytes() == ws.tobytes(), rep


@needs2
def test_c_abi_sharded_async_over_devices():
    """innr_cuda_*_sharded_async: one ticket per call, two calls in flight, every device's part queued without a host
    synchronisation. A run of calls with different queries -- f32 (three metrics), Hamming, u8, interleaved -- equals the
    synchronous sharded entries bit for bit; a third submit and a synchronous sharded call are refused while two are in
    flight; single-device asynchronous calls keep working next to it."""
    import torch
    import innr_b200 as ib
    from innr_b200 import sharded, stream
    n_dev = min(torch.cuda.device_count(), 4)
    n, d = 120_001, 48
    rng = np.random.default_rng(24)
    rows = rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    codes = rng.integers(0, 2**62, size=(n, 3), dtype=np.uint64)
    mat = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
    gp = ib.QuantizationParams.from_range(-1.0, 1.0)
    cuts = [n * r // n_dev for r in range(n_dev + 1)]
    shards, bsh, ush = [], [], []
    for dev, (a, b) in enumerate(zip(cuts, cuts[1:])):
        ib.init(dev)
        shards.append(ib.DeviceBatch.from_rows_flat(rows[a:b].reshape(-1), b - a, d, index_base=a))
        bsh.append(ib.BinaryCorpus.from_words(codes[a:b], b - a, 192, index_base=a))
        ush.append(ib.U8Corpus.from_rows(mat[a:b], gp, index_base=a, dimension=d))
    ib.init(0)
    qs = rng.integers(-3, 4, size=(12, d)).astype(np.float32)
    qc = rng.integers(0, 2**62, size=(12, 3), dtype=np.uint64)
    q8 = rng.uniform(-1, 1, size=(12, d)).astype(np.float32)
    calls = []
    for j in range(12):
        m = ("dot", "cosine", "l2")[j % 3]
        calls.append((lambda j=j, m=m: stream.submit_knn_sharded(m, qs[j], shards, 10), lambda j=j, m=m: sharded.batch_knn_sharded(m, qs[j], shards, 10)))
        calls.append((lambda j=j: stream.submit_hamming_topk_sharded(qc[j], bsh, 100), lambda j=j: sharded.hamming_topk_sharded(qc[j], bsh, 100)))
        calls.append((lambda j=j: stream.submit_knn_u8_sharded(q8[j], ush, 10), lambda j=j: sharded.batch_knn_u8_sharded(q8[j], ush, 10)))
    calls.append((lambda: stream.submit_knn_sharded("cosine", qs[:3], shards, 7), lambda: sharded.batch_knn_sharded("cosine", qs[:3], shards, 7)))
    want = [sync() for _, sync in calls]
    pending, got = None, []
    for submit, _ in calls:
        t = submit()
        if pending is not None:
            got.append(pending.wait())
        pending = t
    got.append(pending.wait())
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g[0], w[0]) and g[1].tobytes() == w[1].tobytes(), i
    # two in flight: a third submit and a synchronous sharded call are refused, then everything drains
    t1 = stream.submit_knn_sharded("dot", qs[0], shards, 10)
    t2 = stream.submit_knn_sharded("dot", qs[1], shards, 10)
    with pytest.raises(ib.InnrCudaError):
        stream.submit_knn_sharded("dot", qs[2], shards, 10)
    with pytest.raises(ib.InnrCudaError):
        sharded.batch_knn_sharded("dot", qs[2], shards, 10)
    a, b = t1.wait(), t2.wait()
    assert np.array_equal(a[0], sharded.batch_knn_sharded("dot", qs[0], shards, 10)[0])
    assert np.array_equal(b[0], sharded.batch_knn_sharded("dot", qs[1], shards, 10)[0])
    # single-device asynchronous calls and sharded ones share the devices' two slots
    t1 = stream.submit_knn_sharded("dot", qs[3], shards, 10)
    t2 = stream.submit_knn("dot", qs[4], shards[0], 10)
    assert np.array_equal(t2.wait()[0], ib.batch_knn_many("dot", qs[4], shards[0], 10)[0])
    assert np.array_equal(t1.wait()[0], sharded.batch_knn_sharded("dot", qs[3], shards, 10)[0])
    # empty result: no ticket inside, empty arrays out; k > 128 is the synchronous entry's business
    idx, sc = stream.submit_knn_sharded("dot", qs[0], shards, 0).wait()
    assert idx.shape == (1, 0) and sc.shape == (1, 0)
    with pytest.raises(NotImplementedError):
        stream.submit_knn_sharded("dot", qs[0], shards, 300)
