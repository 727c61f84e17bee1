"""Host-side logic of the row-sharded top-k on CPU: world_size 2, gloo backend.

Each rank owns a contiguous row range, computes its LOCAL top-k with the oracle (the checker stands in for the
device scan here: no GPU in this container), encodes it as the same 64-bit composite keys the kernels emit
(innr_b200/sharded.py codec == csrc/common.cuh), exchanges them with ONE all_gather and merges. The merged result
must equal the oracle's answer over the whole corpus, for descending (dot/cosine), ascending (L2) and Hamming
keys, including exact ties that straddle the shard boundary (lower GLOBAL index must win)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from innr_b200 import sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import innr_oracle as orc
    n, d, k = 4001, 24, 10
    rng = np.random.default_rng(0)
    rows = rng.integers(-3, 4, size=(n, d)).astype(np.float32)  # integer valued: many exact ties
    rows[n // 2 - 1] = rows[n // 2] = rows[7]                     # a tie group straddling the shard boundary
    q = rng.integers(-3, 4, size=d).astype(np.float32)
    lo, hi = sharded.shard_range(n, rank, world)
    results = {}
    for metric, fn, desc in (("dot", orc.batch_knn_dot, True), ("cosine", orc.batch_knn_cosine, True),
                             ("l2", orc.batch_knn_dot, False)):
        shard = orc.VerticalBatch.from_flat(rows[lo:hi].reshape(-1), hi - lo, d)
        if metric == "l2":
            dist_local = orc.batch_l2_squared(q, shard)
            order = np.lexsort((np.arange(hi - lo), dist_local))[:k]
            idx_l, sc_l = order, dist_local[order]
        else:
            r = fn(q, shard, k)
            idx_l, sc_l = np.array(r.indices), r.scores
        keys = sharded.encode_keys(sc_l, idx_l + lo, desc)          # global index = shard base + local
        local = torch.from_numpy(keys.view(np.int64).copy())
        gathered = torch.empty(world * k, dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, local)               # the ONE collective of the data path
        merged = sharded.merge_keys_host(gathered.numpy().view(np.uint64).reshape(world, k), k)
        results[metric] = sharded.decode_keys(merged, desc)
    # Hamming keys: (distance << 32) | global index
    codes = rng.integers(0, 2**64, size=(n, 2), dtype=np.uint64)
    codes[n // 2] = codes[3]
    qc = rng.integers(0, 2**64, size=2, dtype=np.uint64)
    li, ld = orc.hamming_topk(qc, codes[lo:hi], k)
    keys = (ld.astype(np.uint64) << np.uint64(32)) | (li + np.uint64(lo))
    gathered = torch.empty(world * k, dtype=torch.int64)
    dist.all_gather_into_tensor(gathered, torch.from_numpy(keys.view(np.int64).copy()))
    merged = sharded.merge_keys_host(gathered.numpy().view(np.uint64).reshape(world, k), k)
    results["hamming"] = (merged & np.uint64(0xFFFFFFFF), (merged >> np.uint64(32)).astype(np.uint32))
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), **{f"{m}_{i}": v[i] for m, v in results.items() for i in (0, 1)},
                 rows=rows, q=q, codes=codes, qc=qc)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_rows():
    for n in (0, 1, 7, 10_000_000):
        for world in (1, 2, 3, 8):
            edges = [sharded.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_key_codec_matches_total_cmp():
    s = np.array([np.nan, -np.nan, np.inf, -np.inf, 0.0, -0.0, 1.5, -1.5, 1e-45, -1e-45], np.float32)
    idx = np.arange(s.size)
    asc = np.argsort(sharded.encode_keys(s, idx, False), kind="stable")
    # f32::total_cmp: -NaN < -inf < -1.5 < -1e-45 < -0.0 < +0.0 < 1e-45 < 1.5 < inf < NaN
    assert asc.tolist() == [1, 3, 7, 9, 5, 4, 8, 6, 2, 0]
    desc = np.argsort(sharded.encode_keys(s, idx, True), kind="stable")
    assert desc.tolist() == asc.tolist()[::-1]
    ii, ss = sharded.decode_keys(sharded.encode_keys(s, idx, True), True)
    assert np.array_equal(ii, idx.astype(np.uint64)) and np.array_equal(ss.view(np.uint32), s.view(np.uint32))


def test_two_rank_gloo_merge_equals_global_topk(tmp_path, oracle):
    world, k = 2, 10
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "merged.npz")
    rows, q = z["rows"], z["q"]
    n, d = rows.shape
    ob = oracle.VerticalBatch.from_flat(rows.reshape(-1), n, d)
    for metric, fn in (("dot", oracle.batch_knn_dot), ("cosine", oracle.batch_knn_cosine)):
        w = fn(q, ob, k)
        assert z[f"{metric}_0"].tolist() == w.indices, metric
        assert np.array_equal(z[f"{metric}_1"].view(np.uint32), w.scores.view(np.uint32))
    dist_all = oracle.batch_l2_squared(q, ob)
    order = np.lexsort((np.arange(n), dist_all))[:k]
    assert z["l2_0"].tolist() == order.tolist()
    assert np.array_equal(z["l2_1"].view(np.uint32), dist_all[order].view(np.uint32))
    wi, wd = oracle.hamming_topk(z["qc"], z["codes"], k)
    assert z["hamming_0"].tolist() == wi.tolist() and z["hamming_1"].tolist() == wd.tolist()
