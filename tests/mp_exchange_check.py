"""Run under torchrun with >= 2 ranks, one GPU each (tests/test_gpu_multi.py launches it; bench.py's N > 1 self-check does
the same comparison): the row-sharded top-k through the peer-mapped exchange must equal the NCCL allgather + merge path
bit for bit on every rank, and both must equal the CPU oracle's unsharded answer (rank 0 checks)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import innr_b200 as ib
    from innr_b200 import sharded, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ib.init(local)
    ex = sharded.PeerExchange.for_process_group(dist)
    dev = torch.device(f"cuda:{local}")
    failures = []

    def same(a, b):
        return bool(torch.equal(a[0], b[0])) and bool(torch.equal(a[1].view(torch.int32), b[1].view(torch.int32)))

    # f32 (all three metrics), several queries, k = 10 and 100
    n, d = 50_000, 96
    lo, hi = sharded.shard_range(n, rank, world)
    shard = ib.DeviceBatch.generate("ghash", synth.SALT_CORPUS, lo, hi - lo, d, index_base=lo)
    for metric in ("cosine", "dot", "l2"):
        peer = sharded.ShardedKnn(shard, "f32", metric, exchange=ex)
        nccl = sharded.ShardedKnn(shard, "f32", metric)
        for nq, k in ((1, 10), (7, 10), (3, 100)):
            for rep in range(3):
                qs = synth.ghash_f32(synth.SALT_QUERY, 1000 * rep, nq * d).reshape(nq, d)
                dq = torch.from_numpy(qs).to(dev)
                a = [t.clone() for t in peer.knn_dev(dq, nq, k)]
                b = [t.clone() for t in nccl.knn_dev(dq, nq, k)]
                torch.cuda.synchronize()
                if not same(a, b):
                    failures.append(("f32", metric, nq, k, rep))
                if rank == 0 and rep == 0 and metric != "l2":
                    from oracle import innr_oracle as orc
                    ob = orc.VerticalBatch.from_flat(orc.ghash_f32(synth.SALT_CORPUS, 0, n * d), n, d)
                    for j in range(nq):
                        w = getattr(orc, "batch_knn_" + metric)(qs[j], ob, k)
                        if a[0][j].cpu().tolist() != w.indices or a[1][j].cpu().numpy().tobytes() != w.scores.tobytes():
                            failures.append(("f32-vs-oracle", metric, nq, k, j))
    # pipelined form: the exchange of call i runs on a side stream under the scan of call i + 1; results of call i are
    # read after call i + 1 has been queued (they stay valid until call i + 2 reuses their buffers)
    peer = sharded.ShardedKnn(shard, "f32", "cosine", exchange=ex)
    sync = sharded.ShardedKnn(shard, "f32", "cosine")
    qs = synth.ghash_f32(synth.SALT_QUERY, 5000, 9 * d).reshape(9, d)
    dqs = torch.from_numpy(qs).to(dev)
    main = torch.cuda.current_stream()
    got, pending = [], None
    for i in range(9):
        cur = peer.knn_dev_pipelined(dqs[i], 1, 10)
        if pending is not None:
            main.wait_event(pending[2])
            got.append((pending[0].clone(), pending[1].clone()))
        pending = cur
    main.wait_event(pending[2])
    got.append((pending[0].clone(), pending[1].clone()))
    peer.drain()
    torch.cuda.synchronize()
    for i in range(9):
        want = [t.clone() for t in sync.knn_dev(dqs[i], 1, 10)]
        torch.cuda.synchronize()
        if not same(got[i], want):
            failures.append(("pipelined", i))
    # host-buffer streaming: pinned queries in, pinned results out, one event per call
    hq = torch.from_numpy(qs).pin_memory()
    got, pending = [], None
    for i in range(9):
        cur = peer.knn_dev_pipelined(None, 1, 10, host_queries=hq[i], host_out=True)
        if pending is not None:
            pending[2].synchronize()
            got.append((pending[0].clone(), pending[1].clone()))
        pending = cur
    pending[2].synchronize()
    got.append((pending[0].clone(), pending[1].clone()))
    for i in range(9):
        want = [t.cpu() for t in sync.knn_dev(dqs[i], 1, 10)]
        if not (got[i][0].device.type == "cpu" and same(got[i], want)):
            failures.append(("host-streaming", i))
    # Hamming (heavy ties) and u8
    nb = 200_000
    lo, hi = sharded.shard_range(nb, rank, world)
    codes = ib.BinaryCorpus.generate(synth.SALT_CODES, lo, hi - lo, 1024, index_base=lo)
    qc = torch.from_numpy(synth.ghash_u64(synth.SALT_QUERY, 0, 16).view(np.int64)).to(dev)
    a = [t.clone() for t in sharded.ShardedKnn(codes, "binary", exchange=ex).knn_dev(qc, 1, 100)]
    b = [t.clone() for t in sharded.ShardedKnn(codes, "binary").knn_dev(qc, 1, 100)]
    torch.cuda.synchronize()
    if not (torch.equal(a[0], b[0]) and torch.equal(a[1].long(), b[1].long())):
        failures.append(("hamming",))
    p = ib.QuantizationParams.from_range(-1.0, 1.0)
    u8 = ib.U8Corpus.generate(synth.SALT_CORPUS, lo, hi - lo, 384, p, index_base=lo)
    q8 = torch.from_numpy(synth.ghash_f32(synth.SALT_QUERY, 0, 384)).to(dev)
    a = [t.clone() for t in sharded.ShardedKnn(u8, "u8", exchange=ex).knn_dev(q8, 1, 10)]
    b = [t.clone() for t in sharded.ShardedKnn(u8, "u8").knn_dev(q8, 1, 10)]
    torch.cuda.synchronize()
    if not same(a, b):
        failures.append(("u8",))
    # the same two through the pipelined form (two scan streams, exchange on the side stream), several calls in flight
    for kind, corpus, q, k in (("binary", codes, qc, 100), ("u8", u8, q8, 10)):
        pk, sk = sharded.ShardedKnn(corpus, kind, exchange=ex), sharded.ShardedKnn(corpus, kind)
        want = [t.clone() for t in sk.knn_dev(q, 1, k)]
        outs = []
        for i in range(6):
            r = pk.knn_dev_pipelined(q, 1, k)
            r[2].synchronize()
            outs.append((r[0].clone(), r[1].clone()))
        for i in range(4):
            last = pk.knn_dev_pipelined(q, 1, k)
        pk.drain()
        torch.cuda.synchronize()
        outs.append((last[0].clone(), last[1].clone()))
        for o in outs:
            if not (torch.equal(o[0], want[0]) and torch.equal(o[1].long() if kind == "binary" else o[1], want[1].long() if kind == "binary" else want[1])):
                failures.append(("pipelined", kind))
                break
    if ex.status() != 0:
        failures.append(("timeout",))
    bad = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(bad)
    if failures:
        print(f"rank {rank}: FAILURES {failures}", flush=True)
    dist.destroy_process_group()
    if int(bad.item()):
        sys.exit(1)
    if rank == 0:
        print("mp_exchange_check ok", flush=True)


if __name__ == "__main__":
    main()
