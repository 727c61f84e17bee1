"""Random interleaving of every entry point on the same device context: the library keeps per-device workspaces (partial
lists, tickets, staging buffers, lazily built tensor-core operands), so a wrong reset or a stale buffer shows up only
when calls of different kinds and sizes follow each other. Every answer is compared with a precomputed oracle answer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_random_interleaving_of_all_entry_points(oracle):
    import innr_b200 as ib
    ib.init(0)
    ib.set_option("knn_tc_min_n", 4096)      # let the small corpora take the tensor-core path too
    try:
        rng = np.random.default_rng(2026)
        n, d = 20_000, 64
        rows = rng.standard_normal((n, d)).astype(np.float32)
        gb, ob = ib.VerticalBatch.from_flat(rows.reshape(-1), n, d), oracle.VerticalBatch.from_flat(rows.reshape(-1), n, d)
        small = rng.integers(-2, 3, size=(777, 9)).astype(np.float32)
        gs, os_ = ib.VerticalBatch.from_flat(small.reshape(-1), 777, 9), oracle.VerticalBatch.from_flat(small.reshape(-1), 777, 9)
        codes = rng.integers(0, 2**62, size=(15_000, 4), dtype=np.uint64)
        bc = ib.BinaryCorpus.from_words(codes, 15_000, 256)
        mat = rng.integers(0, 256, size=(9_000, 96), dtype=np.uint8)
        gp, op = ib.QuantizationParams.from_range(-1.0, 1.0), oracle.QuantizationParams.from_range(-1.0, 1.0)
        uc = ib.U8Corpus.from_rows(mat, gp)
        lens = rng.integers(0, 60, size=300)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        toks = rng.standard_normal((int(off[-1]), 128)).astype(np.float32)
        tc = ib.TokenCorpus.from_tokens(toks, off, 128)
        qs = rng.standard_normal((40, d)).astype(np.float32)
        qsmall = rng.integers(-2, 3, size=(5, 9)).astype(np.float32)
        qc = rng.integers(0, 2**62, size=(6, 4), dtype=np.uint64)
        q8 = rng.uniform(-1, 1, size=(5, 96)).astype(np.float32)
        qt = rng.standard_normal((3, 32, 128)).astype(np.float32)
        mask = rng.random(n) < 0.3

        def knn_single(metric, j, k):
            idx, sc = ib.batch_knn_many(metric, qs[j], gb, k)
            widx, wsc = oracle.batch_knn_many(metric, qs[j:j + 1], ob, k)
            assert np.array_equal(bits(sc), bits(wsc)) and (metric == "l2" or np.array_equal(idx, widx)), ("single", metric, k)

        def knn_batch(metric, a, b, k):
            idx, sc = ib.batch_knn_many(metric, qs[a:b], gb, k)
            widx, wsc = oracle.batch_knn_many(metric, qs[a:b], ob, k, n_threads=4)
            assert np.array_equal(bits(sc), bits(wsc)) and (metric == "l2" or np.array_equal(idx, widx)), ("batch", metric, a, b, k)

        def knn_small(j, k):
            g, w = ib.batch_knn_dot(qsmall[j], gs, k), oracle.batch_knn_dot(qsmall[j], os_, k)
            assert list(g.indices) == list(w.indices) and np.array_equal(bits(g.scores), bits(w.scores))

        def filtered(j):
            g = ib.batch_knn_filtered(qs[j], gb, 9, mask)
            w = oracle.batch_knn_filtered(qs[j], ob, 9, lambda i: bool(mask[i]))
            assert list(g.indices) == list(w.indices) and np.array_equal(bits(g.scores), bits(w.scores))

        def pruning(j):
            full = oracle.batch_l2_squared(qs[j], ob)
            thr = float(np.quantile(full, 0.02))
            g, w = ib.batch_l2_squared_pruning(qs[j], gb, thr), oracle.batch_l2_squared_pruning(qs[j], ob, thr)
            assert [i for i, _ in g] == [i for i, _ in w]

        def hamming(a, b, k):
            gi, gd = ib.hamming_topk_many(qc[a:b], bc, k)
            wi, wd = oracle.hamming_topk_many(qc[a:b], codes, k, n_threads=2)
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd)

        def u8(a, b, k):
            gi, gsc = ib.batch_knn_u8_many(q8[a:b], uc, k)
            wi, wsc = oracle.batch_knn_u8_many(q8[a:b], mat, op, k, n_threads=2)
            assert np.array_equal(gi, wi) and np.array_equal(bits(gsc), bits(wsc))

        def maxsim(j, cos):
            got = ib.maxsim_corpus(qt[j], tc, cosine=cos)
            want = oracle.maxsim_corpus(qt[j], toks, off, cosine_flag=cos)
            assert np.allclose(got, want, rtol=2e-5, atol=1e-4)

        def maxsim_batch():
            got = ib.maxsim_corpus_batch(qt, tc, cosine=True)
            for j in range(3):
                assert np.array_equal(bits(got[j]), bits(ib.maxsim_corpus(qt[j], tc, cosine=True)))

        def subset(j):
            cand = rng.choice(n, size=500, replace=False).astype(np.uint64)
            g = ib.batch_knn_subset("cosine", qs[j], gb, cand, 7)
            order = np.sort(cand).astype(np.int64)
            sub = oracle.VerticalBatch.from_flat(rows[order].reshape(-1), len(order), d)
            w = oracle.batch_knn_cosine(qs[j], sub, 7)
            assert [int(i) for i in g.indices] == [int(order[int(t)]) for t in w.indices]

        ops = [lambda: knn_single(rng.choice(["cosine", "dot", "l2"]), int(rng.integers(0, 40)), int(rng.choice([1, 10, 100, 200]))),
               lambda: knn_batch(rng.choice(["cosine", "dot", "l2"]), 0, int(rng.integers(2, 40)), int(rng.choice([1, 10, 64]))),
               lambda: knn_small(int(rng.integers(0, 5)), int(rng.choice([1, 5, 40]))),
               lambda: filtered(int(rng.integers(0, 40))), lambda: pruning(int(rng.integers(0, 40))),
               lambda: hamming(0, int(rng.integers(1, 7)), int(rng.choice([1, 10, 100, 150]))),
               lambda: u8(0, int(rng.integers(1, 6)), int(rng.choice([1, 10, 100]))),
               lambda: maxsim(int(rng.integers(0, 3)), bool(rng.integers(0, 2))), maxsim_batch,
               lambda: subset(int(rng.integers(0, 40)))]
        for _ in range(120):
            ops[int(rng.integers(0, len(ops)))]()
    finally:
        ib.set_option("knn_tc_min_n", 100000)
