"""Seeded random-shape sweeps of the CUDA path against the oracle -- the GPU-side counterpart of the reference's
property tests (tests/property_tests.rs:357-495: batch_l2_matches_individual, batch_dot_matches_individual,
batch_knn_sorted, batch_knn_unique_indices) with the stronger bar of this port: bit-exact scores and indices.
Shapes straddle every tiling boundary (4 vectors per thread, 1024 per tile, 64-float row pitch, 16 / 32 / 128-wide
chunks), with ties (integer-valued rows), zero rows and k around the list sizes (32, 128)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import innr_b200
    innr_b200.init(0)
    return innr_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _shapes(rng, count, n_max, d_max):
    edge_n = [1, 2, 3, 4, 5, 63, 64, 65, 1023, 1024, 1025, 2047, 4096, 4097]
    edge_d = [1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129]
    for j in range(count):
        n = int(rng.choice(edge_n)) if j % 3 == 0 else int(rng.integers(1, n_max))
        d = int(rng.choice(edge_d)) if j % 2 == 0 else int(rng.integers(1, d_max))
        yield n, d


def test_fuzz_f32_knn_scores_filtered_pruning(ib, oracle):
    rng = np.random.default_rng(20261018)
    for n, d in _shapes(rng, 40, 12000, 300):
        ties = rng.random() < 0.5
        rows = (rng.integers(-2, 3, size=(n, d)) if ties else rng.standard_normal((n, d))).astype(np.float32)
        if n > 3:
            rows[rng.integers(0, n)] = 0.0
        q = (rng.integers(-2, 3, size=d) if ties else rng.standard_normal(d)).astype(np.float32)
        gb, ob = ib.VerticalBatch.from_flat(rows.reshape(-1), n, d), oracle.VerticalBatch.from_flat(rows.reshape(-1), n, d)
        tag = (n, d, ties)
        assert np.array_equal(bits(ib.batch_dot(q, gb)), bits(oracle.batch_dot(q, ob))), tag
        assert np.array_equal(bits(ib.batch_l2_squared(q, gb)), bits(oracle.batch_l2_squared(q, ob))), tag
        for k in {1, min(n, 10), min(n, 33), n + 2}:
            for name in ("batch_knn_dot", "batch_knn_cosine"):
                g, w = getattr(ib, name)(q, gb, k), getattr(oracle, name)(q, ob, k)
                assert list(g.indices) == list(w.indices), (tag, name, k)
                assert np.array_equal(bits(g.scores), bits(w.scores)), (tag, name, k)
            g, w = ib.batch_knn(q, gb, k), oracle.batch_knn(q, ob, k)
            assert np.array_equal(bits(g.scores), bits(w.scores)), (tag, "batch_knn", k)
            assert len(set(g.indices)) == len(g.indices)
        mask = rng.random(n) < rng.choice([0.02, 0.5, 0.97])
        g = ib.batch_knn_filtered(q, gb, 7, mask)
        w = oracle.batch_knn_filtered(q, ob, 7, lambda i: bool(mask[i]))
        assert list(g.indices) == list(w.indices) and np.array_equal(bits(g.scores), bits(w.scores)), (tag, "filtered")
        full = oracle.batch_l2_squared(q, ob)
        thr = float(np.quantile(full, rng.choice([0.0, 0.1, 0.9, 1.0])))
        g, w = ib.batch_l2_squared_pruning(q, gb, thr), oracle.batch_l2_squared_pruning(q, ob, thr)
        assert [i for i, _ in g] == [i for i, _ in w], (tag, "pruning", thr)
        assert np.array_equal(bits([s for _, s in g]), bits([s for _, s in w])), (tag, "pruning")
        # variance order, reordered and adaptive kNN (src/batch.rs:441-659)
        assert np.array_equal(bits(ib.batch_dimension_variance(gb)), bits(oracle.batch_dimension_variance(ob))), (tag, "variance")
        g, w = ib.batch_knn_reordered(q, gb, 9), oracle.batch_knn_reordered(q, ob, 9)
        assert list(g.indices) == list(w.indices) and np.array_equal(bits(g.scores), bits(w.scores)), (tag, "reordered")
        wd = int(rng.integers(1, d + 3))
        g, w = ib.batch_knn_adaptive(q, gb, 6, wd), oracle.batch_knn_adaptive(q, ob, 6, wd)
        assert list(g.indices) == list(w.indices) and np.array_equal(bits(g.scores), bits(w.scores)), (tag, "adaptive", wd)


def test_fuzz_hamming_and_u8(ib, oracle):
    rng = np.random.default_rng(7)
    for n, _ in _shapes(rng, 24, 9000, 10):
        dim = int(rng.choice([1, 63, 64, 65, 127, 128, 129, 500, 1024, 1536]))
        words = (dim + 63) // 64
        codes = rng.integers(0, 2**63, size=(n, words), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, words), dtype=np.uint64)
        if dim % 64:
            codes[:, -1] &= np.uint64((1 << (dim % 64)) - 1)
        nq = int(rng.integers(1, 6))
        qs = codes[rng.integers(0, n, size=nq)].copy()
        corpus = ib.BinaryCorpus.from_words(codes, n, dim)
        for k in (1, min(n, 40), n + 1):
            gi, gd = ib.hamming_topk_many(qs, corpus, k)
            wi, wd = oracle.hamming_topk_many(qs, codes, k, n_threads=2)
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd), (n, dim, nq, k)
    gp, op = ib.QuantizationParams.from_range(-0.5, 3.0), oracle.QuantizationParams.from_range(-0.5, 3.0)
    for n, d in _shapes(rng, 24, 6000, 500):
        mat = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
        nq = int(rng.integers(1, 5))
        qs = (rng.standard_normal((nq, d)) * rng.choice([0.1, 1.0, 3.0])).astype(np.float32)
        corpus = ib.U8Corpus.from_rows(mat, gp)
        for k in (1, min(n, 35)):
            gi, gs = ib.batch_knn_u8_many(qs, corpus, k)
            wi, ws = oracle.batch_knn_u8_many(qs, mat, op, k, n_threads=2)
            assert np.array_equal(gi, wi) and np.array_equal(bits(gs), bits(ws)), (n, d, nq, k)


def test_fuzz_maxsim(ib, oracle):
    rng = np.random.default_rng(99)
    for j in range(16):
        dim = int(rng.choice([32, 64, 96, 128, 48, 130]))
        nq = int(rng.choice([1, 7, 32, 33, 64, 70]))
        n_docs = int(rng.integers(1, 400))
        lens = rng.integers(0, int(rng.choice([3, 40, 300])), size=n_docs)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        toks = rng.standard_normal((int(off[-1]), dim)).astype(np.float32)
        q = rng.standard_normal((nq, dim)).astype(np.float32)
        corpus = ib.TokenCorpus.from_tokens(toks, off, dim)
        aq = np.abs(q.astype(np.float64))
        for cos in (False, True):
            got = ib.maxsim_corpus(q, corpus, cosine=cos)
            want = oracle.maxsim_corpus(q, toks, off, cosine_flag=cos)
            scale = np.array([float(np.sum(np.max(aq @ np.abs(toks[off[i]:off[i + 1]].astype(np.float64)).T, axis=1)))
                              if lens[i] else 0.0 for i in range(n_docs)]) if not cos else np.full(n_docs, float(nq))
            err = np.abs(got.astype(np.float64) - want)
            assert np.all(err <= 1e-5 * scale + 1e-6), (dim, nq, n_docs, cos, float(err.max()))
            assert np.all(got[lens == 0] == 0.0)
