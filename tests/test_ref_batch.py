"""Ports of the reference's own batch tests: /root/reference/src/batch.rs:884-1616 and
/root/reference/tests/batch_tests.rs:15-488 (restricted to the functions on the hot path).

Run against the CPU oracle (pins the oracle to the reference) and, on a GPU box, against the CUDA product
through the C-ABI (`api` fixture, conftest.py). Same inputs, same assertions, same tolerances as the Rust tests.
"""
import math

import numpy as np
import pytest


def approx(a, b, tol):
    return abs(float(a) - float(b)) < tol


# ---- src/batch.rs:888-901 / tests/batch_tests.rs:25-36
def test_vertical_batch_creation(api):
    b = api.VerticalBatch.from_rows([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    assert b.num_vectors == 2 and b.dimension == 3
    assert b.get(0, 0) == 1.0 and b.get(0, 1) == 4.0 and b.get(1, 0) == 2.0 and b.get(2, 1) == 6.0


def test_empty_batch(api):  # src/batch.rs:1021-1026, tests/batch_tests.rs:15-22
    b = api.VerticalBatch.from_rows([])
    assert b.num_vectors == 0 and b.dimension == 0


def test_single_vector_batch(api):  # src/batch.rs:1028-1038
    b = api.VerticalBatch.from_rows([[1.0, 2.0, 3.0]])
    assert (b.num_vectors, b.dimension) == (1, 3)
    assert [b.get(d, 0) for d in range(3)] == [1.0, 2.0, 3.0]
    assert b.extract_vector(0).tolist() == [1.0, 2.0, 3.0]


def test_from_flat_matches_from_rows(api):  # src/batch.rs:1044-1066
    rows = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]
    flat = [x for r in rows for x in r]
    a = api.VerticalBatch.from_rows(rows)
    b = api.VerticalBatch.from_flat(flat, 3, 3)
    for d in range(3):
        for v in range(3):
            assert a.get(d, v) == b.get(d, v)


def test_from_flat_single_vector(api):  # src/batch.rs:1068-1075
    b = api.VerticalBatch.from_flat([10.0, 20.0], 1, 2)
    assert (b.num_vectors, b.dimension) == (1, 2)
    assert b.extract_vector(0).tolist() == [10.0, 20.0]


def test_from_rows_ragged_panics(api):  # src/batch.rs:120 assert_eq!(vec.len(), dimension, ...)
    with pytest.raises(AssertionError):
        api.VerticalBatch.from_rows([[1.0, 2.0], [3.0]])


def test_dimension_slice(api):  # src/batch.rs:1081-1090
    b = api.VerticalBatch.from_rows([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    assert list(b.dimension_slice(0)) == [1.0, 3.0, 5.0]
    assert list(b.dimension_slice(1)) == [2.0, 4.0, 6.0]


def test_extract_all_vectors_roundtrip(api):  # src/batch.rs:1599-1615, tests/batch_tests.rs:79-92
    rows = [[1.5, 2.5, 3.5], [4.5, 5.5, 6.5], [7.5, 8.5, 9.5]]
    b = api.VerticalBatch.from_rows(rows)
    for i, r in enumerate(rows):
        assert b.extract_vector(i).tolist() == r


def test_batch_l2_squared(api):  # src/batch.rs:903-928
    b = api.VerticalBatch.from_rows([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    d = api.batch_l2_squared([1.0, 1.0, 0.0], b)
    assert approx(d[0], 2.0, 1e-6) and approx(d[1], 1.0, 1e-6) and approx(d[2], 1.0, 1e-6)


def test_batch_dot(api):  # src/batch.rs:930-951
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    d = api.batch_dot([1.0, 2.0], b)
    assert approx(d[0], 1.0, 1e-6) and approx(d[1], 2.0, 1e-6) and approx(d[2], 3.0, 1e-6)


def test_batch_knn(api):  # src/batch.rs:953-969 (tie between v0 and v1: `contains`)
    b = api.VerticalBatch.from_rows([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [3.0, 0.0]])
    r = api.batch_knn([0.5, 0.0], b, 2)
    assert len(r.indices) == 2 and 0 in r.indices and 1 in r.indices


def test_batch_cosine(api):  # src/batch.rs:998-1016
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    c = api.batch_cosine([1.0, 0.0], b, api.batch_norms(b))
    assert approx(c[0], 1.0, 1e-6) and abs(c[1]) < 1e-6 and approx(c[2], 1 / math.sqrt(2), 0.01)


def test_batch_norms(api):  # src/batch.rs:1096-1113, tests/batch_tests.rs:393-406
    b = api.VerticalBatch.from_rows([[3.0, 4.0], [0.0, 0.0], [1.0, 0.0]])
    n = api.batch_norms(b)
    assert approx(n[0], 5.0, 1e-6) and abs(n[1]) < 1e-6 and approx(n[2], 1.0, 1e-6)


def test_batch_l2_squared_exact_match(api):  # src/batch.rs:1119-1132
    b = api.VerticalBatch.from_rows([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    d = api.batch_l2_squared([3.0, 4.0], b)
    assert abs(d[1]) < 1e-9 and d[0] > 0 and d[2] > 0


def test_batch_dot_zero_query(api):  # src/batch.rs:1138-1146
    b = api.VerticalBatch.from_rows([[1.0, 2.0], [3.0, 4.0]])
    assert list(api.batch_dot([0.0, 0.0], b)) == [0.0, 0.0]


def test_batch_cosine_zero_query(api):  # src/batch.rs:1152-1162, tests/batch_tests.rs:477-488
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0]])
    assert list(api.batch_cosine([0.0, 0.0], b, api.batch_norms(b))) == [0.0, 0.0]


def test_batch_cosine_zero_norm_vector(api):  # src/batch.rs:1164-1174, tests/batch_tests.rs:408-422
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 0.0]])
    c = api.batch_cosine([1.0, 0.0], b, api.batch_norms(b))
    assert approx(c[0], 1.0, 1e-6) and c[1] == 0.0


def test_batch_cosine_norms_len_panics(api):  # src/batch.rs:711 assert_eq!(norms.len(), batch.num_vectors)
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0]])
    with pytest.raises(AssertionError):
        api.batch_cosine([1.0, 0.0], b, [1.0])


def test_query_len_panics(api):  # src/batch.rs:251,285,386,743,778 assert_eq!(query.len(), batch.dimension)
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0]])
    for fn in (api.batch_dot, api.batch_l2_squared):
        with pytest.raises(AssertionError):
            fn([1.0, 0.0, 0.0], b)
    for fn in (api.batch_knn, api.batch_knn_dot, api.batch_knn_cosine):
        with pytest.raises(AssertionError):
            fn([1.0], b, 1)


def test_batch_knn_dot_basic(api):  # src/batch.rs:1184-1197
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])
    r = api.batch_knn_dot([1.0, 0.0], b, 2)
    assert r.indices[0] == 0 and approx(r.scores[0], 1.0, 1e-6)


def test_batch_knn_dot_sorted_descending(api):  # src/batch.rs:1199-1209
    b = api.VerticalBatch.from_rows([[0.5, 0.5], [1.0, 0.0], [0.0, 1.0]])
    r = api.batch_knn_dot([1.0, 0.0], b, 3)
    assert all(r.scores[i] >= r.scores[i + 1] for i in range(2))


def test_batch_knn_cosine_basic(api):  # src/batch.rs:1298-1314
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])
    r = api.batch_knn_cosine([1.0, 0.0], b, 2)
    assert r.indices == [0, 1]
    assert approx(r.scores[0], 1.0, 1e-5) and abs(r.scores[1]) < 1e-5


def test_batch_knn_cosine_empty(api):  # src/batch.rs:1316-1321
    b = api.VerticalBatch.from_rows([])
    assert api.batch_knn_cosine([], b, 5).indices == []


def test_batch_knn_cosine_sorted_descending(api):  # src/batch.rs:1323-1345
    b = api.VerticalBatch.from_rows([[0.1, 1.0], [1.0, 0.0], [0.5, 0.5]])
    r = api.batch_knn_cosine([1.0, 0.0], b, 3)
    assert all(r.scores[i] >= r.scores[i + 1] for i in range(2))
    assert r.indices[0] == 1


def test_batch_knn_k_zero(api):  # src/batch.rs:1433-1442, tests/batch_tests.rs:381-391
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0]])
    for fn in (api.batch_knn, api.batch_knn_dot, api.batch_knn_cosine):
        r = fn([1.0, 0.0], b, 0)
        assert r.indices == [] and len(r.scores) == 0


def test_batch_knn_empty_batch(api):  # src/batch.rs:1444-1449
    b = api.VerticalBatch.from_rows([])
    assert api.batch_knn([], b, 5).indices == []
    assert api.batch_knn_dot([], b, 5).indices == []


def test_batch_knn_k_larger_than_n(api):  # src/batch.rs:1451-1460, tests/batch_tests.rs:369-379
    b = api.VerticalBatch.from_rows([[1.0], [2.0]])
    assert len(api.batch_knn([1.5], b, 10).indices) == 2
    b2 = api.VerticalBatch.from_rows([[1.0, 2.0], [3.0, 4.0]])
    assert len(api.batch_knn([0.0, 0.0], b2, 100).indices) == 2


def test_batch_knn_sorted_by_distance(api):  # src/batch.rs:1462-1480
    b = api.VerticalBatch.from_rows([[10.0, 0.0], [1.0, 0.0], [5.0, 0.0], [0.0, 0.0]])
    r = api.batch_knn([0.0, 0.0], b, 4)
    assert all(r.scores[i] <= r.scores[i + 1] for i in range(3))
    assert r.indices[0] == 3


def test_batch_l2_squared_large(api):  # src/batch.rs:1576-1596
    n, dim = 32, 8
    rows = [[float(i * dim + d) for d in range(dim)] for i in range(n)]
    b = api.VerticalBatch.from_rows(rows)
    d = api.batch_l2_squared(rows[0], b)
    assert abs(d[0]) < 1e-9 and all(x > 0 for x in d[1:])


# ---- tests/batch_tests.rs
def test_l2_squared_identity_symmetric_known(api):  # :98-141
    rows = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]
    b = api.VerticalBatch.from_rows(rows)
    assert abs(api.batch_l2_squared(rows[0], b)[0]) < 1e-6
    d12 = api.batch_l2_squared(rows[0], api.VerticalBatch.from_rows([rows[1]]))[0]
    d21 = api.batch_l2_squared(rows[1], api.VerticalBatch.from_rows([rows[0]]))[0]
    assert approx(d12, d21, 1e-6)
    z = api.VerticalBatch.from_rows([[0.0, 0.0, 0.0]])
    assert approx(api.batch_l2_squared([3.0, 4.0, 0.0], z)[0], 25.0, 1e-6)


def test_dot_product_orthogonal(api):  # :143-159
    b = api.VerticalBatch.from_rows([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    d = api.batch_dot([1.0, 0.0, 0.0], b)
    assert approx(d[0], 1.0, 1e-6) and abs(d[1]) < 1e-6 and abs(d[2]) < 1e-6


def test_cosine_normalized(api):  # :161-187
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [-1.0, 0.0]])
    c = api.batch_cosine([1.0, 0.0], b, api.batch_norms(b))
    assert approx(c[0], 1.0, 1e-6) and abs(c[1]) < 1e-6
    assert approx(c[2], 1.0 / math.sqrt(2.0), 1e-5) and approx(c[3], -1.0, 1e-6)


def test_knn_returns_k_results(api):  # :220-231
    b = api.VerticalBatch.from_rows([[float(i), 0.0] for i in range(100)])
    for k in (1, 5, 10, 50, 100):
        r = api.batch_knn([50.0, 0.0], b, k)
        assert len(r.indices) == k and len(r.scores) == k


def test_knn_results_sorted(api):  # :233-250
    b = api.VerticalBatch.from_rows([[float(i), math.sin(float(i))] for i in range(50)])
    r = api.batch_knn([25.0, 0.0], b, 20)
    assert all(r.scores[i] >= r.scores[i - 1] for i in range(1, len(r.scores)))


def test_knn_finds_exact_match(api):  # :252-267
    b = api.VerticalBatch.from_rows([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    r = api.batch_knn([0.0, 1.0], b, 1)
    assert r.indices[0] == 2 and r.scores[0] < 1e-6


def test_cosine_knn_normalized_matches_dot_knn(api):  # :439-466
    raw = np.array([[math.sin(float(i * 7 + d * 3)) for d in range(8)] for i in range(50)], dtype=np.float32)
    rows = []
    for v in raw:
        n = np.float32(0)
        for x in v:
            n = np.float32(n + x * x)
        n = np.sqrt(n)
        rows.append((v / n).astype(np.float32).tolist())
    q = np.array([math.cos(i * 0.3) for i in range(8)], dtype=np.float32)
    q = (q / np.sqrt(np.sum(q * q, dtype=np.float32))).astype(np.float32)
    b = api.VerticalBatch.from_rows(rows)
    assert api.batch_knn_cosine(q, b, 5).indices == api.batch_knn_dot(q, b, 5).indices


# ---- batch_knn_filtered: src/batch.rs:1353-1420, tests/batch_tests.rs:461-474
def test_filtered_knn_basic(api):
    b = api.VerticalBatch.from_rows([[0.0, 0.0], [1.0, 0.0], [0.1, 0.0], [10.0, 0.0]])
    r = api.batch_knn_filtered([0.0, 0.0], b, 2, lambda i: i % 2 == 0)
    assert list(r.indices) == [0, 2]


def test_filtered_knn_none_pass(api):
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [2.0, 0.0]])
    assert len(api.batch_knn_filtered([0.0, 0.0], b, 2, lambda i: False).indices) == 0


def test_filtered_knn_all_pass(api):
    b = api.VerticalBatch.from_rows([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    f = api.batch_knn_filtered([0.0, 0.0], b, 2, lambda i: True)
    u = api.batch_knn([0.0, 0.0], b, 2)
    assert list(f.indices) == list(u.indices)


def test_filtered_knn_k_larger_than_passing(api):
    b = api.VerticalBatch.from_rows([[1.0], [2.0], [3.0]])
    r = api.batch_knn_filtered([0.0], b, 10, lambda i: i == 0)
    assert list(r.indices) == [0]


def test_filtered_knn_preserves_original_indices(api):
    b = api.VerticalBatch.from_rows([[100.0], [100.0], [0.1], [100.0], [0.2]])
    r = api.batch_knn_filtered([0.0], b, 2, lambda i: i in (2, 4))
    assert list(r.indices) == [2, 4]


def test_filtered_knn_integration(api):
    b = api.VerticalBatch.from_rows([[float(i), 0.0] for i in range(100)])
    r = api.batch_knn_filtered([50.0, 0.0], b, 5, lambda i: i % 2 == 0)
    assert len(r.indices) == 5 and r.indices[0] == 50
    assert all(int(i) % 2 == 0 for i in r.indices)


# ---- batch_l2_squared_pruning: src/batch.rs:971-988, 1475-1505, tests/batch_tests.rs:297-349
def test_batch_pruning(api):
    b = api.VerticalBatch.from_rows([[0.0, 0.0], [1.0, 0.0], [10.0, 0.0]])
    s = api.batch_l2_squared_pruning([0.0, 0.0], b, 2.0)
    assert sorted(i for i, _ in s) == [0, 1]


def test_pruning_threshold_zero(api):
    b = api.VerticalBatch.from_rows([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    s = api.batch_l2_squared_pruning([0.0, 0.0], b, 0.0)
    assert len(s) == 1 and s[0][0] == 0 and abs(s[0][1]) < 1e-9


def test_pruning_all_and_none_survive(api):
    b = api.VerticalBatch.from_rows([[0.1, 0.0], [0.0, 0.1]])
    assert len(api.batch_l2_squared_pruning([0.0, 0.0], b, 100.0)) == 2
    b = api.VerticalBatch.from_rows([[10.0, 0.0], [0.0, 10.0]])
    assert api.batch_l2_squared_pruning([0.0, 0.0], b, 0.5) == []


def test_pruning_filters_far_vectors_and_distances(api):
    b = api.VerticalBatch.from_rows([[0.0, 0.0], [1.0, 0.0], [100.0, 100.0], [2.0, 0.0]])
    s = api.batch_l2_squared_pruning([0.0, 0.0], b, 5.0)
    assert [i for i, _ in s] == [0, 1, 3]
    full = api.batch_l2_squared([0.0, 0.0], b)
    for i, dist in s:
        assert abs(dist - float(full[i])) < 1e-6


def test_pruning_tight_threshold(api):
    b = api.VerticalBatch.from_rows([[float(i), 0.0] for i in range(100)])
    s = api.batch_l2_squared_pruning([50.0, 0.0], b, 4.0)
    idx = [i for i, _ in s]
    assert len(idx) <= 5 and 50 in idx


# ---- examples/batch_demo.rs:77-123 (knn == brute force at 20x8, k=3, generate_embedding)
def test_demo_knn_matches_bruteforce(api, oracle):
    dim, n, k = 8, 20, 3
    corpus = [oracle.generate_embedding(dim, i) for i in range(n)]
    q = oracle.generate_embedding(dim, 999)
    r = api.batch_knn(q, api.VerticalBatch.from_rows(corpus), k)
    naive = sorted(((float(np.sum((q.astype(np.float64) - v) ** 2)), i) for i, v in enumerate(corpus)))[:k]
    assert r.indices == [i for _, i in naive]


# ---- examples/batch_demo.rs:159-225 (checksum: batch vs naive, rel diff < 1e-3 at 10K x 128 x 100 -> reduced
#      query count keeps the CPU suite fast; shape of corpus is the example's)
def test_demo_timing_checksum(api, oracle):
    dim, n, nq = 128, 10_000, 4
    corpus = np.stack([oracle.generate_embedding(dim, i) for i in range(n)])
    b = api.VerticalBatch.from_flat(corpus.reshape(-1), n, dim)
    for j in range(nq):
        q = oracle.generate_embedding(dim, 50_000 + j)
        batch_sum = float(np.sum(api.batch_l2_squared(q, b), dtype=np.float64))
        naive_sum = float(np.sum((corpus.astype(np.float64) - q.astype(np.float64)) ** 2))
        assert abs(batch_sum - naive_sum) / max(abs(naive_sum), 1.0) < 1e-3


# ---- reordered kNN and dimension variance: src/batch.rs:1219-1264, tests/batch_tests.rs:413-426
def _sin_rows(n, d):
    return [[np.float32(math.sin(i * 7 + dd * 3)) for dd in range(d)] for i in range(n)]


def test_batch_knn_reordered_matches_exact(api):  # src/batch.rs:1220-1239
    b = api.VerticalBatch.from_rows(_sin_rows(50, 16))
    q = [np.float32(math.cos(i * 0.1)) for i in range(16)]
    exact, reordered = api.batch_knn(q, b, 5), api.batch_knn_reordered(q, b, 5)
    assert exact.indices == reordered.indices
    for e, r in zip(exact.scores, reordered.scores):
        assert abs(float(e) - float(r)) < 1e-4, f"distance mismatch: exact={e}, reordered={r}"


def test_batch_knn_reordered_empty(api):  # src/batch.rs:1241-1246
    b = api.VerticalBatch.from_rows([])
    assert api.batch_knn_reordered([], b, 5).indices == []


def test_batch_dimension_variance(api):  # src/batch.rs:1248-1264
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [1.0, 5.0], [1.0, 10.0]])
    var = api.batch_dimension_variance(b)
    assert abs(float(var[0])) < 1e-6, "constant dim should have 0 variance"
    assert float(var[1]) > 10.0
    assert float(var[1]) == float(np.float32(50.0) / np.float32(3.0))  # (25 + 0 + 25) / 3 in f32


def test_batch_dimension_variance_degenerate(api):  # src/batch.rs:573-575: <= 1 vector -> zeros
    assert list(api.batch_dimension_variance(api.VerticalBatch.from_rows([[3.0, -4.0, 5.0]]))) == [0.0, 0.0, 0.0]
    assert list(api.batch_dimension_variance(api.VerticalBatch.from_rows([]))) == []


def test_reordered_knn_matches_exact_large(api):  # tests/batch_tests.rs:413-426
    b = api.VerticalBatch.from_rows(_sin_rows(200, 64))
    q = [np.float32(math.cos(i * 0.1)) for i in range(64)]
    assert api.batch_knn(q, b, 10).indices == api.batch_knn_reordered(q, b, 10).indices


def test_reordered_knn_k_clamped_and_zero(api):  # src/batch.rs:624-631
    b = api.VerticalBatch.from_rows([[0.0, 1.0], [2.0, 2.0], [0.5, 0.5]])
    assert len(api.batch_knn_reordered([0.0, 0.0], b, 10).indices) == 3
    assert api.batch_knn_reordered([0.0, 0.0], b, 0).indices == []
    r = api.batch_knn_reordered([0.0, 0.0], b, 2)
    assert r.indices == [2, 0] and list(r.scores) == [0.5, 1.0]


def test_reordered_knn_ties_lower_index_first(api):  # stable sort_by(total_cmp), src/batch.rs:650-651
    b = api.VerticalBatch.from_rows([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0], [0.0, 0.0]])
    r = api.batch_knn_reordered([0.0, 0.0], b, 4)
    assert r.indices == [4, 0, 1, 2]


def test_into_batch_outputs_match_allocating_apis(api):  # tests/batch_tests.rs:188-213, src/batch.rs:915-947, 1008-1016
    b = api.VerticalBatch.from_rows([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0]])
    q = [1.0, 0.5, 0.0]
    out = [9.0] * 8  # stale contents must be cleared
    api.batch_l2_squared_into(q, b, out)
    assert out == list(api.batch_l2_squared(q, b))
    api.batch_dot_into(q, b, out)
    assert out == list(api.batch_dot(q, b))
    norms = api.batch_norms(b)
    norms_into = []
    api.batch_norms_into(b, norms_into)
    assert norms_into == list(norms)
    api.batch_cosine_into(q, b, norms, out)
    assert out == list(api.batch_cosine(q, b, norms))


# ---- batch_knn_adaptive: src/batch.rs:1509-1572, tests/batch_tests.rs:269-292
def test_batch_knn_adaptive_empty(api):
    assert api.batch_knn_adaptive([], api.VerticalBatch.from_rows([]), 5, 2).indices == []


def test_batch_knn_adaptive_k_zero(api):
    assert api.batch_knn_adaptive([1.0, 2.0], api.VerticalBatch.from_rows([[1.0, 2.0]]), 0, 1).indices == []


def test_batch_knn_adaptive_zero_warmup_panics(api):  # src/batch.rs:448 assert!(warmup_dims > 0, ..)
    with pytest.raises(AssertionError):
        api.batch_knn_adaptive([1.0, 2.0], api.VerticalBatch.from_rows([[1.0, 2.0]]), 1, 0)


def test_batch_knn_adaptive_finds_nearest(api):  # src/batch.rs:1527-1545
    b = api.VerticalBatch.from_rows([[0.0] * 4, [100.0] * 4, [0.1] * 4])
    assert api.batch_knn([0.0] * 4, b, 1).indices[0] == 0
    assert api.batch_knn_adaptive([0.0] * 4, b, 1, 2).indices[0] == 0


def test_batch_knn_adaptive_keeps_k_finalized_candidates(api):  # src/batch.rs:1547-1561
    b = api.VerticalBatch.from_rows([[0.0, 1.0]])
    adaptive, exact = api.batch_knn_adaptive([0.0, 0.0], b, 1, 1), api.batch_knn([0.0, 0.0], b, 1)
    assert len(adaptive.indices) == 1 and adaptive.indices == exact.indices
    assert list(adaptive.scores) == list(exact.scores)


def test_batch_knn_adaptive_zero_dimensional_batch_keeps_k(api):  # src/batch.rs:1563-1572
    r = api.batch_knn_adaptive([], api.VerticalBatch.from_rows([[], [], []]), 2, 1)
    assert r.indices == [0, 1] and list(r.scores) == [0.0, 0.0]


def test_knn_adaptive_matches_basic(api):  # tests/batch_tests.rs:269-292
    rows = [[np.float32(i), np.float32(math.sin(np.float32(i) * np.float32(0.1))),
             np.float32(math.cos(np.float32(i) * np.float32(0.1)))] for i in range(100)]
    b = api.VerticalBatch.from_rows(rows)
    basic, adaptive = api.batch_knn([50.0, 0.0, 1.0], b, 10), api.batch_knn_adaptive([50.0, 0.0, 1.0], b, 10, 1)
    for idx in basic.indices:
        assert idx in adaptive.indices, f"Adaptive missing index {idx} from basic top-10"
