"""Ports of /root/reference/src/scalar.rs:399-606 and the bit-exact integer-valued KAT
tests/simd_correctness.rs:365-388 (mixed f32 x u8 dot)."""
import math

import numpy as np
import pytest


def test_quantize_roundtrip(api):  # :399-417
    values = [0.0, 0.5, 1.0, -1.0, 0.25]
    p = api.QuantizationParams.fit(values)
    q = api.quantize_u8(values, p)
    assert q.dimension == 5
    for i, orig in enumerate(values):
        deq = p.alpha * (float(q.data[i]) / 255.0) + p.offset
        assert abs(orig - deq) < p.alpha / 255.0 + 1e-6


def test_quantize_range(api):  # :419-428
    p = api.QuantizationParams.fit([-1.0, 0.0, 1.0])
    q = api.quantize_u8([-1.0, 0.0, 1.0], p)
    assert q.data[0] == 0 and q.data[2] == 255 and abs(int(q.data[1]) - 128) <= 1


def test_asymmetric_dot_matches_exact(api):  # :430-448
    doc, query = [1.0, 2.0, 3.0, 4.0], [0.5] * 4
    exact = sum(d * q for d, q in zip(doc, query))
    p = api.QuantizationParams.fit(doc)
    approx = api.asymmetric_dot_u8(query, api.quantize_u8(doc, p), p)
    assert abs(exact - approx) < p.alpha / 255.0 * len(doc)


def test_mixed_dot_u8_f32_exact(api):  # :465-476 (assert_eq!, bit exact)
    query, codes = [0.5, -2.0, 3.0, 4.5], [2, 7, 11, 13]
    expected = np.float32(0)
    for q, c in zip(query, codes):
        expected = np.float32(expected + np.float32(q) * np.float32(c))
    assert np.float32(api.mixed_dot_u8_f32(query, codes)).tobytes() == expected.tobytes()


def test_mixed_dot_length_mismatch_panics(api):  # :478-482
    with pytest.raises(AssertionError, match="mixed_dot_u8_f32: slice length mismatch"):
        api.mixed_dot_u8_f32([1.0, 2.0], [1])


def test_quantize_empty_constant_params(api):  # :484-507
    p = api.QuantizationParams.fit([])
    q = api.quantize_u8([], p)
    assert q.dimension == 0 and q.memory_bytes() == 0
    p = api.QuantizationParams.fit([5.0] * 10)
    assert api.quantize_u8([5.0] * 10, p).dimension == 10
    p = api.QuantizationParams.from_range(-1.0, 1.0)
    assert abs(p.alpha - 2.0) < 1e-6 and abs(p.offset + 1.0) < 1e-6


def test_memory_bytes(api):  # :518-523
    p = api.QuantizationParams.from_range(0.0, 1.0)
    assert api.quantize_u8([0.5] * 768, p).memory_bytes() == 768


def test_asymmetric_dot_large(api):  # :525-546 (libm sin/cos: tolerance test, not bit test)
    dim = 128
    doc = np.array([math.sin(i * 0.1) for i in range(dim)], dtype=np.float32)
    query = np.array([math.cos(i * 0.3) for i in range(dim)], dtype=np.float32)
    exact = float(np.dot(doc.astype(np.float64), query.astype(np.float64)))
    p = api.QuantizationParams.fit(doc)
    approx = api.asymmetric_dot_u8(query, api.quantize_u8(doc, p), p)
    assert abs(exact - approx) < p.alpha / 255.0 * math.sqrt(dim) + 0.1


def test_asymmetric_dot_dimension_mismatch(api):  # :548-554
    p = api.QuantizationParams.from_range(0.0, 1.0)
    q = api.quantize_u8([0.5, 0.5], p)
    with pytest.raises(AssertionError, match="dimension mismatch"):
        api.asymmetric_dot_u8([1.0, 2.0, 3.0], q, p)


def test_batch_knn_u8(api):  # :581-599
    p = api.QuantizationParams.from_range(-1.0, 1.0)
    corpus = [api.quantize_u8(v, p) for v in ([1.0, 0, 0], [0, 1.0, 0], [-1.0, 0, 0], [0.7, 0.7, 0])]
    r = api.batch_knn_u8([1.0, 0.0, 0.0], corpus, p, 2)
    assert len(r) == 2 and r[0][0] in (0, 3) and r[0][1] >= r[1][1]


def test_batch_knn_u8_empty(api):  # :601-606
    p = api.QuantizationParams.from_range(0.0, 1.0)
    assert api.batch_knn_u8([1.0], [], p, 5) == []


def test_simd_correctness_mixed_dot_exact(api):  # tests/simd_correctness.rs:365-388 -- bit-exact KAT
    for dim in (8, 16, 31, 32, 33, 64, 65, 128):
        for seed in range(5):
            corpus = np.array([(i * 31 + seed * 7) % 256 for i in range(dim)], dtype=np.uint8)
            query = np.array([float((i * 13 + seed * 3) % 8) for i in range(dim)], dtype=np.float32)
            expect = np.float32(0)
            for x, y in zip(query, corpus):
                expect = np.float32(expect + x * np.float32(y))
            got = np.float32(api.mixed_dot_u8_f32(query, corpus))
            assert got.tobytes() == expect.tobytes(), (dim, seed)


def test_quantize_u8_formula_edges(api):  # src/scalar.rs:212-225: round half away, clamp, NaN -> 0
    p = api.QuantizationParams.from_range(0.0, 255.0)   # inv_alpha = 1
    q = api.quantize_u8([0.5, 1.5, 2.5, -3.0, 300.0, 254.5, float("nan")], p)
    assert list(q.data) == [1, 2, 3, 0, 255, 255, 0]


def test_precomputed_matches_direct(api):  # src/scalar.rs:448-462
    doc, query = [1.0, 2.0, 3.0], [0.5, 1.0, 1.5]
    params = api.QuantizationParams.fit(doc)
    quantized = api.quantize_u8(doc, params)
    direct = api.asymmetric_dot_u8(query, quantized, params)
    ctx = api.query_context(query)
    precomputed = api.asymmetric_dot_u8_precomputed(query, quantized, params, ctx)
    assert abs(direct - precomputed) < 1e-6, f"precomputed mismatch: direct={direct}, precomputed={precomputed}"
    assert ctx.query_sum == 3.0
    with pytest.raises(AssertionError):
        api.asymmetric_dot_u8_precomputed(query[:2], quantized, params, ctx)


def test_fit_quantile_clips_outliers(api):  # src/scalar.rs:557-579
    values = [np.float32(i) / np.float32(49.0) - np.float32(1.0) for i in range(98)] + [100.0, -100.0]
    full = api.QuantizationParams.fit(values)
    clipped = api.QuantizationParams.fit_quantile(values, 0.95)
    assert clipped.alpha < full.alpha, f"clipped alpha {clipped.alpha} should be < full alpha {full.alpha}"
    assert clipped.alpha < 10.0


def test_fit_quantile_edges(api):  # src/scalar.rs:104-137: quantile range assert, empty, q = 1 == fit, non-finite filtered
    with pytest.raises(AssertionError):
        api.QuantizationParams.fit_quantile([1.0], 0.0)
    with pytest.raises(AssertionError):
        api.QuantizationParams.fit_quantile([1.0], 1.5)
    p = api.QuantizationParams.fit_quantile([], 0.9)
    assert (p.alpha, p.offset) == (1.0, 0.0)
    vals = [3.0, -1.0, 7.5, 0.25]
    a, b = api.QuantizationParams.fit_quantile(vals, 1.0), api.QuantizationParams.fit(vals)
    assert (a.alpha, a.offset) == (b.alpha, b.offset)
    p = api.QuantizationParams.fit_quantile([float("nan"), float("inf")], 0.5)
    assert (p.alpha, p.offset) == (1.0, 0.0)


def test_fit_quantile_mirror_matches_oracle(oracle):
    """The host mirror's numpy restatement of fit_quantile against the oracle's, over random inputs and quantiles."""
    import innr_b200.scalar as mirror
    rng = np.random.default_rng(11)
    for _ in range(200):
        n = int(rng.integers(1, 400))
        v = (rng.standard_normal(n) * rng.choice([0.01, 1.0, 100.0])).astype(np.float32)
        if rng.random() < 0.3:
            v[rng.integers(0, n)] = rng.choice([np.inf, -np.inf, np.nan, -0.0, 0.0])
        q = float(rng.choice([0.5, 0.9, 0.95, 0.99, 0.999, 1.0, float(rng.uniform(0.01, 1.0))]))
        a, b = mirror.QuantizationParams.fit_quantile(v, q), oracle.QuantizationParams.fit_quantile(v, q)
        assert np.float32(a.alpha).tobytes() == np.float32(b.alpha).tobytes(), (n, q)
        assert np.float32(a.offset).tobytes() == np.float32(b.offset).tobytes(), (n, q)
