"""Ports of /root/reference/src/maxsim.rs:200-381, tests/maxsim_tests.rs:173-188 and
examples/maxsim_colbert.rs:37-152 (hand-computable MaxSim cases). Run on the oracle and, on a GPU box, on the CUDA
product (f32 tolerance of the Rust tests: 1e-6 / 1e-5)."""
import numpy as np
import pytest


def test_maxsim_basic(api):  # src/maxsim.rs:200-215
    s = api.maxsim([[1.0, 0.0], [0.0, 1.0]], [[0.9, 0.1], [0.1, 0.9]])
    assert abs(s - 1.8) < 1e-6


def test_maxsim_empty(api):  # :217-225
    assert api.maxsim([[1.0, 0.0]], []) == 0.0
    assert api.maxsim([], [[1.0, 0.0]]) == 0.0


def test_maxsim_not_commutative(api):  # :227-246
    q, d = [[1.0, 0.0]], [[0.5, 0.5], [0.5, 0.5]]
    a, b = api.maxsim(q, d), api.maxsim(d, q)
    assert abs(a - 0.5) < 1e-6 and abs(b - 1.0) < 1e-6 and abs(a - b) > 0.4


def test_maxsim_cosine_normalized(api):  # :248-261
    q, d = [[1.0, 0.0]], [[1.0, 0.0]]
    assert abs(api.maxsim(q, d) - api.maxsim_cosine(q, d)) < 1e-6


def test_maxsim_cosine_unnormalized(api):  # :263-275
    assert abs(api.maxsim_cosine([[2.0, 0.0]], [[3.0, 0.0]]) - 1.0) < 1e-6


def test_maxsim_single_query_single_doc(api):  # :279-288
    assert abs(api.maxsim([[1.0, 2.0, 3.0]], [[4.0, 5.0, 6.0]]) - 32.0) < 1e-6


def test_maxsim_multiple_query_multiple_doc(api):  # :290-308
    q = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    d = [[0.5, 0.3, 0.0], [0.0, 0.7, 0.9]]
    assert abs(api.maxsim(q, d) - 2.1) < 1e-6


def test_maxsim_identical_embeddings(api):  # :310-319
    v = [1.0, 0.0, 0.0]
    assert abs(api.maxsim([v, v, v], [v, v]) - 3.0) < 1e-6


def test_maxsim_orthogonal_embeddings(api):  # :321-335
    q = [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]]
    d = [[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]
    assert abs(api.maxsim(q, d)) < 1e-6


def test_maxsim_cosine_orthogonal(api):  # :337-346
    assert abs(api.maxsim_cosine([[1.0, 0.0]], [[0.0, 1.0]])) < 1e-6


def test_maxsim_cosine_identical(api):  # :348-356
    v = [3.0, 4.0]
    assert abs(api.maxsim_cosine([v, v], [v]) - 2.0) < 1e-6


def test_maxsim_cosine_empty(api):  # :358-365
    assert api.maxsim_cosine([[1.0, 0.0]], []) == 0.0
    assert api.maxsim_cosine([], [[1.0, 0.0]]) == 0.0


def test_maxsim_higher_dim(api):  # :367-381
    q = [[1.0, 0, 0, 0, 0, 0, 0, 0]]
    d = [[0, 0, 0, 0, 0, 0, 0, 1.0], [0.5, 0.5, 0, 0, 0, 0, 0, 0]]
    assert abs(api.maxsim(q, d) - 0.5) < 1e-6


def test_maxsim_basic_example(api):  # tests/maxsim_tests.rs:173-188
    s = api.maxsim([[1.0, 0.0], [0.0, 1.0]], [[0.9, 0.1], [0.1, 0.9], [0.5, 0.5]])
    assert abs(s - 1.8) < 0.01


def test_maxsim_dim_mismatch_panics(api):  # src/maxsim.rs:103-110
    with pytest.raises(AssertionError):
        api.maxsim([[1.0, 0.0], [1.0]], [[1.0, 0.0]])
    with pytest.raises(AssertionError):
        api.maxsim([[1.0, 0.0]], [[1.0, 0.0, 0.0]])


def test_colbert_demo_basic_and_noncommutative(api):  # examples/maxsim_colbert.rs:37-105
    q = [[1.0, 0, 0, 0], [0, 1.0, 0, 0]]
    d = [[0.9, 0.1, 0, 0], [0.1, 0.8, 0, 0], [0.5, 0.5, 0, 0]]
    assert abs(api.maxsim(q, d) - 1.7) < 1e-5
    q1 = [[1.0, 0, 0, 0]]
    d3 = [[0.5, 0.5, 0, 0], [0.3, 0.7, 0, 0], [0.8, 0.2, 0, 0]]
    assert abs(api.maxsim(q1, d3) - 0.8) < 1e-5 and abs(api.maxsim(d3, q1) - 1.6) < 1e-5


def test_colbert_demo_realistic_scale(api, oracle):  # examples/maxsim_colbert.rs:107-152: 32 x 128 x 128d vs naive
    dim, nq, nd = 128, 32, 128
    q = np.stack([oracle.generate_normalized(dim, i) for i in range(nq)])
    d = np.stack([oracle.generate_normalized(dim, i + 1000) for i in range(nd)])
    naive = float(np.sum(np.max(q.astype(np.float64) @ d.astype(np.float64).T, axis=1)))
    assert abs(api.maxsim(q, d) - naive) < 1e-3
