"""Ports of /root/reference/src/topk.rs:189-347 against the oracle's TopK restatement and the device-backed
`innr_b200.TopK`, plus the same cases pushed through `api.topk_from_distances` (device analogue: N inserts in id order == one fused selection; L2-path tie
rule SURVEY.md 8a row T: exact-tie groups compare as sets)."""
import math

import numpy as np
import pytest


def test_nan_candidate_does_not_poison_topk(api):  # :192-208
    tk = api.TopK(2)
    tk.insert(0, float("nan"))
    tk.insert(1, 1.0)
    tk.insert(2, 0.5)
    ids = [i for i, _ in tk.into_sorted()]
    assert 2 in ids and 1 in ids


def test_basic_top3(api):  # :213-228
    top = api.TopK(3)
    for i, d in enumerate([1.5, 0.3, 2.0, 0.8, 5.0]):
        top.insert(i, d)
    assert len(top) == 3
    r = top.into_sorted()
    assert [(i, np.float32(d)) for i, d in r] == [(1, np.float32(0.3)), (3, np.float32(0.8)), (0, np.float32(1.5))]


def test_threshold_tracking(api):  # :230-251
    top = api.TopK(3)
    assert top.threshold() == math.inf
    top.insert(0, 1.0)
    assert top.threshold() == math.inf
    top.insert(1, 2.0)
    assert top.threshold() == math.inf
    top.insert(2, 3.0)
    assert top.threshold() == 3.0
    top.insert(3, 1.5)
    assert top.threshold() == 2.0
    top.insert(4, 0.5)
    assert top.threshold() == 1.5
    top.insert(5, 10.0)
    assert top.threshold() == 1.5


def test_duplicate_distances(api):  # :254-266
    top = api.TopK(3)
    for i in range(4):
        top.insert(i, 1.0)
    assert len(top) == 3
    r = top.into_sorted()
    assert len(r) == 3 and all(d == 1.0 for _, d in r)


def test_k1_edge_case(api):  # :268-283
    top = api.TopK(1)
    assert top.threshold() == math.inf
    for i, (d, t) in enumerate([(5.0, 5.0), (3.0, 3.0), (10.0, 3.0), (1.0, 1.0)]):
        top.insert(i, d)
        assert top.threshold() == t
    assert top.into_sorted() == [(3, 1.0)]


def test_large_n_k10(api):  # :285-300
    top = api.TopK(10)
    for i in range(10_000):
        top.insert(i, float(i))
    r = top.into_sorted()
    assert [i for i, _ in r] == list(range(10)) and [d for _, d in r] == [float(i) for i in range(10)]


def test_sorted_output_ascending(api):  # :302-316
    top = api.TopK(5)
    for i in reversed(range(5)):
        top.insert(i, float(i))
    r = top.into_sorted()
    assert all(r[i][1] <= r[i + 1][1] for i in range(len(r) - 1))


def test_is_empty_and_len(api):  # :318-333
    top = api.TopK(4)
    assert top.is_empty() and len(top) == 0
    top.insert(0, 1.0)
    assert not top.is_empty() and len(top) == 1
    for i in (1, 2, 3):
        top.insert(i, float(i + 1))
    assert len(top) == 4
    top.insert(4, 5.0)
    assert len(top) == 4


def test_insert_in_sorted_order(api):  # :335-346
    top = api.TopK(4)
    for i in range(4):
        top.insert(i, float(i + 1))
    top.insert(4, 0.5)
    r = top.into_sorted()
    assert r[0] == (4, 0.5) and r[3] == (2, 3.0)


def test_new_zero_panics(api):  # :65 assert!(k > 0, "innr::TopK: k must be >= 1")
    with pytest.raises(AssertionError):
        api.TopK(0)


def test_binary_search_tie_drift_matches_survey(oracle):
    """with the >=1.82 branchless binary_search_by, inserting equal ids 0,1,2,3 into k=4 leaves
    buffer [1,2,3,0] -> into_sorted() ids [0,3,2,1]. Pins the [RECALLED] probe sequence of the restatement."""
    top = oracle.TopK(4)
    for i in range(4):
        top.insert(i, 1.0)
    assert [i for i, _ in top.into_sorted()] == [0, 3, 2, 1]


# ---------------------------------------------------------------- device analogue (also run on the oracle)
def _via_api(api, dists, k):
    return api.topk_from_distances(np.asarray(dists, dtype=np.float32), k)


def _check_topk_sets(dists, k, got):
    """Row-T parity rule: scores exact; indices exact outside tie groups; tied indices must carry that score."""
    dists = np.asarray(dists, dtype=np.float32)
    order = sorted(range(len(dists)), key=lambda i: (_total_key(dists[i]), i))[:k]
    want_scores = [dists[i] for i in order]
    got_ids = [i for i, _ in got]
    got_scores = [np.float32(d) for _, d in got]
    assert len(got) == min(k, len(dists))
    assert len(set(got_ids)) == len(got_ids)
    for w, g in zip(want_scores, got_scores):
        assert w.tobytes() == g.tobytes() or (np.isnan(w) and np.isnan(g))
    for i, s in zip(got_ids, got_scores):
        assert dists[i].tobytes() == s.tobytes() or (np.isnan(s) and np.isnan(dists[i]))


def _total_key(x):
    b = int(np.float32(x).view(np.int32))
    b ^= ((b >> 31) & 0xFFFFFFFF) >> 1
    return b


def test_topk_api_cases(api):
    cases = [
        ([1.5, 0.3, 2.0, 0.8, 5.0], 3),
        ([float("nan"), 1.0, 0.5], 2),
        ([1.0, 1.0, 1.0, 1.0], 3),
        ([5.0, 3.0, 10.0, 1.0], 1),
        ([float(i) for i in range(10_000)], 10),
        ([4.0, 3.0, 2.0, 1.0, 0.0], 5),
        ([0.0, -0.0, 0.0, -0.0], 2),
        ([float("inf"), -float("inf"), 1.0], 3),
    ]
    for dists, k in cases:
        _check_topk_sets(dists, k, _via_api(api, dists, k))
    r = _via_api(api, [1.5, 0.3, 2.0, 0.8, 5.0], 3)
    assert [i for i, _ in r] == [1, 3, 0]
    r = _via_api(api, [float(i) for i in range(10_000)], 10)
    assert [i for i, _ in r] == list(range(10))
