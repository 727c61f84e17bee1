"""Committed fixtures under tests/golden/:

* reference_kats.json -- known answers stated by the reference's own tests (file:line each), run against the oracle on
  the CPU and against the CUDA product on the GPU (`api` fixture);
* configs.npz -- the oracle's answers for the five BASELINE.json configurations at reduced N on the bench's own
  stateless inputs (tests/golden/make_golden.py). The CPU suite recomputes them (any drift of the oracle shows up as a
  diff against the committed file); the GPU suite checks the CUDA path against the committed file WITHOUT calling the
  oracle -- bit-exact for the scans, <= 1e-5 relative for MaxSim.
"""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
GOLD = np.load(os.path.join(HERE, "golden", "configs.npz"))

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------------ reference KATs
@pytest.mark.parametrize("case", KATS["maxsim"], ids=lambda c: c["ref"])
def test_kat_maxsim(api, case):
    got = getattr(api, case["fn"])([np.float32(t) for t in case["query"]], [np.float32(t) for t in case["doc"]])
    assert abs(got - case["expect"]) <= case["tol"], (case["ref"], got)


@pytest.mark.parametrize("case", KATS["batch_scores"], ids=lambda c: c["ref"])
def test_kat_batch_scores(api, case):
    b = api.VerticalBatch.from_rows(case["rows"])
    if case["fn"] == "batch_cosine":
        got = api.batch_cosine(case["query"], b, api.batch_norms(b))
    else:
        got = getattr(api, case["fn"])(case["query"], b)
    assert np.all(np.abs(np.asarray(got, np.float64) - np.asarray(case["expect"])) <= max(case["tol"], 1e-6)), (case["ref"], got)


@pytest.mark.parametrize("case", KATS["batch_knn"], ids=lambda c: c["ref"])
def test_kat_batch_knn(api, case):
    r = getattr(api, case["fn"])(case["query"], api.VerticalBatch.from_rows(case["rows"]), case["k"])
    assert sorted(r.indices) == sorted(case["expect_set"]), case["ref"]


def test_kat_pruning_and_variance(api):
    for case in KATS["pruning"]:
        b = api.VerticalBatch.from_rows(case["rows"])
        got = api.batch_l2_squared_pruning(case["query"], b, case["threshold"])
        assert [i for i, _ in got] == case["expect_indices"], case["ref"]
    for case in KATS["variance"]:
        var = api.batch_dimension_variance(api.VerticalBatch.from_rows(case["rows"]))
        assert np.all(np.abs(np.asarray(var, np.float64) - np.asarray(case["expect"])) <= case["tol"]), case["ref"]


def test_kat_binary(api):
    for case in KATS["binary"]:
        dim = case["dim"]
        if "a_words" in case:
            a = api.PackedBinary(np.array(case["a_words"], np.uint64), dim)
            b = api.PackedBinary(np.array(case["b_words"], np.uint64), dim)
        else:
            a, b = api.PackedBinary.zeros(dim), api.PackedBinary.zeros(dim)
            for i in case["a_bits"]:
                a.set(i, True)
            for i in case["b_bits"]:
                b.set(i, True)
        assert api.binary_hamming(a, b) == case["hamming"], case["ref"]
        if "dot" in case:
            assert api.binary_dot(a, b) == case["dot"], case["ref"]
            assert abs(api.binary_jaccard(a, b) - case["jaccard"]) <= case["tol"], case["ref"]


def test_kat_mixed_dot_exact(api):
    for case in KATS["mixed_dot_exact"]:
        if "query" in case:
            assert api.mixed_dot_u8_f32(case["query"], np.array(case["codes"], np.uint8)) == case["expect"], case["ref"]
            continue
        for dim in case["dims"]:          # every product and partial sum is an exactly representable integer
            for seed in case["seeds"]:
                corpus = np.array([(i * 31 + seed * 7) % 256 for i in range(dim)], np.uint8)
                query = np.array([(i * 13 + seed * 3) % 8 for i in range(dim)], np.float32)
                expect = float(sum(int(q) * int(c) for q, c in zip(query, corpus)))
                assert api.mixed_dot_u8_f32(query, corpus) == expect, (case["ref"], dim, seed)


def test_kat_topk(api):
    for case in KATS["topk"]:
        t = api.TopK(case["k"])
        for i, d in case["inserts"]:
            t.insert(i, d)
        assert [(i, np.float32(d)) for i, d in t.into_sorted()] == [(i, np.float32(d)) for i, d in case["expect"]], case["ref"]


# ------------------------------------------------------------------------------------------------ config vectors
def test_golden_configs_match_oracle(oracle):
    """The committed vectors are what the oracle computes today (run make_golden.py after an intended oracle change)."""
    _spec.loader.exec_module(make_golden)
    fresh = make_golden.build()
    assert sorted(fresh) == sorted(GOLD.files)
    for key in GOLD.files:
        if key.startswith("c3_"):
            assert np.array_equal(bits(fresh[key]), bits(GOLD[key])), key
        else:
            assert np.array_equal(fresh[key], GOLD[key]), key


@pytest.mark.gpu
def test_golden_configs_cuda():
    """The CUDA path against the committed vectors, inputs generated on the device by the bench's generators."""
    import innr_b200 as ib
    ib.init(0)
    _spec.loader.exec_module(make_golden)
    S, SC, SQ, SCO = make_golden.SHAPES, make_golden.SALT_CORPUS, make_golden.SALT_QUERY, make_golden.SALT_CODES

    from innr_b200 import synth   # the bench's host-side twin of the device generators (numpy, no oracle)

    def ghash_f32(salt, count):
        return synth.ghash_f32(salt, 0, count)

    # C1
    s = S["c1"]
    b = ib.DeviceBatch.generate("gref", 0, 0, s["n"], s["d"])
    qs = np.stack([ib.DeviceBatch.generate("gref", 0, 50_000 + j, 1, s["d"]).extract_vector(0) for j in range(s["nq"])])  # row = seed
    idx, sc = ib.batch_knn_many("dot", qs, b, s["k"])
    assert np.array_equal(idx.astype(np.uint32), GOLD["c1_idx"]) and np.array_equal(bits(sc), GOLD["c1_score_bits"])
    # C2
    s = S["c2"]
    b = ib.DeviceBatch.generate("ghash", SC, 0, s["n"], s["d"])
    qs = ghash_f32(SQ, s["nq"] * s["d"]).reshape(s["nq"], s["d"])
    for metric in ("cosine", "dot"):
        idx, sc = ib.batch_knn_many(metric, qs, b, s["k"])
        assert np.array_equal(idx.astype(np.uint32), GOLD[f"c2_{metric}_idx"]), metric
        assert np.array_equal(bits(sc), GOLD[f"c2_{metric}_score_bits"]), metric
    idx, sc = ib.batch_knn_many("l2", qs, b, s["k"])   # continuous data: no exact ties, TopK order is determined
    assert np.array_equal(idx.astype(np.uint32), GOLD["c2_l2_idx"]) and np.array_equal(bits(sc), GOLD["c2_l2_score_bits"])
    assert np.array_equal(bits(ib.batch_dimension_variance(b)), GOLD["c2_variance_bits"])
    r = ib.batch_knn_reordered(qs[0], b, s["k"])
    assert np.array_equal(np.array(r.indices, np.uint32), GOLD["c2_reordered_idx"])
    assert np.array_equal(bits(r.scores), GOLD["c2_reordered_score_bits"])
    r = ib.batch_knn_adaptive(qs[0], b, s["k"], 32)
    assert np.array_equal(np.array(r.indices, np.uint32), GOLD["c2_adaptive_idx"])
    assert np.array_equal(bits(r.scores), GOLD["c2_adaptive_score_bits"])
    # C3 (f32 tolerance of north_star: 1e-5 relative)
    s = S["c3"]
    corpus = ib.TokenCorpus.generate(SC, 0, s["n_docs"], s["nt"], s["dim"])
    q = ghash_f32(SQ, s["nq"] * s["dim"]).reshape(s["nq"], s["dim"])
    for cos, key in ((False, "c3_maxsim"), (True, "c3_maxsim_cosine")):
        got, want = ib.maxsim_corpus(q, corpus, cosine=cos).astype(np.float64), GOLD[key].astype(np.float64)
        assert float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-30))) < 1e-5, key
    # C4
    s = S["c4"]
    codes = ib.BinaryCorpus.generate(SCO, 0, s["n"], s["dim"])
    qw = synth.ghash_u64(SQ, 0, s["nq"] * 16).reshape(s["nq"], 16)
    idx, dist = ib.hamming_topk_many(qw, codes, s["k"])
    assert np.array_equal(idx.astype(np.uint32), GOLD["c4_idx"]) and np.array_equal(dist.astype(np.uint32), GOLD["c4_dist"])
    # C5
    s = S["c5"]
    p = ib.QuantizationParams.from_range(-1.0, 1.0)
    c8 = ib.U8Corpus.generate(SC, 0, s["n"], s["d"], p)
    qs = ghash_f32(SQ, s["nq"] * s["d"]).reshape(s["nq"], s["d"])
    idx, sc = ib.batch_knn_u8_many(qs, c8, s["k"])
    assert np.array_equal(idx.astype(np.uint32), GOLD["c5_idx"]) and np.array_equal(bits(sc), GOLD["c5_score_bits"])


def test_host_generators_match_oracle(oracle):
    """innr_b200/synth.py (numpy twin of the device generators, used for bench and test queries) against the oracle's
    C++ generators: same splitmix64 stream, same 24-bit f32 mapping, for arbitrary salts and offsets (wrap-around too)."""
    from innr_b200 import synth
    for salt, first, count in ((synth.SALT_CORPUS, 0, 1000), (synth.SALT_QUERY, 12345, 777), (synth.SALT_CODES, 2**40 + 5, 64),
                               (0xFFFFFFFFFFFFFFF0, 0, 64), (0, 0, 3)):
        assert np.array_equal(synth.ghash_u64(salt, first, count), oracle.ghash_u64(salt, first, count)), (salt, first)
        a, b = synth.ghash_f32(salt, first, count), oracle.ghash_f32(salt, first, count)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (salt, first)
        assert float(a.min()) >= -1.0 and float(a.max()) < 1.0
