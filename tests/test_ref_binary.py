"""Ports of /root/reference/src/binary.rs:217-488 (the binary_hamming / PackedBinary / encode_binary cases) and the
caller composition examples/binary_demo.rs:174-180 (stable sort_by_key, take k)."""
import numpy as np
import pytest

U64 = 0xFFFFFFFFFFFFFFFF


def test_new_masks_padding_bits(api):  # :217-225
    dirty = api.PackedBinary([U64], 8)
    zeros = api.PackedBinary([0], 8)
    assert api.binary_hamming(dirty, zeros) == 8


def test_new_len_panics(api):  # :50-57
    with pytest.raises(AssertionError):
        api.PackedBinary([0, 0], 8)


def test_binary_ops(api):  # :229-245 (hamming part)
    a, b = api.PackedBinary.zeros(4), api.PackedBinary.zeros(4)
    a.set(0, True); a.set(1, True); b.set(1, True); b.set(2, True)
    assert api.binary_hamming(a, b) == 2


def test_zeros_new_memory(api):  # :251-275
    v = api.PackedBinary.zeros(128)
    assert v.dimension == 128 and len(v.data) == 2 and not any(v.get(i) for i in range(128))
    w = api.PackedBinary([0xFF], 8)
    assert all(w.get(i) for i in range(8))
    assert api.PackedBinary.zeros(256).memory_bytes() == 32


def test_set_get_bounds(api):  # :281-315
    v = api.PackedBinary.zeros(64)
    v.set(0, True); assert v.get(0)
    v.set(0, False); assert not v.get(0)
    s = api.PackedBinary.zeros(4)
    s.set(100, True)
    assert not s.get(100) and not s.get(4) and not s.get(1000)
    v.set(63, True)
    assert v.get(63) and not v.get(62)


def test_multi_word_hamming(api):  # :321-338
    a, b = api.PackedBinary.zeros(128), api.PackedBinary.zeros(128)
    a.set(0, True); b.set(0, True); a.set(64, True); b.set(65, True)
    assert api.binary_hamming(a, b) == 2


def test_encode_binary_cases(api):  # :384-431, :478-487
    p = api.encode_binary([1.0, 2.0, 3.0, 4.0], 0.0)
    assert all(p.get(i) for i in range(4))
    p = api.encode_binary([-1.0, -2.0, -3.0, -4.0], 0.0)
    assert not any(p.get(i) for i in range(4))
    assert not api.encode_binary([0.0], 0.0).get(0)          # strict >
    assert api.encode_binary([], 0.0).dimension == 0
    v = [1.0 if i % 2 == 0 else -1.0 for i in range(768)]
    p = api.encode_binary(v, 0.0)
    assert p.dimension == 768 and len(p.data) == 12
    assert all(p.get(i) == (i % 2 == 0) for i in range(768))
    p = api.encode_binary([0.1, 0.5, 0.9, 1.5], 0.5)
    assert [p.get(i) for i in range(4)] == [False, False, True, True]


def test_hamming_identical_complement(api):  # :437-448
    v = api.encode_binary([1.0, -1.0, 1.0, -1.0], 0.0)
    assert api.binary_hamming(v, v) == 0
    a = api.encode_binary([1.0] * 4, 0.0)
    b = api.encode_binary([-1.0] * 4, 0.0)
    assert api.binary_hamming(a, b) == 4


def test_hamming_doc_example(api):  # src/binary.rs:145-151 doc test
    a = api.encode_binary([1.0, -1.0, 1.0, -1.0], 0.0)
    b = api.encode_binary([1.0, 1.0, -1.0, -1.0], 0.0)
    assert api.binary_hamming(a, b) == 2


def test_hamming_dimension_mismatch_panics(api):  # :155-159
    with pytest.raises(AssertionError):
        api.binary_hamming(api.PackedBinary.zeros(64), api.PackedBinary.zeros(128))


def test_hamming_topk_composition_small(api, oracle):  # examples/binary_demo.rs:144-180 at reduced size
    dim, n, k = 256, 500, 10
    docs = [api.encode_binary(oracle.generate_normalized(dim, i), 0.0) for i in range(n)]
    q = api.encode_binary(oracle.generate_normalized(dim, 100_000), 0.0)
    codes = np.stack([np.asarray(d.data, dtype=np.uint64) for d in docs])
    idx, dist = api.hamming_topk(np.asarray(q.data, dtype=np.uint64), codes, k)
    full = [api.binary_hamming(q, d) for d in docs]
    want = sorted(range(n), key=lambda i: (full[i], i))[:k]       # stable sort_by_key == (h, index)
    assert [int(i) for i in idx] == want
    assert [int(d) for d in dist] == [full[i] for i in want]


# ---- binary_dot / binary_jaccard: src/binary.rs:228-244, 330-372, 444-474, doc tests :170-174, :190-196
def test_dot_jaccard_basic(api):
    a, b = api.PackedBinary.zeros(4), api.PackedBinary.zeros(4)
    a.set(0, True); a.set(1, True)
    b.set(1, True); b.set(2, True)
    assert api.binary_hamming(a, b) == 2
    assert api.binary_dot(a, b) == 1
    assert abs(api.binary_jaccard(a, b) - 1.0 / 3.0) < 1e-6


def test_multi_word_dot_jaccard(api):
    a, b = api.PackedBinary.zeros(128), api.PackedBinary.zeros(128)
    for i in (0, 64, 65):
        a.set(i, True)
    for i in (0, 64, 100):
        b.set(i, True)
    assert api.binary_dot(a, b) == 2
    assert abs(api.binary_jaccard(a, b) - 0.5) < 1e-6


def test_dot_self_jaccard_identical_disjoint_empty(api):
    v = api.encode_binary([1.0, -1.0, 1.0, -1.0, 1.0], 0.0)
    assert api.binary_dot(v, v) == 3
    v = api.encode_binary([1.0, -1.0, 1.0], 0.0)
    assert abs(api.binary_jaccard(v, v) - 1.0) < 1e-6
    a, b = api.encode_binary([1.0, -1.0], 0.0), api.encode_binary([-1.0, 1.0], 0.0)
    assert abs(api.binary_jaccard(a, b)) < 1e-6
    assert abs(api.binary_jaccard(api.PackedBinary.zeros(4), api.PackedBinary.zeros(4)) - 1.0) < 1e-6


def test_dot_jaccard_doc_examples(api):
    a = api.encode_binary([1.0, -1.0, 1.0, -1.0], 0.0)
    b = api.encode_binary([1.0, 1.0, -1.0, -1.0], 0.0)
    assert api.binary_dot(a, b) == 1
    assert abs(api.binary_jaccard(a, b) - 1.0 / 3.0) < 1e-6
