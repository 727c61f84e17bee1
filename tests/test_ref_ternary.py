"""Ports of the reference's ternary tests (/root/reference/src/ternary.rs:336-520) against the oracle and, on a GPU box,
the CUDA product through the C-ABI (same scan shape, other codecs)."""
import numpy as np
import pytest


def test_new_masks_padding_pairs(api):  # :338-350
    dirty = api.PackedTernary(np.array([0xFFFFFFFFFFFF5555], np.uint64), 8)
    clean = api.PackedTernary.zeros(8)
    for i in range(8):
        clean.set(i, 1)
    assert dirty.nnz() == clean.nnz()
    assert np.array_equal(dirty.data, clean.data)


def test_encode_decode(api):  # :354-365
    p = api.encode_ternary([0.5, -0.5, 0.1, -0.1, 0.8, -0.8], 0.3)
    assert [p.get(i) for i in range(6)] == [1, -1, 0, 0, 1, -1]


def test_ternary_dot_same_opposite_orthogonal(api):  # :367-410
    a = api.PackedTernary.zeros(4)
    a.set(0, 1); a.set(1, -1); a.set(2, 0); a.set(3, 1)
    assert api.ternary_dot(a, a) == 3
    a, b = api.PackedTernary.zeros(4), api.PackedTernary.zeros(4)
    a.set(0, 1); a.set(1, -1); b.set(0, -1); b.set(1, 1)
    assert api.ternary_dot(a, b) == -2
    a, b = api.PackedTernary.zeros(4), api.PackedTernary.zeros(4)
    a.set(0, 1); b.set(1, 1)
    assert api.ternary_dot(a, b) == 0


def test_large_vector(api):  # :412-435
    values = [((i / 768.0) - 0.5) * (2.0 if i % 3 == 0 else 0.5) for i in range(768)]
    p = api.encode_ternary(np.array(values, np.float32), 0.3)
    assert p.data.size == 24 and p.memory_bytes() == 192
    assert api.ternary_dot(p, p) == p.nnz()


def test_asymmetric_dot(api):  # :437-450
    t = api.PackedTernary.zeros(4)
    t.set(0, 1); t.set(1, -1); t.set(2, 0); t.set(3, 1)
    assert abs(api.ternary_asymmetric_dot([0.5, 0.5, 0.5, 0.5], t) - 0.5) < 1e-6


def test_hamming(api):  # :452-468
    a, b = api.PackedTernary.zeros(4), api.PackedTernary.zeros(4)
    a.set(0, 1); a.set(1, -1); a.set(2, 1)
    b.set(0, 1); b.set(1, 1); b.set(2, -1)
    assert api.ternary_hamming(a, b) == 2


def test_accessors(api):  # :474-515
    v = api.PackedTernary.zeros(100)
    assert all(v.get(i) == 0 for i in range(100)) and v.nnz() == 0
    v = api.PackedTernary.zeros(3)
    v.set(0, 1); v.set(1, -1); v.set(2, 0)
    assert [v.get(i) for i in range(3)] == [1, -1, 0]
    v = api.PackedTernary.zeros(1)
    v.set(0, 1); assert v.get(0) == 1
    v.set(0, -1); assert v.get(0) == -1
    v.set(0, 0); assert v.get(0) == 0
    v = api.PackedTernary.zeros(4)
    v.set(100, 1)
    assert v.get(4) == 0 and v.get(1000) == 0


def test_word_boundary(api):  # :521-540
    v = api.PackedTernary.zeros(64)
    v.set(31, 1); v.set(32, -1)
    assert v.get(31) == 1 and v.get(32) == -1 and v.get(30) == 0 and v.get(33) == 0
    assert api.ternary_dot(v, v) == 2


def test_dimension_mismatch_panics(api):  # :192-196
    with pytest.raises(AssertionError):
        api.ternary_dot(api.PackedTernary.zeros(32), api.PackedTernary.zeros(64))
