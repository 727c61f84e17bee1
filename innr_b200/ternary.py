"""Host-side mirror of innr::ternary (src/ternary.rs) over the CUDA C-ABI: `PackedTernary`, `encode_ternary`,
`ternary_dot`, `ternary_hamming`, `ternary_asymmetric_dot` (= ternary::asymmetric_dot), `ternary_sparsity`, plus the
corpus-level entries of the device path (`TernaryCorpus`, `ternary_scores_all`, `ternary_topk`). SURVEY.md 8f row 4."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .batch import _Handle

TERNARY_DOT, TERNARY_HAMMING, TERNARY_ASYMMETRIC_DOT = 0, 1, 2
_OPS = {"dot": TERNARY_DOT, "hamming": TERNARY_HAMMING, "asymmetric_dot": TERNARY_ASYMMETRIC_DOT}


class PackedTernary:  # src/ternary.rs:50-157
    def __init__(self, data, dimension: int):
        data = np.array(data, dtype=np.uint64).reshape(-1)
        expect = (dimension + 31) // 32
        assert data.size == expect, (
            f"PackedTernary: data length {data.size} doesn't match dimension {dimension} (expected {expect} words)")
        rem = dimension % 32
        if rem and data.size:  # mask padding pairs past `dimension` (src/ternary.rs:72-79)
            data[-1] &= np.uint64((1 << (rem * 2)) - 1)
        self.data = data
        self.dimension = int(dimension)

    @classmethod
    def zeros(cls, dimension: int):
        return cls(np.zeros((dimension + 31) // 32, np.uint64), dimension)

    def set(self, idx: int, val: int):
        if idx >= self.dimension:
            return
        w, b = idx // 32, (idx % 32) * 2
        cur = int(self.data[w]) & ~(0b11 << b) & 0xFFFFFFFFFFFFFFFF
        bits = 0b01 if val == 1 else (0b10 if val == -1 else 0)
        self.data[w] = np.uint64(cur | (bits << b))

    def get(self, idx: int) -> int:
        if idx >= self.dimension:
            return 0
        bits = (int(self.data[idx // 32]) >> ((idx % 32) * 2)) & 0b11
        return 1 if bits == 0b01 else (-1 if bits == 0b10 else 0)

    def nnz(self) -> int:
        return sum(1 for i in range(self.dimension) if self.get(i) != 0)

    def memory_bytes(self) -> int:
        return self.data.size * 8


def encode_ternary(values, threshold: float) -> PackedTernary:  # src/ternary.rs:163-173, on the device
    v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
    out = np.zeros((v.size + 31) // 32, np.uint64)
    L.call("innr_cuda_encode_ternary", v.ctypes.data_as(L.f32p), v.size, C.c_float(threshold),
           out.ctypes.data_as(L.u64p))
    return PackedTernary(out, v.size)


class TernaryCorpus:
    """Device-resident set of packed ternary codes (chunk-major layout, ternary.cu)."""

    def __init__(self, handle: _Handle, n: int, dimension: int, index_base: int = 0):
        self._handle = handle
        self.num_codes, self.dimension, self.index_base = int(n), int(dimension), int(index_base)

    @property
    def h(self):
        return self._handle.h

    @classmethod
    def from_words(cls, words, n: int, dimension: int, index_base: int = 0):
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1)
        assert w.size == n * ((dimension + 31) // 32)
        h = C.c_void_p()
        L.call("innr_cuda_upload_ternary", w.ctypes.data_as(L.u64p), n, dimension, index_base, C.byref(h))
        return cls(_Handle(h), n, dimension, index_base)

    @classmethod
    def from_codes(cls, codes, index_base: int = 0):
        codes = list(codes)
        if not codes:
            return cls.from_words(np.zeros(0, np.uint64), 0, 0, index_base)
        dim = codes[0].dimension
        for c in codes:
            assert c.dimension == dim
        return cls.from_words(np.stack([c.data for c in codes]) if dim else np.zeros(0, np.uint64), len(codes), dim,
                              index_base)

    @classmethod
    def from_f32(cls, batch, threshold: float):
        """encode_ternary of every vector of a device-resident f32 corpus, on the device."""
        from .batch import _dev
        dev = _dev(batch)
        h = C.c_void_p()
        L.call("innr_cuda_ternary_from_f32", dev.h, C.c_float(threshold), C.byref(h))
        return cls(_Handle(h), dev.num_vectors, dev.dimension, dev.index_base)


def _query_arg(op: int, query):
    if op == TERNARY_ASYMMETRIC_DOT:
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        return q, q.size
    assert isinstance(query, PackedTernary)
    return np.ascontiguousarray(query.data, dtype=np.uint64), query.dimension


def ternary_scores_all(op: str, query, corpus: TernaryCorpus) -> np.ndarray:
    """op: 'dot' / 'hamming' (query: PackedTernary; int32 result) or 'asymmetric_dot' (query: f32; float32 result)."""
    o = _OPS[op]
    q, qdim = _query_arg(o, query)
    if o == TERNARY_ASYMMETRIC_DOT:
        out = np.zeros(corpus.num_codes, np.float32)
        L.call("innr_cuda_ternary_scores_all", corpus.h, o, C.c_void_p(q.ctypes.data), qdim, out.ctypes.data_as(L.f32p), None)
        return out
    out = np.zeros(corpus.num_codes, np.int32)
    L.call("innr_cuda_ternary_scores_all", corpus.h, o, C.c_void_p(q.ctypes.data), qdim, None,
           out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def ternary_topk(op: str, query, corpus: TernaryCorpus, k: int):
    o = _OPS[op]
    q, qdim = _query_arg(o, query)
    kk = max(k, 1)
    idx = np.zeros(kk, np.uint64)
    sc = np.zeros(kk, np.float32)
    cnt = C.c_size_t(0)
    L.call("innr_cuda_ternary_topk", corpus.h, o, C.c_void_p(q.ctypes.data), qdim, k, idx.ctypes.data_as(L.u64p),
           sc.ctypes.data_as(L.f32p), C.byref(cnt))
    return idx[:cnt.value], sc[:cnt.value]


def ternary_dot(a: PackedTernary, b: PackedTernary) -> int:  # src/ternary.rs:191 (pairwise; 1-code corpus)
    assert a.dimension == b.dimension, f"innr::ternary_dot: dimension mismatch ({a.dimension} vs {b.dimension})"
    if a.dimension == 0:
        return 0
    return int(ternary_scores_all("dot", a, TernaryCorpus.from_codes([b]))[0])


def ternary_hamming(a: PackedTernary, b: PackedTernary) -> int:  # src/ternary.rs:301
    assert a.dimension == b.dimension
    if a.dimension == 0:
        return 0
    return int(ternary_scores_all("hamming", a, TernaryCorpus.from_codes([b]))[0])


def ternary_asymmetric_dot(query, t: PackedTernary) -> float:  # src/ternary.rs:286 (ternary::asymmetric_dot)
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
    assert q.size == t.dimension
    if t.dimension == 0:
        return 0.0
    return float(ternary_scores_all("asymmetric_dot", q, TernaryCorpus.from_codes([t]))[0])


def ternary_sparsity(v: PackedTernary) -> float:  # src/ternary.rs:327
    if v.dimension == 0:
        return 0.0
    return float(np.float32(1.0) - np.float32(v.nnz()) / np.float32(v.dimension))


# the reference's own names inside `innr::ternary` (src/ternary.rs:286, :327)
asymmetric_dot = ternary_asymmetric_dot
sparsity = ternary_sparsity
