"""ctypes/numpy front-end for the CPU oracle (oracle/libinnr_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this module; nothing under
innr_b200/ does.  Function names follow the reference (innr 0.6.3) so the parity
tests read like the reference's own tests; each wrapper cites the same file:line as
the C++ restatement it calls (see innr_ref.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libinnr_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (g++, -ffp-contract=off)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("innr_ref.cpp", "innr_ref.h", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None

_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        sz, f32, i = C.c_size_t, C.c_float, C.c_int
        sig = {
            "innr_ref_host_has_avx512": (i, []),
            "innr_ref_host_has_avx2_fma": (i, []),
            "innr_ref_set_simd_mode": (None, [i]),
            "innr_ref_dense_backend": (C.c_char_p, [sz]),
            "innr_ref_from_flat": (None, [_f32p, sz, sz, _f32p]),
            "innr_ref_extract_vector": (None, [_f32p, sz, sz, sz, _f32p]),
            "innr_ref_batch_l2_squared": (None, [_f32p, _f32p, sz, sz, _f32p]),
            "innr_ref_batch_dot": (None, [_f32p, _f32p, sz, sz, _f32p]),
            "innr_ref_batch_norms": (None, [_f32p, sz, sz, _f32p]),
            "innr_ref_batch_cosine": (None, [_f32p, _f32p, sz, sz, _f32p, _f32p]),
            "innr_ref_batch_knn": (sz, [_f32p, _f32p, sz, sz, sz, _u64p, _f32p]),
            "innr_ref_batch_knn_dot": (sz, [_f32p, _f32p, sz, sz, sz, _u64p, _f32p]),
            "innr_ref_batch_knn_cosine": (sz, [_f32p, _f32p, sz, sz, sz, _u64p, _f32p]),
            "innr_ref_batch_knn_filtered": (sz, [_f32p, _f32p, sz, sz, sz, _u8p, _u64p, _f32p]),
            "innr_ref_batch_l2_squared_pruning": (sz, [_f32p, _f32p, sz, sz, f32, _u64p, _f32p]),
            "innr_ref_batch_knn_adaptive": (sz, [_f32p, _f32p, sz, sz, sz, sz, _u64p, _f32p]),
            "innr_ref_batch_dimension_variance": (None, [_f32p, sz, sz, _f32p]),
            "innr_ref_variance_order": (None, [_f32p, sz, _u64p]),
            "innr_ref_batch_knn_reordered": (sz, [_f32p, _f32p, sz, sz, sz, _u64p, _f32p]),
            "innr_ref_qparams_fit_quantile": (i, [_f32p, sz, f32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
            "innr_ref_asymmetric_dot_u8_precomputed": (f32, [_f32p, _u8p, sz, f32, f32, f32]),
            "innr_ref_topk_new": (C.c_void_p, [sz]),
            "innr_ref_topk_free": (None, [C.c_void_p]),
            "innr_ref_topk_insert": (None, [C.c_void_p, C.c_uint32, f32]),
            "innr_ref_topk_threshold": (f32, [C.c_void_p]),
            "innr_ref_topk_len": (sz, [C.c_void_p]),
            "innr_ref_topk_into_sorted": (sz, [C.c_void_p, _u32p, _f32p]),
            "innr_ref_maxsim": (f32, [_f32p, sz, _f32p, sz, sz]),
            "innr_ref_maxsim_cosine": (f32, [_f32p, sz, _f32p, sz, sz]),
            "innr_ref_maxsim_corpus": (None, [_f32p, sz, _f32p, _u64p, sz, sz, i, _f32p, i]),
            "innr_ref_packed_binary_mask": (None, [_u64p, sz]),
            "innr_ref_encode_binary": (None, [_f32p, sz, f32, _u64p]),
            "innr_ref_binary_hamming": (C.c_uint32, [_u64p, _u64p, sz]),
            "innr_ref_binary_dot": (C.c_uint32, [_u64p, _u64p, sz]),
            "innr_ref_packed_ternary_mask": (None, [_u64p, sz]),
            "innr_ref_encode_ternary": (None, [_f32p, sz, f32, _u64p]),
            "innr_ref_ternary_dot": (C.c_int32, [_u64p, _u64p, sz]),
            "innr_ref_ternary_hamming": (C.c_uint32, [_u64p, _u64p, sz]),
            "innr_ref_ternary_asymmetric_dot": (C.c_float, [_f32p, _u64p, sz]),
            "innr_ref_binary_jaccard": (C.c_float, [_u64p, _u64p, sz]),
            "innr_ref_hamming_topk": (sz, [_u64p, _u64p, sz, sz, sz, _u64p, _u32p]),
            "innr_ref_qparams_from_range": (None, [f32, f32, _f32p, _f32p]),
            "innr_ref_qparams_fit": (None, [_f32p, sz, _f32p, _f32p]),
            "innr_ref_quantize_u8": (None, [_f32p, sz, f32, f32, _u8p]),
            "innr_ref_query_sum": (f32, [_f32p, sz]),
            "innr_ref_mixed_dot_u8_f32": (f32, [_f32p, _u8p, sz]),
            "innr_ref_mixed_dot_u8_f32_portable": (f32, [_f32p, _u8p, sz]),
            "innr_ref_dot_u8_f32_avx2_intrin": (f32, [_f32p, _u8p, sz]),
            "innr_ref_dot_u8_f32_avx2_emul": (f32, [_f32p, _u8p, sz]),
            "innr_ref_asymmetric_dot_u8": (f32, [_f32p, _u8p, sz, f32, f32]),
            "innr_ref_batch_knn_u8": (sz, [_f32p, _u8p, sz, sz, f32, f32, sz, _u64p, _f32p]),
            "innr_ref_generate_embedding": (None, [sz, C.c_uint64, _f32p]),
            "innr_ref_generate_normalized": (None, [sz, C.c_uint64, _f32p]),
            "innr_ref_splitmix64": (C.c_uint64, [C.c_uint64]),
            "innr_ref_ghash_f32": (None, [C.c_uint64, C.c_uint64, sz, _f32p]),
            "innr_ref_ghash_u64": (None, [C.c_uint64, C.c_uint64, sz, _u64p]),
            "innr_ref_ghash_f32_mt": (None, [C.c_uint64, C.c_uint64, sz, _f32p, i]),
            "innr_ref_ghash_u64_mt": (None, [C.c_uint64, C.c_uint64, sz, _u64p, i]),
            "innr_ref_ghash_pdx_mt": (None, [C.c_uint64, C.c_uint64, sz, sz, _f32p, i]),
            "innr_ref_ghash_u8_mt": (None, [C.c_uint64, C.c_uint64, sz, sz, f32, f32, _u8p, i]),
            "innr_ref_batch_knn_many": (sz, [i, _f32p, sz, _f32p, sz, sz, sz, _u64p, _f32p, i]),
            "innr_ref_hamming_topk_many": (sz, [_u64p, sz, _u64p, sz, sz, sz, _u64p, _u32p, i]),
            "innr_ref_batch_knn_u8_many": (sz, [_f32p, sz, _u8p, sz, sz, f32, f32, sz, _u64p, _f32p, i]),
        }
        for name in ("dot", "cosine", "dot_portable", "cosine_portable", "dot_avx512_intrin", "dot_avx512_emul",
                     "dot_avx2_intrin", "dot_avx2_emul", "cosine_avx512_intrin", "cosine_avx512_emul",
                     "cosine_avx2_intrin", "cosine_avx2_emul"):
            sig["innr_ref_" + name] = (f32, [_f32p, _f32p, sz])
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


# --------------------------------------------------------------------------- helpers
def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


def host_has_avx512() -> bool:
    return bool(lib().innr_ref_host_has_avx512())


def set_simd_mode(mode: int) -> None:
    lib().innr_ref_set_simd_mode(int(mode))


def dense_backend(n: int) -> str:
    return lib().innr_ref_dense_backend(n).decode()


def _pair(name):
    def f(a, b):
        a, b = _f32(a), _f32(b)
        assert a.shape == b.shape and a.ndim == 1
        return float(getattr(lib(), "innr_ref_" + name)(_p(a, _f32p), _p(b, _f32p), a.size))

    f.__name__ = name
    return f


dot = _pair("dot")
cosine = _pair("cosine")
dot_portable = _pair("dot_portable")
cosine_portable = _pair("cosine_portable")
dot_avx512_intrin = _pair("dot_avx512_intrin")
dot_avx512_emul = _pair("dot_avx512_emul")
dot_avx2_intrin = _pair("dot_avx2_intrin")
dot_avx2_emul = _pair("dot_avx2_emul")
cosine_avx512_intrin = _pair("cosine_avx512_intrin")
cosine_avx512_emul = _pair("cosine_avx512_emul")
cosine_avx2_intrin = _pair("cosine_avx2_intrin")
cosine_avx2_emul = _pair("cosine_avx2_emul")


# --------------------------------------------------------------------------- VerticalBatch
class VerticalBatch:
    """src/batch.rs:88-220. `data` is the dimension-major buffer data[d*N + i]."""

    def __init__(self, data: np.ndarray, num_vectors: int, dimension: int):
        self.data = _f32(data).reshape(-1)
        self.num_vectors = int(num_vectors)
        self.dimension = int(dimension)
        assert self.data.size == self.num_vectors * self.dimension

    @classmethod
    def from_rows(cls, rows) -> "VerticalBatch":  # :103
        if len(rows) == 0:
            return cls(np.zeros(0, np.float32), 0, 0)
        d = len(rows[0])
        for r in rows:
            if len(r) != d:
                raise AssertionError("Inconsistent vector dimension")
        flat = _f32(np.array(rows, dtype=np.float32).reshape(len(rows), d))
        return cls.from_flat(flat.reshape(-1), len(rows), d)

    from_slices = from_rows  # :138

    @classmethod
    def from_flat(cls, data, num_vectors: int, dimension: int) -> "VerticalBatch":  # :167
        data = _f32(data).reshape(-1)
        assert data.size == num_vectors * dimension
        out = np.zeros(num_vectors * dimension, np.float32)
        if data.size:
            lib().innr_ref_from_flat(_p(data, _f32p), num_vectors, dimension, _p(out, _f32p))
        return cls(out, num_vectors, dimension)

    def get(self, dim: int, vec_idx: int) -> float:
        return float(self.data[dim * self.num_vectors + vec_idx])

    def dimension_slice(self, dim: int) -> np.ndarray:
        return self.data[dim * self.num_vectors:(dim + 1) * self.num_vectors]

    def extract_vector(self, i: int) -> np.ndarray:
        out = np.zeros(self.dimension, np.float32)
        if self.dimension:
            lib().innr_ref_extract_vector(_p(self.data, _f32p), self.num_vectors, self.dimension, i, _p(out, _f32p))
        return out


def _scan(fn_name, query, batch: VerticalBatch):
    q = _f32(query)
    assert q.size == batch.dimension  # assert_eq!(query.len(), batch.dimension)
    out = np.zeros(batch.num_vectors, np.float32)
    getattr(lib(), fn_name)(_p(q, _f32p), _p(batch.data, _f32p), batch.num_vectors, batch.dimension, _p(out, _f32p))
    return out


def batch_l2_squared(query, batch):  # src/batch.rs:236
    return _scan("innr_ref_batch_l2_squared", query, batch)


def batch_dot(query, batch):  # src/batch.rs:270
    return _scan("innr_ref_batch_dot", query, batch)


def batch_norms(batch):  # src/batch.rs:663
    out = np.zeros(batch.num_vectors, np.float32)
    lib().innr_ref_batch_norms(_p(batch.data, _f32p), batch.num_vectors, batch.dimension, _p(out, _f32p))
    return out


def batch_cosine(query, batch, norms):  # src/batch.rs:690
    q, norms = _f32(query), _f32(norms)
    assert norms.size == batch.num_vectors
    assert q.size == batch.dimension
    out = np.zeros(batch.num_vectors, np.float32)
    lib().innr_ref_batch_cosine(_p(q, _f32p), _p(batch.data, _f32p), batch.num_vectors, batch.dimension,
                                _p(norms, _f32p), _p(out, _f32p))
    return out


class BatchKnnResult:  # src/batch.rs:368-377
    def __init__(self, indices, scores):
        self.indices = list(int(i) for i in indices)
        self.scores = np.asarray(scores, dtype=np.float32)

    def __repr__(self):
        return f"BatchKnnResult(indices={self.indices}, scores={self.scores.tolist()})"


def _knn(fn_name, query, batch, k):
    q = _f32(query)
    assert q.size == batch.dimension
    kk = max(1, min(k, max(batch.num_vectors, 1)))
    idx = np.zeros(kk, np.uint64)
    sc = np.zeros(kk, np.float32)
    m = getattr(lib(), fn_name)(_p(q, _f32p), _p(batch.data, _f32p), batch.num_vectors, batch.dimension, k,
                                _p(idx, _u64p), _p(sc, _f32p))
    return BatchKnnResult(idx[:m], sc[:m])


def batch_knn(query, batch, k):  # src/batch.rs:385
    return _knn("innr_ref_batch_knn", query, batch, k)


def batch_knn_dot(query, batch, k):  # src/batch.rs:742
    return _knn("innr_ref_batch_knn_dot", query, batch, k)


def batch_knn_cosine(query, batch, k):  # src/batch.rs:777
    return _knn("innr_ref_batch_knn_cosine", query, batch, k)


def batch_knn_filtered(query, batch, k, predicate):  # src/batch.rs:820
    q = _f32(query)
    assert q.size == batch.dimension
    mask = np.array([1 if predicate(i) else 0 for i in range(batch.num_vectors)], dtype=np.uint8)
    kk = max(1, min(k, max(batch.num_vectors, 1)))
    idx = np.zeros(kk, np.uint64)
    sc = np.zeros(kk, np.float32)
    m = lib().innr_ref_batch_knn_filtered(_p(q, _f32p), _p(batch.data, _f32p), batch.num_vectors, batch.dimension, k,
                                          _p(mask, _u8p), _p(idx, _u64p), _p(sc, _f32p))
    return BatchKnnResult(idx[:m], sc[:m])


def batch_l2_squared_pruning(query, batch, threshold):  # src/batch.rs:320
    q = _f32(query)
    assert q.size == batch.dimension
    idx = np.zeros(max(batch.num_vectors, 1), np.uint64)
    ds = np.zeros(max(batch.num_vectors, 1), np.float32)
    m = lib().innr_ref_batch_l2_squared_pruning(_p(q, _f32p), _p(batch.data, _f32p), batch.num_vectors,
                                                batch.dimension, threshold, _p(idx, _u64p), _p(ds, _f32p))
    return [(int(idx[j]), float(ds[j])) for j in range(m)]


def batch_knn_adaptive(query, batch, k, warmup_dims):  # src/batch.rs:441
    q = _f32(query)
    assert q.size == batch.dimension
    assert warmup_dims > 0, "warmup_dims must be > 0"
    kk = max(1, min(k, max(batch.num_vectors, 1)))
    idx = np.zeros(kk, np.uint64)
    sc = np.zeros(kk, np.float32)
    m = lib().innr_ref_batch_knn_adaptive(_p(q, _f32p), _p(batch.data, _f32p), batch.num_vectors, batch.dimension, k,
                                          warmup_dims, _p(idx, _u64p), _p(sc, _f32p))
    return BatchKnnResult(idx[:m], sc[:m])


def _into(out, values):  # `out.clear(); out.resize(n, ..)` then filled
    out.clear()
    out.extend(float(x) for x in values)


def batch_l2_squared_into(query, batch, out):  # src/batch.rs:250
    _into(out, batch_l2_squared(query, batch))


def batch_dot_into(query, batch, out):  # src/batch.rs:284
    _into(out, batch_dot(query, batch))


def batch_norms_into(batch, out):  # src/batch.rs:672
    _into(out, batch_norms(batch))


def batch_cosine_into(query, batch, norms, out):  # src/batch.rs:705
    _into(out, batch_cosine(query, batch, norms))


def batch_dimension_variance(batch):  # src/batch.rs:572
    out = np.zeros(batch.dimension, np.float32)
    lib().innr_ref_batch_dimension_variance(_p(batch.data, _f32p), batch.num_vectors, batch.dimension, _p(out, _f32p))
    return out


def variance_order(variances):  # src/batch.rs:599
    v = _f32(variances)
    order = np.zeros(v.size, np.uint64)
    lib().innr_ref_variance_order(_p(v, _f32p), v.size, _p(order, _u64p))
    return [int(i) for i in order]


def batch_knn_reordered(query, batch, k):  # src/batch.rs:621
    return _knn("innr_ref_batch_knn_reordered", query, batch, k)


def batch_knn_many(metric: str, queries, batch, k, n_threads=1):
    qs = _f32(queries).reshape(-1, batch.dimension)
    nq = qs.shape[0]
    idx = np.zeros((nq, max(k, 1)), np.uint64)
    sc = np.zeros((nq, max(k, 1)), np.float32)
    m = lib().innr_ref_batch_knn_many({"dot": 0, "cosine": 1, "l2": 2}[metric], _p(qs, _f32p), nq,
                                      _p(batch.data, _f32p), batch.num_vectors, batch.dimension, k,
                                      _p(idx, _u64p), _p(sc, _f32p), n_threads)
    return idx[:, :m], sc[:, :m]


# --------------------------------------------------------------------------- TopK
class TopK:  # src/topk.rs:47-187
    def __init__(self, k: int):
        assert k > 0, "innr::TopK: k must be >= 1"
        self._h = lib().innr_ref_topk_new(k)
        self.k = k

    def __del__(self):
        if getattr(self, "_h", None):
            lib().innr_ref_topk_free(self._h)
            self._h = None

    def insert(self, id_: int, distance: float):
        lib().innr_ref_topk_insert(self._h, id_, C.c_float(distance))

    def threshold(self) -> float:
        return float(lib().innr_ref_topk_threshold(self._h))

    def __len__(self):
        return int(lib().innr_ref_topk_len(self._h))

    def is_empty(self):
        return len(self) == 0

    def into_sorted(self):
        ids = np.zeros(self.k, np.uint32)
        ds = np.zeros(self.k, np.float32)
        m = lib().innr_ref_topk_into_sorted(self._h, _p(ids, _u32p), _p(ds, _f32p))
        return [(int(ids[j]), float(ds[j])) for j in range(m)]


def topk_from_distances(distances, k):
    """N inserts in id order (what batch_knn does, src/batch.rs:401-404) then into_sorted()."""
    if k == 0 or len(distances) == 0:
        return []
    t = TopK(min(k, len(distances)))
    for i, d in enumerate(np.asarray(distances, dtype=np.float32)):
        t.insert(i, float(d))
    return t.into_sorted()


# --------------------------------------------------------------------------- MaxSim
def _tokens(t):
    if len(t) == 0:
        return np.zeros((0, 0), np.float32)
    dim = len(t[0])
    for x in t:
        if len(x) != dim:
            raise AssertionError("dimension mismatch")
    return _f32(np.array(t, dtype=np.float32).reshape(len(t), dim))


def maxsim(query_tokens, doc_tokens) -> float:  # src/maxsim.rs:96
    q, d = _tokens(query_tokens), _tokens(doc_tokens)
    if q.shape[0] == 0 or d.shape[0] == 0:
        return 0.0
    assert q.shape[1] == d.shape[1], "dimension mismatch (doc)"
    return float(lib().innr_ref_maxsim(_p(q, _f32p), q.shape[0], _p(d, _f32p), d.shape[0], q.shape[1]))


def maxsim_cosine(query_tokens, doc_tokens) -> float:  # src/maxsim.rs:168
    q, d = _tokens(query_tokens), _tokens(doc_tokens)
    if q.shape[0] == 0 or d.shape[0] == 0:
        return 0.0
    assert q.shape[1] == d.shape[1], "dimension mismatch (doc)"
    return float(lib().innr_ref_maxsim_cosine(_p(q, _f32p), q.shape[0], _p(d, _f32p), d.shape[0], q.shape[1]))


def maxsim_corpus(query_tokens, tokens, doc_offsets, cosine_flag=False, n_threads=1):
    """Caller composition examples/maxsim_colbert.rs:171-174: one score per doc."""
    q = _f32(query_tokens)
    t = _f32(tokens)
    off = np.ascontiguousarray(doc_offsets, dtype=np.uint64)
    n_docs = off.size - 1
    dim = q.shape[1] if q.ndim == 2 and q.shape[0] else (t.shape[1] if t.ndim == 2 else 0)
    out = np.zeros(max(n_docs, 0), np.float32)
    if n_docs > 0:
        lib().innr_ref_maxsim_corpus(_p(q, _f32p), q.shape[0], _p(t, _f32p), _p(off, _u64p), n_docs, dim,
                                     1 if cosine_flag else 0, _p(out, _f32p), n_threads)
    return out


# --------------------------------------------------------------------------- binary
class PackedBinary:  # src/binary.rs:37-117
    def __init__(self, data, dimension: int):
        data = np.ascontiguousarray(data, dtype=np.uint64).copy()
        expect = (dimension + 63) // 64
        assert data.size == expect, (
            f"PackedBinary: data length {data.size} doesn't match dimension {dimension} (expected {expect} words)")
        if data.size:
            lib().innr_ref_packed_binary_mask(_p(data, _u64p), dimension)
        self.data = data
        self.dimension = dimension

    @classmethod
    def zeros(cls, dimension):
        return cls(np.zeros((dimension + 63) // 64, np.uint64), dimension)

    def set(self, idx, val):
        if idx >= self.dimension:
            return
        w, b = idx // 64, idx % 64
        if val:
            self.data[w] |= np.uint64(1 << b)
        else:
            self.data[w] &= np.uint64(~(1 << b) & 0xFFFFFFFFFFFFFFFF)

    def get(self, idx):
        if idx >= self.dimension:
            return False
        return bool((int(self.data[idx // 64]) >> (idx % 64)) & 1)

    def memory_bytes(self):
        return self.data.size * 8


def encode_binary(values, threshold: float) -> PackedBinary:  # src/binary.rs:133
    v = _f32(values)
    out = np.zeros((v.size + 63) // 64, np.uint64)
    if v.size:
        lib().innr_ref_encode_binary(_p(v, _f32p), v.size, C.c_float(threshold), _p(out, _u64p))
    return PackedBinary(out, v.size)


def binary_hamming(a: PackedBinary, b: PackedBinary) -> int:  # src/binary.rs:154
    assert a.dimension == b.dimension, (
        f"innr::binary_hamming: dimension mismatch ({a.dimension} vs {b.dimension})")
    return int(lib().innr_ref_binary_hamming(_p(a.data, _u64p), _p(b.data, _u64p), a.data.size))


def binary_dot(a: PackedBinary, b: PackedBinary) -> int:  # src/binary.rs:178
    assert a.dimension == b.dimension
    return int(lib().innr_ref_binary_dot(_p(a.data, _u64p), _p(b.data, _u64p), a.data.size))


def binary_jaccard(a: PackedBinary, b: PackedBinary) -> float:  # src/binary.rs:198
    assert a.dimension == b.dimension
    return float(lib().innr_ref_binary_jaccard(_p(a.data, _u64p), _p(b.data, _u64p), a.data.size))


class PackedTernary:  # src/ternary.rs:50-157
    def __init__(self, data, dimension: int):
        data = np.ascontiguousarray(data, dtype=np.uint64).copy()
        expect = (dimension + 31) // 32
        assert data.size == expect, (
            f"PackedTernary: data length {data.size} doesn't match dimension {dimension} (expected {expect} words)")
        if data.size:
            lib().innr_ref_packed_ternary_mask(_p(data, _u64p), dimension)
        self.data = data
        self.dimension = dimension

    @classmethod
    def zeros(cls, dimension):
        return cls(np.zeros((dimension + 31) // 32, np.uint64), dimension)

    def set(self, idx, val):
        if idx >= self.dimension:
            return
        w, b = idx // 32, (idx % 32) * 2
        cur = int(self.data[w]) & ~(0b11 << b) & 0xFFFFFFFFFFFFFFFF
        bits = 0b01 if val == 1 else (0b10 if val == -1 else 0)
        self.data[w] = np.uint64(cur | (bits << b))

    def get(self, idx):
        if idx >= self.dimension:
            return 0
        bits = (int(self.data[idx // 32]) >> ((idx % 32) * 2)) & 0b11
        return 1 if bits == 0b01 else (-1 if bits == 0b10 else 0)

    def nnz(self):
        return sum(1 for i in range(self.dimension) if self.get(i) != 0)

    def memory_bytes(self):
        return self.data.size * 8


def encode_ternary(values, threshold: float) -> PackedTernary:  # src/ternary.rs:163
    v = _f32(values)
    out = np.zeros((v.size + 31) // 32, np.uint64)
    if v.size:
        lib().innr_ref_encode_ternary(_p(v, _f32p), v.size, C.c_float(threshold), _p(out, _u64p))
    return PackedTernary(out, v.size)


def ternary_dot(a: PackedTernary, b: PackedTernary) -> int:  # src/ternary.rs:191
    assert a.dimension == b.dimension, f"innr::ternary_dot: dimension mismatch ({a.dimension} vs {b.dimension})"
    return int(lib().innr_ref_ternary_dot(_p(a.data, _u64p), _p(b.data, _u64p), a.data.size))


def ternary_hamming(a: PackedTernary, b: PackedTernary) -> int:  # src/ternary.rs:301
    assert a.dimension == b.dimension
    return int(lib().innr_ref_ternary_hamming(_p(a.data, _u64p), _p(b.data, _u64p), a.data.size))


def ternary_asymmetric_dot(query, t: PackedTernary) -> float:  # src/ternary.rs:286 (ternary::asymmetric_dot)
    q = _f32(query)
    assert q.size == t.dimension
    return float(lib().innr_ref_ternary_asymmetric_dot(_p(q, _f32p), _p(t.data, _u64p), t.dimension))


def ternary_sparsity(v: PackedTernary) -> float:  # src/ternary.rs:327
    if v.dimension == 0:
        return 0.0
    return float(np.float32(1.0) - np.float32(v.nnz()) / np.float32(v.dimension))


def hamming_topk(query_words, codes, k):
    """examples/binary_demo.rs:174-180: all distances, stable sort_by_key, take(k)."""
    q = np.ascontiguousarray(query_words, dtype=np.uint64)
    c = np.ascontiguousarray(codes, dtype=np.uint64).reshape(-1, q.size)
    n = c.shape[0]
    kk = max(1, min(k, max(n, 1)))
    idx = np.zeros(kk, np.uint64)
    ds = np.zeros(kk, np.uint32)
    m = lib().innr_ref_hamming_topk(_p(q, _u64p), _p(c, _u64p), n, q.size, k, _p(idx, _u64p), _p(ds, _u32p))
    return idx[:m], ds[:m]


def hamming_topk_many(queries, codes, k, n_threads=1):
    qs = np.ascontiguousarray(queries, dtype=np.uint64)
    words = qs.shape[1]
    c = np.ascontiguousarray(codes, dtype=np.uint64).reshape(-1, words)
    nq = qs.shape[0]
    idx = np.zeros((nq, max(k, 1)), np.uint64)
    ds = np.zeros((nq, max(k, 1)), np.uint32)
    m = lib().innr_ref_hamming_topk_many(_p(qs, _u64p), nq, _p(c, _u64p), c.shape[0], words, k, _p(idx, _u64p),
                                         _p(ds, _u32p), n_threads)
    return idx[:, :m], ds[:, :m]


# --------------------------------------------------------------------------- scalar u8
class QuantizationParams:  # src/scalar.rs:44-163
    def __init__(self, alpha: float, offset: float):
        self.alpha = float(np.float32(alpha))
        self.offset = float(np.float32(offset))

    @classmethod
    def from_range(cls, mn, mx):
        a, o = C.c_float(), C.c_float()
        lib().innr_ref_qparams_from_range(C.c_float(mn), C.c_float(mx), C.byref(a), C.byref(o))
        return cls(a.value, o.value)

    @classmethod
    def fit(cls, values):
        v = _f32(values)
        a, o = C.c_float(), C.c_float()
        lib().innr_ref_qparams_fit(_p(v, _f32p), v.size, C.byref(a), C.byref(o))
        return cls(a.value, o.value)

    @classmethod
    def fit_quantile(cls, values, quantile):  # :104-137
        v = _f32(values)
        a, o = C.c_float(), C.c_float()
        rc = lib().innr_ref_qparams_fit_quantile(_p(v, _f32p), v.size, C.c_float(quantile), C.byref(a), C.byref(o))
        assert rc == 0, "quantile must be in (0.0, 1.0]"
        return cls(a.value, o.value)


class QuantizedU8:  # src/scalar.rs:171-208
    def __init__(self, data, dimension):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.size == dimension, f"QuantizedU8: data length {data.size} doesn't match dimension {dimension}"
        self.data = data
        self.dimension = dimension

    def memory_bytes(self):
        return self.data.size


def quantize_u8(values, params: QuantizationParams) -> QuantizedU8:  # src/scalar.rs:212
    v = _f32(values)
    out = np.zeros(v.size, np.uint8)
    if v.size:
        lib().innr_ref_quantize_u8(_p(v, _f32p), v.size, C.c_float(params.alpha), C.c_float(params.offset),
                                   _p(out, _u8p))
    return QuantizedU8(out, v.size)


def query_sum(query) -> float:  # src/scalar.rs:236
    q = _f32(query)
    return float(lib().innr_ref_query_sum(_p(q, _f32p), q.size))


def mixed_dot_u8_f32(a, b) -> float:  # src/scalar.rs:314
    a = _f32(a)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    assert a.size == b.size, f"mixed_dot_u8_f32: slice length mismatch ({a.size} vs {b.size})"
    return float(lib().innr_ref_mixed_dot_u8_f32(_p(a, _f32p), _p(b, _u8p), a.size))


def dot_u8_f32_variant(name, a, b) -> float:
    a = _f32(a)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    return float(getattr(lib(), "innr_ref_" + name)(_p(a, _f32p), _p(b, _u8p), a.size))


def asymmetric_dot_u8(query, quantized: QuantizedU8, params: QuantizationParams) -> float:  # src/scalar.rs:261
    q = _f32(query)
    assert q.size == quantized.dimension, (
        f"asymmetric_dot_u8: dimension mismatch ({q.size} vs {quantized.dimension})")
    return float(lib().innr_ref_asymmetric_dot_u8(_p(q, _f32p), _p(quantized.data, _u8p), q.size,
                                                  C.c_float(params.alpha), C.c_float(params.offset)))


class QueryContext:  # src/scalar.rs:228-232
    def __init__(self, query_sum: float):
        self.query_sum = float(np.float32(query_sum))


def query_context(query) -> QueryContext:  # src/scalar.rs:236
    return QueryContext(query_sum(query))


def asymmetric_dot_u8_precomputed(query, quantized: QuantizedU8, params: QuantizationParams, ctx: QueryContext) -> float:
    q = _f32(query)  # src/scalar.rs:286
    assert q.size == quantized.dimension, (
        f"asymmetric_dot_u8_precomputed: dimension mismatch ({q.size} vs {quantized.dimension})")
    return float(lib().innr_ref_asymmetric_dot_u8_precomputed(_p(q, _f32p), _p(quantized.data, _u8p), q.size,
                                                              C.c_float(params.alpha), C.c_float(params.offset),
                                                              C.c_float(ctx.query_sum)))


def batch_knn_u8(query, corpus, params: QuantizationParams, k):  # src/scalar.rs:370
    """corpus: list[QuantizedU8] or an (n, d) uint8 array."""
    q = _f32(query)
    if isinstance(corpus, np.ndarray):
        mat = np.ascontiguousarray(corpus, dtype=np.uint8)
    else:
        mat = np.stack([c.data for c in corpus]) if len(corpus) else np.zeros((0, q.size), np.uint8)
    n = mat.shape[0]
    kk = max(1, min(k, max(n, 1)))
    idx = np.zeros(kk, np.uint64)
    sc = np.zeros(kk, np.float32)
    m = lib().innr_ref_batch_knn_u8(_p(q, _f32p), _p(mat, _u8p), n, q.size, C.c_float(params.alpha),
                                    C.c_float(params.offset), k, _p(idx, _u64p), _p(sc, _f32p))
    return [(int(idx[j]), float(sc[j])) for j in range(m)]


def batch_knn_u8_many(queries, mat, params, k, n_threads=1):
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    qs = _f32(queries).reshape(-1, mat.shape[1])
    nq = qs.shape[0]
    idx = np.zeros((nq, max(k, 1)), np.uint64)
    sc = np.zeros((nq, max(k, 1)), np.float32)
    m = lib().innr_ref_batch_knn_u8_many(_p(qs, _f32p), nq, _p(mat, _u8p), mat.shape[0], mat.shape[1],
                                         C.c_float(params.alpha), C.c_float(params.offset), k, _p(idx, _u64p),
                                         _p(sc, _f32p), n_threads)
    return idx[:, :m], sc[:, :m]


# --------------------------------------------------------------------------- generators
def generate_embedding(dim: int, seed: int) -> np.ndarray:  # examples/batch_demo.rs:233-242
    out = np.zeros(dim, np.float32)
    lib().innr_ref_generate_embedding(dim, seed, _p(out, _f32p))
    return out


def generate_normalized(dim: int, seed: int) -> np.ndarray:  # examples/maxsim_colbert.rs:212-228
    out = np.zeros(dim, np.float32)
    lib().innr_ref_generate_normalized(dim, seed, _p(out, _f32p))
    return out


def splitmix64(x: int) -> int:
    return int(lib().innr_ref_splitmix64(C.c_uint64(x & 0xFFFFFFFFFFFFFFFF)))


def ghash_f32(salt: int, first_idx: int, count: int) -> np.ndarray:
    out = np.zeros(count, np.float32)
    if count:
        lib().innr_ref_ghash_f32(C.c_uint64(salt), C.c_uint64(first_idx), count, _p(out, _f32p))
    return out


def ghash_u64(salt: int, first_idx: int, count: int) -> np.ndarray:
    out = np.zeros(count, np.uint64)
    if count:
        lib().innr_ref_ghash_u64(C.c_uint64(salt), C.c_uint64(first_idx), count, _p(out, _u64p))
    return out


# multi-threaded generators for bench.py's CPU legs (full BASELINE-size corpora built on the host in their final layout)
def ghash_f32_mt(salt: int, first_idx: int, count: int, n_threads: int) -> np.ndarray:
    out = np.empty(count, np.float32)
    if count:
        lib().innr_ref_ghash_f32_mt(C.c_uint64(salt), C.c_uint64(first_idx), count, _p(out, _f32p), n_threads)
    return out


def ghash_u64_mt(salt: int, first_idx: int, count: int, n_threads: int) -> np.ndarray:
    out = np.empty(count, np.uint64)
    if count:
        lib().innr_ref_ghash_u64_mt(C.c_uint64(salt), C.c_uint64(first_idx), count, _p(out, _u64p), n_threads)
    return out


def ghash_vertical_batch(salt: int, first_row: int, n: int, d: int, n_threads: int) -> "VerticalBatch":
    """VerticalBatch.from_flat(ghash rows [first_row, first_row+n) x d) without the row-major intermediate."""
    data = np.empty(n * d, np.float32)
    if n * d:
        lib().innr_ref_ghash_pdx_mt(C.c_uint64(salt), C.c_uint64(first_row), n, d, _p(data, _f32p), n_threads)
    return VerticalBatch(data, n, d)


def ghash_u8_rows(salt: int, first_row: int, n: int, d: int, params: "QuantizationParams", n_threads: int) -> np.ndarray:
    """quantize_u8 of ghash rows: (n, d) uint8, row-major."""
    out = np.empty((n, d), np.uint8)
    if n * d:
        lib().innr_ref_ghash_u8_mt(C.c_uint64(salt), C.c_uint64(first_row), n, d, params.alpha, params.offset,
                                   _p(out, _u8p), n_threads)
    return out
