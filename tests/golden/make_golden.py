#!/usr/bin/env python
"""Regenerates tests/golden/configs.npz: the oracle's answers for the five BASELINE.json configurations at reduced N,
on the bench's own stateless synthetic inputs (G-ref lattice for C1, G-hash for C2-C5, fixed salts), so
that the GPU parity tests can also check the CUDA path against COMMITTED vectors (no oracle involved at test time) and
the CPU suite can detect any drift of the oracle itself.

The reference is a Rust crate and cannot be built or imported in this image (no cargo/rustc, no network): these vectors
come from the C++ restatement under oracle/, which is pinned to the reference by the ports of the reference's own tests
(tests/test_ref_*.py) and by tests/golden/reference_kats.json (known answers transcribed from the reference's tests).

    python tests/golden/make_golden.py        # rewrites tests/golden/configs.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import innr_oracle as o  # noqa: E402

SALT_CORPUS, SALT_QUERY, SALT_CODES = 0x5EED0000, 0x5EED0001, 0x5EED0002
SHAPES = {  # reduced N, full D / k / query counts of BASELINE.json's configs
    "c1": dict(n=10_000, d=128, nq=100, k=10),          # the reference's own fixture size (examples/batch_demo.rs:167-170)
    "c2": dict(n=100_000, d=768, nq=4, k=10),
    "c3": dict(n_docs=3_000, nt=180, dim=128, nq=32),
    "c4": dict(n=500_000, dim=1024, nq=2, k=100),
    "c5": dict(n=200_000, d=384, nq=2, k=10),
}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def build():
    o.build()
    out = {}
    # C1: batch_knn_dot over the G-ref lattice, 100 queries
    s = SHAPES["c1"]
    rows = np.stack([o.generate_embedding(s["d"], i) for i in range(s["n"])])
    b = o.VerticalBatch.from_flat(rows.reshape(-1), s["n"], s["d"])
    qs = np.stack([o.generate_embedding(s["d"], 50_000 + j) for j in range(s["nq"])])
    idx, sc = o.batch_knn_many("dot", qs, b, s["k"], n_threads=8)
    out["c1_idx"], out["c1_score_bits"] = idx.astype(np.uint32), bits(sc)
    # C2: cosine / dot / L2 kNN over G-hash rows, D = 768
    s = SHAPES["c2"]
    rows = o.ghash_f32(SALT_CORPUS, 0, s["n"] * s["d"])
    b = o.VerticalBatch.from_flat(rows, s["n"], s["d"])
    qs = o.ghash_f32(SALT_QUERY, 0, s["nq"] * s["d"]).reshape(s["nq"], s["d"])
    for metric in ("cosine", "dot", "l2"):
        idx, sc = o.batch_knn_many(metric, qs, b, s["k"], n_threads=8)
        out[f"c2_{metric}_idx"], out[f"c2_{metric}_score_bits"] = idx.astype(np.uint32), bits(sc)
    out["c2_variance_bits"] = bits(o.batch_dimension_variance(b))
    r = o.batch_knn_reordered(qs[0], b, s["k"])
    out["c2_reordered_idx"], out["c2_reordered_score_bits"] = np.array(r.indices, np.uint32), bits(r.scores)
    r = o.batch_knn_adaptive(qs[0], b, s["k"], 32)
    out["c2_adaptive_idx"], out["c2_adaptive_score_bits"] = np.array(r.indices, np.uint32), bits(r.scores)
    # C3: maxsim / maxsim_cosine, 32 x 128 query tokens, 180-token documents
    s = SHAPES["c3"]
    toks = o.ghash_f32(SALT_CORPUS, 0, s["n_docs"] * s["nt"] * s["dim"]).reshape(-1, s["dim"])
    q = o.ghash_f32(SALT_QUERY, 0, s["nq"] * s["dim"]).reshape(s["nq"], s["dim"])
    off = np.arange(0, s["n_docs"] * s["nt"] + 1, s["nt"], dtype=np.uint64)
    out["c3_maxsim"] = o.maxsim_corpus(q, toks, off, cosine_flag=False, n_threads=8).astype(np.float32)
    out["c3_maxsim_cosine"] = o.maxsim_corpus(q, toks, off, cosine_flag=True, n_threads=8).astype(np.float32)
    # C4: Hamming top-100 over 1024-bit codes
    s = SHAPES["c4"]
    codes = o.ghash_u64(SALT_CODES, 0, s["n"] * 16).reshape(s["n"], 16)
    qc = o.ghash_u64(SALT_QUERY, 0, s["nq"] * 16).reshape(s["nq"], 16)
    idx, dist = o.hamming_topk_many(qc, codes, s["k"], n_threads=8)
    out["c4_idx"], out["c4_dist"] = idx.astype(np.uint32), dist.astype(np.uint32)
    # C5: batch_knn_u8 over rows quantised with from_range(-1, 1)
    s = SHAPES["c5"]
    p = o.QuantizationParams.from_range(-1.0, 1.0)
    vals = o.ghash_f32(SALT_CORPUS, 0, s["n"] * s["d"])
    mat = o.quantize_u8(vals, p).data.reshape(s["n"], s["d"])
    qs = o.ghash_f32(SALT_QUERY, 0, s["nq"] * s["d"]).reshape(s["nq"], s["d"])
    idx, sc = o.batch_knn_u8_many(qs, mat, p, s["k"], n_threads=8)
    out["c5_idx"], out["c5_score_bits"] = idx.astype(np.uint32), bits(sc)
    return out


if __name__ == "__main__":
    data = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, {k: v.shape for k, v in data.items()})
