/*
 * innr_ref.h -- CPU oracle for the innr batch similarity-search hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library. The
 * product path (innr_b200/, libinnr_cuda.so) never links or calls it.
 *
 * What it is: a C++ restatement of innr 0.6.3 (/root/reference, Rust) for the
 * functions in SURVEY.md section 8(a). The reference cannot be compiled in this
 * image (no cargo/rustc, no network), so this is a "port" oracle, pinned by the
 * reference's own hand-computable unit tests, examples and the bit-exact
 * integer-valued mixed-dot KAT (tests/test_oracle_*.py port them one by one).
 * Every function cites the reference file:line it follows.
 *
 * Arithmetic rules (so results equal the Rust build bit for bit):
 *  - compiled with -ffp-contract=off: Rust never fuses a*b+c;
 *  - explicit SIMD kernels (dot_avx512, cosine_avx512, dot_avx2, cosine_avx2,
 *    dot_u8_f32_avx2) are restated twice: with the same intrinsics in the same
 *    order, and as a scalar "virtual lane chain" emulation built on fmaf().
 *    The two are bit-identical (checked in tests); the emulation is what runs
 *    on a host without AVX-512 so the oracle always answers as the reference
 *    would on an AVX-512 host (north_star: "innr's AVX-512 build").
 *
 * Not verifiable offline (recalled from Rust std source, stated in DESIGN.md):
 *  - slice::binary_search_by probe sequence (>= 1.82 branchless form), used by
 *    TopK::find_insert_pos (src/topk.rs:173-186); it only matters for exact
 *    ties on the L2 path;
 *  - _mm512_reduce_add_ps lowering in rustc/LLVM (halving tree 8,4,2,1).
 */
#ifndef INNR_REF_H
#define INNR_REF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- host capability / forcing the emulated SIMD path (tests) ---------- */
int innr_ref_host_has_avx512(void);
int innr_ref_host_has_avx2_fma(void);
/* mode: 0 = auto (intrinsics when the host has them, else emulation),
 *       1 = force scalar emulation of the AVX-512/AVX2 kernels */
void innr_ref_set_simd_mode(int mode);

/* ---- crate constants: src/lib.rs:170-184, src/dense.rs:26 -------------- */
#define INNR_REF_MIN_DIM_SIMD 16
#define INNR_REF_MIN_DIM_AVX512 64
#define INNR_REF_NORM_EPSILON 1e-9f

/* ---- backend introspection: src/backend.rs:18-67 ----------------------- */
/* returns "avx512" | "avx2+fma" | "portable" for THIS host (Display strings
 * src/backend.rs:31-41). */
const char* innr_ref_dense_backend(size_t len);

/* ---- dense pairwise f32: src/dense.rs:56-125, 243-346 ------------------ */
float innr_ref_dot(const float* a, const float* b, size_t n);          /* dispatching dot() as on an AVX-512 host */
float innr_ref_cosine(const float* a, const float* b, size_t n);       /* dispatching cosine() */
float innr_ref_dot_portable(const float* a, const float* b, size_t n); /* src/dense.rs:103-125 */
float innr_ref_cosine_portable(const float* a, const float* b, size_t n); /* src/dense.rs:287-346 */
/* explicit kernels, both forms (for oracle self-checks) */
float innr_ref_dot_avx512_intrin(const float* a, const float* b, size_t n);  /* src/arch/x86_64.rs:31-106; needs AVX-512F host */
float innr_ref_dot_avx512_emul(const float* a, const float* b, size_t n);
float innr_ref_dot_avx2_intrin(const float* a, const float* b, size_t n);    /* src/arch/x86_64.rs:183-265 */
float innr_ref_dot_avx2_emul(const float* a, const float* b, size_t n);
float innr_ref_cosine_avx512_intrin(const float* a, const float* b, size_t n); /* src/arch/x86_64.rs:681-786 */
float innr_ref_cosine_avx512_emul(const float* a, const float* b, size_t n);
float innr_ref_cosine_avx2_intrin(const float* a, const float* b, size_t n);   /* src/arch/x86_64.rs:799-915 */
float innr_ref_cosine_avx2_emul(const float* a, const float* b, size_t n);

/* ---- VerticalBatch: src/batch.rs:88-220 -------------------------------- */
/* rows: row-major n x d  ->  pdx: dimension-major data[dd*n + i] (from_flat :167) */
void innr_ref_from_flat(const float* rows, size_t n, size_t d, float* pdx);
void innr_ref_extract_vector(const float* pdx, size_t n, size_t d, size_t i, float* out); /* :217 */

/* ---- batch scans: src/batch.rs:236-297, 663-728 ------------------------ */
void innr_ref_batch_l2_squared(const float* q, const float* pdx, size_t n, size_t d, float* out);
void innr_ref_batch_dot(const float* q, const float* pdx, size_t n, size_t d, float* out);
void innr_ref_batch_norms(const float* pdx, size_t n, size_t d, float* out);
void innr_ref_batch_cosine(const float* q, const float* pdx, size_t n, size_t d,
                           const float* norms, float* out);

/* ---- batch kNN: src/batch.rs:385-411, 742-800 -------------------------- */
/* each returns the number of results written (min(k, n); 0 if n==0||k==0) */
size_t innr_ref_batch_knn(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                          uint64_t* out_idx, float* out_score);          /* L2 via TopK */
size_t innr_ref_batch_knn_dot(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                              uint64_t* out_idx, float* out_score);
size_t innr_ref_batch_knn_cosine(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                                 uint64_t* out_idx, float* out_score);
/* src/batch.rs:820-882; mask[i] != 0 <=> predicate(i) */
size_t innr_ref_batch_knn_filtered(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                                   const uint8_t* mask, uint64_t* out_idx, float* out_score);
/* src/batch.rs:320-365; returns count of survivors (idx ascending) */
size_t innr_ref_batch_l2_squared_pruning(const float* q, const float* pdx, size_t n, size_t d,
                                         float threshold, uint64_t* out_idx, float* out_dist);

/* src/batch.rs:441-564; returns (size_t)-1 when warmup == 0 (the reference's assertion) */
size_t innr_ref_batch_knn_adaptive(const float* q, const float* pdx, size_t n, size_t d, size_t k, size_t warmup,
                                   uint64_t* out_idx, float* out_score);
/* src/batch.rs:572-592 (out[d]), :599-603 (order[d]), :621-659 */
void innr_ref_batch_dimension_variance(const float* pdx, size_t n, size_t d, float* out);
void innr_ref_variance_order(const float* variances, size_t d, uint64_t* order);
size_t innr_ref_batch_knn_reordered(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                                    uint64_t* out_idx, float* out_score);

/* ---- TopK: src/topk.rs:47-187 ------------------------------------------ */
typedef struct innr_ref_topk innr_ref_topk;
innr_ref_topk* innr_ref_topk_new(size_t k);           /* k == 0 -> NULL (reference panics, :65) */
void innr_ref_topk_free(innr_ref_topk* t);
void innr_ref_topk_insert(innr_ref_topk* t, uint32_t id, float distance);
float innr_ref_topk_threshold(const innr_ref_topk* t);
size_t innr_ref_topk_len(const innr_ref_topk* t);
/* writes len() pairs ascending by distance; consumes nothing (can be called once logically) */
size_t innr_ref_topk_into_sorted(const innr_ref_topk* t, uint32_t* out_id, float* out_dist);

/* ---- MaxSim: src/maxsim.rs:96-194 -------------------------------------- */
/* tokens are contiguous row-major: q[nq][dim], d[nd][dim] (the Rust API takes &[&[f32]];
 * equal dims are asserted there, here they are implied by the layout) */
float innr_ref_maxsim(const float* q, size_t nq, const float* d, size_t nd, size_t dim);
float innr_ref_maxsim_cosine(const float* q, size_t nq, const float* d, size_t nd, size_t dim);
/* corpus composition (examples/maxsim_colbert.rs:171-174): score every doc; doc j owns token rows
 * [doc_offsets[j], doc_offsets[j+1]) of tokens[.][dim]; n_threads >= 1 splits docs over std::threads */
void innr_ref_maxsim_corpus(const float* q, size_t nq, const float* tokens, const uint64_t* doc_offsets,
                            size_t n_docs, size_t dim, int cosine, float* out_scores, int n_threads);

/* ---- binary: src/binary.rs:37-165 -------------------------------------- */
/* masks padding bits of the last word in place (PackedBinary::new :50-68); words = ceil(dim/64) */
void innr_ref_packed_binary_mask(uint64_t* words, size_t dim_bits);
void innr_ref_encode_binary(const float* v, size_t n, float threshold, uint64_t* out_words); /* :133-141 */
uint32_t innr_ref_binary_hamming(const uint64_t* a, const uint64_t* b, size_t words);         /* :154-165 */
uint32_t innr_ref_binary_dot(const uint64_t* a, const uint64_t* b, size_t words);             /* :178-185 */
float innr_ref_binary_jaccard(const uint64_t* a, const uint64_t* b, size_t words);            /* :198-213 */
/* ---- src/ternary.rs: 2 bits per value (01 = +1, 10 = -1), 32 values per u64 ---- */
void innr_ref_packed_ternary_mask(uint64_t* words, size_t dimension);                            /* :72-79 */
void innr_ref_encode_ternary(const float* v, size_t n, float threshold, uint64_t* out_words);   /* :163-173 */
int32_t innr_ref_ternary_dot(const uint64_t* a, const uint64_t* b, size_t words);                /* :191-281 */
uint32_t innr_ref_ternary_hamming(const uint64_t* a, const uint64_t* b, size_t words);           /* :301-324 */
float innr_ref_ternary_asymmetric_dot(const float* q, const uint64_t* t, size_t dimension);      /* :286-296 */
/* caller composition examples/binary_demo.rs:174-180: all distances, stable sort_by_key, take k */
size_t innr_ref_hamming_topk(const uint64_t* q, const uint64_t* codes, size_t n, size_t words, size_t k,
                             uint64_t* out_idx, uint32_t* out_dist);

/* ---- scalar u8: src/scalar.rs:44-393 ----------------------------------- */
void innr_ref_qparams_from_range(float mn, float mx, float* alpha, float* offset); /* :54-60 */
void innr_ref_qparams_fit(const float* v, size_t n, float* alpha, float* offset);  /* :68-88 */
int innr_ref_qparams_fit_quantile(const float* v, size_t n, float quantile, float* alpha, float* offset); /* :104-137; 1 = assert */
float innr_ref_asymmetric_dot_u8_precomputed(const float* q, const uint8_t* codes, size_t n, float alpha, float offset,
                                             float query_sum);                                    /* :286-300 */
void innr_ref_quantize_u8(const float* v, size_t n, float alpha, float offset, uint8_t* out); /* :212-225 */
float innr_ref_query_sum(const float* q, size_t n);                                 /* :236-240 */
float innr_ref_mixed_dot_u8_f32(const float* a, const uint8_t* b, size_t n);        /* :314-358 dispatch */
float innr_ref_mixed_dot_u8_f32_portable(const float* a, const uint8_t* b, size_t n);
float innr_ref_dot_u8_f32_avx2_intrin(const float* a, const uint8_t* b, size_t n);  /* src/arch/x86_64.rs:928-1020 */
float innr_ref_dot_u8_f32_avx2_emul(const float* a, const uint8_t* b, size_t n);
float innr_ref_asymmetric_dot_u8(const float* q, const uint8_t* codes, size_t n, float alpha, float offset); /* :261-300 */
/* corpus: contiguous row-major n x d bytes (the Rust API takes &[QuantizedU8]) */
size_t innr_ref_batch_knn_u8(const float* q, const uint8_t* corpus, size_t n, size_t d,
                             float alpha, float offset, size_t k,
                             uint64_t* out_idx, float* out_score);                  /* :370-393 */

/* ---- deterministic generators ------------------------------------------ */
void innr_ref_generate_embedding(size_t dim, uint64_t seed, float* out);   /* examples/batch_demo.rs:233-242 */
void innr_ref_generate_normalized(size_t dim, uint64_t seed, float* out);  /* examples/maxsim_colbert.rs:212-228 */
/* G-hash: u = splitmix64(salt + idx); f32 = float(u >> 40) * 2^-23 - 1 */
uint64_t innr_ref_splitmix64(uint64_t x);
void innr_ref_ghash_f32(uint64_t salt, uint64_t first_idx, size_t count, float* out);
void innr_ref_ghash_u64(uint64_t salt, uint64_t first_idx, size_t count, uint64_t* out);
/* multi-threaded forms for bench.py's CPU legs (same values), and generators that write the final layouts directly */
void innr_ref_ghash_f32_mt(uint64_t salt, uint64_t first_idx, size_t count, float* out, int n_threads);
void innr_ref_ghash_u64_mt(uint64_t salt, uint64_t first_idx, size_t count, uint64_t* out, int n_threads);
/* rows [first_row, first_row+n) x d as a VerticalBatch (pdx[dd*n + i]); == from_flat of the row-major generator output */
void innr_ref_ghash_pdx_mt(uint64_t salt, uint64_t first_row, size_t n, size_t d, float* pdx, int n_threads);
/* quantize_u8 of G-hash rows: row-major n x d codes */
void innr_ref_ghash_u8_mt(uint64_t salt, uint64_t first_row, size_t n, size_t d, float alpha, float offset, uint8_t* out,
                          int n_threads);

/* ---- multi-threaded drivers for the CPU baseline (north_star: queries spread over all cores,
 *      one query per thread, shared read-only corpus) ---------------------- */
/* metric: 0 = dot, 1 = cosine, 2 = l2. queries: nq x d row-major. outputs nq x k (row-major, padded
 * rows when k > n are left untouched); returns per-query result count. */
size_t innr_ref_batch_knn_many(int metric, const float* queries, size_t nq, const float* pdx,
                               size_t n, size_t d, size_t k, uint64_t* out_idx, float* out_score,
                               int n_threads);
size_t innr_ref_hamming_topk_many(const uint64_t* queries, size_t nq, const uint64_t* codes, size_t n,
                                  size_t words, size_t k, uint64_t* out_idx, uint32_t* out_dist,
                                  int n_threads);
size_t innr_ref_batch_knn_u8_many(const float* queries, size_t nq, const uint8_t* corpus, size_t n,
                                  size_t d, float alpha, float offset, size_t k, uint64_t* out_idx,
                                  float* out_score, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* INNR_REF_H */
