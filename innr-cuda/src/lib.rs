//! innr-cuda: thin sys + safe crate over libinnr_cuda (include/innr_cuda.h).
//! UNVERIFIED SOURCE: written against the C header, never compiled here (no cargo/rustc in this environment,
//! SURVEY F2). The same symbols are bound and exercised through Python ctypes (innr_b200/_lib.py, tests/).

#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};
#[repr(C)] pub struct innr_cuda_corpus { _p: [u8; 0] }

extern "C" {
    pub fn innr_cuda_device_count(out: *mut c_int) -> c_int;
    pub fn innr_cuda_init(device: c_int) -> c_int;
    pub fn innr_cuda_shutdown() -> c_int;
    pub fn innr_cuda_last_error() -> *const c_char;
    pub fn innr_cuda_backend_name() -> *const c_char;
    pub fn innr_cuda_dense_backend(len: usize, out_is_cuda: *mut c_int) -> c_int;
    pub fn innr_cuda_set_option(name: *const c_char, value: f64) -> c_int;
    pub fn innr_cuda_knn_tc_last_stats(filter_ms: *mut f32, total_ms: *mut f32, filter_flops: *mut f64, candidates: *mut u64, exact_scan_queries: *mut u32, passes: *mut c_int) -> c_int;
    pub fn innr_cuda_launch_count(out: *mut u64) -> c_int;
    pub fn innr_cuda_last_kernel_ms(out: *mut f32) -> c_int;
    // f32 PDX corpus ------------------------------------------------------ src/batch.rs:88-220
    pub fn innr_cuda_upload_f32_pdx(pdx: *const f32, n: usize, d: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_upload_f32_rows(rows: *const f32, n: usize, d: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_wrap_f32_pdx_dev(dev: *const f32, n: usize, d: usize, ld: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_generate_f32_pdx(generator: c_int, salt: u64, first_row: u64, n: usize, d: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_free(c: *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_corpus_info(c: *const innr_cuda_corpus, kind: *mut c_int, n: *mut usize, d: *mut usize, ld: *mut usize, index_base: *mut u64, bytes: *mut usize) -> c_int;
    pub fn innr_cuda_extract_vector(c: *const innr_cuda_corpus, i: usize, out: *mut f32) -> c_int;
    // scans ------------------------------------------------------------------ src/batch.rs:236-297, 663-728
    pub fn innr_cuda_batch_dot(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, out: *mut f32) -> c_int;
    pub fn innr_cuda_batch_l2_squared(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, out: *mut f32) -> c_int;
    pub fn innr_cuda_batch_norms(c: *const innr_cuda_corpus, out: *mut f32) -> c_int;
    pub fn innr_cuda_batch_cosine(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, norms: *const f32, norms_len: usize, out: *mut f32) -> c_int;
    // kNN --------------------------------------------------------------------- src/batch.rs:385-411, 742-800
    pub fn innr_cuda_batch_knn(c: *const innr_cuda_corpus, metric: c_int, queries: *const f32, n_queries: usize, q_len: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_batch_knn_keys_dev(c: *const innr_cuda_corpus, metric: c_int, dev_q: *const f32, n_queries: usize, k: usize, dev_keys: *mut u64, stream: *mut c_void) -> c_int;
    pub fn innr_cuda_merge_keys_dev(dev_in: *const u64, n_lists: usize, n_queries: usize, k: usize, metric: c_int, dev_keys_out: *mut u64, dev_idx: *mut u64, dev_score: *mut f32, stream: *mut c_void) -> c_int;
    pub fn innr_cuda_topk_from_distances(d: *const f32, n: usize, k: usize, out_id: *mut u32, out_dist: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_batch_knn_subset(c: *const innr_cuda_corpus, metric: c_int, q: *const f32, q_len: usize, candidates: *const u64, n_candidates: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_binary_from_f32(f32_corpus: *const innr_cuda_corpus, threshold: f32, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_u8_from_f32(f32_corpus: *const innr_cuda_corpus, alpha: f32, offset: f32, out: *mut *mut innr_cuda_corpus) -> c_int;
    // filtered kNN, pruning --------------------------------------------------- src/batch.rs:320-365, 820-882
    pub fn innr_cuda_batch_knn_filtered(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, k: usize, mask_words: *const u64, mask_len_words: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_batch_l2_squared_pruning(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, threshold: f32, out_idx: *mut u64, out_dist: *mut f32, capacity: usize, out_count: *mut usize) -> c_int;
    // dimension variance, reordered kNN ----------------------------------------- src/batch.rs:572-659
    pub fn innr_cuda_batch_dimension_variance(c: *const innr_cuda_corpus, out: *mut f32, out_len: usize) -> c_int;
    pub fn innr_cuda_batch_knn_adaptive(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, k: usize, warmup_dims: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_batch_knn_reordered(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    // binary -------------------------------------------------------------------- src/binary.rs:37-165
    pub fn innr_cuda_upload_binary(words: *const u64, n: usize, dim_bits: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_generate_binary(salt: u64, first_row: u64, n: usize, dim_bits: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_hamming_all(c: *const innr_cuda_corpus, q: *const u64, q_dim_bits: usize, out: *mut u32) -> c_int;
    pub fn innr_cuda_binary_dot_all(c: *const innr_cuda_corpus, q: *const u64, q_dim_bits: usize, out: *mut u32) -> c_int;      // src/binary.rs:178
    pub fn innr_cuda_binary_jaccard_all(c: *const innr_cuda_corpus, q: *const u64, q_dim_bits: usize, out: *mut f32) -> c_int;  // src/binary.rs:198
    pub fn innr_cuda_hamming_topk(c: *const innr_cuda_corpus, q: *const u64, n_queries: usize, q_dim_bits: usize, k: usize, out_idx: *mut u64, out_dist: *mut u32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_hamming_topk_keys_dev(c: *const innr_cuda_corpus, dev_q: *const u64, n_queries: usize, k: usize, dev_keys: *mut u64, stream: *mut c_void) -> c_int;
    pub fn innr_cuda_encode_binary(v: *const f32, n: usize, threshold: f32, out_words: *mut u64) -> c_int;
    // scalar u8 ------------------------------------------------------------------- src/scalar.rs:44-393
    pub fn innr_cuda_upload_u8(rows: *const u8, n: usize, d: usize, alpha: f32, offset: f32, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_generate_u8(salt: u64, first_row: u64, n: usize, d: usize, alpha: f32, offset: f32, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_quantize_u8(v: *const f32, n: usize, alpha: f32, offset: f32, out: *mut u8) -> c_int;
    pub fn innr_cuda_mixed_dot_u8_all(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, out: *mut f32) -> c_int;
    pub fn innr_cuda_asymmetric_dot_u8_all(c: *const innr_cuda_corpus, q: *const f32, q_len: usize, out: *mut f32) -> c_int;
    pub fn innr_cuda_batch_knn_u8(c: *const innr_cuda_corpus, queries: *const f32, n_queries: usize, q_len: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_batch_knn_u8_keys_dev(c: *const innr_cuda_corpus, dev_q: *const f32, n_queries: usize, k: usize, dev_keys: *mut u64, stream: *mut c_void) -> c_int;
    // MaxSim ---------------------------------------------------------------------- src/maxsim.rs:96-194
    pub fn innr_cuda_upload_tokens(tokens: *const f32, doc_offsets: *const u64, n_docs: usize, dim: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_generate_tokens(salt: u64, first_doc: u64, n_docs: usize, tokens_per_doc: usize, dim: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_maxsim(c: *const innr_cuda_corpus, q_tokens: *const f32, n_q: usize, q_dim: usize, cosine_flag: c_int, out_scores: *mut f32) -> c_int;
    pub fn innr_cuda_maxsim_dev(c: *const innr_cuda_corpus, dev_q: *const f32, n_q: usize, cosine_flag: c_int, dev_scores: *mut f32, stream: *mut c_void) -> c_int;
    // sharded entries, prefix views, MaxSim batches, ternary codes, binary top-k ------- include/innr_cuda.h
    pub fn innr_cuda_batch_knn_sharded(shards: *const *const innr_cuda_corpus, n_shards: usize, metric: c_int, queries: *const f32, n_queries: usize, q_len: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_batch_knn_u8_sharded(shards: *const *const innr_cuda_corpus, n_shards: usize, queries: *const f32, n_queries: usize, q_len: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_binary_topk(c: *const innr_cuda_corpus, op: c_int, query_words: *const u64, query_dim_bits: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_encode_ternary(v: *const f32, n: usize, threshold: f32, out_words: *mut u64) -> c_int;
    pub fn innr_cuda_hamming_topk_sharded(shards: *const *const innr_cuda_corpus, n_shards: usize, q: *const u64, n_queries: usize, q_dim_bits: usize, k: usize, out_idx: *mut u64, out_dist: *mut u32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_maxsim_batch(c: *const innr_cuda_corpus, q_tokens: *const f32, n_queries: usize, n_q: usize, q_dim: usize, cosine_flag: c_int, out_scores: *mut f32) -> c_int;
    pub fn innr_cuda_maxsim_batch_dev(c: *const innr_cuda_corpus, dev_q: *const f32, n_queries: usize, n_q: usize, cosine_flag: c_int, dev_scores: *mut f32, stream: *mut c_void) -> c_int;
    pub fn innr_cuda_prefix_view(c: *const innr_cuda_corpus, prefix_dim: usize, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_ternary_from_f32(f32_corpus: *const innr_cuda_corpus, threshold: f32, out: *mut *mut innr_cuda_corpus) -> c_int;
    pub fn innr_cuda_ternary_scores_all(c: *const innr_cuda_corpus, op: c_int, query: *const c_void, query_dim: usize, out_f32: *mut f32, out_i32: *mut i32) -> c_int;
    pub fn innr_cuda_ternary_topk(c: *const innr_cuda_corpus, op: c_int, query: *const c_void, query_dim: usize, k: usize, out_idx: *mut u64, out_score: *mut f32, out_count: *mut usize) -> c_int;
    pub fn innr_cuda_upload_ternary(words: *const u64, n: usize, dimension: usize, index_base: u64, out: *mut *mut innr_cuda_corpus) -> c_int;
}

// ---- safe wrappers: same panics, same results ----
pub struct DeviceBatch { h: *mut innr_cuda_corpus, n: usize, d: usize }
unsafe impl Send for DeviceBatch {} unsafe impl Sync for DeviceBatch {}   // immutable after upload
impl Drop for DeviceBatch { fn drop(&mut self) { unsafe { innr_cuda_free(self.h); } } }

impl DeviceBatch {
    /// Upload once; `batch.data()` (src/batch.rs:212) is exactly the upload format.
    pub fn from_batch(batch: &innr::batch::VerticalBatch) -> Self {
        let mut h = core::ptr::null_mut();
        check(unsafe { innr_cuda_upload_f32_pdx(batch.data().as_ptr(), batch.num_vectors(), batch.dimension(), 0, &mut h) });
        Self { h, n: batch.num_vectors(), d: batch.dimension() }
    }
}

fn knn(q: &[f32], b: &DeviceBatch, k: usize, metric: c_int) -> innr::batch::BatchKnnResult {
    assert_eq!(q.len(), b.d);                                          // src/batch.rs:386 / :743 / :778
    let kk = k.min(b.n);
    let (mut idx, mut sc, mut cnt) = (vec![0u64; kk.max(1)], vec![0f32; kk.max(1)], 0usize);
    check(unsafe { innr_cuda_batch_knn(b.h, metric, q.as_ptr(), 1, q.len(), k, idx.as_mut_ptr(), sc.as_mut_ptr(), &mut cnt) });
    idx.truncate(cnt); sc.truncate(cnt);
    innr::batch::BatchKnnResult { indices: idx.into_iter().map(|i| i as usize).collect(), scores: sc }
}
pub fn batch_knn(q: &[f32], b: &DeviceBatch, k: usize) -> innr::batch::BatchKnnResult { knn(q, b, k, 2) }
pub fn batch_knn_dot(q: &[f32], b: &DeviceBatch, k: usize) -> innr::batch::BatchKnnResult { knn(q, b, k, 0) }
pub fn batch_knn_cosine(q: &[f32], b: &DeviceBatch, k: usize) -> innr::batch::BatchKnnResult { knn(q, b, k, 1) }

pub fn batch_l2_squared_into(q: &[f32], b: &DeviceBatch, out: &mut Vec<f32>) {
    assert_eq!(q.len(), b.d);                                          // src/batch.rs:251
    out.clear(); out.resize(b.n, 0.0);
    check(unsafe { innr_cuda_batch_l2_squared(b.h, q.as_ptr(), q.len(), out.as_mut_ptr()) });
}
/// The reference takes `predicate: impl Fn(usize) -> bool` and materialises `mask: Vec<bool>` itself (src/batch.rs:839);
/// the shim evaluates the closure into a bitmask on the host and the device never reads the rows of rejected vectors.
pub fn batch_knn_filtered<F: Fn(usize) -> bool>(q: &[f32], b: &DeviceBatch, k: usize, predicate: F) -> innr::batch::BatchKnnResult {
    assert_eq!(q.len(), b.d);                                          // src/batch.rs:829
    let mut mask = vec![0u64; (b.n + 63) / 64];
    for i in 0..b.n { if predicate(i) { mask[i / 64] |= 1u64 << (i % 64); } }
    let kk = k.min(b.n).max(1);
    let (mut idx, mut sc, mut cnt) = (vec![0u64; kk], vec![0f32; kk], 0usize);
    check(unsafe { innr_cuda_batch_knn_filtered(b.h, q.as_ptr(), q.len(), k, mask.as_ptr(), mask.len(), idx.as_mut_ptr(), sc.as_mut_ptr(), &mut cnt) });
    idx.truncate(cnt); sc.truncate(cnt);
    innr::batch::BatchKnnResult { indices: idx.into_iter().map(|i| i as usize).collect(), scores: sc }
}
pub fn batch_l2_squared_pruning(q: &[f32], b: &DeviceBatch, threshold: f32) -> Vec<(usize, f32)> {
    assert_eq!(q.len(), b.d);                                          // src/batch.rs:325
    let (mut idx, mut ds, mut cnt) = (vec![0u64; b.n.max(1)], vec![0f32; b.n.max(1)], 0usize);
    check(unsafe { innr_cuda_batch_l2_squared_pruning(b.h, q.as_ptr(), q.len(), threshold, idx.as_mut_ptr(), ds.as_mut_ptr(), b.n, &mut cnt) });
    idx.into_iter().zip(ds).take(cnt).map(|(i, d)| (i as usize, d)).collect()
}
pub fn batch_knn_adaptive(q: &[f32], b: &DeviceBatch, k: usize, warmup_dims: usize) -> innr::batch::BatchKnnResult {
    assert_eq!(q.len(), b.d);                                          // src/batch.rs:447
    assert!(warmup_dims > 0, "warmup_dims must be > 0");               // src/batch.rs:448
    let kk = k.min(b.n).max(1);
    let (mut idx, mut sc, mut cnt) = (vec![0u64; kk], vec![0f32; kk], 0usize);
    check(unsafe { innr_cuda_batch_knn_adaptive(b.h, q.as_ptr(), q.len(), k, warmup_dims, idx.as_mut_ptr(), sc.as_mut_ptr(), &mut cnt) });
    idx.truncate(cnt); sc.truncate(cnt);
    innr::batch::BatchKnnResult { indices: idx.into_iter().map(|i| i as usize).collect(), scores: sc }
}
pub fn batch_dimension_variance(b: &DeviceBatch) -> Vec<f32> {          // src/batch.rs:572
    let mut out = vec![0f32; b.d];
    check(unsafe { innr_cuda_batch_dimension_variance(b.h, out.as_mut_ptr(), out.len()) });
    out
}
pub fn batch_knn_reordered(q: &[f32], b: &DeviceBatch, k: usize) -> innr::batch::BatchKnnResult {
    assert_eq!(q.len(), b.d);                                          // src/batch.rs:622
    let kk = k.min(b.n).max(1);
    let (mut idx, mut sc, mut cnt) = (vec![0u64; kk], vec![0f32; kk], 0usize);
    check(unsafe { innr_cuda_batch_knn_reordered(b.h, q.as_ptr(), q.len(), k, idx.as_mut_ptr(), sc.as_mut_ptr(), &mut cnt) });
    idx.truncate(cnt); sc.truncate(cnt);
    innr::batch::BatchKnnResult { indices: idx.into_iter().map(|i| i as usize).collect(), scores: sc }
}
/// &[QuantizedU8] / &[PackedBinary] / &[&[f32]] token lists are scattered heap objects in Rust: the shim packs
/// them contiguously once, at upload (`innr_cuda_upload_u8` / `_upload_binary` / `_upload_tokens`).
pub fn batch_knn_u8(q: &[f32], c: &DeviceU8, k: usize) -> Vec<(usize, f32)> { /* innr_cuda_batch_knn_u8, zip */ }

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(innr_cuda_last_error()) }.to_string_lossy().into_owned();
        panic!("innr-cuda: {msg}");       // INNR_EINVAL mirrors the reference's assert_eq! panics
    }
}
