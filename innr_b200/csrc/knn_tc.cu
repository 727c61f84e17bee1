// knn_tc.cu -- multi-query batch_knn_dot / batch_knn_cosine with the tensor cores as an exact-result FILTER.
//
// For large query batches (BASELINE C2b: 1024 queries over 10M x 768) scoring is a dense contraction. The reference
// (src/batch.rs:742-800) scores every (query, vector) pair in f32 and fully sorts; results must stay bit-exact with
// it, which a TF32 contraction cannot deliver directly. So the tensor cores only *prune*:
//
//   1. sample pass   exact top-k of every query over a prefix of the corpus (the bit-exact scan kernel of
//                    scan_f32.cu) -> L_q = k-th best exact score = a lower bound of the true k-th best score;
//   2. filter pass   S = X * [Qhi ; Qlo]^T on tcgen05 (3-term TF32 split, f32 accumulate in TMEM), reading the PDX
//                    corpus directly as the MN-major A operand (TMA, SWIZZLE_128B_ATOM_32B); every pair with
//                    S >= L_q - margin is appended to a per-query candidate list (a few hundred per query);
//   3. rescore       candidates are re-scored with the reference's exact sequential f32 arithmetic and the top-k is
//                    selected on the same 64-bit keys as the scan kernel -> indices and scores bit-identical to
//                    batch_knn_dot / batch_knn_cosine. A query whose list overflows falls back to the exact scan.
//
// The margin (2e-5 of ||q||*max||v||, 2e-5 absolute for cosine) is ~10x the error bound of the split
// (2^-21 relative to sum|q_i v_i| plus f32 accumulation); tests compare against the exact path on i.i.d. data and on
// the reference's near-tie lattice.
//
// Filter kernel: persistent, one CTA per SM, 320 threads: TMA producer warp, MMA issuer warp, 4 converter warps
// (Xlo -> TMEM as the A operand of the third product), 4 epilogue warps (TMEM -> registers, threshold, append).
// Work unit = 128 vectors x 128 queries, K loop over 32-dimension blocks through a 4-stage shared-memory ring
// (X 16 KB + Qhi 16 KB + Qlo 16 KB per stage); 12 MMAs (M128 N128 K8) per block.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace innr {

namespace {

using namespace tc;

constexpr int KT_THREADS = 320;
constexpr int VT = 128;    // vectors per work unit (UMMA M)
constexpr int QT = 128;    // queries per work unit (UMMA N)
constexpr int KB = 32;     // dimensions per K block
constexpr int KSTAGES = 4;
constexpr int X_BYTES = KB * VT * 4;   // 16 KB: 4 boxes of [32 dims][32 vectors]
constexpr int Q_BYTES = QT * KB * 4;   // 16 KB: [128 queries][32 dims], K-major SW128
constexpr int KSTAGE_BYTES = X_BYTES + 2 * Q_BYTES;
constexpr int ACC_COL0 = 0;      // 2 accumulators x 128 columns
constexpr int XLO_COL0 = 256;    // 2 Xlo buffers x 32 columns
constexpr float NORM_EPS = 1e-9f;

struct KtShared {
  uint64_t full[KSTAGES], empty[KSTAGES], lo_ready[2], lo_free[2], acc_full[2], acc_empty[2];
  float thr[2][QT];
  uint32_t tmem_base;
};

struct KtArgs {
  unsigned n, d, n_vtiles, n_qgroups, kblocks;
  unsigned index_base;
  int cosine;
  const float* inv_norms;   // n floats: 1/||v|| (0 when ||v|| <= eps), cosine only
  const float* thr;         // n_qgroups*QT thresholds (+inf for padded queries)
  unsigned* cand_count;     // nq_pad counters
  unsigned* cand;           // nq_pad x cap local vector indices
  unsigned cap;
};

__global__ void __launch_bounds__(KT_THREADS, 1) knn_tc_filter_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                       const __grid_constant__ CUtensorMap tm_qhi,
                                                                       const __grid_constant__ CUtensorMap tm_qlo,
                                                                       const KtArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  KtShared* st = reinterpret_cast<KtShared*>(smem + KSTAGES * KSTAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // contiguous range of vector tiles for this CTA; every query group is processed for a tile before moving on, so
  // the tile's rows are re-read from L2, not from HBM
  const unsigned vt_lo = (unsigned)((unsigned long long)a.n_vtiles * blockIdx.x / gridDim.x);
  const unsigned vt_hi = (unsigned)((unsigned long long)a.n_vtiles * (blockIdx.x + 1) / gridDim.x);
  const unsigned n_units = (vt_hi - vt_lo) * a.n_qgroups;
  const unsigned n_iters = n_units * a.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < KSTAGES; ++s) {
      mbar_init(&st->full[s], 1);
      mbar_init(&st->empty[s], 129);  // MMAs done with the stage (1 commit) + 128 converter threads done reading X
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&st->lo_ready[b], 128);
      mbar_init(&st->lo_free[b], 1);
      mbar_init(&st->acc_full[b], 1);
      mbar_init(&st->acc_empty[b], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_qhi);
    tma_prefetch_desc(&tm_qlo);
  }
  if (warp == 9) tmem_alloc<512>(&st->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = st->tmem_base;

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      for (unsigned it = 0; it < n_iters; ++it) {
        const unsigned unit = it / a.kblocks, kb = it % a.kblocks;
        const unsigned vt = vt_lo + unit / a.n_qgroups, qg = unit % a.n_qgroups;
        const int s = it % KSTAGES;
        mbar_wait(&st->empty[s], ((it / KSTAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&st->full[s], KSTAGE_BYTES);
        uint8_t* sb = smem + s * KSTAGE_BYTES;
        for (int mb = 0; mb < 4; ++mb) tma_load_2d(sb + mb * 4096, &tm_x, &st->full[s], (int)(vt * VT + mb * 32), (int)(kb * KB));
        tma_load_2d(sb + X_BYTES, &tm_qhi, &st->full[s], (int)(kb * KB), (int)(qg * QT));
        tma_load_2d(sb + X_BYTES + Q_BYTES, &tm_qlo, &st->full[s], (int)(kb * KB), (int)(qg * QT));
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc_mn = make_idesc_tf32(VT, QT, true);   // A = X tile in shared memory, MN-major
      const uint32_t idesc_ts = make_idesc_tf32(VT, QT, false);  // A = Xlo in tensor memory
      auto issue_hi = [&](unsigned it) {  // Xhi.Qhi + Xhi.Qlo
        const unsigned unit = it / a.kblocks, kb = it % a.kblocks;
        const int s = it % KSTAGES;
        const uint32_t xb = smem_u32(smem + s * KSTAGE_BYTES);
        const uint32_t acc = tmem + ACC_COL0 + (unit & 1) * QT;
#pragma unroll
        for (int ks = 0; ks < KB / 8; ++ks) {
          const uint64_t ad = make_smem_desc_mnmajor_sw128_32b(xb + ks * 1024, 4096, 512);
          const uint64_t bh = make_smem_desc_kmajor_sw128(xb + X_BYTES + ks * 32);
          const uint64_t bl = make_smem_desc_kmajor_sw128(xb + X_BYTES + Q_BYTES + ks * 32);
          umma_tf32(acc, ad, bh, idesc_mn, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(acc, ad, bl, idesc_mn, 1u);
        }
      };
      auto issue_lo = [&](unsigned it) {  // Xlo.Qhi
        const unsigned unit = it / a.kblocks;
        const int s = it % KSTAGES, b = it & 1;
        const uint32_t xb = smem_u32(smem + s * KSTAGE_BYTES);
        const uint32_t acc = tmem + ACC_COL0 + (unit & 1) * QT;
#pragma unroll
        for (int ks = 0; ks < KB / 8; ++ks) {
          const uint64_t bh = make_smem_desc_kmajor_sw128(xb + X_BYTES + ks * 32);
          umma_tf32_ts(acc, tmem + XLO_COL0 + b * KB + ks * 8, bh, idesc_ts, 1u);
        }
      };
      unsigned nh = 0, nl = 0;
      while (nl < n_iters) {
        if (nl < nh) {
          const int b = nl & 1;
          if (mbar_try_wait(&st->lo_ready[b], (nl >> 1) & 1)) {
            tc_fence_after_sync();
            issue_lo(nl);
            umma_commit(&st->empty[nl % KSTAGES]);  // the stage (X and Q) is no longer read by the tensor core
            umma_commit(&st->lo_free[b]);
            if (nl % a.kblocks == a.kblocks - 1) umma_commit(&st->acc_full[(nl / a.kblocks) & 1]);
            ++nl;
          }
        }
        if (nh < n_iters && nh < nl + KSTAGES) {
          const int s = nh % KSTAGES;
          const unsigned unit = nh / a.kblocks;
          bool ok = mbar_try_wait(&st->full[s], (nh / KSTAGES) & 1);
          if (ok && nh % a.kblocks == 0) ok = mbar_try_wait(&st->acc_empty[unit & 1], ((unit >> 1) & 1) ^ 1);
          if (ok) {
            tc_fence_after_sync();
            issue_hi(nh);
            ++nh;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== converters: Xlo -> TMEM (A operand of the third product) ===========================
    const int m = threadIdx.x - 128;  // vector (TMEM lane) owned by this thread
    const int mb = m >> 5, e = m & 31;
    for (unsigned it = 0; it < n_iters; ++it) {
      const int s = it % KSTAGES, b = it & 1;
      mbar_wait(&st->full[s], (it / KSTAGES) & 1);
      mbar_wait(&st->lo_free[b], ((it >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      const uint8_t* box = smem + s * KSTAGE_BYTES + mb * 4096;
      uint32_t lo[32];
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        // SWIZZLE_128B_ATOM_32B: 32-byte chunk index XOR (row % 4)
        const float x = *reinterpret_cast<const float*>(box + k * 128 + ((((e >> 3) ^ (k & 3)) << 5) | ((e & 7) << 2)));
        lo[k] = __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
      }
      mbar_arrive(&st->empty[s]);
      tmem_st_32x32b_x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + XLO_COL0 + b * KB, lo);
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(&st->lo_ready[b]);
    }
  } else {
    // =========================== epilogue: threshold + append ===========================
    for (unsigned unit = 0; unit < n_units; ++unit) {
      const unsigned vt = vt_lo + unit / a.n_qgroups, qg = unit % a.n_qgroups;
      const int ab = unit & 1;
      const unsigned v = vt * VT + warp * 32 + lane;  // local vector index of this lane
      st->thr[ab][warp * 32 + lane] = a.thr[qg * QT + warp * 32 + lane];
      float rn = 1.0f;
      if (a.cosine) rn = v < a.n ? a.inv_norms[v] : 0.0f;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // thresholds of this unit visible to the 4 epilogue warps
      mbar_wait(&st->acc_full[ab], (unit >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ACC_COL0 + ab * QT;
#pragma unroll 1
      for (int c0 = 0; c0 < QT; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        tmem_ld_wait();
        if (c0 + 32 == QT) {
          tc_fence_before_sync();
          if (lane == 0) mbar_arrive(&st->acc_empty[ab]);
        }
        if (v < a.n) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sc = __uint_as_float(r[j]) * rn;
            if (sc >= st->thr[ab][c0 + j]) {
              const unsigned q = qg * QT + c0 + j;
              const unsigned pos = atomicAdd(&a.cand_count[q], 1u);
              if (pos < a.cap) a.cand[(size_t)q * a.cap + pos] = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// ---- Qhi / Qlo operand matrices (row-major [nq_pad][d_pad]) ---------------------------------------------------------
__global__ void knn_tc_prep_queries_kernel(const float* __restrict__ q, unsigned nq, unsigned d, unsigned nq_pad,
                                           unsigned d_pad, int cosine, float* __restrict__ qhi, float* __restrict__ qlo,
                                           float* __restrict__ qnorm) {
  const unsigned row = blockIdx.x;
  __shared__ float s_scale;
  if (threadIdx.x == 0) {
    float scale = 1.0f, qn = 0.0f;
    if (row < nq) {
      float ss = 0.0f;  // sequential f32 sum, as batch_cosine_into (src/batch.rs:714)
      for (unsigned k = 0; k < d; ++k) ss = __fadd_rn(ss, __fmul_rn(q[(size_t)row * d + k], q[(size_t)row * d + k]));
      qn = __fsqrt_rn(ss);
      if (cosine) scale = qn > NORM_EPS ? 1.0f / qn : 0.0f;
    }
    s_scale = scale;
    if (row < nq_pad) qnorm[row] = qn;
  }
  __syncthreads();
  for (unsigned k = threadIdx.x; k < d_pad; k += blockDim.x) {
    float v = (row < nq && k < d) ? q[(size_t)row * d + k] * s_scale : 0.0f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    qhi[(size_t)row * d_pad + k] = hi;
    qlo[(size_t)row * d_pad + k] = v - hi;
  }
}

// ---- 1/||v|| and max ||v|| from the exact norms -----------------------------------------------------------------------
__global__ void knn_tc_inv_norms_kernel(const float* __restrict__ norms, unsigned n, float* __restrict__ inv,
                                        unsigned* __restrict__ max_bits) {
  float mx = 0.0f;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float nv = norms[i];
    inv[i] = nv > NORM_EPS ? 1.0f / nv : 0.0f;
    if (nv == nv) mx = fmaxf(mx, nv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits, __float_as_uint(mx));  // non-negative floats order like uints
}

// ---- thresholds from the sample pass: thr_q = (k-th best exact score of the sample) - margin ----------------------------
__global__ void knn_tc_threshold_kernel(const uint64_t* __restrict__ sample_keys, unsigned nq, unsigned nq_pad, unsigned k,
                                        int cosine, const float* __restrict__ qnorm, const unsigned* __restrict__ max_bits,
                                        float* __restrict__ thr) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq_pad) return;
  float t = INFINITY;  // padded queries accept nothing
  if (q < nq) {
    const uint64_t key = sample_keys[(size_t)q * k + (k - 1)];
    if (key == KEY_SENTINEL) {
      t = -INFINITY;  // sample smaller than k: no bound
    } else {
      const float lb = __uint_as_float(order_bits_to_f32_bits(~(uint32_t)(key >> 32)));
      const float scale = cosine ? 1.0f : qnorm[q] * __uint_as_float(*max_bits);
      t = (lb == lb) ? lb - 2e-5f * scale : -INFINITY;
    }
  }
  thr[q] = t;
}

// ---- exact rescoring + selection: one CTA per query ---------------------------------------------------------------------
constexpr int RS_THREADS = 256;

template <int R>
__global__ void __launch_bounds__(RS_THREADS) knn_tc_rescore_kernel(const float* __restrict__ data, size_t ld, unsigned n,
                                                                    unsigned d, unsigned index_base,
                                                                    const float* __restrict__ queries, int cosine,
                                                                    const unsigned* __restrict__ cand_count,
                                                                    const unsigned* __restrict__ cand, unsigned cap,
                                                                    int k, uint64_t* __restrict__ out_keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);
  uint64_t* smem_keys = reinterpret_cast<uint64_t*>(sq + ((d + 3) & ~3u));
  __shared__ float s_qn;
  const unsigned q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned i = threadIdx.x; i < d; i += blockDim.x) sq[i] = queries[(size_t)q * d + i];
  if (threadIdx.x == 0) {
    float ss = 0.0f;
    for (unsigned i = 0; i < d; ++i) ss = __fadd_rn(ss, __fmul_rn(queries[(size_t)q * d + i], queries[(size_t)q * d + i]));
    s_qn = __fsqrt_rn(ss);
  }
  __syncthreads();
  const float qn = s_qn;
  unsigned cnt = cand_count[q];
  if (cnt > cap) cnt = cap;
  WarpList<R> list;
  list.init();
  uint64_t thr = KEY_SENTINEL;
  const unsigned cnt_round = (cnt + 31u) / 32u * 32u;
  for (unsigned c = threadIdx.x; c < cnt_round; c += blockDim.x) {
    const bool valid = c < cnt;
    uint64_t key = KEY_SENTINEL;
    if (valid) {
      const unsigned i = cand[(size_t)q * cap + c];
      const float* p = data + i;
      float acc = 0.0f, ss = 0.0f;
      for (unsigned dd = 0; dd < d; ++dd) {  // the reference's sequential unfused sums (src/batch.rs:290-296, 676-681)
        const float v = __ldg(p + (size_t)dd * ld);
        acc = __fadd_rn(acc, __fmul_rn(sq[dd], v));
        ss = __fadd_rn(ss, __fmul_rn(v, v));
      }
      float s = acc;
      if (cosine) {
        const float nrm = __fsqrt_rn(ss);
        s = (!(qn < NORM_EPS) && nrm > NORM_EPS) ? __fdiv_rn(acc, __fmul_rn(qn, nrm)) : 0.0f;
      }
      key = make_key_desc(s, index_base + i);
    }
    list.offer(key, valid, thr, k, lane);
  }
  block_tree_merge<R>(list, k, smem_keys);
  if (warp == 0) list.store(out_keys + (size_t)q * k, k, lane);
}

}  // namespace

bool make_pdx_tmap(CUtensorMap* m, const float* dev_pdx, size_t n, size_t d, size_t ld) {
  if (n == 0 || d == 0) return false;
  return make_tmap_f32_rows(m, dev_pdx, d, n, KB, ld, /*atom32=*/true);  // box = 32 vectors x 32 dims
}

size_t knn_tc_cap() { return 4096; }

// workspace layout helpers -----------------------------------------------------------------------------------------------
struct KnnTcPlan {
  unsigned nq_pad, d_pad, n_s;
  size_t off_qhi, off_qlo, off_qnorm, off_thr, off_cnt, off_cand, off_skeys, total;
};

static KnnTcPlan make_plan(size_t n, size_t d, size_t nq, size_t k) {
  KnnTcPlan p{};
  p.nq_pad = (unsigned)((nq + QT - 1) / QT * QT);
  p.d_pad = (unsigned)((d + KB - 1) / KB * KB);
  size_t ns = n / 128;
  if (ns < 64 * k) ns = 64 * k;
  if (ns < 8192) ns = 8192;
  if (ns > n) ns = n;
  p.n_s = (unsigned)ns;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  p.off_qhi = take((size_t)p.nq_pad * p.d_pad * 4);
  p.off_qlo = take((size_t)p.nq_pad * p.d_pad * 4);
  p.off_qnorm = take((size_t)p.nq_pad * 4);
  p.off_thr = take((size_t)p.nq_pad * 4);
  p.off_cnt = take((size_t)p.nq_pad * 4);
  p.off_cand = take((size_t)p.nq_pad * knn_tc_cap() * 4);
  p.off_skeys = take((size_t)nq * k * 8);
  p.total = o;
  return p;
}

size_t knn_tc_workspace_bytes(size_t n, size_t d, size_t nq, size_t k) { return make_plan(n, d, nq, k).total; }

bool knn_tc_supported(const PdxView& v, int mode, size_t nq, size_t k) {
  return (mode == PDX_DOT || mode == PDX_COSINE_FUSED) && nq >= 1 && k >= 1 && k <= 32 && v.d >= 1 && v.n >= 4096 &&
         v.n < 0x7FFFFF00ull && v.ld % 4 == 0;
}

cudaError_t launch_pdx_knn_tc(const PdxView& v, const CUtensorMap& tm_x, int mode, const float* dev_queries, size_t nq,
                              size_t k, uint64_t* dev_keys, const float* dev_inv_norms, const unsigned* dev_max_norm_bits,
                              void* workspace, unsigned* host_counts, Workspace& ws, cudaStream_t s,
                              uint64_t* launches, std::vector<unsigned>* overflow_queries) {
  const int cosine = mode == PDX_COSINE_FUSED;
  const KnnTcPlan p = make_plan(v.n, v.d, nq, k);
  uint8_t* w = (uint8_t*)workspace;
  float* qhi = (float*)(w + p.off_qhi);
  float* qlo = (float*)(w + p.off_qlo);
  float* qnorm = (float*)(w + p.off_qnorm);
  float* thr = (float*)(w + p.off_thr);
  unsigned* cnt = (unsigned*)(w + p.off_cnt);
  unsigned* cand = (unsigned*)(w + p.off_cand);
  uint64_t* skeys = (uint64_t*)(w + p.off_skeys);
  cudaError_t e;
  static const bool trace = getenv("INNR_KNN_TC_TRACE") != nullptr;
  cudaEvent_t ev[6];
  if (trace)
    for (auto& x : ev) cudaEventCreate(&x);
  auto mark = [&](int i) { if (trace) cudaEventRecord(ev[i], s); };
  mark(0);

  // 1. sample pass on a prefix of the corpus (exact, bit-identical scores)
  PdxView prefix = v;
  prefix.n = p.n_s;
  e = launch_pdx_knn(prefix, mode, dev_queries, nq, k, skeys, ws, s, launches);
  if (e != cudaSuccess) return e;
  mark(1);
  // 2. operands and thresholds
  knn_tc_prep_queries_kernel<<<p.nq_pad, 128, 0, s>>>(dev_queries, (unsigned)nq, (unsigned)v.d, p.nq_pad, p.d_pad, cosine,
                                                      qhi, qlo, qnorm);
  knn_tc_threshold_kernel<<<(p.nq_pad + 127) / 128, 128, 0, s>>>(skeys, (unsigned)nq, p.nq_pad, (unsigned)k, cosine, qnorm,
                                                                 dev_max_norm_bits, thr);
  e = cudaMemsetAsync(cnt, 0, (size_t)p.nq_pad * 4, s);
  if (e != cudaSuccess) return e;
  *launches += 2;
  CUtensorMap tm_qhi, tm_qlo;
  if (!make_tmap_f32_rows(&tm_qhi, qhi, p.nq_pad, p.d_pad, QT) || !make_tmap_f32_rows(&tm_qlo, qlo, p.nq_pad, p.d_pad, QT))
    return cudaErrorInvalidValue;
  mark(2);
  // 3. tensor-core filter
  static bool attr_set = false;
  const size_t smem = (size_t)KSTAGES * KSTAGE_BYTES + sizeof(KtShared);
  if (!attr_set) {
    e = cudaFuncSetAttribute(knn_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  KtArgs a{};
  a.n = (unsigned)v.n;
  a.d = (unsigned)v.d;
  a.n_vtiles = (unsigned)((v.n + VT - 1) / VT);
  a.n_qgroups = p.nq_pad / QT;
  a.kblocks = p.d_pad / KB;
  a.index_base = v.index_base;
  a.cosine = cosine;
  a.inv_norms = dev_inv_norms;
  a.thr = thr;
  a.cand_count = cnt;
  a.cand = cand;
  a.cap = (unsigned)knn_tc_cap();
  unsigned grid = (unsigned)ws.num_sms;
  if (grid > a.n_vtiles) grid = a.n_vtiles;
  knn_tc_filter_kernel<<<grid, KT_THREADS, smem, s>>>(tm_x, tm_qhi, tm_qlo, a);
  ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  mark(3);
  // 4. exact rescoring + selection
  const size_t rs_smem = ((v.d + 3) & ~(size_t)3) * 4 + (size_t)(RS_THREADS / 32) * k * 8;
  knn_tc_rescore_kernel<1><<<(unsigned)nq, RS_THREADS, rs_smem, s>>>(v.data, v.ld, (unsigned)v.n, (unsigned)v.d, v.index_base,
                                                                     dev_queries, cosine, cnt, cand, a.cap, (int)k, dev_keys);
  ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  mark(4);
  // 5. overflowed candidate lists -> the caller re-runs those queries on the exact scan
  if (host_counts && overflow_queries) {
    e = cudaMemcpyAsync(host_counts, cnt, nq * 4, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    overflow_queries->clear();
    for (size_t q = 0; q < nq; ++q)
      if (host_counts[q] > a.cap) overflow_queries->push_back((unsigned)q);
    if (trace) {
      float t[4];
      for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
      unsigned long long tot = 0, mx = 0;
      for (size_t q = 0; q < nq; ++q) { tot += host_counts[q]; if (host_counts[q] > mx) mx = host_counts[q]; }
      fprintf(stderr, "[knn_tc] sample(n_s=%u) %.2f ms | prep %.2f | filter %.2f | rescore %.2f | candidates total %llu max %llu overflow %zu\n",
              p.n_s, t[0], t[1], t[2], t[3], tot, mx, overflow_queries->size());
      for (auto& x : ev) cudaEventDestroy(x);
    }
  }
  return cudaSuccess;
}

cudaError_t launch_knn_tc_inv_norms(const float* dev_norms, size_t n, float* dev_inv, unsigned* dev_max_bits, cudaStream_t s,
                                    uint64_t* launches) {
  cudaError_t e = cudaMemsetAsync(dev_max_bits, 0, 4, s);
  if (e != cudaSuccess) return e;
  knn_tc_inv_norms_kernel<<<148 * 4, 256, 0, s>>>(dev_norms, (unsigned)n, dev_inv, dev_max_bits);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace innr
