// maxsim_tc.cu -- ColBERT MaxSim on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces the reference's per-pair AVX-512 loops (maxsim_avx512 src/arch/x86_64.rs:119-143, cosine_avx512
// :681-786, driven per document by examples/maxsim_colbert.rs:171-174) with one dense contraction per 128-token
// tile:   S[128 tokens x 64] = X[128 x 128] * [Qhi ; Qlo]^T        (kind::tf32, f32 accumulate in TMEM)
// issued twice per tile: once with X as loaded by TMA into shared memory (the tensor core reads the top 19 bits =
// Xhi) and once with Xlo = X - trunc_tf32(X) as the A operand in TENSOR MEMORY (written there by the converter
// warps with tcgen05.st, so the split costs no shared-memory write and no second operand read: the kernel is
// shared-memory-bandwidth sensitive, 4 x 64 KB per tile). Adding columns j and 32+j gives
// (Qhi+Qlo)_j . Xhi + Qhi_j . Xlo -- the classic 3-term split whose error is ~2^-21 relative to sum|q.x| (the f32 tolerance of the
// north_star, 1e-5, is 2^-16.6). The row-max over a document's tokens and the sum over query tokens are fused in
// the TMEM->register epilogue; one f32 per document leaves the SM.
//
// Shape of the machine (one persistent CTA per SM, 448 threads, ~225 KB shared memory, 192 TMEM columns):
//   warp 12     TMA producer: 4 x cp.async.bulk.tensor (128 rows x 128 B, SWIZZLE_128B) per tile -> 3-stage ring
//   warp 13     MMA issuer (one thread): hi(i) then lo(i-1), 16 tcgen05.mma (M128 N64 K8) each, commits to mbarriers
//   warps 4-11  converters: after hi(i) has been read, Xlo in place + per-token sum of squares (cosine)
//   warps 0-3   epilogue: tcgen05.ld 32 lanes x 64 columns, hi+lo add, cosine scale, max over token lanes
//               (halving butterfly; segmented scan when a document boundary falls inside the 32-token chunk),
//               warp 0 stitches chunk summaries in token order and writes one score per document.
// HBM-bound by design: tensor time is ~40 % of the tile's HBM time.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace innr {

namespace {

using namespace tc;

constexpr int TC_THREADS = 320;   // 4 epilogue + 4 converter + TMA + MMA warps
constexpr int TILE_M = 128;                 // tokens per tile (UMMA M)
constexpr int DIM = 128;                    // K
constexpr int NQ = 32;                      // query tokens (padded)
constexpr int UMMA_N = 64;                  // [Qhi ; Qlo]
constexpr int STAGES = 3;
constexpr int PANEL_BYTES = TILE_M * 128;   // 16 KB: [128 rows][32 floats], 128-byte swizzle
constexpr int STAGE_BYTES = 4 * PANEL_BYTES;
constexpr int QPANEL_BYTES = UMMA_N * 128;  // 8 KB
constexpr int QBYTES = 4 * QPANEL_BYTES;
constexpr float EPS_SQ = 1e-9f * 1e-9f;
constexpr int NO_DOC = 0x7FFFFFFF;
constexpr int BB_COL0 = 192;                // TMEM columns [192, 200): ring of per-token sum of squares (8 tiles)
constexpr int LO_COL0 = 256;                // TMEM columns [256, 512): two Xlo buffers of 128 columns

struct __align__(8) Summary {  // per 32-token chunk, written by its epilogue warp, read by the stitcher (warp 0)
  float head[NQ];
  float tail[NQ];
  int first_doc, last_doc;
};

struct SharedTail {  // everything after the operand buffers
  uint64_t full[STAGES], empty[STAGES], lo_ready[2], lo_free[2], tmem_full[STAGES], tmem_empty[STAGES];
  Summary sum[4];
  uint32_t tmem_base;
  int range[4];  // doc_lo, doc_hi (+ token range as two u32 halves are kept in registers)
};

struct TcArgs {
  const uint64_t* doc_offsets;
  unsigned long long uniform_tokens, total_tokens;
  unsigned n_docs, n_q;
  const float* q;
  int cosine;
  int debug_mode;  // 0 normal; 1 = TMA streaming only; 2 = hi pass only; 4 = no epilogue math (profiling aids)
  float* out;
};

__device__ __forceinline__ unsigned long long doc_begin(const TcArgs& a, unsigned d) {
  return a.uniform_tokens ? (unsigned long long)d * a.uniform_tokens : a.doc_offsets[d];
}

// swizzled byte offset of element (row, k) inside a [rows][32 floats] panel group (4 panels, panel stride `pstride`)
__device__ __forceinline__ uint32_t sw128_offset(int row, int k, int pstride) {
  const int p = k >> 5, c = (k & 31) >> 2, e = k & 3;
  return (uint32_t)(p * pstride + row * 128 + ((c ^ (row & 7)) << 4) + (e << 2));
}

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1) maxsim_tc_kernel(const __grid_constant__ CUtensorMap tm_tokens,
                                                                   const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_tok = smem;                                  // STAGES x 64 KB
  uint8_t* s_q = smem + STAGES * STAGE_BYTES;             // 32 KB
  SharedTail* st = reinterpret_cast<SharedTail*>(s_q + QBYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- this CTA's document range: contiguous, token-balanced, starts and ends on document boundaries ----
  unsigned doc_lo, doc_hi;
  {
    const unsigned long long t_lo = a.total_tokens * blockIdx.x / gridDim.x;
    const unsigned long long t_hi = a.total_tokens * (blockIdx.x + 1) / gridDim.x;
    auto first_doc_at_or_after = [&](unsigned long long t) -> unsigned {  // smallest d with begin(d) >= t
      if (a.uniform_tokens) return (unsigned)((t + a.uniform_tokens - 1) / a.uniform_tokens);
      unsigned lo = 0, hi = a.n_docs;
      while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (a.doc_offsets[mid] >= t) hi = mid; else lo = mid + 1;
      }
      return lo;
    };
    doc_lo = blockIdx.x == 0 ? 0u : first_doc_at_or_after(t_lo);
    doc_hi = blockIdx.x == gridDim.x - 1 ? a.n_docs : first_doc_at_or_after(t_hi);
    if (doc_lo > a.n_docs) doc_lo = a.n_docs;
    if (doc_hi > a.n_docs) doc_hi = a.n_docs;
    if (doc_hi < doc_lo) doc_hi = doc_lo;
  }
  const unsigned long long tok_lo = doc_begin(a, doc_lo), tok_hi = doc_begin(a, doc_hi);
  const unsigned n_tiles = (unsigned)((tok_hi - tok_lo + TILE_M - 1) / TILE_M);

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&st->full[s], 1);
      mbar_init(&st->empty[s], 129);  // hi MMAs done reading (1 commit) + 128 converter threads done reading
      mbar_init(&st->tmem_full[s], 1);
      mbar_init(&st->tmem_empty[s], 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&st->lo_ready[b], 128);
      mbar_init(&st->lo_free[b], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_tokens);
  }
  if (warp == 9) tmem_alloc<512>(&st->tmem_base);
  // B operand: rows 0..31 = Qhi, rows 32..63 = Qlo (cosine: rows pre-scaled by 1/||q||), K-major SW128 panels
  for (int idx = threadIdx.x; idx < NQ * DIM; idx += blockDim.x) {
    const int r = idx / DIM, k = idx % DIM;
    float v = (r < (int)a.n_q) ? a.q[(size_t)r * DIM + k] : 0.0f;
    if (a.cosine && r < (int)a.n_q) {
      float aa = 0.0f;
      const float* qp = a.q + (size_t)r * DIM;
      for (int kk = 0; kk < DIM; ++kk) aa = fmaf(qp[kk], qp[kk], aa);
      v = aa > EPS_SQ ? v * rsqrtf(aa) : 0.0f;  // query with ~zero norm -> cosine 0 for every token
    }
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const float lo = v - hi;
    *reinterpret_cast<float*>(s_q + sw128_offset(r, k, QPANEL_BYTES)) = hi;
    *reinterpret_cast<float*>(s_q + sw128_offset(32 + r, k, QPANEL_BYTES)) = lo;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = st->tmem_base;

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      for (unsigned i = 0; i < n_tiles; ++i) {
        const int s = i % STAGES;
        mbar_wait(&st->empty[s], ((i / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&st->full[s], STAGE_BYTES);
        const int row0 = (int)(tok_lo + (unsigned long long)i * TILE_M);
        for (int p = 0; p < 4; ++p) tma_load_2d(s_tok + s * STAGE_BYTES + p * PANEL_BYTES, &tm_tokens, &st->full[s], p * 32, row0);
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TILE_M, UMMA_N);
      const uint32_t q_base = smem_u32(s_q);
      auto issue_hi = [&](int s) {  // A = X tile in shared memory (tensor core reads the TF32 part = Xhi)
        const uint32_t a_base = smem_u32(s_tok + s * STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < DIM / 8; ++kk) {
          const uint64_t ad = make_smem_desc_kmajor_sw128(a_base + (kk >> 2) * PANEL_BYTES + (kk & 3) * 32);
          const uint64_t bd = make_smem_desc_kmajor_sw128(q_base + (kk >> 2) * QPANEL_BYTES + (kk & 3) * 32);
          umma_tf32(tmem + s * UMMA_N, ad, bd, idesc, kk > 0 ? 1u : 0u);
        }
      };
      // A = Xlo in tensor memory (128 lanes x 128 columns); B = the Qhi rows only (N = 32): Qlo.Xlo is below
      // 2^-22 of the product and is not worth a quarter of the tensor work (the kernel runs under the power cap).
      const uint32_t idesc_lo = make_idesc_tf32(TILE_M, NQ);
      auto issue_lo = [&](int t, int b) {
#pragma unroll
        for (int kk = 0; kk < DIM / 8; ++kk) {
          const uint64_t bd = make_smem_desc_kmajor_sw128(q_base + (kk >> 2) * QPANEL_BYTES + (kk & 3) * 32);
          umma_tf32_ts(tmem + t * UMMA_N, tmem + LO_COL0 + b * DIM + kk * 8, bd, idesc_lo, 1u);
        }
      };
      if (a.debug_mode == 1) {
        for (unsigned i = 0; i < n_tiles; ++i) {
          mbar_wait(&st->full[i % STAGES], (i / STAGES) & 1);
          for (int r = 0; r < 129; ++r) mbar_arrive(&st->empty[i % STAGES]);
        }
      }
      if (a.debug_mode == 2) {  // profiling aid: hi pass only (no Xlo), results are TF32-accurate only
        for (unsigned i = 0; i < n_tiles; ++i) {
          const int s = i % STAGES;
          mbar_wait(&st->full[s], (i / STAGES) & 1);
          mbar_wait(&st->tmem_empty[s], ((i / STAGES) & 1) ^ 1);
          tc_fence_after_sync();
          issue_hi(s);
          umma_commit(&st->empty[s]);
          umma_commit(&st->tmem_full[s]);
        }
      }
      // hi(i) and lo(j) are issued in whatever order their inputs become ready (never block on one while the
      // other could run); lo(j) always follows hi(j) because both accumulate into the same TMEM columns.
      unsigned nh = 0, nl = 0;
      while (a.debug_mode != 1 && a.debug_mode != 2 && nl < n_tiles) {
        if (nl < nh) {  // lo(nl): the converters have written Xlo of tile nl to TMEM buffer nl % 2
          const int b = nl & 1;
          if (mbar_try_wait(&st->lo_ready[b], (nl >> 1) & 1)) {
            tc_fence_after_sync();
            issue_lo(nl % STAGES, b);
            umma_commit(&st->tmem_full[nl % STAGES]);  // accumulator complete for the epilogue
            umma_commit(&st->lo_free[b]);              // Xlo buffer may be overwritten
            ++nl;
          }
        }
        if (nh < n_tiles && nh < nl + STAGES) {  // hi(nh): X as loaded
          const int s = nh % STAGES;
          if (mbar_try_wait(&st->full[s], (nh / STAGES) & 1) &&
              mbar_try_wait(&st->tmem_empty[s], ((nh / STAGES) & 1) ^ 1)) {
            tc_fence_after_sync();
            issue_hi(s);
            umma_commit(&st->empty[s]);  // 1 of 129: the tensor core has finished reading the stage
            ++nh;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== converters: Xlo -> TMEM, token sum of squares ===========================
    // One token row per thread = one TMEM lane per thread (warps 4-7 own lane quadrants 0-3). A panel row (8 chunks
    // of 16 B) is loaded at once, split, and written as 32 TMEM columns with one tcgen05.st.
    const int row = threadIdx.x - 128;
    for (unsigned i = 0; a.debug_mode != 1 && i < n_tiles; ++i) {
      const int s = i % STAGES, b = i & 1;
      mbar_wait(&st->full[s], (i / STAGES) & 1);
      if (a.debug_mode == 2) { mbar_arrive(&st->empty[s]); continue; }
      mbar_wait(&st->lo_free[b], ((i >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      const uint8_t* base = s_tok + s * STAGE_BYTES + row * 128;
      const uint32_t tdst = tmem + ((uint32_t)((warp & 3) * 32) << 16) + LO_COL0 + b * DIM;
      float ss = 0.0f;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const uint8_t* pbase = base + p * PANEL_BYTES;
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(pbase + ((c ^ (row & 7)) << 4));
        uint32_t lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (a.cosine) {
            ss = fmaf(v[c].x, v[c].x, ss); ss = fmaf(v[c].y, v[c].y, ss);
            ss = fmaf(v[c].z, v[c].z, ss); ss = fmaf(v[c].w, v[c].w, ss);
          }
          lo[4 * c + 0] = __float_as_uint(v[c].x - __uint_as_float(__float_as_uint(v[c].x) & 0xFFFFE000u));
          lo[4 * c + 1] = __float_as_uint(v[c].y - __uint_as_float(__float_as_uint(v[c].y) & 0xFFFFE000u));
          lo[4 * c + 2] = __float_as_uint(v[c].z - __uint_as_float(__float_as_uint(v[c].z) & 0xFFFFE000u));
          lo[4 * c + 3] = __float_as_uint(v[c].w - __uint_as_float(__float_as_uint(v[c].w) & 0xFFFFE000u));
        }
        if (p == 3) mbar_arrive(&st->empty[s]);  // this thread has read its whole row: 1 of 129
        tmem_st_32x32b_x32(tdst + 32 * p, lo);
      }
      tmem_st_32x32b_x1(tmem + ((uint32_t)((warp & 3) * 32) << 16) + BB_COL0 + (i & 7), __float_as_uint(ss));
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(&st->lo_ready[b]);
    }
  } else {
    // =========================== epilogue (warps 0-3 = TMEM lane quadrants 0-3) ===========================
    float carry = -INFINITY;  // warp 0 / lane j: running max of query token j for the document being stitched
    int carry_doc = -1;
    unsigned cur_doc = doc_lo;  // warp-uniform cursor: document containing this warp's chunk start
    const float lane_is_query = lane < (int)a.n_q ? 1.0f : 0.0f;
    auto finalize = [&](int doc, float m) {  // warp 0: sum over query tokens of the per-token maxima
      float v = lane_is_query != 0.0f ? m : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
      if (lane == 0) a.out[doc] = v;
    };
    for (unsigned i = 0; a.debug_mode != 1 && i < n_tiles; ++i) {
      const int t = i % STAGES;
      const unsigned long long c0 = tok_lo + (unsigned long long)i * TILE_M + warp * 32;
      const unsigned long long g = c0 + lane;
      const bool valid = g < tok_hi;
      mbar_wait(&st->tmem_full[t], (i / STAGES) & 1);
      tc_fence_after_sync();
      uint32_t rh[32], rl[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + t * UMMA_N;
      tmem_ld_32x32b_x32(taddr, rh);
      tmem_ld_32x32b_x32(taddr + 32, rl);
      const uint32_t bb_bits = tmem_ld_32x32b_x1(tmem + ((uint32_t)(warp * 32) << 16) + BB_COL0 + (i & 7));
      tmem_ld_wait();
      float bbv = __uint_as_float(bb_bits);
      tc_fence_before_sync();
      if (lane == 0) mbar_arrive(&st->tmem_empty[t]);
      if (a.debug_mode == 4) { if (rh[0] == 0x12345u && rl[0] == 0x54321u) a.out[0] = bbv; continue; }
      float sc[NQ];
      const float rt = a.cosine ? (bbv > EPS_SQ ? rsqrtf(bbv) : 0.0f) : 1.0f;
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        float v = (__uint_as_float(rh[j]) + __uint_as_float(rl[j])) * rt;
        sc[j] = (valid && v == v) ? v : -INFINITY;  // NaN never replaces the max (x86_64.rs:135)
      }
      // ---- document of every lane's token ----
      Summary& sm = st->sum[warp];
      named_bar_sync(2, 128);  // warp 0 has finished stitching the previous tile: the summary slots are free
      if (c0 < tok_hi) {  // warp-uniform
        int doc = NO_DOC;
        if (a.uniform_tokens) {
          if (valid) doc = (int)((unsigned)g / (unsigned)a.uniform_tokens);  // tokens < 2^31
        } else {
          while (cur_doc + 1 < doc_hi && a.doc_offsets[cur_doc + 1] <= c0) ++cur_doc;  // uniform
          if (valid) {
            unsigned d = cur_doc;
            while (a.doc_offsets[d + 1] <= g) ++d;
            doc = (int)d;
          }
        }
        const int first_doc = __shfl_sync(FULL_MASK, doc, 0);
        const unsigned vmask = __ballot_sync(FULL_MASK, valid);
        const int last_lane = 31 - __clz(vmask);
        const int last_doc = __shfl_sync(FULL_MASK, doc, last_lane);
        if (first_doc == last_doc && vmask == FULL_MASK) {
          // fast path: the whole chunk lies inside one document. Halving butterfly: 31 shuffles for 32 columns;
          // afterwards lane l holds the max over all 32 tokens of column bitrev-free index `col` below.
          float v16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float mine = (lane & 16) ? sc[j + 16] : sc[j];
            const float send = (lane & 16) ? sc[j] : sc[j + 16];
            v16[j] = fmaxf(mine, __shfl_xor_sync(FULL_MASK, send, 16));
          }
          float v8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float mine = (lane & 8) ? v16[j + 8] : v16[j];
            const float send = (lane & 8) ? v16[j] : v16[j + 8];
            v8[j] = fmaxf(mine, __shfl_xor_sync(FULL_MASK, send, 8));
          }
          float v4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float mine = (lane & 4) ? v8[j + 4] : v8[j];
            const float send = (lane & 4) ? v8[j] : v8[j + 4];
            v4[j] = fmaxf(mine, __shfl_xor_sync(FULL_MASK, send, 4));
          }
          float v2[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float mine = (lane & 2) ? v4[j + 2] : v4[j];
            const float send = (lane & 2) ? v4[j] : v4[j + 2];
            v2[j] = fmaxf(mine, __shfl_xor_sync(FULL_MASK, send, 2));
          }
          const float mine = (lane & 1) ? v2[1] : v2[0];
          const float send = (lane & 1) ? v2[0] : v2[1];
          const float m = fmaxf(mine, __shfl_xor_sync(FULL_MASK, send, 1));
          // column owned by this lane: bit b of the column index was fixed by (lane & b)
          const int col = lane;  // (lane&16)->+16, (lane&8)->+8, ... composes to the lane index itself
          sm.head[col] = m;
          if (lane == 0) {
            sm.first_doc = first_doc;
            sm.last_doc = last_doc;
          }
        } else {
          // slow path: document boundaries (or the end of the range) inside the chunk -> segmented max-scan
          const int prev_doc = __shfl_up_sync(FULL_MASK, doc, 1);
          (void)prev_doc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int odoc = __shfl_up_sync(FULL_MASK, doc, o);
            const bool take = lane >= o && odoc == doc;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              const float ov = __shfl_up_sync(FULL_MASK, sc[j], o);
              sc[j] = take ? fmaxf(sc[j], ov) : sc[j];
            }
          }
          // lane that ends a segment: next lane has a different document (or is invalid / beyond the warp)
          const int next_doc = __shfl_down_sync(FULL_MASK, doc, 1);
          const bool seg_end = valid && (lane == 31 || next_doc != doc);
          if (seg_end) {
            if (doc == first_doc) {
#pragma unroll
              for (int j = 0; j < NQ; ++j) sm.head[j] = sc[j];
            } else if (doc == last_doc) {
#pragma unroll
              for (int j = 0; j < NQ; ++j) sm.tail[j] = sc[j];
            } else {  // a document entirely inside this chunk: finish it here, summing in query order
              float total = 0.0f;
#pragma unroll
              for (int j = 0; j < NQ; ++j)
                if (j < (int)a.n_q) total += sc[j];
              a.out[doc] = total;
            }
          }
          if (lane == 0) {
            sm.first_doc = first_doc;
            sm.last_doc = last_doc;
          }
        }
      } else if (lane == 0) {
        sm.first_doc = NO_DOC;  // chunk entirely beyond this CTA's range
        sm.last_doc = NO_DOC;
      }
      named_bar_sync(1, 128);  // the 4 epilogue warps: summaries of tile i are in shared memory
      if (warp == 0) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const Summary& s2 = st->sum[w];
          const int fd = s2.first_doc, ld = s2.last_doc;
          if (fd == NO_DOC) continue;
          const float head = s2.head[lane];
          if (carry_doc != fd) {
            if (carry_doc >= 0) finalize(carry_doc, carry);
            carry = head;
            carry_doc = fd;
          } else {
            carry = fmaxf(carry, head);
          }
          if (ld != fd) {
            finalize(fd, carry);
            carry = s2.tail[lane];
            carry_doc = ld;
          }
        }
      }
    }
    if (warp == 0 && carry_doc >= 0) finalize(carry_doc, carry);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

}  // namespace

bool make_token_tmap(CUtensorMap* m, const float* dev_tokens, size_t total_tokens, size_t dim) {
  if (dim != DIM || total_tokens == 0) return false;
  return make_tmap_f32_rows(m, dev_tokens, total_tokens, dim, TILE_M);
}

size_t maxsim_tc_smem_bytes() { return (size_t)STAGES * STAGE_BYTES + QBYTES + sizeof(SharedTail); }

bool maxsim_tc_supported(const TokView& v, size_t n_q) {
  return v.dim == DIM && n_q >= 1 && n_q <= NQ && v.total_tokens > 0 && v.tmap_valid &&
         v.total_tokens < 0x7FFFFFFFull;
}

cudaError_t launch_maxsim_tc(const TokView& v, const float* dev_q, size_t n_q, int cosine, float* dev_scores,
                             int num_sms, cudaStream_t s, uint64_t* launches) {
  static bool attr_set = false;
  const size_t smem = maxsim_tc_smem_bytes();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(maxsim_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  // empty documents never see a token: their score is 0.0 (src/maxsim.rs:97-99)
  cudaError_t e = cudaMemsetAsync(dev_scores, 0, v.n_docs * sizeof(float), s);
  if (e != cudaSuccess) return e;
  TcArgs a{};
  a.doc_offsets = v.doc_offsets;
  a.uniform_tokens = v.uniform_tokens;
  a.total_tokens = v.total_tokens;
  a.n_docs = (unsigned)v.n_docs;
  a.n_q = (unsigned)n_q;
  a.q = dev_q;
  a.cosine = cosine;
  a.out = dev_scores;
  static const int dbg = getenv("INNR_MAXSIM_DEBUG") ? atoi(getenv("INNR_MAXSIM_DEBUG")) : 0;
  a.debug_mode = dbg;
  unsigned grid = (unsigned)num_sms;
  const unsigned long long tiles = (v.total_tokens + TILE_M - 1) / TILE_M;
  if (grid > tiles) grid = (unsigned)tiles;
  if (grid > v.n_docs) grid = (unsigned)v.n_docs;
  if (grid == 0) grid = 1;
  maxsim_tc_kernel<<<grid, TC_THREADS, smem, s>>>(v.tmap, a);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace innr
