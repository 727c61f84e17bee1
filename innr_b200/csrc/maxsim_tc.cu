// maxsim_tc.cu -- ColBERT MaxSim on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces the reference's per-pair AVX-512 loops (maxsim_avx512 src/arch/x86_64.rs:119-143, cosine_avx512
// :681-786, driven per document by examples/maxsim_colbert.rs:171-174) with one dense contraction per 128-token
// tile:   S[128 tokens x 64] = X[128 x 128] * [Qhi ; Qlo]^T        (kind::tf32, f32 accumulate in TMEM)
// issued twice per tile: once with X as loaded by TMA into shared memory (the tensor core reads the top 19 bits =
// Xhi, N = 64) and once with Xlo = X - trunc_tf32(X) as the A operand in TENSOR MEMORY against Qhi only (N = 32;
// written there by the converter warps with tcgen05.st, so the split costs no shared-memory write and no second
// operand read). Adding columns j and 32+j gives (Qhi+Qlo)_j . Xhi + Qhi_j . Xlo -- the 3-term split whose error is
// ~2^-21 relative to sum|q.x| (the f32 tolerance of the north_star, 1e-5, is 2^-16.6). The max over a document's
// tokens and the sum over query tokens are fused in the TMEM->register epilogue; one f32 per document leaves the SM.
//
// Shape of the machine (one persistent CTA per SM, 320 threads, ~225 KB shared memory, all 512 TMEM columns: accumulator
// stages in [0, 256), two Xlo buffers in [256, 512)):
//   The CTA's contiguous document range is cut into FOUR document-aligned token streams. A 128-row tile is made of
//   the next 32 tokens of each stream (rows 32w..32w+31 = stream w), so TMEM lane quadrant w -- the only lanes warp w
//   of a warpgroup may read -- always holds consecutive tokens of one stream and every epilogue warp carries its
//   running per-document maxima in registers across tiles: no cross-warp stitching, no barriers between them.
//   warp 8      TMA producer: 16 x cp.async.bulk.tensor (32 rows x 128 B, SWIZZLE_128B) per tile -> 3-stage ring
//   warp 9      MMA issuer (one thread): hi(i) / lo(j) in readiness order, 16 tcgen05.mma (K = 8) each
//   warps 4-7   converters: Xlo -> TMEM (LOP3 + packed FADD2), one token row per thread
//   warps 0-3   epilogue: tcgen05.ld 32 lanes x 64 columns, hi+lo add, cosine scale by the token's cached 1/||x||,
//               max over the 32 token lanes with redux.sync.max.f32 (CREDUX, 4.4 cycles per column measured vs 10 for the
//               shuffle butterfly: dev/credux_probe.cu), running max per document, sum in query order at its end.
// Per-token 1/||x|| is computed once at upload (the reference recomputes both norms for each of the 5760 pairs).
// The kernel runs under the 1 kW power cap (HBM at 7 TB/s + tensor + converters): instructions per tile, not
// bytes, decide the sustained rate, hence the lean epilogue and the suspended (not spinning) mbarrier waits.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace innr {

namespace {

using namespace tc;

constexpr int TC_THREADS = 320;   // 4 epilogue + 4 converter + TMA + MMA warps
constexpr int TILE_M = 128;                 // tokens per tile (UMMA M)
constexpr int CHUNK = 32;                   // tokens per stream per tile (one TMEM lane quadrant)
constexpr int NQ = 32;                      // query tokens per group (one TMEM column block)
constexpr int ACC_MAX = 3;                  // TMEM accumulator stages: 3 x 64 columns (G = 1) or 2 x 128 (G = 2)
constexpr int MAX_STAGES = 8;
constexpr int PANEL_BYTES = TILE_M * 128;   // 16 KB: [128 rows][32 floats], 128-byte swizzle
constexpr int BOX_BYTES = CHUNK * 128;      // one TMA box: 32 rows x 128 B
constexpr float EPS_SQ = 1e-9f * 1e-9f;
constexpr int LO_COL0 = 256;                // TMEM columns [256, 512): two Xlo buffers of DIM (<= 128) columns

// P = dim / 32 panels of 32 floats (dim 32, 64, 96, 128); G = groups of 32 query tokens scored in ONE corpus pass
// (B operand = [Qhi ; Qlo] of 64 G rows, accumulator 64 G columns). The shared-memory ring takes what the operands leave.
// KH = K halves per tile (token dimensions 129..256): a tile's 128 tokens arrive as KH stages of 128 columns each, and
// both passes accumulate over the halves into the same TMEM columns (KH = 2 only with G = 1: the B operand of a 256-d
// query group is 64 KB, a query pair's footprint).
template <int P, int G, int KH = 1>
struct Shape {
  static constexpr int DIM = 32 * P;                    // columns per stage
  static constexpr int UMMA_N = 64 * G;                 // [Qhi ; Qlo]
  static constexpr int ACC = G == 1 ? 3 : 2;
  static constexpr int STAGE_BYTES = P * PANEL_BYTES;
  static constexpr int QPANEL_BYTES = UMMA_N * 128;     // one K panel of the B operand
  static constexpr int QBYTES = P * KH * QPANEL_BYTES;
  static constexpr int FIT = (227 * 1024 - 1024 - QBYTES) / STAGE_BYTES;
  static constexpr int STAGES = FIT > MAX_STAGES ? MAX_STAGES : FIT;
};

struct SharedTail {  // everything after the operand buffers
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], lo_ready[2], lo_free[2], tmem_full[ACC_MAX], tmem_empty[ACC_MAX];
  unsigned long long s_tok[5];  // token boundaries of the four streams
  unsigned s_doc[5];            // document boundaries of the four streams
  float q_inv[2 * NQ];          // cosine: 1/||q_r|| per query token row
  float part[4][CHUNK + 1];     // G = 2: per epilogue warp, first-group partial sums of the documents ending in the chunk
  uint32_t tmem_base;
};

struct TcArgs {
  const uint64_t* doc_offsets;
  const float* inv_norms;  // per token 1/||x|| (0 for ||x||^2 <= 1e-18), cosine only
  unsigned long long uniform_tokens, total_tokens;
  unsigned n_docs, n_q;
  unsigned dim;  // token dimension (a multiple of 4, <= 32 P): the columns up to 32 P are zero-filled by TMA / the query staging
  const float* q;
  int debug_mode;  // 0 normal; 1 = TMA streaming only; 2 = hi pass only; 4 = no epilogue math (profiling aids)
  int accumulate;  // 0: out[doc] = sum; 1: out[doc] += sum (second and later groups of 32 query tokens)
  float* out;
  // G = 2 only: the two column groups are two DIFFERENT queries of <= 32 tokens each (batch of queries sharing one corpus
  // pass): group 1 reads q_b / n_q_b and writes out_b; sums are kept apart
  int split;
  const float* q_b;
  unsigned n_q_b;
  float* out_b;
};

__device__ __forceinline__ unsigned long long doc_begin(const TcArgs& a, unsigned d) {
  return a.uniform_tokens ? (unsigned long long)d * a.uniform_tokens : a.doc_offsets[d];
}

// swizzled byte offset of element (row, k) inside a [rows][32 floats] panel group (4 panels, panel stride `pstride`)
__device__ __forceinline__ uint32_t sw128_offset(int row, int k, int pstride) {
  const int p = k >> 5, c = (k & 31) >> 2, e = k & 3;
  return (uint32_t)(p * pstride + row * 128 + ((c ^ (row & 7)) << 4) + (e << 2));
}

// waits that suspend the thread in hardware (up to the hint) instead of spinning on the issue port
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
  }
}

__device__ __forceinline__ float redux_max(float v) {  // CREDUX.MAX.F32: NaN inputs are ignored unless all are NaN
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}
// packed f32x2 (sm_100): two IEEE operations per instruction
__device__ __forceinline__ void sub2(float& x0, float& x1, float y0, float y1) {
  uint64_t x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(y0), "f"(y1));
  asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
}
__device__ __forceinline__ void addmul2(float& x0, float& x1, float y0, float y1, float s, bool scale) {
  uint64_t x, y, z;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(y0), "f"(y1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
  if (scale) {
    asm("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(s));
    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(z));
  }
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
}

template <bool COSINE, int P, int G, int KH = 1>
__global__ void __launch_bounds__(TC_THREADS, 1) maxsim_tc_kernel(const __grid_constant__ CUtensorMap tm_tokens,
                                                                   const TcArgs a) {
  using SH = Shape<P, G, KH>;
  constexpr int DIM = SH::DIM, STAGES = SH::STAGES, STAGE_BYTES = SH::STAGE_BYTES, QBYTES = SH::QBYTES,
                UMMA_N = SH::UMMA_N, ACC = SH::ACC, QPANEL_BYTES = SH::QPANEL_BYTES, NQG = NQ * G, DIMT = DIM * KH;
  static_assert(STAGES >= 2, "ring too shallow");
  static_assert(KH == 1 || (G == 1 && P == 4), "K halves: full 128-column stages, one query group");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_tok = smem;                                  // STAGES x (P x 16 KB)
  uint8_t* s_q = smem + STAGES * STAGE_BYTES;             // P x 8 KB
  SharedTail* st = reinterpret_cast<SharedTail*>(s_q + QBYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    // this CTA's document range: contiguous, token-balanced, starts and ends on document boundaries; then the same
    // cut once more into four streams
    auto first_doc_at_or_after = [&](unsigned long long t) -> unsigned {  // smallest d with begin(d) >= t
      if (a.uniform_tokens) {
        const unsigned long long d = (t + a.uniform_tokens - 1) / a.uniform_tokens;
        return d > a.n_docs ? a.n_docs : (unsigned)d;
      }
      unsigned lo = 0, hi = a.n_docs;
      while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (a.doc_offsets[mid] >= t) hi = mid; else lo = mid + 1;
      }
      return lo;
    };
    const unsigned long long t_lo = a.total_tokens * blockIdx.x / gridDim.x;
    const unsigned long long t_hi = a.total_tokens * (blockIdx.x + 1) / gridDim.x;
    unsigned doc_lo = blockIdx.x == 0 ? 0u : first_doc_at_or_after(t_lo);
    unsigned doc_hi = blockIdx.x == gridDim.x - 1 ? a.n_docs : first_doc_at_or_after(t_hi);
    if (doc_hi < doc_lo) doc_hi = doc_lo;
    const unsigned long long tok_lo = doc_begin(a, doc_lo), tok_hi = doc_begin(a, doc_hi);
    unsigned prev = doc_lo;
    for (int w = 0; w <= 4; ++w) {
      unsigned d = w == 0 ? doc_lo : (w == 4 ? doc_hi : first_doc_at_or_after(tok_lo + (tok_hi - tok_lo) * w / 4));
      if (d < prev) d = prev;
      if (d > doc_hi) d = doc_hi;
      st->s_doc[w] = d;
      st->s_tok[w] = doc_begin(a, d);
      prev = d;
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&st->full[s], 1);
      mbar_init(&st->empty[s], 5);  // hi MMAs done reading (1 commit) + 4 converter warps done reading (one elected
                                    // arrive per warp: 128 per-thread arrives on one mbarrier serialise in the LSU and
                                    // showed up as a quarter of the kernel's shared-memory wavefronts)
    }
    for (int t = 0; t < ACC; ++t) {
      mbar_init(&st->tmem_full[t], 1);
      mbar_init(&st->tmem_empty[t], 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&st->lo_ready[b], 4);  // one elected arrive per converter warp
      mbar_init(&st->lo_free[b], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_tokens);
  }
  if (warp == 9) tmem_alloc<512>(&st->tmem_base);
  // B operand: rows [0, 32G) = Qhi, rows [32G, 64G) = Qlo (cosine: rows pre-scaled by 1/||q||), K-major SW128 panels
  auto q_row = [&](int r, bool& qvalid) -> const float* {
    const bool second = G > 1 && a.split && r >= NQ;
    qvalid = (G > 1 && a.split) ? (second ? r - NQ < (int)a.n_q_b : r < (int)a.n_q) : r < (int)a.n_q;
    return second ? a.q_b + (size_t)(r - NQ) * a.dim : a.q + (size_t)r * a.dim;
  };
  if (COSINE) {  // 1/||q_r|| once per query token (one thread per row), not once per element
    if ((int)threadIdx.x < NQG) {
      bool qvalid;
      const float* qp = q_row((int)threadIdx.x, qvalid);
      float aa = 0.0f;
      if (qvalid)
        for (int kk = 0; kk < (int)a.dim; ++kk) aa = fmaf(qp[kk], qp[kk], aa);
      // a query token with ~zero norm scores cosine 0 against every token (x86_64.rs:781-785)
      st->q_inv[threadIdx.x] = (qvalid && aa > EPS_SQ) ? 1.0f / sqrtf(aa) : 0.0f;
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < NQG * DIMT; idx += blockDim.x) {
    const int r = idx / DIMT, k = idx % DIMT;
    bool qvalid;
    const float* qsrc = q_row(r, qvalid);
    float v = (qvalid && k < (int)a.dim) ? qsrc[k] : 0.0f;
    if (COSINE) v *= st->q_inv[r];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const float lo = v - hi;
    *reinterpret_cast<float*>(s_q + sw128_offset(r, k, QPANEL_BYTES)) = hi;
    *reinterpret_cast<float*>(s_q + sw128_offset(NQG + r, k, QPANEL_BYTES)) = lo;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = st->tmem_base;
  unsigned n_tiles = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const unsigned t = (unsigned)((st->s_tok[w + 1] - st->s_tok[w] + CHUNK - 1) / CHUNK);
    n_tiles = t > n_tiles ? t : n_tiles;
  }
  const unsigned n_it = n_tiles * KH;  // stages: iteration it = K half (it % KH) of tile it / KH

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      long long row[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) row[w] = (long long)st->s_tok[w];
      // an exhausted stream keeps loading its last box (valid memory, masked in the epilogue)
      const long long last_row = a.total_tokens > CHUNK ? (long long)a.total_tokens - CHUNK : 0;
      for (unsigned i = 0; i < n_it; ++i) {
        const int s = i % STAGES;
        const int half = KH == 1 ? 0 : (int)(i % KH);
        mbar_wait_sleepy(&st->empty[s], ((i / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&st->full[s], STAGE_BYTES);
        uint8_t* dst = s_tok + s * STAGE_BYTES;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int r0 = (int)(row[w] < (long long)st->s_tok[w + 1] ? row[w] : last_row);  // rows past the matrix end are zero-filled
#pragma unroll
          for (int p = 0; p < P; ++p)  // columns past `dim` (last panels of the last half) are zero-filled by TMA
            tma_load_2d(dst + p * PANEL_BYTES + w * BOX_BYTES, &tm_tokens, &st->full[s], half * DIM + p * 32, r0);
          if (half == KH - 1) row[w] += CHUNK;
        }
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the control flow (warp-uniform values -> descriptors live in uniform registers, no R2UR per
    // MMA); one elected lane issues. A profile of the previous version showed this thread, not the tensor pipe, as the
    // limiter: 686 instructions per tile at IPC 0.2.
    const uint32_t idesc = make_idesc_tf32(TILE_M, UMMA_N);
    const uint32_t idesc_lo = make_idesc_tf32(TILE_M, NQG);
    const uint64_t q_desc = make_smem_desc_kmajor_sw128(smem_u32(s_q));
    const uint64_t a_desc0 = make_smem_desc_kmajor_sw128(smem_u32(s_tok));
    auto issue_hi = [&](int s, int t, int half) {  // A = X stage in shared memory (tensor core reads the TF32 part = Xhi)
      const uint64_t ad0 = desc_advance(a_desc0, (uint32_t)s * STAGE_BYTES);
      const uint64_t qd0 = desc_advance(q_desc, (uint32_t)half * P * QPANEL_BYTES);  // this half's K panels of the B operand
      const uint32_t acc = tmem + t * UMMA_N;
      if (KH == 1 || half == 0) umma_tf32_c<false>(acc, ad0, qd0, idesc);
      else umma_tf32_c<true>(acc, ad0, qd0, idesc);
#pragma unroll
      for (int kk = 1; kk < DIM / 8; ++kk)
        umma_tf32_c<true>(acc, desc_advance(ad0, (kk >> 2) * PANEL_BYTES + (kk & 3) * 32),
                          desc_advance(qd0, (kk >> 2) * QPANEL_BYTES + (kk & 3) * 32), idesc);
    };
    // A = Xlo in tensor memory (128 lanes x 128 columns); B = the Qhi rows only (N = 32): Qlo.Xlo is below
    // 2^-22 of the product and is not worth a quarter of the tensor work (the kernel runs under the power cap).
    auto issue_lo = [&](int t, int b, int half) {
      const uint32_t acc = tmem + t * UMMA_N, src = tmem + LO_COL0 + b * DIM;
      const uint64_t qd0 = desc_advance(q_desc, (uint32_t)half * P * QPANEL_BYTES);
#pragma unroll
      for (int kk = 0; kk < DIM / 8; ++kk)
        umma_tf32_ts_c<true>(acc, src + kk * 8, desc_advance(qd0, (kk >> 2) * QPANEL_BYTES + (kk & 3) * 32), idesc_lo);
    };
    if (a.debug_mode == 1) {
      for (unsigned i = 0; i < n_it; ++i) {
        mbar_wait(&st->full[i % STAGES], (i / STAGES) & 1);
        if (lane == 0)
          for (int r = 0; r < 5; ++r) mbar_arrive(&st->empty[i % STAGES]);
        __syncwarp();
      }
    }
    if (a.debug_mode == 2) {  // profiling aid: hi pass only (no Xlo), results are TF32-accurate only
      for (unsigned i = 0; i < n_it; ++i) {
        const unsigned tile = i / KH;
        const int s = i % STAGES, t = tile % ACC, half = (int)(i % KH);
        mbar_wait(&st->full[s], (i / STAGES) & 1);
        if (half == 0) mbar_wait(&st->tmem_empty[t], ((tile / ACC) & 1) ^ 1);
        tc_fence_after_sync();
        if (elect_one_sync()) {
          issue_hi(s, t, half);
          umma_commit(&st->empty[s]);
          if (half == KH - 1) umma_commit(&st->tmem_full[t]);
        }
        __syncwarp();
      }
    }
    // hi(i) and lo(j) are issued in whatever order their inputs become ready (never block on one while the
    // other could run); lo(j) always follows hi(j) because both accumulate into the same TMEM columns.
    unsigned nh = 0, nl = 0;  // counted in stages (K halves of tiles)
    while (a.debug_mode != 1 && a.debug_mode != 2 && nl < n_it) {
      bool progressed = false;
      if (nl < nh) {  // lo(nl): the converters have written Xlo of stage nl to TMEM buffer nl % 2
        const int b = nl & 1;
        if (mbar_try_wait(&st->lo_ready[b], (nl >> 1) & 1)) {
          const unsigned tile = nl / KH;
          const int half = (int)(nl % KH);
          tc_fence_after_sync();
          if (elect_one_sync()) {
            issue_lo(tile % ACC, b, half);
            if (half == KH - 1) umma_commit(&st->tmem_full[tile % ACC]);  // accumulator complete for the epilogue
            umma_commit(&st->lo_free[b]);                                   // Xlo buffer may be overwritten
          }
          __syncwarp();
          ++nl;
          progressed = true;
        }
      }
      if (nh < n_it && nh < nl + ACC * KH) {  // hi(nh): X as loaded
        const unsigned tile = nh / KH;
        const int s = nh % STAGES, t = tile % ACC, half = (int)(nh % KH);
        if (mbar_try_wait(&st->full[s], (nh / STAGES) & 1) &&
            (half > 0 || mbar_try_wait(&st->tmem_empty[t], ((tile / ACC) & 1) ^ 1))) {
          tc_fence_after_sync();
          if (elect_one_sync()) {
            issue_hi(s, t, half);
            umma_commit(&st->empty[s]);  // 1 of 5: the tensor core has finished reading the stage
          }
          __syncwarp();
          ++nh;
          progressed = true;
        }
      }
      if (!progressed) __nanosleep(64);
    }
  } else if (warp >= 4) {
    // =========================== converters: Xlo -> TMEM ===========================
    // One token row per thread = one TMEM lane per thread (warps 4-7 own lane quadrants 0-3). A panel row (8 chunks
    // of 16 B) is loaded at once, split, and written as 32 TMEM columns with one tcgen05.st.
    const int row = threadIdx.x - 128;
    for (unsigned i = 0; a.debug_mode != 1 && i < n_it; ++i) {
      const int s = i % STAGES, b = i & 1;
      mbar_wait_sleepy(&st->full[s], (i / STAGES) & 1);
      if (a.debug_mode == 2) {
        if (lane == 0) mbar_arrive(&st->empty[s]);
        continue;
      }
      mbar_wait_sleepy(&st->lo_free[b], ((i >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      const uint8_t* base = s_tok + s * STAGE_BYTES + row * 128;
      const uint32_t tdst = tmem + ((uint32_t)((warp & 3) * 32) << 16) + LO_COL0 + b * DIM;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const uint8_t* pbase = base + p * PANEL_BYTES;
        ulonglong2 v[8];  // two packed f32 pairs per 16-byte chunk: the pairs stay in 64-bit registers end to end
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const ulonglong2*>(pbase + ((c ^ (row & 7)) << 4));
        if (p == P - 1) {  // every lane of this warp has read its whole row: 1 of 5
          __syncwarp();
          if (lane == 0) mbar_arrive(&st->empty[s]);
        }
        uint32_t lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint64_t l01 = sub2_rn(v[c].x, v[c].x & 0xFFFFE000FFFFE000ull);  // x - trunc_tf32(x), two lanes
          const uint64_t l23 = sub2_rn(v[c].y, v[c].y & 0xFFFFE000FFFFE000ull);
          lo[4 * c + 0] = (uint32_t)l01;
          lo[4 * c + 1] = (uint32_t)(l01 >> 32);
          lo[4 * c + 2] = (uint32_t)l23;
          lo[4 * c + 3] = (uint32_t)(l23 >> 32);
        }
        tmem_st_32x32b_x32(tdst + 32 * p, lo);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->lo_ready[b]);
    }
  } else {
    // =========================== epilogue (warp w = TMEM lane quadrant w = stream w) ===========================
    const unsigned long long s_lo = st->s_tok[warp], s_hi = st->s_tok[warp + 1];
    unsigned cur_doc = st->s_doc[warp];
    unsigned long long cur_end = 0;  // end token of cur_doc; 0 forces the first lookup
    bool have_doc = false;
    float carry[G][NQ];  // per-lane running max of query token 32 g + j over this lane's tokens of the current document
#pragma unroll
    for (int gq = 0; gq < G; ++gq)
#pragma unroll
      for (int j = 0; j < NQ; ++j) carry[gq][j] = -INFINITY;
    const int n_q = (int)a.n_q;
    float* part = st->part[warp];
    // the token's cached 1/||x|| is fetched one tile ahead: the load was 12 % of this kernel's stall samples when it was
    // issued right in front of its use (the accumulator is usually ready when the epilogue comes for it)
    auto load_rt = [&](unsigned long long c) -> float {
      if (!COSINE || c >= s_hi) return 1.0f;
      return c + lane < s_hi ? __ldg(a.inv_norms + c + lane) : 0.0f;
    };
    float rt_next = load_rt(s_lo);
    for (unsigned i = 0; a.debug_mode != 1 && i < n_tiles; ++i) {
      const int t = i % ACC;
      const unsigned long long c0 = s_lo + (unsigned long long)i * CHUNK;
      const unsigned long long g = c0 + lane;
      const bool active = c0 < s_hi;  // warp-uniform: this stream still has tokens in tile i
      const float rt = rt_next;
      rt_next = load_rt(c0 + CHUNK);
      mbar_wait_sleepy(&st->tmem_full[t], (i / ACC) & 1);
      tc_fence_after_sync();
      if (!active) {
        tc_fence_before_sync();
        if (lane == 0) mbar_arrive(&st->tmem_empty[t]);
        continue;
      }
      const bool rt_zero = COSINE && __any_sync(0xFFFFFFFFu, rt == 0.0f);
      const unsigned long long c1 = c0 + CHUNK < s_hi ? c0 + CHUNK : s_hi;
      // the query groups are reduced one after the other (registers: one group's scores at a time); both walk the same
      // document segments of the chunk, so the cursor is rewound for the second group
      const unsigned doc0 = cur_doc;
      const unsigned long long end0 = cur_end;
      const bool have0 = have_doc;
#pragma unroll
      for (int gq = 0; gq < G; ++gq) {
        uint32_t rh[32], rl[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + t * UMMA_N + gq * NQ;
        tmem_ld_32x32b_x32(taddr, rh);            // Xhi.Qhi + Xlo.Qhi
        tmem_ld_32x32b_x32(taddr + NQ * G, rl);   // Xhi.Qlo
        tmem_ld_wait();
        if (gq == G - 1) {  // the accumulator is in registers: the MMA warp may overwrite it
          tc_fence_before_sync();
          if (lane == 0) mbar_arrive(&st->tmem_empty[t]);
        }
        if (a.debug_mode == 4) { if (rh[0] == 0x12345u && rl[0] == 0x54321u) a.out[0] = rt; continue; }
        float sc[NQ];
#pragma unroll
        for (int j = 0; j < NQ; j += 2) {
          float x0 = __uint_as_float(rh[j]), x1 = __uint_as_float(rh[j + 1]);
          addmul2(x0, x1, __uint_as_float(rl[j]), __uint_as_float(rl[j + 1]), rt, COSINE);
          sc[j] = x0;
          sc[j + 1] = x1;
        }
        if (rt_zero) {          // rare: a token below the norm guard scores exactly 0.0 against every query token
          if (rt == 0.0f) {     // (x86_64.rs:781-785), also when it holds NaN
#pragma unroll
            for (int j = 0; j < NQ; ++j) sc[j] = 0.0f;
          }
        }
        if (G > 1 && gq > 0) {
          cur_doc = doc0;
          cur_end = end0;
          have_doc = have0;
          __syncwarp();  // part[] of the first group is visible
        }
        // ---- segments of this chunk: [pos, seg_end) lies inside one document ----
        unsigned long long pos = c0;
        int seg = 0;
        while (pos < c1) {  // warp-uniform
          if (!have_doc || cur_end <= pos) {  // next non-empty document (empty ones keep the memset 0.0)
            if (have_doc) ++cur_doc;
            have_doc = true;
            cur_end = doc_begin(a, cur_doc + 1);
            while (cur_end <= pos) {
              ++cur_doc;
              cur_end = doc_begin(a, cur_doc + 1);
            }
          }
          const unsigned long long seg_end = cur_end < c1 ? cur_end : c1;
          // the running maxima stay per lane (one FMNMX per score); lanes are only combined when the document ends
          if (pos == c0 && seg_end == c0 + CHUNK) {  // the whole chunk lies inside one document
#pragma unroll
            for (int j = 0; j < NQ; ++j) carry[gq][j] = fmaxf(carry[gq][j], sc[j]);
          } else {
            const bool in = g >= pos && g < seg_end;
#pragma unroll
            for (int j = 0; j < NQ; ++j) carry[gq][j] = fmaxf(carry[gq][j], in ? sc[j] : -INFINITY);
          }
          if (seg_end == cur_end) {  // the document ends here: sum of the maxima in query order from 0.0 (x86_64.rs:139)
            const bool split = G > 1 && a.split;
            const int lim = split ? (gq ? (int)a.n_q_b : n_q) : n_q - gq * NQ;  // query tokens of this group
            float total = (G > 1 && gq > 0 && !split) ? part[seg] : 0.0f;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              if (j < lim) total += redux_max(carry[gq][j]);
              carry[gq][j] = -INFINITY;
            }
            if (split) {
              if (lane == 0) (gq ? a.out_b : a.out)[cur_doc] = total;
            } else if (gq == G - 1) {
              if (lane == 0) a.out[cur_doc] = a.accumulate ? a.out[cur_doc] + total : total;
            } else if (lane == 0) {
              part[seg] = total;
            }
            ++seg;
          }
          pos = seg_end;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// per-token 1/||x|| for maxsim_cosine: one warp per token, 0 when ||x||^2 <= 1e-18 (cosine_avx512's guard)
__global__ void token_inv_norms_kernel(const float* __restrict__ tokens, size_t total, unsigned dim,
                                       float* __restrict__ inv) {
  const size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= total) return;
  const float* p = tokens + t * dim;
  float ss = 0.0f;
  for (unsigned k = lane; k < dim; k += 32) ss = fmaf(p[k], p[k], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
  // ss NaN fails the guard like the reference's `bb > 1e-18` (score 0.0); ss = +inf passes it there and yields a NaN
  // cosine, which f32::max skips -> NaN here so that the epilogue's max skips it too
  if (lane == 0) inv[t] = ss > EPS_SQ ? (ss < INFINITY ? 1.0f / sqrtf(ss) : __int_as_float(0x7FC00000)) : 0.0f;
}

}  // namespace

bool make_token_tmap(CUtensorMap* m, const float* dev_tokens, size_t total_tokens, size_t dim) {
  // TMA needs a 16-byte row pitch; a row that is not a whole number of 32-column panels is completed with zeros by the
  // out-of-bounds fill of the last box
  if (dim == 0 || dim % 4 != 0 || dim > 256 || total_tokens == 0) return false;
  return make_tmap_f32_rows(m, dev_tokens, total_tokens, dim, CHUNK);
}

cudaError_t launch_token_inv_norms(const float* dev_tokens, size_t total_tokens, size_t dim, float* dev_inv,
                                   cudaStream_t s, LaunchCounter* launches) {
  if (total_tokens == 0) return cudaSuccess;
  token_inv_norms_kernel<<<(unsigned)((total_tokens + 7) / 8), 256, 0, s>>>(dev_tokens, total_tokens, (unsigned)dim, dev_inv);
  ++*launches;
  return cudaGetLastError();
}

// dim <= 256, a multiple of 4 (panels of 32 columns, the last ones zero-filled by TMA; 129..256: two K halves per
// tile); up to 64 query tokens per corpus pass (32 above dim 128), more in several passes (the sum over query tokens
// is additive across passes)
bool maxsim_tc_supported(const TokView& v, size_t n_q) {
  return v.dim >= 4 && v.dim <= 256 && v.dim % 4 == 0 && n_q >= 1 && v.total_tokens > 0 &&
         v.tmap_valid && v.inv_norms != nullptr && v.total_tokens < 0x7FFFFF00ull;
}

namespace {
template <bool COSINE, int P, int G, int KH = 1>
cudaError_t launch_shape(const CUtensorMap& tm, const TcArgs& a, unsigned grid, cudaStream_t s) {
  using SH = Shape<P, G, KH>;
  constexpr size_t smem = (size_t)SH::STAGES * SH::STAGE_BYTES + SH::QBYTES + sizeof(SharedTail);
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr_set_dev[16] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(maxsim_tc_kernel<COSINE, P, G, KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  maxsim_tc_kernel<COSINE, P, G, KH><<<grid, TC_THREADS, smem, s>>>(tm, a);
  return cudaGetLastError();
}
template <bool COSINE, int G>
cudaError_t launch_dim(size_t dim, const CUtensorMap& tm, const TcArgs& a, unsigned grid, cudaStream_t s) {
  if (dim > 128) {  // 129..256 columns: two K halves of 128 columns per tile, one query group per pass
    if (G != 1) return cudaErrorInvalidValue;
    return launch_shape<COSINE, 4, 1, 2>(tm, a, grid, s);
  }
  switch ((dim + 31) / 32) {  // panels of 32 columns; TMA zero-fills the columns of the last panel past `dim`
    case 1: return launch_shape<COSINE, 1, G>(tm, a, grid, s);
    case 2: return launch_shape<COSINE, 2, G>(tm, a, grid, s);
    case 3: return launch_shape<COSINE, 3, G>(tm, a, grid, s);
    case 4: return launch_shape<COSINE, 4, G>(tm, a, grid, s);
  }
  return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t launch_maxsim_tc(const TokView& v, const float* dev_q, size_t n_q, int cosine, float* dev_scores,
                             int num_sms, cudaStream_t s, LaunchCounter* launches) {
  // empty documents never see a token: their score is 0.0 (src/maxsim.rs:97-99)
  cudaError_t e = cudaMemsetAsync(dev_scores, 0, v.n_docs * sizeof(float), s);
  if (e != cudaSuccess) return e;
  TcArgs a{};
  a.doc_offsets = v.doc_offsets;
  a.inv_norms = v.inv_norms;
  a.uniform_tokens = v.uniform_tokens;
  a.total_tokens = v.total_tokens;
  a.n_docs = (unsigned)v.n_docs;
  a.dim = (unsigned)v.dim;
  a.out = dev_scores;
  static const int dbg = getenv("INNR_MAXSIM_DEBUG") ? atoi(getenv("INNR_MAXSIM_DEBUG")) : 0;
  a.debug_mode = dbg;
  unsigned grid = (unsigned)num_sms;
  const unsigned long long tiles = (v.total_tokens + TILE_M - 1) / TILE_M;
  if (grid > tiles) grid = (unsigned)tiles;
  if (grid > v.n_docs) grid = (unsigned)v.n_docs;
  if (grid == 0) grid = 1;
  // one corpus pass per 64 query tokens (two column groups per accumulator), a last pass of <= 32 with one group
  for (size_t q0 = 0; q0 < n_q;) {
    const size_t rem = n_q - q0;
    const bool two = rem > NQ && v.dim <= 128;  // two column groups per pass need the B operand to fit: dim <= 128
    const size_t cap = two ? 2 * NQ : NQ;
    const size_t take = rem < cap ? rem : cap;
    a.n_q = (unsigned)take;
    a.q = dev_q + q0 * v.dim;
    a.accumulate = q0 > 0;
    if (two) e = cosine ? launch_dim<true, 2>(v.dim, v.tmap, a, grid, s) : launch_dim<false, 2>(v.dim, v.tmap, a, grid, s);
    else e = cosine ? launch_dim<true, 1>(v.dim, v.tmap, a, grid, s) : launch_dim<false, 1>(v.dim, v.tmap, a, grid, s);
    if (e != cudaSuccess) return e;
    ++*launches;
    q0 += take;
  }
  return cudaSuccess;
}

// A batch of queries (each <= 32 tokens) against the document set: two queries share every corpus pass (their tokens are
// the two column groups of one accumulator); dev_scores is n_queries x n_docs.
cudaError_t launch_maxsim_tc_batch(const TokView& v, const float* dev_q, size_t n_queries, size_t n_q, int cosine,
                                   float* dev_scores, int num_sms, cudaStream_t s, LaunchCounter* launches) {
  cudaError_t e = cudaMemsetAsync(dev_scores, 0, n_queries * v.n_docs * sizeof(float), s);
  if (e != cudaSuccess) return e;
  TcArgs a{};
  a.doc_offsets = v.doc_offsets;
  a.inv_norms = v.inv_norms;
  a.uniform_tokens = v.uniform_tokens;
  a.total_tokens = v.total_tokens;
  a.n_docs = (unsigned)v.n_docs;
  a.dim = (unsigned)v.dim;
  unsigned grid = (unsigned)num_sms;
  const unsigned long long tiles = (v.total_tokens + TILE_M - 1) / TILE_M;
  if (grid > tiles) grid = (unsigned)tiles;
  if (grid > v.n_docs) grid = (unsigned)v.n_docs;
  if (grid == 0) grid = 1;
  const size_t per_pass = v.dim <= 128 ? 2 : 1;  // dim > 128: the B operand of one 256-d query fills a pair's space
  for (size_t i = 0; i < n_queries; i += per_pass) {
    const bool pair = per_pass == 2 && i + 1 < n_queries;
    a.q = dev_q + i * n_q * v.dim;
    a.n_q = (unsigned)n_q;
    a.out = dev_scores + i * v.n_docs;
    a.split = pair ? 1 : 0;
    a.q_b = pair ? dev_q + (i + 1) * n_q * v.dim : nullptr;
    a.n_q_b = (unsigned)n_q;
    a.out_b = pair ? dev_scores + (i + 1) * v.n_docs : nullptr;
    if (pair) e = cosine ? launch_dim<true, 2>(v.dim, v.tmap, a, grid, s) : launch_dim<false, 2>(v.dim, v.tmap, a, grid, s);
    else e = cosine ? launch_dim<true, 1>(v.dim, v.tmap, a, grid, s) : launch_dim<false, 1>(v.dim, v.tmap, a, grid, s);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  return cudaSuccess;
}

}  // namespace innr
