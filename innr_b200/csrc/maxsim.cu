// maxsim.cu -- ColBERT late-interaction scoring of one query token set against a device-resident document set.
//
// Replaces (reference, innr 0.6.3):
//   maxsim          src/maxsim.rs:96-137  -> maxsim_avx512 src/arch/x86_64.rs:119-143 (sum_i max_j dot(q_i, d_j))
//   maxsim_cosine   src/maxsim.rs:168-194 -> cosine_avx512 src/arch/x86_64.rs:681-786 per pair
//   caller loop     examples/maxsim_colbert.rs:171-174 (one maxsim call per document)
//
// v1 (this file): CUDA-core register-tiled contraction, f32 FFMA, one CTA per document, document tokens staged
// through shared memory, per-query-token running max in registers, the 32 maxima summed in query order
// (total starts at 0.0 and adds in query order, x86_64.rs:139). Not bit-exact with the CPU's 64-lane FMA chains
// (different summation order); parity bar is 1e-5 relative, condition-aware.
// Cosine: per-token sum of squares computed once per token (the reference recomputes both norms for every
// pair), then ab / (sqrt(aa) * sqrt(bb)) guarded by aa > 1e-18 && bb > 1e-18 (x86_64.rs:781-785).
#include "common.cuh"
#include "kernels.cuh"

namespace innr {

namespace {

constexpr int MS_THREADS = 256;
constexpr int MS_WARPS = MS_THREADS / 32;
constexpr int MS_QR = 4;              // query rows per thread  (warp w owns q = w, w+8, w+16, w+24)
constexpr int MS_QPASS = MS_WARPS * MS_QR;  // 32 query tokens per pass
constexpr int MS_TT = 64;             // doc tokens per smem tile (lane owns tokens lane, lane+32)
constexpr float EPS_SQ = 1e-9f * 1e-9f;

struct MsArgs {
  const float* tokens;
  const uint64_t* doc_offsets;
  unsigned long long uniform_tokens;
  unsigned n_docs, dim, n_q;
  unsigned dk, q_rows;  // dimension chunk (multiple of 4); query rows held in shared memory (all padded rows, or 32)
  const float* q;
  int cosine;
  float* out;
};

__global__ void __launch_bounds__(MS_THREADS) maxsim_kernel(const MsArgs a) {
  extern __shared__ __align__(16) float smem[];
  // The contraction runs over dimension chunks of `dk` columns (the whole row when dim <= MS_DK), so shared memory does
  // not grow with dim. The query rows stay resident when one chunk covers the row and they fit (a.q_rows = all padded
  // query rows); otherwise the 32 rows of the current pass are re-staged per (token tile, chunk) from L2.
  const unsigned dim = a.dim, dk = a.dk, stride = dk + 4;  // padded row: conflict-free LDS.128
  const unsigned n_chunks = (dim + dk - 1) / dk;
  const unsigned n_pass = (a.n_q + MS_QPASS - 1) / MS_QPASS;
  const unsigned nq_pad = n_pass * MS_QPASS;
  const bool q_resident = a.q_rows == nq_pad && n_chunks == 1;
  float* sQ = smem;                                    // q_rows x stride
  float* sT = sQ + (size_t)a.q_rows * stride;          // 64 x stride
  float* s_aa = sT + (size_t)MS_TT * stride;           // nq_pad query sums of squares
  float* s_bb = s_aa + nq_pad;                         // 64 token sums of squares
  float* s_qmax = s_bb + MS_TT;                        // nq_pad per-query maxima
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  auto stage_q = [&](unsigned row0, unsigned rows, unsigned c0) {  // query rows [row0, row0 + rows), columns [c0, c0 + dk)
    for (unsigned idx = threadIdx.x; idx < rows * stride; idx += blockDim.x) {
      const unsigned r = idx / stride, c = idx % stride, gq = row0 + r, col = c0 + c;
      sQ[idx] = (gq < a.n_q && c < dk && col < dim) ? a.q[(size_t)gq * dim + col] : 0.0f;
    }
  };
  if (q_resident) stage_q(0, nq_pad, 0);
  if (a.cosine)
    for (unsigned r = threadIdx.x; r < nq_pad; r += blockDim.x) {
      float ss = 0.0f;
      if (r < a.n_q)
        for (unsigned c = 0; c < dim; ++c) {
          const float x = a.q[(size_t)r * dim + c];
          ss = fmaf(x, x, ss);
        }
      s_aa[r] = ss;
    }
  __syncthreads();

  for (unsigned doc = blockIdx.x; doc < a.n_docs; doc += gridDim.x) {
    const unsigned long long t0 = a.uniform_tokens ? (unsigned long long)doc * a.uniform_tokens : a.doc_offsets[doc];
    const unsigned long long t1 = a.uniform_tokens ? t0 + a.uniform_tokens : a.doc_offsets[doc + 1];
    const unsigned nt = (unsigned)(t1 - t0);
    if (nt == 0) {  // empty doc -> 0.0 (src/maxsim.rs:97-99)
      if (threadIdx.x == 0) a.out[doc] = 0.0f;
      continue;
    }
    for (unsigned pass = 0; pass < n_pass; ++pass) {
      float qmax[MS_QR];
#pragma unroll
      for (int r = 0; r < MS_QR; ++r) qmax[r] = -INFINITY;
      for (unsigned tt0 = 0; tt0 < nt; tt0 += MS_TT) {
        const unsigned tn = min((unsigned)MS_TT, nt - tt0);
        float acc[MS_QR][2];
#pragma unroll
        for (int r = 0; r < MS_QR; ++r) acc[r][0] = acc[r][1] = 0.0f;
        float bb = 0.0f;  // threads < 64: running sum of squares of token `threadIdx.x` across the chunks
        for (unsigned kc = 0; kc < n_chunks; ++kc) {
          const unsigned c0 = kc * dk, cw = min(dk, dim - c0);
          __syncthreads();  // previous tile / chunk fully consumed
          const float* src = a.tokens + (size_t)(t0 + tt0) * dim + c0;
          if ((dim & 3) == 0) {  // then c0 and cw are multiples of 4 as well
            const unsigned s4 = stride / 4, w4 = cw / 4;
            for (unsigned idx = threadIdx.x; idx < MS_TT * s4; idx += blockDim.x) {
              const unsigned r = idx / s4, c = idx % s4;
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r < tn && c < w4) v = ldg_stream_f4(src + (size_t)r * dim + 4 * c);
              *reinterpret_cast<float4*>(sT + r * stride + 4 * c) = v;
            }
          } else {
            for (unsigned idx = threadIdx.x; idx < MS_TT * stride; idx += blockDim.x) {
              const unsigned r = idx / stride, c = idx % stride;
              sT[idx] = (r < tn && c < cw) ? src[(size_t)r * dim + c] : 0.0f;
            }
          }
          if (!q_resident) stage_q(pass * MS_QPASS, MS_QPASS, c0);
          __syncthreads();
          if (a.cosine && threadIdx.x < MS_TT) {
            const float* tp = sT + threadIdx.x * stride;
            for (unsigned c = 0; c < cw; ++c) bb = fmaf(tp[c], tp[c], bb);
          }
          const float* tp0 = sT + lane * stride;
          const float* tp1 = sT + (lane + 32) * stride;
          const float* qp = sQ + (size_t)((q_resident ? pass * MS_QPASS : 0) + warp) * stride;
          for (unsigned c = 0; c < cw; c += 4) {  // rows are zero padded to a multiple of 4
            const float4 x0 = *reinterpret_cast<const float4*>(tp0 + c);
            const float4 x1 = *reinterpret_cast<const float4*>(tp1 + c);
#pragma unroll
            for (int r = 0; r < MS_QR; ++r) {
              const float4 qv = *reinterpret_cast<const float4*>(qp + (size_t)r * MS_WARPS * stride + c);
              acc[r][0] = fmaf(qv.x, x0.x, acc[r][0]);
              acc[r][0] = fmaf(qv.y, x0.y, acc[r][0]);
              acc[r][0] = fmaf(qv.z, x0.z, acc[r][0]);
              acc[r][0] = fmaf(qv.w, x0.w, acc[r][0]);
              acc[r][1] = fmaf(qv.x, x1.x, acc[r][1]);
              acc[r][1] = fmaf(qv.y, x1.y, acc[r][1]);
              acc[r][1] = fmaf(qv.z, x1.z, acc[r][1]);
              acc[r][1] = fmaf(qv.w, x1.w, acc[r][1]);
            }
          }
        }
        if (a.cosine) {
          if (threadIdx.x < MS_TT) s_bb[threadIdx.x] = bb;
          __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < MS_QR; ++r) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const unsigned tok = lane + 32 * h;
            float sc = acc[r][h];
            if (a.cosine) {
              const float aa = s_aa[pass * MS_QPASS + warp + r * MS_WARPS], tb = s_bb[tok];
              sc = (aa > EPS_SQ && tb > EPS_SQ) ? __fdiv_rn(sc, __fmul_rn(__fsqrt_rn(aa), __fsqrt_rn(tb))) : 0.0f;
            }
            // `if score > max_score` (x86_64.rs:135) / f32::max (maxsim.rs:190): NaN never replaces the max
            if (tok < tn && sc > qmax[r]) qmax[r] = sc;
          }
        }
      }
      // max over the 32 lanes (tokens), then publish per query token
#pragma unroll
      for (int r = 0; r < MS_QR; ++r) {
        float m = qmax[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
        if (lane == 0) s_qmax[pass * MS_QPASS + warp + r * MS_WARPS] = m;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float total = 0.0f;  // total_score = 0.0; total_score += max_score in query order
      for (unsigned i = 0; i < a.n_q; ++i) total = __fadd_rn(total, s_qmax[i]);
      a.out[doc] = total;
    }
  }
}

constexpr unsigned MS_DK = 256;  // widest dimension chunk

size_t maxsim_smem(size_t dk, size_t q_rows, size_t nq_pad) {
  return ((q_rows + MS_TT) * (dk + 4) + 2 * nq_pad + MS_TT) * sizeof(float);
}

}  // namespace

cudaError_t launch_maxsim(const TokView& v, const float* dev_q, size_t n_q, int cosine, float* dev_scores,
                          cudaStream_t s, LaunchCounter* launches) {
  if (v.n_docs == 0) return cudaSuccess;
  const size_t dim4 = (v.dim + 3) / 4 * 4;
  const size_t dk = dim4 < MS_DK ? dim4 : MS_DK;
  const size_t nq_pad = (n_q + MS_QPASS - 1) / MS_QPASS * MS_QPASS;
  // all query rows resident when the row is one chunk and they fit beside the token tile; else 32 rows per pass
  size_t q_rows = nq_pad;
  if (dim4 > dk || maxsim_smem(dk, nq_pad, nq_pad) > 200 * 1024) q_rows = MS_QPASS;
  const size_t smem = maxsim_smem(dk, q_rows, nq_pad);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;  // only for absurd query-token counts (the per-token tables)
  MsArgs a{};
  a.dk = (unsigned)dk;
  a.q_rows = (unsigned)q_rows;
  a.tokens = v.tokens;
  a.doc_offsets = v.doc_offsets;
  a.uniform_tokens = v.uniform_tokens;
  a.n_docs = (unsigned)v.n_docs;
  a.dim = (unsigned)v.dim;
  a.n_q = (unsigned)n_q;
  a.q = dev_q;
  a.cosine = cosine;
  a.out = dev_scores;
  cudaError_t e = cudaFuncSetAttribute(maxsim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, maxsim_kernel, MS_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  unsigned grid = (unsigned)occ * 148u;
  if (grid > v.n_docs) grid = (unsigned)v.n_docs;
  maxsim_kernel<<<grid, MS_THREADS, smem, s>>>(a);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace innr
