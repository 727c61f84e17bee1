// u8.cu -- HBM-bound asymmetric f32 x u8 scan with fused top-k (bit-exact with the reference's AVX2 kernel).
//
// Replaces (reference, innr 0.6.3):
//   dot_u8_f32_avx2                  src/arch/x86_64.rs:928-1020 (reached via src/scalar.rs:327-349 for n >= 16,
//                                    on AVX-512 hosts too -- no AVX-512 variant exists)
//   mixed_dot_u8_f32_portable        src/scalar.rs:353-358 (n < 16)
//   asymmetric_dot_u8[_precomputed]  src/scalar.rs:261-300:  (alpha/255)*mixed + offset*query_sum, unfused
//   query_context                    src/scalar.rs:236-240
//   quantize_u8                      src/scalar.rs:212-225
//   batch_knn_u8                     src/scalar.rs:370-393 (stable descending sort, truncate)
//
// Bit-exact recipe (verified against the intrinsics in tests/test_oracle_simd.py): the AVX2 kernel
// is 32 independent FMA chains -- chain c accumulates elements idx == c (mod 32), ascending, with fused
// multiply-add; the four 8-lane accumulators combine as (a0+a1)+(a2+a3), then lanes j+(j+4), (0+2),(1+3), and
// the sum of those two. Then an 8-wide remainder (own 8 chains, same horizontal sum, added), then a scalar tail
// with separately rounded multiply and add. fmaf on the GPU is IEEE-exact, so one thread holding the 32 chains
// of one vector reproduces the CPU result bit for bit.
//
// Device layout: 16-byte chunks (16 dimensions), chunk-major: codes[c * ld + i]. One thread owns one vector; a
// warp load is 512 contiguous bytes. u8 -> f32 is exact via the 2^23 magic number (PRMT + FADD, no I2F).
#include "common.cuh"
#include "kernels.cuh"

namespace innr {

namespace {

constexpr int U8_THREADS = 256;

// The 2^23 magic constant is kept in a register the compiler cannot see through, so that PRMT takes the byte
// selector as its immediate operand (with a literal constant the selector is re-materialised by a MOV per byte).
// (An asm mov is folded by ptxas too, so the value travels as a kernel argument: U8Args::magic.)
#ifndef INNR_U8_I2F
#define INNR_U8_I2F 0
#endif
__device__ __forceinline__ float byte_to_f32(unsigned word, int k, unsigned magic) {
#if INNR_U8_I2F
  (void)magic;
  return __uint2float_rn((word >> (8 * k)) & 0xFFu);  // I2F.F32.U8 with a byte selector
#else
  // bytes: [b_k, 0x00, 0x00, 0x4B] = 2^23 + b_k as f32; subtracting 2^23 is exact
  unsigned bits = __byte_perm(word, magic, 0x7440u | (unsigned)k);
  return __fsub_rn(__uint_as_float(bits), 8388608.0f);
#endif
}

__device__ __forceinline__ void fma16(const uint4 v, const float* __restrict__ q, float* acc, unsigned magic) {
  const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 qv = *reinterpret_cast<const float4*>(q + 4 * g);  // 16-byte aligned by construction
    acc[4 * g + 0] = fmaf(qv.x, byte_to_f32(w[g], 0, magic), acc[4 * g + 0]);
    acc[4 * g + 1] = fmaf(qv.y, byte_to_f32(w[g], 1, magic), acc[4 * g + 1]);
    acc[4 * g + 2] = fmaf(qv.z, byte_to_f32(w[g], 2, magic), acc[4 * g + 2]);
    acc[4 * g + 3] = fmaf(qv.w, byte_to_f32(w[g], 3, magic), acc[4 * g + 3]);
  }
}

__device__ __forceinline__ float hsum8(const float* v) {  // src/arch/x86_64.rs:982-987
  float s0 = __fadd_rn(v[0], v[4]), s1 = __fadd_rn(v[1], v[5]);
  float s2 = __fadd_rn(v[2], v[6]), s3 = __fadd_rn(v[3], v[7]);
  return __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
}

// mixed_dot_u8_f32 of the smem query against the vector whose chunk 0 is at p
__device__ __forceinline__ float mixed_dot(const uint4* __restrict__ p, size_t ld, unsigned d, unsigned chunks,
                                           const float* __restrict__ sq, const unsigned magic) {
  if (d == 0) return 0.0f;
  if (d < 16) {  // portable path: sequential unfused sum (src/scalar.rs:353-358)
    uint4 v = ldg_stream_u4(p);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    float s = 0.0f;
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if ((unsigned)e < d) s = __fadd_rn(s, __fmul_rn(sq[e], byte_to_f32(w[e >> 2], e & 3, magic)));
    return s;
  }
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
  const unsigned chunks32 = d / 32;
  unsigned b = 0;
  for (; b + 2 <= chunks32; b += 2) {  // 64 dimensions per iteration, 4 loads in flight
    uint4 v0 = ldg_stream_u4(p + (size_t)(2 * b) * ld);
    uint4 v1 = ldg_stream_u4(p + (size_t)(2 * b + 1) * ld);
    uint4 v2 = ldg_stream_u4(p + (size_t)(2 * b + 2) * ld);
    uint4 v3 = ldg_stream_u4(p + (size_t)(2 * b + 3) * ld);
    fma16(v0, sq + 32 * b, acc, magic);
    fma16(v1, sq + 32 * b + 16, acc + 16, magic);
    fma16(v2, sq + 32 * b + 32, acc, magic);
    fma16(v3, sq + 32 * b + 48, acc + 16, magic);
  }
  for (; b < chunks32; ++b) {
    uint4 v0 = ldg_stream_u4(p + (size_t)(2 * b) * ld);
    uint4 v1 = ldg_stream_u4(p + (size_t)(2 * b + 1) * ld);
    fma16(v0, sq + 32 * b, acc, magic);
    fma16(v1, sq + 32 * b + 16, acc + 16, magic);
  }
  float all[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    all[j] = __fadd_rn(__fadd_rn(acc[j], acc[8 + j]), __fadd_rn(acc[16 + j], acc[24 + j]));
  float result = hsum8(all);

  // remainder: up to 31 elements in chunks 2*chunks32 (+1); zero padded in memory
  const unsigned rs = chunks32 * 32, remaining = d - rs, n8 = (remaining / 8) * 8;
  float rem[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) rem[j] = 0.0f;
  unsigned w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (remaining > 0) {
    uint4 v0 = ldg_stream_u4(p + (size_t)(2 * chunks32) * ld);
    w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
    if (remaining > 16) {
      uint4 v1 = ldg_stream_u4(p + (size_t)(2 * chunks32 + 1) * ld);
      w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    }
  }
#pragma unroll
  for (int e = 0; e < 24; ++e)  // 8-wide chunks (at most 3)
    if ((unsigned)e < n8) rem[e & 7] = fmaf(sq[rs + e], byte_to_f32(w[e >> 2], e & 3, magic), rem[e & 7]);
  result = __fadd_rn(result, hsum8(rem));  // always added, even when empty (src/arch/x86_64.rs:1004-1009)
#pragma unroll
  for (int e = 0; e < 31; ++e)  // scalar tail, unfused (src/arch/x86_64.rs:1013-1017)
    if ((unsigned)e >= n8 && (unsigned)e < remaining)
      result = __fadd_rn(result, __fmul_rn(sq[rs + e], byte_to_f32(w[e >> 2], e & 3, magic)));
  return result;
}

struct U8Args {
  const uint4* data;
  unsigned long long ld;
  unsigned n, d, chunks, n_tiles, index_base;
  float alpha, offset;
  unsigned magic;      // 0x4B000000 (2^23 as f32 bits), see byte_to_f32
  const float* query;  // device, d floats
  int k, mode;         // mode (scores kernel): 0 raw mixed dot, 1 asymmetric score
  uint64_t* partials;
  uint64_t* out_keys;
  uint64_t* group_partials;
  unsigned* tickets;
  float* scores_out;
};

#ifndef INNR_U8_MINB
#define INNR_U8_MINB 3
#endif
template <int R, bool KNN>
__global__ void __launch_bounds__(U8_THREADS, INNR_U8_MINB) u8_scan_kernel(const U8Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned d_pad = (a.d + 31) / 32 * 32 + 32;
  float* sq = reinterpret_cast<float*>(smem_raw);
  float* s_misc = sq + d_pad;  // [0] = query_sum
  uint64_t* smem_keys = reinterpret_cast<uint64_t*>(s_misc + 4);
  const int lane = threadIdx.x & 31;
  for (unsigned i = threadIdx.x; i < d_pad; i += blockDim.x) sq[i] = i < a.d ? a.query[i] : 0.0f;
  if (threadIdx.x == 0) {  // query_context: query.iter().sum() sequential (src/scalar.rs:236-240)
    float s = 0.0f;
    for (unsigned i = 0; i < a.d; ++i) s = __fadd_rn(s, a.query[i]);
    s_misc[0] = s;
  }
  __syncthreads();
  const float scale = __fdiv_rn(a.alpha, 255.0f);             // params.alpha / 255.0
  const float bias = __fmul_rn(a.offset, s_misc[0]);          // params.offset * ctx.query_sum

  WarpList<R> lists[1];
  uint64_t thrs[1];
  lists[0].init();
  thrs[0] = KEY_SENTINEL;

  for (unsigned tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const unsigned i = tile * U8_THREADS + threadIdx.x;
    const bool valid = i < a.n;
    float score = 0.0f;
    if (valid) {
      float mixed = mixed_dot(a.data + i, a.ld, a.d, a.chunks, sq, a.magic);
      score = (KNN || a.mode == 1) ? __fadd_rn(__fmul_rn(scale, mixed), bias) : mixed;  // src/scalar.rs:299
    }
    if (KNN) lists[0].offer(make_key_desc(score, a.index_base + i), valid, thrs[0], a.k, lane);
    else if (valid) a.scores_out[i] = score;
  }
  if (KNN) block_finish<R, 1>(lists, 1, a.k, smem_keys, a.partials, a.group_partials, a.out_keys, a.tickets);
}

__global__ void u8_pack_kernel(const uint8_t* __restrict__ rows, unsigned n, unsigned d, uint4* __restrict__ codes,
                               size_t ld, unsigned chunks) {
  const size_t total = (size_t)chunks * ld;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const unsigned c = (unsigned)(t / ld);
    const size_t i = t % ld;
    unsigned w[4] = {0, 0, 0, 0};
    if (i < n) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        unsigned dd = 16 * c + e;
        if (dd < d) w[e >> 2] |= (unsigned)rows[i * d + dd] << (8 * (e & 3));
      }
    }
    codes[t] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ unsigned quantize_one(float v, float offset, float inv_alpha) {
  // let normalized = (v - params.offset) * inv_alpha; normalized.round().clamp(0.0, 255.0) as u8
  float r = roundf(__fmul_rn(__fsub_rn(v, offset), inv_alpha));  // f32::round: half away from zero
  if (r != r) return 0u;                                         // NaN as u8 == 0
  r = fminf(fmaxf(r, 0.0f), 255.0f);
  return (unsigned)r;
}

__global__ void generate_u8_kernel(uint64_t salt, uint64_t first_row, unsigned n, unsigned d, float alpha,
                                   float offset, uint4* __restrict__ codes, size_t ld, unsigned chunks) {
  const size_t total = (size_t)chunks * ld;
  const float inv_alpha = __fdiv_rn(255.0f, alpha);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const unsigned c = (unsigned)(t / ld);
    const size_t i = t % ld;
    unsigned w[4] = {0, 0, 0, 0};
    if (i < n) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        unsigned dd = 16 * c + e;
        if (dd < d) {
          float v = ghash_value(salt, (first_row + i) * d + dd);
          w[e >> 2] |= quantize_one(v, offset, inv_alpha) << (8 * (e & 3));
        }
      }
    }
    codes[t] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void quantize_u8_kernel(const float* __restrict__ values, size_t n, float alpha, float offset,
                                   uint8_t* __restrict__ out) {
  const float inv_alpha = __fdiv_rn(255.0f, alpha);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
    out[t] = (uint8_t)quantize_one(values[t], offset, inv_alpha);
}

size_t u8_smem(size_t d, size_t k, bool knn) {
  size_t d_pad = (d + 31) / 32 * 32 + 32;
  return (d_pad + 4) * sizeof(float) + (knn ? (size_t)(U8_THREADS / 32) * k * sizeof(uint64_t) : 0);
}

template <int R, bool KNN>
cudaError_t launch_u8(const U8Args& a, size_t smem, int num_sms, cudaStream_t s) {
  auto kern = u8_scan_kernel<R, KNN>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, U8_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  unsigned grid = KNN ? balanced_grid(a.n_tiles, (unsigned)occ * (unsigned)num_sms) : (a.n_tiles ? a.n_tiles : 1);
  kern<<<grid, U8_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}

U8Args make_args(const U8View& v, const float* q) {
  U8Args a{};
  a.data = v.data;
  a.ld = v.ld;
  a.n = (unsigned)v.n;
  a.d = (unsigned)v.d;
  a.chunks = (unsigned)v.chunks;
  a.n_tiles = (unsigned)((v.n + U8_THREADS - 1) / U8_THREADS);
  a.index_base = v.index_base;
  a.alpha = v.alpha;
  a.offset = v.offset;
  a.magic = 0x4B000000u;
  a.query = q;
  return a;
}

}  // namespace

cudaError_t launch_u8_pack(const uint8_t* dev_rows, size_t n, size_t d, uint4* dev_codes, size_t ld,
                           cudaStream_t s, uint64_t* launches) {
  if (n == 0 || d == 0) return cudaSuccess;
  u8_pack_kernel<<<148 * 8, 256, 0, s>>>(dev_rows, (unsigned)n, (unsigned)d, dev_codes, ld, (unsigned)((d + 15) / 16));
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_generate_u8(uint64_t salt, uint64_t first_row, size_t n, size_t d, float alpha, float offset,
                               uint4* dev_codes, size_t ld, cudaStream_t s, uint64_t* launches) {
  if (n == 0 || d == 0) return cudaSuccess;
  generate_u8_kernel<<<148 * 16, 256, 0, s>>>(salt, first_row, (unsigned)n, (unsigned)d, alpha, offset, dev_codes, ld,
                                              (unsigned)((d + 15) / 16));
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_quantize_u8(const float* dev_values, size_t n, float alpha, float offset, uint8_t* dev_out,
                               cudaStream_t s, uint64_t* launches) {
  if (n == 0) return cudaSuccess;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  quantize_u8_kernel<<<grid, 256, 0, s>>>(dev_values, n, alpha, offset, dev_out);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_u8_scores(const U8View& v, int mode, const float* dev_query, float* dev_out,
                             cudaStream_t s, uint64_t* launches) {
  if (v.n == 0) return cudaSuccess;
  U8Args a = make_args(v, dev_query);
  a.mode = mode;
  a.scores_out = dev_out;
  size_t smem = u8_smem(v.d, 0, false);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = launch_u8<1, false>(a, smem, 148, s);
  if (e == cudaSuccess) ++*launches;
  return e;
}

cudaError_t launch_u8_knn(const U8View& v, const float* dev_queries, size_t nq, size_t k, uint64_t* dev_keys,
                          Workspace& ws, cudaStream_t s, uint64_t* launches) {
  if (k > 128) return cudaErrorInvalidValue;
  size_t smem = u8_smem(v.d, k, true);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  for (size_t q = 0; q < nq; ++q) {
    U8Args a = make_args(v, dev_queries + q * v.d);
    a.k = (int)k;
    a.partials = ws.partials;
    a.group_partials = ws.group_partials;
    a.tickets = ws.tickets;
    a.out_keys = dev_keys + q * k;
    cudaError_t e = (k <= 32) ? launch_u8<1, true>(a, smem, ws.num_sms, s) : launch_u8<4, true>(a, smem, ws.num_sms, s);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  return cudaSuccess;
}

}  // namespace innr
