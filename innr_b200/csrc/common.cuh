// common.cuh -- shared device helpers for libinnr_cuda (sm_100a).
//
//  * composite 64-bit selection keys: (order_bits(score) << 32) | global_index, so that
//    "stable sort, truncate k" (src/batch.rs:756-758, src/scalar.rs:390-391, examples/binary_demo.rs:178)
//    becomes "k smallest distinct keys";
//  * WarpList<R>: a per-warp register-resident sorted list of 32*R keys (the fused top-k tracker that
//    replaces TopK::insert, src/topk.rs:96-121, and the full sorts);
//  * block merge + last-CTA merge so one launch yields the final top-k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace innr {

constexpr uint64_t KEY_SENTINEL = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned FULL_MASK = 0xFFFFFFFFu;

// f32::total_cmp order -> unsigned ascending order (Rust core::f32::total_cmp transform, then sign flip)
__device__ __forceinline__ uint32_t f32_order_bits(float x) {
  uint32_t b = __float_as_uint(x);
  b ^= ((uint32_t)((int32_t)b >> 31)) >> 1;
  return b ^ 0x80000000u;
}
__host__ __device__ __forceinline__ uint32_t order_bits_to_f32_bits(uint32_t u) {
  uint32_t o = u ^ 0x80000000u;
  o ^= ((uint32_t)((int32_t)o >> 31)) >> 1;
  return o;
}
// ascending score (L2, Hamming-as-float never used): smaller score first, ties -> lower index
__device__ __forceinline__ uint64_t make_key_asc(float score, uint32_t gidx) {
  return ((uint64_t)f32_order_bits(score) << 32) | gidx;
}
// descending score (dot, cosine, u8): larger score first, ties -> lower index
__device__ __forceinline__ uint64_t make_key_desc(float score, uint32_t gidx) {
  return ((uint64_t)(~f32_order_bits(score)) << 32) | gidx;
}
__device__ __forceinline__ uint64_t make_key_u32(uint32_t dist, uint32_t gidx) {
  return ((uint64_t)dist << 32) | gidx;
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(FULL_MASK, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(FULL_MASK, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int delta) {
  uint32_t lo = __shfl_up_sync(FULL_MASK, (uint32_t)v, delta);
  uint32_t hi = __shfl_up_sync(FULL_MASK, (uint32_t)(v >> 32), delta);
  return ((uint64_t)hi << 32) | lo;
}

// streaming 128-bit global load: read-only path, do not allocate in L1 (each corpus byte is used once)
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- stateless synthetic generators; the oracle has the same functions on the CPU ----
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// G-hash: float(u >> 40) * 2^-23 - 1 : 24-bit integer -> exact f32, exact scale, exact subtract
__device__ __forceinline__ float ghash_value(uint64_t salt, uint64_t idx) {
  uint64_t u = splitmix64(salt + idx);
  return __fadd_rn(__fmul_rn(__uint2float_rn((unsigned)(u >> 40)), 1.0f / 8388608.0f), -1.0f);
}
// G-ref: generate_embedding (examples/batch_demo.rs:233-242), bit-identical to the Rust code
__device__ __forceinline__ float gref_value(uint64_t seed, uint64_t j) {
  uint64_t x = seed * 6364136223846793005ull + j * 1442695040888963407ull;
  float f = __uint2float_rn((unsigned)(x >> 33));  // (x >> 33) as f32
  f = __fmul_rn(f, 1.0f / 2147483648.0f);          // / (1u64 << 31) as f32  (exact)
  return __fadd_rn(__fmul_rn(f, 2.0f), -1.0f);     // * 2.0 - 1.0
}

// Sorted ascending list of 32*R keys held in registers: element p = r*32 + lane lives in v[r] of `lane`.
template <int R>
struct WarpList {
  uint64_t v[R];

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = KEY_SENTINEL;
  }
  // element at position p (uniform), broadcast to the warp. The register is picked with uniform branches
  // (a select chain is turned into an indexed local-memory load by the compiler).
  __device__ __forceinline__ uint64_t at(int p) const {
    const int l = p & 31;
    if (R == 1) return shfl_u64(v[0], l);
    uint64_t x;
    switch (p >> 5) {
      case 0: x = shfl_u64(v[0], l); break;
      case 1: x = shfl_u64(v[R > 1 ? 1 : 0], l); break;
      case 2: x = shfl_u64(v[R > 2 ? 2 : 0], l); break;
      default: x = shfl_u64(v[R > 3 ? 3 : 0], l); break;
    }
    return x;
  }
  // insert a warp-uniform key x (distinct from all present keys); the largest element falls off
  __device__ __forceinline__ void insert(uint64_t x, int lane) {
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      uint64_t cur = v[r];
      uint64_t up = shfl_up_u64(cur, 1);
      bool prev_lt = true;
      if (r > 0) {
        uint64_t prev_last = shfl_u64(v[r - 1], 31);
        if (lane == 0) up = prev_last;
        prev_lt = up < x;
      } else {
        prev_lt = (lane == 0) ? true : (up < x);
      }
      v[r] = (cur < x) ? cur : (prev_lt ? x : up);
    }
  }
  // offer one candidate per lane; thr = current k-th key (uniform), updated in place
  __device__ __forceinline__ void offer(uint64_t key, bool valid, uint64_t& thr, int k, int lane) {
    unsigned m = __ballot_sync(FULL_MASK, valid && key < thr);
    while (m) {
      int src = __ffs(m) - 1;
      uint64_t x = shfl_u64(key, src);
      insert(x, lane);
      thr = at(k - 1);
      if (lane == src) valid = false;
      m = __ballot_sync(FULL_MASK, valid && key < thr);
    }
  }
  // write the first k keys to dst[0..k)
  __device__ __forceinline__ void store(uint64_t* dst, int k, int lane) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int p = r * 32 + lane;
      if (p < k) dst[p] = v[r];
    }
  }
};

// Merge `n_lists` sorted key lists of length k (stride `stride` u64) from `src` into `list` (warp-cooperative).
// GLOBAL: src is global memory written by other CTAs of this launch -> read through L2 (ld.global.cg).
template <int R, bool GLOBAL>
__device__ __forceinline__ void warp_merge_lists(WarpList<R>& list, uint64_t& thr, const uint64_t* src,
                                                 int first, int step, int n_lists, size_t stride, int k,
                                                 int lane) {
  for (int g = first; g < n_lists; g += step) {
    const uint64_t* lp = src + (size_t)g * stride;
    for (int j = 0; j < k; j += 32) {
      int p = j + lane;
      uint64_t key = KEY_SENTINEL;
      if (p < k) key = GLOBAL ? __ldcg((const unsigned long long*)(lp + p)) : lp[p];
      // lists are sorted: once the first key of a 32-chunk misses the threshold the rest do too
      uint64_t first_key = shfl_u64(key, 0);
      if (first_key >= thr) break;
      list.offer(key, p < k && key != KEY_SENTINEL, thr, k, lane);
    }
  }
}

// Block-level finish shared by all selection kernels.
//  1. every warp publishes its list to shared memory, warp 0 merges them -> CTA top-k
//  2. CTA top-k goes to partials[(blockIdx.x * nq + q) * k ...]
//  3. the last CTA to arrive (ticket) merges all partials and writes out_keys[q * k ...]
// `lists`/`thrs` are per-warp arrays of QB lists. smem must hold (warps * k) u64.
template <int R, int QB>
__device__ __forceinline__ void block_finish(WarpList<R> (&lists)[QB], uint64_t (&thrs)[QB], int nq_valid,
                                             int k, uint64_t* smem_keys, uint64_t* partials,
                                             uint64_t* out_keys, unsigned* ticket) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  __shared__ unsigned s_is_last;
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    if (q >= nq_valid) break;
    __syncthreads();
    lists[q].store(smem_keys + (size_t)warp * k, k, lane);
    __syncthreads();
    if (warp == 0) {
      warp_merge_lists<R, false>(lists[q], thrs[q], smem_keys, 1, 1, n_warps, (size_t)k, k, lane);
      lists[q].store(partials + ((size_t)blockIdx.x * nq_valid + q) * k, k, lane);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ticket, 1u);
    s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    if (q >= nq_valid) break;
    WarpList<R> fin;
    fin.init();
    uint64_t thr = KEY_SENTINEL;
    // each warp merges a strided subset of the CTA partials for query q
    warp_merge_lists<R, true>(fin, thr, partials + (size_t)q * k, warp, n_warps, (int)gridDim.x,
                        (size_t)nq_valid * k, k, lane);
    __syncthreads();
    fin.store(smem_keys + (size_t)warp * k, k, lane);
    __syncthreads();
    if (warp == 0) {
      warp_merge_lists<R, false>(fin, thr, smem_keys, 1, 1, n_warps, (size_t)k, k, lane);
      fin.store(out_keys + (size_t)q * k, k, lane);
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch on this stream
}

}  // namespace innr
