// umma_probe.cu -- one-tile device test of the tcgen05 path used by the MaxSim kernel:
// D[128 x 64] = A[128 x 128] * B[64 x 128]^T, kind::tf32, A and B K-major, TMA SWIZZLE_128B panels.
// Checks descriptor encodings, the TMEM accumulator layout and tcgen05.ld 32x32b before the real kernel uses them.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../tc_common.cuh"
using namespace innr;
using namespace innr::tc;

constexpr int M = 128, N = 64, K = 128;

// A given as a PDX matrix At[k][m] (MN-major for the MMA): 16 TMA boxes of [32 k][32 m], SWIZZLE_128B.
__global__ void __launch_bounds__(128) probe_mn(const __grid_constant__ CUtensorMap tmAt,
                                                const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                          // box (kb, mb) at (kb*4+mb)*4096
  float* sB = reinterpret_cast<float*>(smem + 4 * 16384);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<64>(&tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_full, (M + N) * K * 4);
    for (int kb = 0; kb < 4; ++kb)
      for (int mb = 0; mb < 4; ++mb) tma_load_2d(sA + (kb * 4 + mb) * 4096, &tmAt, &bar_full, mb * 32, kb * 32);
    for (int p = 0; p < 4; ++p) tma_load_2d(sB + p * (N * 32), &tmB, &bar_full, p * 32, 0);
    mbar_wait(&bar_full, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(M, N, true);
    for (int kk = 0; kk < K / 8; ++kk) {
      const uint32_t a_addr = smem_u32(sA) + (kk / 4) * 16384 + (kk % 4) * 1024;
      const uint32_t b_addr = smem_u32(sB + (kk / 4) * (N * 32)) + (kk % 4) * 32;
      umma_tf32(tmem, make_smem_desc_mnmajor_sw128_32b(a_addr, 4096, 512), make_smem_desc_kmajor_sw128(b_addr), idesc, kk > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();
  uint32_t r[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA,
                                             const __grid_constant__ CUtensorMap tmB, float* out, int from_tmem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sA = reinterpret_cast<float*>(smem);                  // 4 panels x [128][32]
  float* sB = reinterpret_cast<float*>(smem + 4 * 16384);      // 4 panels x [64][32]
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_full, (M + N) * K * 4);
    for (int p = 0; p < 4; ++p) {
      tma_load_2d(sA + p * (M * 32), &tmA, &bar_full, p * 32, 0);
      tma_load_2d(sB + p * (N * 32), &tmB, &bar_full, p * 32, 0);
    }
    mbar_wait(&bar_full, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(M, N);
    for (int kk = 0; kk < K / 8; ++kk) {
      const uint32_t a_addr = smem_u32(sA + (kk / 4) * (M * 32)) + (kk % 4) * 32;
      const uint32_t b_addr = smem_u32(sB + (kk / 4) * (N * 32)) + (kk % 4) * 32;
      umma_tf32(tmem, make_smem_desc_kmajor_sw128(a_addr), make_smem_desc_kmajor_sw128(b_addr), idesc, kk > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();
  if (from_tmem) {
    // second pass: A staged into TMEM columns [64, 192) by the owning lanes, D2 = A(tmem) * B^T into columns [192, 256)
    __shared__ uint64_t bar2;
    if (threadIdx.x == 0) { mbar_init(&bar2, 1); fence_barrier_init(); }
    for (int c0 = 0; c0 < K; c0 += 32) {
      uint32_t v[32];
      const int row = threadIdx.x;
      for (int j = 0; j < 32; ++j) {
        const int k = c0 + j, p = k >> 5, c = (k & 31) >> 2, e = k & 3;
        v[j] = __float_as_uint(*reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sA) + p * (M * 128) + row * 128 +
                                                         ((c ^ (row & 7)) << 4) + (e << 2)));
      }
      tmem_st_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c0, v);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_tf32(M, N);
      for (int kk = 0; kk < K / 8; ++kk) {
        const uint32_t b_addr = smem_u32(sB + (kk / 4) * (N * 32)) + (kk % 4) * 32;
        umma_tf32_ts(tmem + 192, tmem + 64 + kk * 8, make_smem_desc_kmajor_sw128(b_addr), idesc, kk > 0);
      }
      umma_commit(&bar2);
    }
    mbar_wait(&bar2, 0);
    tc_fence_after_sync();
  }
  const uint32_t dcol = from_tmem ? 192 : 0;
  uint32_t r[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + dcol + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

int main() {
  std::vector<float> A(M * K), B(N * K);
  srand(1);
  for (auto& x : A) x = (float)((rand() % 33) - 16) / 8.0f;   // exactly representable in tf32
  for (auto& x : B) x = (float)((rand() % 33) - 16) / 16.0f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, M * N * 4);
  CUtensorMap tmA, tmB;
  if (!make_tmap_f32_rows(&tmA, dA, M, K, M) || !make_tmap_f32_rows(&tmB, dB, N, K, N)) { printf("tensor map failed\n"); return 2; }
  size_t smem = 4 * 16384 + 4 * 8192 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int rc = 0;
  for (int from_tmem = 0; from_tmem < 2; ++from_tmem) {
    cudaMemset(dO, 0, M * N * 4);
    probe<<<1, 128, smem>>>(tmA, tmB, dO, from_tmem);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<float> O(M * N);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[i * K + k] * B[j * K + k];
      double err = fabs(ref - O[i * N + j]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
    }
    printf("umma_probe (A from %s): max abs err %.3g, bad %d / %d  (O[0][0]=%f O[5][7]=%f)\n", from_tmem ? "TMEM" : "smem", maxerr, bad, M * N, O[0], O[5 * N + 7]);
    rc |= bad ? 1 : 0;
  }
  {  // MN-major A
    std::vector<float> At(K * M);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) At[k * M + i] = A[i * K + k];
    float* dAt; cudaMalloc(&dAt, At.size() * 4);
    cudaMemcpy(dAt, At.data(), At.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tmAt;
    if (!make_tmap_f32_rows(&tmAt, dAt, K, M, 32, 0, true)) { printf("tensor map At failed\n"); return 2; }
    cudaFuncSetAttribute(probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(dO, 0, M * N * 4);
    probe_mn<<<1, 128, smem>>>(tmAt, tmB, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error (mn): %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<float> O(M * N);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[i * K + k] * B[j * K + k];
      double err = fabs(ref - O[i * N + j]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
    }
    printf("umma_probe (A MN-major from PDX): max abs err %.3g, bad %d / %d  (O[0][0]=%f O[5][7]=%f)\n", maxerr, bad, M * N, O[0], O[5 * N + 7]);
    rc |= bad ? 1 : 0;
  }
  return rc;
}
