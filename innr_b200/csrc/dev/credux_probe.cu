// credux_probe.cu -- microbenchmark behind the MaxSim epilogue design: cycles per 32-column warp max-reduction with
//   (a) redux.sync.max.f32 (SASS CREDUX.MAX.F32, result in a uniform register) + FMNMX into a per-thread running max
//   (b) the halving shuffle butterfly (31 SHFL + 31 FMNMX + selects)
// and of packed f32x2 add / mul. One warp per SM sub-partition (4 warps per CTA), one CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o credux_probe credux_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float redux_max(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}

template <int MODE>
__global__ void __launch_bounds__(128) probe(const float* in, float* out, long long* cycles, int iters) {
  const int lane = threadIdx.x & 31;
  float v[32], carry[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] = in[(threadIdx.x * 32 + j) & 1023];
    carry[j] = -INFINITY;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) carry[j] = fmaxf(carry[j], redux_max(v[j]));
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += 1.0f;  // keeps the reductions from being hoisted
    } else if (MODE == 1) {
      float a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float mine = (lane & 16) ? v[j + 16] : v[j], send = (lane & 16) ? v[j] : v[j + 16];
        a[j] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 16));
      }
      float b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float mine = (lane & 8) ? a[j + 8] : a[j], send = (lane & 8) ? a[j] : a[j + 8];
        b[j] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 8));
      }
      float c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float mine = (lane & 4) ? b[j + 4] : b[j], send = (lane & 4) ? b[j] : b[j + 4];
        c[j] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 4));
      }
      float d[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float mine = (lane & 2) ? c[j + 2] : c[j], send = (lane & 2) ? c[j] : c[j + 2];
        d[j] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 2));
      }
      const float mine = (lane & 1) ? d[1] : d[0], send = (lane & 1) ? d[0] : d[1];
      carry[0] = fmaxf(carry[0], fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 1)));
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += 1.0f;
    } else {  // only the 32 FADD of the loop body (baseline to subtract)
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += 1.0f;
    }
  }
  const long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += carry[j] + v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[MODE] = t1 - t0;
}

int main() {
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  cudaMalloc(&out, 148 * 128 * 4);
  cudaMallocManaged(&cyc, 3 * sizeof(long long));
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    probe<0><<<148, 128>>>(in, out, cyc, iters);
    probe<1><<<148, 128>>>(in, out, cyc, iters);
    probe<2><<<148, 128>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
  }
  printf("cycles per 32-column warp reduction (1 warp per SMSP): credux+fmnmx %.1f | shuffle butterfly %.1f | (loop baseline %.1f)\n",
         (double)cyc[0] / iters, (double)cyc[1] / iters, (double)cyc[2] / iters);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
