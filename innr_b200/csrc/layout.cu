// layout.cu -- upload-time layout kernels and on-device synthetic generators.
//
//  * rows -> PDX transpose: VerticalBatch::from_flat / from_rows (src/batch.rs:103-183) done on the device;
//  * G-ref generator: generate_embedding (examples/batch_demo.rs:233-242), bit-identical to the Rust code
//    (u64 wrapping arithmetic, u64->f32 RNE conversion, exact power-of-two scaling, one rounded subtract);
//  * G-hash generator: splitmix64(salt + row*d + j) -> 24-bit uniform in [-1, 1).
// Generators are stateless so any row can be re-derived on the CPU by the oracle without holding the corpus.
#include "common.cuh"
#include "kernels.cuh"

namespace innr {

namespace {

// rows: n x d row-major (a chunk of the corpus); pdx: first column of the chunk inside the PDX matrix (row pitch ld);
// columns [0, cols) are written (cols >= n: the tail of the last chunk zero-fills the pitch padding)
__global__ void transpose_rows_to_pdx_kernel(const float* __restrict__ rows, unsigned n, unsigned d,
                                             float* __restrict__ pdx, size_t ld, unsigned cols) {
  __shared__ float tile[32][33];
  const unsigned i0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  // read rows[i][d0 + tx] coalesced along d
  for (unsigned r = threadIdx.y; r < 32; r += blockDim.y) {
    unsigned i = i0 + r, dd = d0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && dd < d) ? rows[(size_t)i * d + dd] : 0.0f;
  }
  __syncthreads();
  // write pdx[dd][i0 + tx] coalesced along i
  for (unsigned r = threadIdx.y; r < 32; r += blockDim.y) {
    unsigned dd = d0 + r, i = i0 + threadIdx.x;
    if (dd < d && i < cols) pdx[(size_t)dd * ld + i] = tile[threadIdx.x][r];
  }
}

__global__ void generate_f32_pdx_kernel(int generator, uint64_t salt, uint64_t first_row, unsigned n, unsigned d,
                                        float* __restrict__ pdx, size_t ld) {
  const size_t total = (size_t)d * ld;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t dd = t / ld, i = t % ld;
    float v = 0.0f;
    if (i < n) {
      const uint64_t row = first_row + i;
      v = generator == 0 ? ghash_value(salt, row * d + dd) : gref_value(salt + row, dd);
    }
    pdx[t] = v;
  }
}

__global__ void generate_tokens_kernel(uint64_t salt, uint64_t first_row, size_t n_rows, unsigned dim,
                                       float* __restrict__ out) {
  const size_t total = n_rows * dim;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    out[t] = ghash_value(salt, first_row * dim + t);
}

}  // namespace

cudaError_t launch_transpose_rows_to_pdx(const float* dev_rows, size_t n, size_t d, float* dev_pdx, size_t ld,
                                         size_t cols, cudaStream_t s, LaunchCounter* launches) {
  if (cols == 0 || d == 0) return cudaSuccess;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((d + 31) / 32));
  transpose_rows_to_pdx_kernel<<<grid, dim3(32, 8), 0, s>>>(dev_rows, (unsigned)n, (unsigned)d, dev_pdx, ld, (unsigned)cols);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_generate_f32_pdx(int generator, uint64_t salt, uint64_t first_row, size_t n, size_t d,
                                    float* dev_pdx, size_t ld, cudaStream_t s, LaunchCounter* launches) {
  if (n == 0 || d == 0) return cudaSuccess;
  generate_f32_pdx_kernel<<<148 * 16, 256, 0, s>>>(generator, salt, first_row, (unsigned)n, (unsigned)d, dev_pdx, ld);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_generate_tokens(uint64_t salt, uint64_t first_row, size_t n_rows, size_t dim, float* dev_tokens,
                                   cudaStream_t s, LaunchCounter* launches) {
  if (n_rows == 0 || dim == 0) return cudaSuccess;
  generate_tokens_kernel<<<148 * 16, 256, 0, s>>>(salt, first_row, n_rows, (unsigned)dim, dev_tokens);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace innr
