// exchange.cu -- the cross-GPU step of a row-sharded top-k without a collective library:
// every rank PUBLISHES its local top-k keys into a small mailbox that lives in every peer's memory (peer-mapped over
// NVLink: cudaIpcOpenMemHandle between processes, cudaDeviceEnablePeerAccess inside one process), raises a flag with
// system-scope release semantics, spins on the flags of its own mailbox and MERGES the n_ranks lists -- one launch, no
// NCCL call, no host round trip. Every rank ends up with the same merged lists ("lower global index wins" holds
// across shards because the keys carry global indices).
//
// Mailbox of one rank (device memory of that rank):
//   keys  [2 parities][n_ranks][slot_keys] u64      slot (p, r) is written only by rank r, in calls of parity p
//   flags [2 parities][n_ranks][EX_MAX_CTAS] u64    flag (p, r, b) = number of the last call rank r's CTA b published
// Calls are numbered 1, 2, ... identically on all ranks (same sequence of calls, like any collective). Two parities
// make slot reuse safe: rank r can only publish call e + 2 after it finished call e + 1, which needed every peer's
// call-(e + 1) flag, which a peer raises only after its own call-e kernel (and its reads of slot (e & 1, r)) completed
// -- calls of one rank are ordered on its stream.
#include "common.cuh"
#include "kernels.cuh"

namespace innr {

namespace {

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct ExArgs {
  const uint64_t* local_keys;  // nq x k sorted keys of this rank (sentinel padded)
  uint64_t* const* peers;      // device table: mailbox base of every rank as mapped in this process
  uint64_t* mailbox;           // this rank's own mailbox
  int nq, k, n_ranks, rank;
  size_t slot_keys;
  uint64_t call;               // number of this call (>= 1)
  int publish_only;            // test hook / non-root ranks of a rooted exchange: publish and return
  int descending;
  uint64_t* keys_out;          // any of the three may be null
  uint64_t* idx_out;
  float* score_out;
  uint32_t* dist_out;          // high half of the key as is (Hamming distance)
  unsigned* status;            // set to 1 when a peer's flag did not arrive within timeout_ns
  uint64_t timeout_ns;
};

__device__ __forceinline__ size_t key_off(const ExArgs& a, int parity, int r) {
  return ((size_t)parity * a.n_ranks + r) * a.slot_keys;
}
__device__ __forceinline__ size_t flag_off(const ExArgs& a, int parity, int r, int cta) {
  return 2 * (size_t)a.n_ranks * a.slot_keys + ((size_t)parity * a.n_ranks + r) * EX_MAX_CTAS + cta;
}

// CTA b owns the queries q = b, b + gridDim.x, ... on EVERY rank (the grid is a function of nq only), so it depends on
// the CTAs b of the peers and on nothing else.
template <int R>
__global__ void __launch_bounds__(EX_THREADS) exchange_merge_kernel(ExArgs a) {
  __shared__ int s_timeout;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = EX_THREADS / 32;
  const int parity = (int)(a.call & 1);
  const int my_q = (a.nq - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // queries of this CTA
  if (threadIdx.x == 0) s_timeout = 0;
  // ---- publish: this rank's lists of this CTA's queries into slot (parity, rank) of every mailbox ----
  const int per_peer = my_q * a.k;
  for (int t = threadIdx.x; t < per_peer * a.n_ranks; t += EX_THREADS) {
    const int p = t / per_peer, rem = t - p * per_peer;
    const int q = (int)blockIdx.x + (rem / a.k) * (int)gridDim.x, j = rem % a.k;
    const size_t e = (size_t)q * a.k + j;
    a.peers[p][key_off(a, parity, a.rank) + e] = a.local_keys[e];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.n_ranks)
    st_release_sys(a.peers[threadIdx.x] + flag_off(a, parity, a.rank, (int)blockIdx.x), a.call);
  if (a.publish_only) return;
  // ---- wait for the same CTA of every rank ----
  if (threadIdx.x < a.n_ranks) {
    const uint64_t* f = a.mailbox + flag_off(a, parity, (int)threadIdx.x, (int)blockIdx.x);
    const uint64_t t0 = global_timer_ns();
    while (ld_acquire_sys(f) < a.call) {
      if (global_timer_ns() - t0 > a.timeout_ns) {
        s_timeout = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (s_timeout) {
    if (threadIdx.x == 0) atomicExch(a.status, 1u);
    return;
  }
  // ---- merge: one warp per query ----
  for (int jq = warp; jq < my_q; jq += n_warps) {
    const int q = (int)blockIdx.x + jq * (int)gridDim.x;
    WarpList<R> list;
    list.init();
    for (int r = 0; r < a.n_ranks; ++r)
      list.template merge_from<true>(a.mailbox + key_off(a, parity, r) + (size_t)q * a.k, a.k, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int p = r * 32 + lane;
      if (p < a.k) {
        const uint64_t key = list.v[r];
        const size_t o = (size_t)q * a.k + p;
        if (a.keys_out) a.keys_out[o] = key;
        if (a.idx_out) a.idx_out[o] = key & 0xFFFFFFFFull;
        uint32_t hi = (uint32_t)(key >> 32);
        if (a.dist_out) a.dist_out[o] = hi;
        if (a.score_out) {
          if (a.descending) hi = ~hi;
          a.score_out[o] = __uint_as_float(order_bits_to_f32_bits(hi));
        }
      }
    }
  }
}

}  // namespace

size_t exchange_mailbox_bytes(int n_ranks, size_t slot_keys) {
  return (2 * (size_t)n_ranks * slot_keys + 2 * (size_t)n_ranks * EX_MAX_CTAS) * sizeof(uint64_t);
}

cudaError_t launch_exchange_merge(const ExchangeView& x, const uint64_t* dev_local_keys, size_t nq, size_t k, uint64_t call,
                                  int publish_only, int descending, uint64_t* dev_keys_out, uint64_t* dev_idx,
                                  float* dev_score, uint32_t* dev_dist, cudaStream_t s, LaunchCounter* launches) {
  if (k == 0 || k > 128 || nq == 0 || nq * k > x.slot_keys) return cudaErrorInvalidValue;
  ExArgs a;
  a.local_keys = dev_local_keys;
  a.peers = x.dev_peer_table;
  a.mailbox = x.mailbox;
  a.nq = (int)nq;
  a.k = (int)k;
  a.n_ranks = x.n_ranks;
  a.rank = x.rank;
  a.slot_keys = x.slot_keys;
  a.call = call;
  a.publish_only = publish_only;
  a.descending = descending;
  a.keys_out = dev_keys_out;
  a.idx_out = dev_idx;
  a.score_out = dev_score;
  a.dist_out = dev_dist;
  a.status = x.dev_status;
  a.timeout_ns = x.timeout_ns;
  unsigned grid = (unsigned)((nq + 7) / 8);
  if (grid > EX_MAX_CTAS) grid = EX_MAX_CTAS;
  if (k <= 32) exchange_merge_kernel<1><<<grid, EX_THREADS, 0, s>>>(a);
  else exchange_merge_kernel<4><<<grid, EX_THREADS, 0, s>>>(a);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace innr
