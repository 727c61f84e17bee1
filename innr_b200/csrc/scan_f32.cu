// scan_f32.cu -- HBM-bound scan of a device-resident VerticalBatch (PDX layout) with fused exact top-k.
//
// Replaces (reference, innr 0.6.3):
//   batch_dot_into          src/batch.rs:284-297      batch_l2_squared_into  src/batch.rs:250-266
//   batch_norms_into        src/batch.rs:672-686      batch_cosine_into      src/batch.rs:705-728
//   batch_knn (L2 + TopK)   src/batch.rs:385-411      batch_knn_dot          src/batch.rs:742-764
//   batch_knn_cosine        src/batch.rs:777-800      TopK::insert           src/topk.rs:96-121
//
// Bit-exactness: the reference's batch loops are `acc[i] += q[d] * v[d][i]` with d
// ascending and NO fused multiply-add, so one thread owning vector i and computing
// __fadd_rn(acc, __fmul_rn(q, v)) for d = 0..D-1 reproduces every score bit for bit. The PDX layout makes that
// mapping perfectly coalesced: a thread owns 4 consecutive vectors (one 128-bit load per dimension row), a warp
// reads 512 contiguous bytes per row, a CTA tile covers TILE = 4 * blockDim vectors.
//
// batch_knn_cosine recomputes all corpus norms on every call (src/batch.rs:788); here sum(v*v) is accumulated
// in the same pass over the same registers (same sequential order => same bits) at zero extra HBM bytes.
//
// Selection never materialises N scores: scores become 64-bit composite keys (common.cuh) and flow into a
// per-warp register list, a CTA merge and a last-CTA merge -- one launch per query group.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace innr {

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int VPT = 4;                        // vectors per thread (one float4)
constexpr int TILE = SCAN_THREADS * VPT;      // vectors per CTA tile
constexpr float NORM_EPS = 1e-9f;             // crate::NORM_EPSILON, src/lib.rs:178

struct PdxArgs {
  const float* data;
  unsigned long long ld;
  unsigned n, d, ld4;     // ld4 = ld (u32 copy for bounds checks; ld < 2^32)
  unsigned n_tiles;
  unsigned index_base;
  const float* queries;   // nq_total x d (device); blockIdx.y selects the group of QB queries
  int nq_valid;           // queries in this launch (all y groups together)
  int k;
  uint64_t* partials;
  uint64_t* out_keys;
  uint64_t* group_partials;
  unsigned* tickets;
  float* scores_out;      // scores mode: out[q * ld + i]
  const float* norms_in;  // PDX_COSINE_NORMS
  const uint32_t* mask;   // MASKED: bit i set = vector i passes the predicate (batch_knn_filtered)
  const uint32_t* perm;   // PDX_L2_PERM: the order in which the dimension rows are accumulated (batch_knn_reordered)
  float threshold;        // PDX_L2_PRUNE
  float one;              // 1.0f, opaque to the compiler (see add2_unfusable)
};

template <int MODE>
__device__ __forceinline__ void accumulate(float q, float v, float& acc) {
  if (MODE == PDX_L2 || MODE == PDX_L2_PRUNE || MODE == PDX_L2_PERM) {
    float diff = __fsub_rn(q, v);                 // let diff = q_d - v_d;
    acc = __fadd_rn(acc, __fmul_rn(diff, diff));  // *dist += diff * diff;
  } else {
    acc = __fadd_rn(acc, __fmul_rn(q, v));        // *prod += q_d * v_d;
  }
}

// the same for two vectors at once (acc, v: packed pairs; q2 = {q, q}): unfused mul + add per lane, bit-identical
template <int MODE>
__device__ __forceinline__ void accumulate2(uint64_t q2, uint64_t v2, uint64_t& acc2, uint64_t one2) {
  if (MODE == PDX_L2 || MODE == PDX_L2_PRUNE || MODE == PDX_L2_PERM) {
    const uint64_t diff = sub2_rn(q2, v2);
    acc2 = add2_unfusable(mul2_rn(diff, diff), acc2, one2);
  } else {
    acc2 = add2_unfusable(mul2_rn(q2, v2), acc2, one2);
  }
}

template <int MODE, int QB, int R, bool KNN, bool MASKED = false, int UX = 0>
__global__ void __launch_bounds__(SCAN_THREADS) pdx_scan_kernel(const PdxArgs a_in) {
  static_assert(MODE != PDX_L2_PRUNE || (QB == 1 && !KNN), "pruning is a single-query scores scan");
  static_assert(MODE != PDX_L2_PERM || (QB == 1 && !MASKED), "the reordered scan is a single-query scan");
  constexpr bool PERM = (MODE == PDX_L2_PERM);
  constexpr bool NEED_SS = (MODE == PDX_COSINE_FUSED || MODE == PDX_NORMS);
  constexpr bool NEED_DOT = (MODE != PDX_NORMS);
  // dimension rows in flight per thread. The sum over d is sequential by contract, so a thread makes D / U round trips
  // to memory: 8 (4 with eight queries' accumulators) saturate HBM when every SM holds several CTAs; launches with at
  // most a couple of CTAs per SM (small corpora, small shards, C1) are bound by that latency chain instead and use the
  // deep variants (UX = 32 / 16).
  constexpr int U = UX ? UX : ((QB == 1) ? 8 : 4);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [d][QB] interleaved queries (padded so the d-loop can read whole float4 groups), then qnorm[QB], then keys
  const unsigned d_pad = (a_in.d + 31u) / 32u * 32u;  // same for every U (scan_smem_bytes)
  float* sq = reinterpret_cast<float*>(smem_raw);
  float* s_qn = sq + (size_t)d_pad * QB;
  unsigned* s_perm = reinterpret_cast<unsigned*>(s_qn + ((QB + 3) & ~3));  // PERM: d_pad row indices (d_pad % 8 == 0)
  uint64_t* smem_keys = reinterpret_cast<uint64_t*>(s_perm + (PERM ? d_pad : 0u));

  const int lane = threadIdx.x & 31;

  // query group of this CTA (grid.y): shifts the query, output and merge-workspace pointers
  PdxArgs a = a_in;
  {
    const unsigned y = blockIdx.y;
    const unsigned n_groups = (gridDim.x + FINISH_GROUP - 1) / FINISH_GROUP;
    a.queries += (size_t)y * QB * a.d;
    a.nq_valid = min(QB, a_in.nq_valid - (int)(y * QB));
    a.out_keys += (size_t)y * QB * a.k;
    a.partials += (size_t)y * gridDim.x * QB * a.k;
    a.group_partials += (size_t)y * n_groups * QB * a.k;
    a.tickets += (size_t)y * (1 + n_groups);
  }

  if (NEED_DOT) {
    for (unsigned idx = threadIdx.x; idx < d_pad * QB; idx += blockDim.x) {
      unsigned dd = idx / QB, q = idx % QB;
      if (PERM) {  // step j of the reordered scan pairs query[order[j]] with row order[j]
        const unsigned row = dd < a.d ? a.perm[dd] : 0u;
        s_perm[dd] = row;
        sq[idx] = dd < a.d ? a.queries[row] : 0.0f;
      } else {
        sq[idx] = (dd < a.d && (int)q < a.nq_valid) ? a.queries[(size_t)q * a.d + dd] : 0.0f;
      }
    }
  }
  __syncthreads();
  constexpr bool NEED_QN = NEED_DOT && (MODE == PDX_COSINE_FUSED || MODE == PDX_COSINE_NORMS);
  if (NEED_QN) {
    // query_norm = query.iter().map(|x| x * x).sum::<f32>().sqrt()   (src/batch.rs:714) -- sequential, so one thread per
    // query; read from the shared-memory copy (a dependent chain of global loads cost ~20 us per launch). The chain is
    // d dependent multiply-adds (~3 us at d = 768) and its result is only needed in the first tile's epilogue, so the
    // CTA does NOT wait for it here: the last warp computes it and joins the scan late, the other warps start streaming
    // at once, and the barrier in front of the first epilogue (below) finds it long finished.
    if (threadIdx.x >= SCAN_THREADS - 32 && threadIdx.x - (SCAN_THREADS - 32) < QB) {
      const unsigned q = threadIdx.x - (SCAN_THREADS - 32);
      float ss = 0.0f;
      if ((int)q < a.nq_valid) {
        for (unsigned dd = 0; dd < a.d; ++dd) {
          const float x = sq[(size_t)dd * QB + q];
          ss = __fadd_rn(ss, __fmul_rn(x, x));
        }
      }
      s_qn[q] = __fsqrt_rn(ss);
    }
  }
  bool qn_visible = !NEED_QN;

  WarpList<R> lists[QB];
  uint64_t thrs[QB];
  if (KNN) {
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      lists[q].init();
      thrs[q] = KEY_SENTINEL;
    }
  }

  for (unsigned tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const unsigned i0 = tile * TILE + threadIdx.x * VPT;
    bool active = i0 < a.ld4;  // ld is a multiple of 4: the whole float4 is in bounds
    unsigned nib = 0xFu;       // MASKED: predicate bits of this thread's four vectors (i0 % 4 == 0: one word)
    if (MASKED) {
      nib = active ? (a.mask[i0 >> 5] >> (i0 & 31)) & 0xFu : 0u;
      active = nib != 0;       // predicate pushdown: rows of rejected vectors are not even read
    }
    // accumulators as packed pairs (vectors 0,1 and 2,3 of this thread): the unfused multiply and add of the reference
    // run as FMUL2 / FADD2 -- two IEEE operations per instruction, same bits, half the FP32 issue slots (what bounds
    // the multi-query kernel)
    uint64_t acc2[QB][2];
    uint64_t ss2[2];
    float mx[VPT];   // PDX_L2_PRUNE: running max of the non-NaN partial distances
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ss2[h] = 0ull;  // {+0.0f, +0.0f}
#pragma unroll
      for (int q = 0; q < QB; ++q) acc2[q][h] = 0ull;
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) mx[j] = -INFINITY;  // the initial 0.0 is not a partial the reference tests
    const unsigned amask = (MODE == PDX_L2_PRUNE) ? __ballot_sync(FULL_MASK, active) : 0u;
    const uint64_t one2 = pack2(a.one, a.one);
    auto step = [&](const float4& v, unsigned dq) {
      const uint64_t v01 = pack2(v.x, v.y), v23 = pack2(v.z, v.w);
      if (NEED_SS) {
        ss2[0] = add2_unfusable(mul2_rn(v01, v01), ss2[0], one2);
        ss2[1] = add2_unfusable(mul2_rn(v23, v23), ss2[1], one2);
      }
      if (NEED_DOT) {
#pragma unroll
        for (int q = 0; q < QB; ++q) {
          const float qv = sq[(size_t)dq * QB + q];
          const uint64_t q2 = pack2(qv, qv);
          accumulate2<MODE>(q2, v01, acc2[q][0], one2);
          accumulate2<MODE>(q2, v23, acc2[q][1], one2);
        }
      }
      if (MODE == PDX_L2_PRUNE) {
        // the reference prunes vector i at the first dimension whose partial distance exceeds the threshold
        // (src/batch.rs:347-351). Partial sums of squares never decrease until one becomes NaN (a NaN compares
        // false and stays alive), so "some partial > threshold" == "max of the non-NaN partials > threshold".
        float p0, p1, p2, p3;
        unpack2(acc2[0][0], p0, p1);
        unpack2(acc2[0][1], p2, p3);
        mx[0] = fmaxf(mx[0], p0);
        mx[1] = fmaxf(mx[1], p1);
        mx[2] = fmaxf(mx[2], p2);
        mx[3] = fmaxf(mx[3], p3);
      }
    };
    if (active) {
      const float* p = a.data + i0;
      unsigned dd = 0;
      for (; dd + U <= a.d; dd += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream_f4(PERM ? p + (size_t)s_perm[dd + u] * a.ld : p + (size_t)u * a.ld);
        if (!PERM) p += (size_t)U * a.ld;
#pragma unroll
        for (int u = 0; u < U; ++u) step(v[u], dd + u);
        if (MODE == PDX_L2_PRUNE) {  // every vector this warp owns in the tile is pruned: stop reading its rows
          const bool dead = mx[0] > a.threshold && mx[1] > a.threshold && mx[2] > a.threshold && mx[3] > a.threshold;
          if (__all_sync(amask, dead)) { dd = a.d; break; }
        }
      }
      for (; dd < a.d; ++dd) {  // D % U tail
        const float4 v = ldg_stream_f4(PERM ? p + (size_t)s_perm[dd] * a.ld : p);
        if (!PERM) p += a.ld;
        step(v, dd);
      }
    }
    float acc[QB][VPT];
    float ss[VPT];
    unpack2(ss2[0], ss[0], ss[1]);
    unpack2(ss2[1], ss[2], ss[3]);
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      unpack2(acc2[q][0], acc[q][0], acc[q][1]);
      unpack2(acc2[q][1], acc[q][2], acc[q][3]);
    }
    if (MODE == PDX_L2_PRUNE) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) ss[j] = mx[j];
    }

    // epilogue: scores (and keys)
    if (!qn_visible) {  // first tile of this CTA: the query norm written by the last warp (CTA-uniform branch)
      __syncthreads();
      qn_visible = true;
    }
    float nrm[VPT];
    if (MODE == PDX_COSINE_FUSED || MODE == PDX_NORMS) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) nrm[j] = __fsqrt_rn(ss[j]);  // *norm = norm.sqrt()  (src/batch.rs:683-685)
    }
    if (MODE == PDX_COSINE_NORMS) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) nrm[j] = (active && i0 + j < a.n) ? a.norms_in[i0 + j] : 0.0f;
    }
    auto score = [&](int q, int j) -> float {
      if (MODE == PDX_COSINE_FUSED || MODE == PDX_COSINE_NORMS) {
        const float qn = s_qn[q];
        // src/batch.rs:716-727: qn < eps -> 0 for all; norm > eps ? dot / (qn * norm) : 0
        float c = 0.0f;
        if (!(qn < NORM_EPS) && nrm[j] > NORM_EPS) c = __fdiv_rn(acc[q][j], __fmul_rn(qn, nrm[j]));
        return c;
      } else if (MODE == PDX_NORMS) {
        return nrm[j];
      } else if (MODE == PDX_L2_PRUNE) {
        return ss[j] > a.threshold ? -1.0f : acc[q][j];  // distances are never negative: -1.0 marks "pruned"
      }
      return acc[q][j];
    };
    auto make_key = [&](float sc, unsigned i) -> uint64_t {
      return (MODE == PDX_L2 || MODE == PDX_L2_PERM) ? make_key_asc(sc, a.index_base + i) : make_key_desc(sc, a.index_base + i);
    };
    if (KNN && QB > 1) {
      // vector j of this thread against all QB queries at once: the QB lists are updated in lock step (offer_multi)
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const unsigned i = i0 + j;
        uint64_t key[QB];
#pragma unroll
        for (int q = 0; q < QB; ++q) key[q] = make_key(score(q, j), i);
        offer_multi<R, QB>(lists, key, active && i < a.n && (!MASKED || ((nib >> j) & 1u)), thrs, a.nq_valid, a.k, lane);
      }
    } else {
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        if (KNN) {
          if (q < a.nq_valid) {
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
              const unsigned i = i0 + j;
              lists[q].offer(make_key(score(q, j), i), active && i < a.n && (!MASKED || ((nib >> j) & 1u)), thrs[q], a.k, lane);
            }
          }
        } else if (active && q < a.nq_valid) {
          *reinterpret_cast<float4*>(a.scores_out + (size_t)q * a.ld + i0) =
              make_float4(score(q, 0), score(q, 1), score(q, 2), score(q, 3));
        }
      }
    }
  }

  if (KNN) block_finish<R, QB>(lists, a.nq_valid, a.k, smem_keys, a.partials, a.group_partials, a.out_keys, a.tickets);
}

template <int MODE, int QB, int R, bool KNN, bool MASKED = false, int UX = 0>
cudaError_t launch_one(const PdxArgs& a, size_t smem, int ny, int num_sms, cudaStream_t s) {
  auto kern = pdx_scan_kernel<MODE, QB, R, KNN, MASKED, UX>;
  static size_t smem_set_dev[16] = {};
  size_t& smem_set = smem_set_dev[current_device_slot()];
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  int occ = 0;
  cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SCAN_THREADS, smem);
  if (eo != cudaSuccess) return eo;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  unsigned grid = balanced_grid(a.n_tiles, (unsigned)occ * (unsigned)num_sms);
  if (!KNN) grid = a.n_tiles ? a.n_tiles : 1;  // no cross-tile state: one CTA per tile, hardware scheduler balances
  kern<<<dim3(grid, ny > 0 ? ny : 1), SCAN_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}

// grid.x the KNN launch will use (needed to size the merge workspace per query group)
template <int MODE, int QB, int R, int UX = 0>
unsigned knn_grid_x(unsigned n_tiles, size_t smem, int num_sms) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pdx_scan_kernel<MODE, QB, R, true, false, UX>, SCAN_THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
  return balanced_grid(n_tiles, (unsigned)occ * (unsigned)num_sms);
}

size_t scan_smem_bytes(size_t d, int qb, int k, bool knn, bool perm = false) {
  size_t d_pad = (d + 31) / 32 * 32;
  size_t b = d_pad * qb * sizeof(float) + ((qb + 3) & ~3) * sizeof(float);
  if (perm) b += d_pad * sizeof(unsigned);
  if (knn) b += (size_t)(SCAN_THREADS / 32) * k * sizeof(uint64_t) * qb;  // QB > 1: every warp parks QB lists (block_finish)
  return b;
}

}  // namespace

unsigned pdx_max_grid(int num_sms) { return (unsigned)num_sms * 8u; }

namespace {
// the arguments every scan launch shares. Columns actually scanned: n rounded up to a whole float4 (the row pitch is a
// multiple of 4, so the last float4 is in bounds); a prefix view (n << ld) scans only its own columns.
PdxArgs base_args(const PdxView& v, const Workspace& ws, size_t k) {
  PdxArgs a{};
  a.data = v.data;
  a.ld = v.ld;
  a.ld4 = (unsigned)std::min<size_t>(v.ld, (v.n + 3) / 4 * 4);
  a.n = (unsigned)v.n;
  a.d = (unsigned)v.d;
  a.n_tiles = (unsigned)(((size_t)a.ld4 + TILE - 1) / TILE);
  a.index_base = v.index_base;
  a.one = 1.0f;
  a.k = (int)k;
  a.partials = ws.partials;
  a.group_partials = ws.group_partials;
  a.tickets = ws.tickets;
  a.nq_valid = 1;
  return a;
}
}  // namespace

cudaError_t launch_pdx_knn(const PdxView& v, int mode, const float* dev_queries, size_t nq, size_t k,
                           uint64_t* dev_keys, Workspace& ws, cudaStream_t s, LaunchCounter* launches) {
  PdxArgs a = base_args(v, ws, k);
  const bool big_k = k > 32;
  // query blocking: 8 queries share one pass over the corpus when their lists fit in registers; several groups of 8
  // run as grid.y of ONE launch as long as their merge workspaces fit (small corpora / many queries: C1, sample pass)
  size_t done = 0;
  while (done < nq) {
    int qb = (nq - done >= 2 && !big_k) ? 8 : 1;
    size_t smem = scan_smem_bytes(v.d, qb, (int)k, true);
    if (smem > 200 * 1024 && qb == 8) {
      qb = 1;
      smem = scan_smem_bytes(v.d, 1, (int)k, true);
    }
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    int ny = 1;
    // at most ~2 CTAs per SM in the whole launch: the per-thread latency chain binds, use the deep-prefetch variants
    const size_t groups_left8 = (nq - done + 7) / 8;
    const bool deep = !big_k && (size_t)a.n_tiles * (qb == 8 ? groups_left8 : 1) <= 2 * (size_t)ws.num_sms;
    // not even one CTA per SM with eight queries per CTA (C1): four per CTA = twice the warps to hide latency with
    if (qb == 8 && deep && (size_t)a.n_tiles * groups_left8 <= (size_t)ws.num_sms) {
      qb = 4;
      smem = scan_smem_bytes(v.d, 4, (int)k, true);
    }
    if (qb > 1) {
      unsigned gx = 1;
      if (qb == 4) {
        if (mode == PDX_DOT) gx = knn_grid_x<PDX_DOT, 4, 1, 16>(a.n_tiles, smem, ws.num_sms);
        else if (mode == PDX_L2) gx = knn_grid_x<PDX_L2, 4, 1, 16>(a.n_tiles, smem, ws.num_sms);
        else gx = knn_grid_x<PDX_COSINE_FUSED, 4, 1, 16>(a.n_tiles, smem, ws.num_sms);
      } else if (deep) {
        if (mode == PDX_DOT) gx = knn_grid_x<PDX_DOT, 8, 1, 16>(a.n_tiles, smem, ws.num_sms);
        else if (mode == PDX_L2) gx = knn_grid_x<PDX_L2, 8, 1, 16>(a.n_tiles, smem, ws.num_sms);
        else gx = knn_grid_x<PDX_COSINE_FUSED, 8, 1, 16>(a.n_tiles, smem, ws.num_sms);
      } else if (mode == PDX_DOT) gx = knn_grid_x<PDX_DOT, 8, 1>(a.n_tiles, smem, ws.num_sms);
      else if (mode == PDX_L2) gx = knn_grid_x<PDX_L2, 8, 1>(a.n_tiles, smem, ws.num_sms);
      else gx = knn_grid_x<PDX_COSINE_FUSED, 8, 1>(a.n_tiles, smem, ws.num_sms);
      const size_t n_groups = (gx + FINISH_GROUP - 1) / FINISH_GROUP;
      size_t fit = ws.partials_cap / ((size_t)gx * qb * k);
      fit = std::min(fit, ws.group_cap / (n_groups * qb * k));
      fit = std::min(fit, ws.tickets_cap / (1 + n_groups));
      const size_t groups_left = (nq - done + qb - 1) / qb;
      ny = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(fit, groups_left), 65535));
    }
    const size_t nq_launch = std::min<size_t>(nq - done, (size_t)ny * qb);
    a.queries = dev_queries + done * v.d;
    a.nq_valid = (int)nq_launch;
    a.out_keys = dev_keys + done * k;
    cudaError_t e;
#define INNR_DISPATCH(MODE)                                                                          \
  if (qb == 4) e = launch_one<MODE, 4, 1, true, false, 16>(a, smem, ny, ws.num_sms, s);              \
  else if (qb == 8 && deep) e = launch_one<MODE, 8, 1, true, false, 16>(a, smem, ny, ws.num_sms, s); \
  else if (qb == 8) e = launch_one<MODE, 8, 1, true>(a, smem, ny, ws.num_sms, s);                    \
  else if (deep) e = launch_one<MODE, 1, 1, true, false, 32>(a, smem, 1, ws.num_sms, s);             \
  else if (!big_k) e = launch_one<MODE, 1, 1, true>(a, smem, 1, ws.num_sms, s);                      \
  else e = launch_one<MODE, 1, 4, true>(a, smem, 1, ws.num_sms, s);
    if (mode == PDX_DOT) { INNR_DISPATCH(PDX_DOT) }
    else if (mode == PDX_L2) { INNR_DISPATCH(PDX_L2) }
    else if (mode == PDX_COSINE_FUSED) { INNR_DISPATCH(PDX_COSINE_FUSED) }
    else return cudaErrorInvalidValue;
#undef INNR_DISPATCH
    if (e != cudaSuccess) return e;
    ++*launches;
    done += nq_launch;
  }
  return cudaSuccess;
}

cudaError_t launch_pdx_knn_filtered(const PdxView& v, const float* dev_query, const uint32_t* dev_mask, size_t k,
                                    uint64_t* dev_keys, Workspace& ws, cudaStream_t s, LaunchCounter* launches) {
  PdxArgs a = base_args(v, ws, k);
  a.queries = dev_query;
  a.out_keys = dev_keys;
  a.mask = dev_mask;
  const size_t smem = scan_smem_bytes(v.d, 1, (int)k, true);
  if (smem > 227 * 1024 || k > 128) return cudaErrorInvalidValue;
  cudaError_t e = (k <= 32) ? launch_one<PDX_L2, 1, 1, true, true>(a, smem, 1, ws.num_sms, s)
                            : launch_one<PDX_L2, 1, 4, true, true>(a, smem, 1, ws.num_sms, s);
  if (e == cudaSuccess) ++*launches;
  return e;
}

cudaError_t launch_pdx_knn_reordered(const PdxView& v, const float* dev_query, const uint32_t* dev_perm, size_t k,
                                     uint64_t* dev_keys, Workspace& ws, cudaStream_t s, LaunchCounter* launches) {
  PdxArgs a = base_args(v, ws, k);
  a.queries = dev_query;
  a.out_keys = dev_keys;
  a.perm = dev_perm;
  const size_t smem = scan_smem_bytes(v.d, 1, (int)k, true, true);
  if (smem > 227 * 1024 || k > 128 || !dev_perm) return cudaErrorInvalidValue;
  cudaError_t e = (k <= 32) ? launch_one<PDX_L2_PERM, 1, 1, true>(a, smem, 1, ws.num_sms, s)
                            : launch_one<PDX_L2_PERM, 1, 4, true>(a, smem, 1, ws.num_sms, s);
  if (e == cudaSuccess) ++*launches;
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// batch_dimension_variance (src/batch.rs:572-592). Both sums of a dimension row are sequential f32 sums over all n
// vectors, so a row is one chain of n dependent adds whatever the hardware: one warp per row, rows in parallel (d warps);
// the chain, not HBM, bounds it (~4 cycles per vector per pass).
// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr int VAR_WARPS = 4;

constexpr int VAR_UV = 4;  // 128-value groups per step (512 values: 2 KB of shared memory per warp)

// One sequential f32 sum over a dimension row: sum of x (CENTERED = false) or of (x - mean) * (x - mean), unfused.
// The warp loads 512 consecutive values per step (the next step's loads are issued before this step's chain runs),
// parks them in shared memory, and every lane replays the chain from broadcast 128-bit reads.
template <bool CENTERED>
__device__ __forceinline__ float row_chain(const float* row, unsigned n, float mean, int lane, float4* buf) {
  float acc = 0.0f;
  constexpr unsigned STEP = 128u * VAR_UV;
  const unsigned n_steps = n / STEP;
  float4 cur[VAR_UV], nxt[VAR_UV];
  if (n_steps) {
#pragma unroll
    for (int u = 0; u < VAR_UV; ++u) cur[u] = ldg_stream_f4(row + 128u * u + 4u * lane);
  }
  for (unsigned st = 0; st < n_steps; ++st) {
    if (st + 1 < n_steps) {
      const float* p = row + (size_t)(st + 1) * STEP + 4u * lane;
#pragma unroll
      for (int u = 0; u < VAR_UV; ++u) nxt[u] = ldg_stream_f4(p + 128u * u);
    }
#pragma unroll
    for (int u = 0; u < VAR_UV; ++u) {
      float4 v = cur[u];
      if (CENTERED) {  // computed once per value by the lane that loaded it
        v.x = __fsub_rn(v.x, mean); v.y = __fsub_rn(v.y, mean); v.z = __fsub_rn(v.z, mean); v.w = __fsub_rn(v.w, mean);
        v.x = __fmul_rn(v.x, v.x); v.y = __fmul_rn(v.y, v.y); v.z = __fmul_rn(v.z, v.z); v.w = __fmul_rn(v.w, v.w);
      }
      buf[u * 32 + lane] = v;
    }
    __syncwarp();
#pragma unroll 16
    for (int j = 0; j < VAR_UV * 32; ++j) {
      const float4 x = buf[j];
      acc = __fadd_rn(acc, x.x);
      acc = __fadd_rn(acc, x.y);
      acc = __fadd_rn(acc, x.z);
      acc = __fadd_rn(acc, x.w);
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < VAR_UV; ++u) cur[u] = nxt[u];
  }
#pragma unroll 8
  for (unsigned i = n_steps * STEP; i < n; ++i) {  // tail (< 512 values): every lane reads the same value
    float x = __ldg(row + i);
    if (CENTERED) { x = __fsub_rn(x, mean); x = __fmul_rn(x, x); }
    acc = __fadd_rn(acc, x);
  }
  return acc;
}

__global__ void __launch_bounds__(VAR_WARPS * 32) dimension_variance_kernel(const float* __restrict__ data, size_t ld,
                                                                            unsigned n, unsigned d, float* __restrict__ out) {
  __shared__ float4 s_buf[VAR_WARPS][VAR_UV * 32];
  const unsigned dd = blockIdx.x * VAR_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (dd >= d) return;
  float4* buf = s_buf[threadIdx.x >> 5];
  const float* row = data + (size_t)dd * ld;
  const float nf = (float)n;                                     // let n = batch.num_vectors as f32;
  const float mean = __fdiv_rn(row_chain<false>(row, n, 0.0f, lane, buf), nf);
  const float var = __fdiv_rn(row_chain<true>(row, n, mean, lane, buf), nf);
  if (lane == 0) out[dd] = var;
}

}  // namespace

cudaError_t launch_dimension_variance(const PdxView& v, float* dev_out, cudaStream_t s, LaunchCounter* launches) {
  if (v.d == 0) return cudaSuccess;
  if (v.n <= 1) return cudaMemsetAsync(dev_out, 0, v.d * sizeof(float), s);  // src/batch.rs:573-575
  dimension_variance_kernel<<<(unsigned)((v.d + VAR_WARPS - 1) / VAR_WARPS), VAR_WARPS * 32, 0, s>>>(v.data, v.ld, (unsigned)v.n,
                                                                                                  (unsigned)v.d, dev_out);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_pdx_scores(const PdxView& v, int mode, const float* dev_query, const float* dev_norms,
                              float* dev_out, Workspace& ws, cudaStream_t s, LaunchCounter* launches, float threshold,
                              const uint32_t* dev_perm) {
  PdxArgs a = base_args(v, ws, 0);
  a.perm = dev_perm;
  a.queries = dev_query;
  a.scores_out = dev_out;
  a.norms_in = dev_norms;
  a.threshold = threshold;
  size_t smem = scan_smem_bytes(v.d, 1, 0, false, mode == PDX_L2_PERM);
  if (smem > 227 * 1024 || (mode == PDX_L2_PERM && !dev_perm)) return cudaErrorInvalidValue;
  cudaError_t e;
  // few tiles (<= 2 per SM): the D / U round trips of each thread bind, not HBM -> deep-prefetch instantiations
  const bool deep = (size_t)a.n_tiles <= 2 * (size_t)ws.num_sms;
  if (deep && (mode == PDX_DOT || mode == PDX_L2 || mode == PDX_NORMS || mode == PDX_COSINE_NORMS || mode == PDX_COSINE_FUSED)) {
    switch (mode) {
      case PDX_DOT: e = launch_one<PDX_DOT, 1, 1, false, false, 32>(a, smem, 1, ws.num_sms, s); break;
      case PDX_L2: e = launch_one<PDX_L2, 1, 1, false, false, 32>(a, smem, 1, ws.num_sms, s); break;
      case PDX_NORMS: e = launch_one<PDX_NORMS, 1, 1, false, false, 32>(a, smem, 1, ws.num_sms, s); break;
      case PDX_COSINE_NORMS: e = launch_one<PDX_COSINE_NORMS, 1, 1, false, false, 32>(a, smem, 1, ws.num_sms, s); break;
      default: e = launch_one<PDX_COSINE_FUSED, 1, 1, false, false, 32>(a, smem, 1, ws.num_sms, s); break;
    }
    if (e == cudaSuccess) ++*launches;
    return e;
  }
  switch (mode) {
    case PDX_L2_PERM: e = launch_one<PDX_L2_PERM, 1, 1, false>(a, smem, 1, ws.num_sms, s); break;
    case PDX_DOT: e = launch_one<PDX_DOT, 1, 1, false>(a, smem, 1, ws.num_sms, s); break;
    case PDX_L2: e = launch_one<PDX_L2, 1, 1, false>(a, smem, 1, ws.num_sms, s); break;
    case PDX_NORMS: e = launch_one<PDX_NORMS, 1, 1, false>(a, smem, 1, ws.num_sms, s); break;
    case PDX_COSINE_NORMS: e = launch_one<PDX_COSINE_NORMS, 1, 1, false>(a, smem, 1, ws.num_sms, s); break;
    case PDX_L2_PRUNE: e = launch_one<PDX_L2_PRUNE, 1, 1, false>(a, smem, 1, ws.num_sms, s); break;
    case PDX_COSINE_FUSED: e = launch_one<PDX_COSINE_FUSED, 1, 1, false>(a, smem, 1, ws.num_sms, s); break;
    default: return cudaErrorInvalidValue;
  }
  if (e == cudaSuccess) ++*launches;
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// merge of key lists (K10: after the allgather of per-shard top-k) and the TopK analogue
// ---------------------------------------------------------------------------------------------------------
namespace {

template <int R>
__global__ void __launch_bounds__(32) merge_keys_kernel(const uint64_t* in, int n_lists, int nq, int k,
                                                        int descending, uint64_t* keys_out, uint64_t* idx_out,
                                                        float* score_out) {
  const int q = blockIdx.x, lane = threadIdx.x;
  WarpList<R> list;
  list.init();
  warp_merge_lists<R, true>(list, in + (size_t)q * k, 0, 1, n_lists, (size_t)nq * k, k, lane);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int p = r * 32 + lane;
    if (p < k) {
      uint64_t key = list.v[r];
      if (keys_out) keys_out[(size_t)q * k + p] = key;
      if (idx_out) idx_out[(size_t)q * k + p] = key & 0xFFFFFFFFull;
      if (score_out) {
        uint32_t hi = (uint32_t)(key >> 32);
        if (descending) hi = ~hi;
        score_out[(size_t)q * k + p] = __uint_as_float(order_bits_to_f32_bits(hi));
      }
    }
  }
}

// Selection over a device score vector. KIND 0: f32 ascending (TopK analogue / L2), 1: f32 descending (dot, cosine,
// u8), 2: u32 ascending (Hamming). Keys carry index_base + i. `floor_key` (device, may be null) makes the launch one
// ROUND of a larger selection: only keys strictly greater than *floor_key take part, so consecutive rounds of <= 128
// keys each, every one starting above the last key of the round before, enumerate the k smallest keys for any k
// (keys are unique: the index is their low half).
template <int R, int KIND>
__global__ void __launch_bounds__(SCAN_THREADS) topk_scores_kernel(const void* scores, unsigned n, unsigned index_base,
                                                                  const uint32_t* ids,  // null: id = index_base + i
                                                                  const uint32_t* mask, // non-null: only entries whose bit is set
                                                                  unsigned seg_len, unsigned seg_stride,  // KIND 3 (see below)
                                                                  int k, const uint64_t* floor_key, uint64_t* partials,
                                                                  uint64_t* group_partials, uint64_t* out_keys,
                                                                  unsigned* tickets) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* smem_keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int lane = threadIdx.x & 31;
  WarpList<R> lists[1];
  uint64_t thrs[1];
  lists[0].init();
  thrs[0] = KEY_SENTINEL;
  const bool has_floor = floor_key != nullptr;
  const uint64_t floor = has_floor ? *floor_key : 0ull;
  // grid-stride over whole warps so every lane reaches the ballots together
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned n_round = (n + 31u) / 32u * 32u;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    bool valid = i < n && (!mask || ((mask[i >> 5] >> (i & 31)) & 1u));
    uint64_t key = KEY_SENTINEL;
    if (valid) {
      const unsigned id = ids ? ids[i] : index_base + i;
      if (KIND == 3) key = static_cast<const uint64_t*>(scores)[(size_t)(i / seg_len) * seg_stride + (i % seg_len)];
      else if (KIND == 4) key = make_key_u32(~static_cast<const uint32_t*>(scores)[i], id);
      else if (KIND == 2) key = make_key_u32(static_cast<const uint32_t*>(scores)[i], id);
      else if (KIND == 1) key = make_key_desc(static_cast<const float*>(scores)[i], id);
      else key = make_key_asc(static_cast<const float*>(scores)[i], id);
      valid = (KIND != 3 || key != KEY_SENTINEL) && (!has_floor || (key > floor && floor != KEY_SENTINEL));
    }
    lists[0].offer(key, valid, thrs[0], k, lane);
  }
  block_finish<R, 1>(lists, 1, k, smem_keys, partials, group_partials, out_keys, tickets);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// compaction of a pruned distance vector: survivors (entries != -1.0) in ascending index order, the order of the
// reference's `alive.iter().enumerate().filter(..).collect()` (src/batch.rs:358-364)
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr int CP_THREADS = 256, CP_ITEMS = 4, CP_BLOCK = CP_THREADS * CP_ITEMS;

__device__ __forceinline__ bool survivor(float d) { return !(d == -1.0f); }  // NaN distances survive (never pruned)

__global__ void __launch_bounds__(CP_THREADS) compact_count_kernel(const float* __restrict__ dist, unsigned n,
                                                                   unsigned* __restrict__ block_counts) {
  const unsigned base = blockIdx.x * CP_BLOCK;
  unsigned c = 0;
#pragma unroll
  for (int it = 0; it < CP_ITEMS; ++it) {
    const unsigned i = base + it * CP_THREADS + threadIdx.x;
    c += __popc(__ballot_sync(FULL_MASK, i < n && survivor(dist[i])));
  }
  __shared__ unsigned s_c[CP_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;  // every lane of the warp holds the warp total
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w) t += s_c[w];
    block_counts[blockIdx.x] = t;
  }
}

// in place: counts -> exclusive prefix; offsets[n_blocks] = total. One CTA (n_blocks <= 2^32 / 1024).
__global__ void __launch_bounds__(1024) compact_scan_kernel(unsigned* __restrict__ offsets, unsigned n_blocks) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned b0 = 0; b0 < n_blocks; b0 += 1024) {
    const unsigned b = b0 + threadIdx.x;
    const unsigned v = b < n_blocks ? offsets[b] : 0u;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(FULL_MASK, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned w = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(FULL_MASK, w, o);
        if (lane >= o) w += y;
      }
      s_w[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned carry = s_carry;
    const unsigned incl = carry + (warp ? s_w[warp - 1] : 0u) + x;
    if (b < n_blocks) offsets[b] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n_blocks] = s_carry;
}

__global__ void __launch_bounds__(CP_THREADS) compact_scatter_kernel(const float* __restrict__ dist, unsigned n,
                                                                     unsigned long long index_base,
                                                                     const unsigned* __restrict__ block_offsets,
                                                                     uint64_t* __restrict__ out_idx,
                                                                     float* __restrict__ out_dist) {
  const unsigned base = blockIdx.x * CP_BLOCK;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ unsigned s_c[CP_ITEMS][CP_THREADS / 32];
  float d[CP_ITEMS];
  unsigned bal[CP_ITEMS];
#pragma unroll
  for (int it = 0; it < CP_ITEMS; ++it) {
    const unsigned i = base + it * CP_THREADS + threadIdx.x;
    d[it] = i < n ? dist[i] : -1.0f;
    bal[it] = __ballot_sync(FULL_MASK, survivor(d[it]));
    if (lane == 0) s_c[it][warp] = __popc(bal[it]);
  }
  __syncthreads();
  // position of (it, warp) inside the block: items are laid out it-major (index = base + it*256 + tid)
  unsigned before = block_offsets[blockIdx.x];
#pragma unroll
  for (int it = 0; it < CP_ITEMS; ++it) {
    unsigned pre = 0;
    for (int w = 0; w < warp; ++w) pre += s_c[it][w];
    if (survivor(d[it])) {
      const unsigned pos = before + pre + __popc(bal[it] & ((1u << lane) - 1u));
      out_idx[pos] = index_base + base + it * CP_THREADS + threadIdx.x;
      out_dist[pos] = d[it];
    }
    unsigned tot = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w) tot += s_c[it][w];
    before += tot;
  }
}
}  // namespace

size_t compact_blocks(size_t n) { return (n + CP_BLOCK - 1) / CP_BLOCK; }

cudaError_t launch_compact_count(const float* dev_dist, size_t n, unsigned* dev_block_offsets, cudaStream_t s,
                                 LaunchCounter* launches) {
  const unsigned nb = (unsigned)compact_blocks(n);
  if (nb == 0) return cudaMemsetAsync(dev_block_offsets, 0, sizeof(unsigned), s);
  compact_count_kernel<<<nb, CP_THREADS, 0, s>>>(dev_dist, (unsigned)n, dev_block_offsets);
  compact_scan_kernel<<<1, 1024, 0, s>>>(dev_block_offsets, nb);
  *launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_compact_scatter(const float* dev_dist, size_t n, uint64_t index_base, const unsigned* dev_block_offsets,
                                   uint64_t* dev_idx, float* dev_out, cudaStream_t s, LaunchCounter* launches) {
  const unsigned nb = (unsigned)compact_blocks(n);
  if (nb == 0) return cudaSuccess;
  compact_scatter_kernel<<<nb, CP_THREADS, 0, s>>>(dev_dist, (unsigned)n, index_base, dev_block_offsets, dev_idx, dev_out);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_merge_keys(const uint64_t* dev_in, size_t n_lists, size_t nq, size_t k, int descending,
                              uint64_t* dev_keys_out, uint64_t* dev_idx, float* dev_score, cudaStream_t s,
                              LaunchCounter* launches) {
  if (k > 128) return cudaErrorInvalidValue;
  if (k <= 32)
    merge_keys_kernel<1><<<(unsigned)nq, 32, 0, s>>>(dev_in, (int)n_lists, (int)nq, (int)k, descending,
                                                      dev_keys_out, dev_idx, dev_score);
  else
    merge_keys_kernel<4><<<(unsigned)nq, 32, 0, s>>>(dev_in, (int)n_lists, (int)nq, (int)k, descending,
                                                      dev_keys_out, dev_idx, dev_score);
  ++*launches;
  return cudaGetLastError();
}

namespace {
// Exact scores of a SUBSET of the corpus (re-rank stage of a two-stage search: binary / u8 first pass -> exact f32,
// src/scalar.rs:366-368, examples/binary_demo.rs:235-237). One thread per candidate, the reference's sequential unfused
// arithmetic (same bits as the full scan); out-of-range candidates score the metric's worst value.
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) subset_scores_kernel(const float* __restrict__ data, size_t ld, unsigned n,
                                                                    unsigned d, unsigned index_base,
                                                                    const float* __restrict__ query,
                                                                    const uint32_t* __restrict__ cand, unsigned m,
                                                                    float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);
  __shared__ float s_qn;
  for (unsigned i = threadIdx.x; i < d; i += blockDim.x) sq[i] = query[i];
  if (threadIdx.x == 0) {
    float ss = 0.0f;
    for (unsigned i = 0; i < d; ++i) ss = __fadd_rn(ss, __fmul_rn(query[i], query[i]));
    s_qn = __fsqrt_rn(ss);
  }
  __syncthreads();
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const unsigned gid = cand[j];
  const unsigned i = gid - index_base;  // candidates carry global ids
  if (gid < index_base || i >= n) {
    out[j] = MODE == PDX_L2 ? __int_as_float(0x7FC00000) : __int_as_float(0xFFC00000);  // sorts last under total_cmp
    return;
  }
  const float* p = data + i;
  float acc = 0.0f, ss = 0.0f;
  for (unsigned dd = 0; dd < d; ++dd) {
    const float v = __ldg(p + (size_t)dd * ld);
    accumulate<MODE == PDX_COSINE_FUSED ? PDX_DOT : MODE>(sq[dd], v, acc);
    if (MODE == PDX_COSINE_FUSED) ss = __fadd_rn(ss, __fmul_rn(v, v));
  }
  float sc = acc;
  if (MODE == PDX_COSINE_FUSED) {
    const float nrm = __fsqrt_rn(ss), qn = s_qn;
    sc = (!(qn < NORM_EPS) && nrm > NORM_EPS) ? __fdiv_rn(acc, __fmul_rn(qn, nrm)) : 0.0f;
  }
  out[j] = sc;
}
}  // namespace

cudaError_t launch_subset_scores(const PdxView& v, int mode, const float* dev_query, const uint32_t* dev_cand, size_t m,
                                 float* dev_out, cudaStream_t s, LaunchCounter* launches) {
  if (m == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((m + SCAN_THREADS - 1) / SCAN_THREADS);
  const size_t smem = ((v.d + 3) & ~(size_t)3) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
#define INNR_SUBSET(MODE)                                                                                           \
  {                                                                                                                 \
    auto kern = subset_scores_kernel<MODE>;                                                                         \
    if (smem > 48 * 1024) {                                                                                         \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
      if (e != cudaSuccess) return e;                                                                               \
    }                                                                                                               \
    kern<<<grid, SCAN_THREADS, smem, s>>>(v.data, v.ld, (unsigned)v.n, (unsigned)v.d, v.index_base, dev_query,      \
                                          dev_cand, (unsigned)m, dev_out);                                          \
  }
  if (mode == PDX_DOT) INNR_SUBSET(PDX_DOT)
  else if (mode == PDX_L2) INNR_SUBSET(PDX_L2)
  else if (mode == PDX_COSINE_FUSED) INNR_SUBSET(PDX_COSINE_FUSED)
  else return cudaErrorInvalidValue;
#undef INNR_SUBSET
  ++*launches;
  return cudaGetLastError();
}

// k keys of a score vector, any k: rounds of <= 128 keys, each bounded below by the last key of the round before
// (read on the device: no host synchronisation between rounds)
cudaError_t launch_topk_from_scores(const void* dev_scores, int kind, size_t n, uint32_t index_base, size_t k,
                                    uint64_t* dev_keys, Workspace& ws, cudaStream_t s, LaunchCounter* launches,
                                    const uint32_t* dev_ids, const uint32_t* dev_mask, unsigned seg_len, unsigned seg_stride) {
  unsigned grid = (unsigned)((n + SCAN_THREADS - 1) / SCAN_THREADS);
  unsigned cap = (unsigned)ws.num_sms * 4u;
  if (grid > cap) grid = cap;
  if (grid == 0) grid = 1;
  for (size_t done = 0; done < k; done += 128) {
    const int kr = (int)std::min<size_t>(128, k - done);
    const uint64_t* floor = done ? dev_keys + done - 1 : nullptr;
    uint64_t* out = dev_keys + done;
    const size_t smem = (size_t)(SCAN_THREADS / 32) * kr * sizeof(uint64_t);
#define INNR_TOPK_LAUNCH(R, KIND)                                                                                   \
  topk_scores_kernel<R, KIND><<<grid, SCAN_THREADS, smem, s>>>(dev_scores, (unsigned)n, index_base, dev_ids, dev_mask, seg_len, seg_stride, kr, floor, \
                                                               ws.partials, ws.group_partials, out, ws.tickets)
    if (kr <= 32) {
      if (kind == 0) INNR_TOPK_LAUNCH(1, 0); else if (kind == 1) INNR_TOPK_LAUNCH(1, 1);
      else if (kind == 2) INNR_TOPK_LAUNCH(1, 2); else if (kind == 3) INNR_TOPK_LAUNCH(1, 3); else INNR_TOPK_LAUNCH(1, 4);
    } else {
      if (kind == 0) INNR_TOPK_LAUNCH(4, 0); else if (kind == 1) INNR_TOPK_LAUNCH(4, 1);
      else if (kind == 2) INNR_TOPK_LAUNCH(4, 2); else if (kind == 3) INNR_TOPK_LAUNCH(4, 3); else INNR_TOPK_LAUNCH(4, 4);
    }
#undef INNR_TOPK_LAUNCH
    ++*launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

namespace {
__global__ void decode_keys_kernel(const uint64_t* __restrict__ keys, size_t total, int descending, uint64_t* __restrict__ idx,
                                   float* __restrict__ score) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const uint64_t key = keys[t];
  if (idx) idx[t] = key & 0xFFFFFFFFull;
  if (score) {
    uint32_t hi = (uint32_t)(key >> 32);
    if (descending) hi = ~hi;
    score[t] = __uint_as_float(order_bits_to_f32_bits(hi));
  }
}
}  // namespace

// merge for k > 128: per query, selection rounds of <= 128 over the n_lists * k gathered keys, then one decode launch
cudaError_t launch_merge_keys_big(const uint64_t* dev_in, size_t n_lists, size_t nq, size_t k, int descending,
                                  uint64_t* dev_keys_out, uint64_t* dev_idx, float* dev_score, Workspace& ws, cudaStream_t s,
                                  LaunchCounter* launches) {
  for (size_t q = 0; q < nq; ++q) {
    cudaError_t e = launch_topk_from_scores(dev_in + q * k, 3, n_lists * k, 0, k, dev_keys_out + q * k, ws, s, launches, nullptr,
                                            nullptr, (unsigned)k, (unsigned)(nq * k));
    if (e != cudaSuccess) return e;
  }
  if (dev_idx || dev_score) {
    const size_t total = nq * k;
    decode_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dev_keys_out, total, descending, dev_idx, dev_score);
    ++*launches;
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// batch_knn_adaptive (src/batch.rs:441-564). The reference walks the dimensions one by one and, inside a dimension, the
// vectors in index order, pruning a candidate when its partial distance exceeds a threshold that is refreshed after
// every dimension d with d % 32 == 0 -- and never pruning once only k candidates are left. Between two refreshes the
// threshold is constant, so an "epoch" (the dimensions up to and including the next refresh point) is one pass: every
// live vector continues its sequential sum and notes the first dimension at which it exceeded the threshold. Unless
// fewer than k candidates would remain, all of those are pruned, in any order; otherwise the host resolves the
// reference's (dimension, index) order from the noted dimensions (api.cu). Rows of dead vectors are not read.
// ---------------------------------------------------------------------------------------------------------
namespace {

struct AdaptArgs {
  const float* data;
  unsigned long long ld;
  unsigned ld4;
  const float* query;
  unsigned d0, d1;          // dimensions [d0, d1) of this epoch
  float* dist;              // partial distances (in/out)
  const uint32_t* mask_in;  // bit i: vector i is a live candidate
  uint32_t* mask_out;
  const float* thr;         // thr[0]: the epoch's threshold (device)
  uint32_t* ev;             // ev[i]: first dimension at which vector i exceeded the threshold (written when it did)
  unsigned* pruned;         // += number of vectors that exceeded it
  int no_prune;             // only k candidates left: accumulate, never prune (src/batch.rs:520 `alive_count > k`)
};

__global__ void __launch_bounds__(SCAN_THREADS) adaptive_epoch_kernel(const AdaptArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);
  const unsigned nd = a.d1 - a.d0;
  for (unsigned j = threadIdx.x; j < nd; j += blockDim.x) sq[j] = a.query[a.d0 + j];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned i0 = (blockIdx.x * SCAN_THREADS + threadIdx.x) * VPT;
  const bool in = i0 < a.ld4;
  unsigned nib = in ? (a.mask_in[i0 >> 5] >> (i0 & 31)) & 0xFu : 0u;
  unsigned n_dead = 0;
  if (nib) {
    float4 acc = *reinterpret_cast<const float4*>(a.dist + i0);
    const float T = a.thr[0];
    unsigned ex = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    const float* p = a.data + (size_t)a.d0 * a.ld + i0;
    auto step = [&](const float4& v, unsigned j) {
      const float q = sq[j];
      float df;
      df = __fsub_rn(q, v.x); acc.x = __fadd_rn(acc.x, __fmul_rn(df, df));
      df = __fsub_rn(q, v.y); acc.y = __fadd_rn(acc.y, __fmul_rn(df, df));
      df = __fsub_rn(q, v.z); acc.z = __fadd_rn(acc.z, __fmul_rn(df, df));
      df = __fsub_rn(q, v.w); acc.w = __fadd_rn(acc.w, __fmul_rn(df, df));
      if (!(ex & 1u) && acc.x > T) { ex |= 1u; e0 = a.d0 + j; }   // *dist > threshold  (src/batch.rs:520)
      if (!(ex & 2u) && acc.y > T) { ex |= 2u; e1 = a.d0 + j; }
      if (!(ex & 4u) && acc.z > T) { ex |= 4u; e2 = a.d0 + j; }
      if (!(ex & 8u) && acc.w > T) { ex |= 8u; e3 = a.d0 + j; }
    };
    constexpr int U = 8;
    unsigned j = 0;
    for (; j + U <= nd; j += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg_stream_f4(p + (size_t)u * a.ld);
      p += (size_t)U * a.ld;
#pragma unroll
      for (int u = 0; u < U; ++u) step(v[u], j + u);
    }
    for (; j < nd; ++j) {
      const float4 v = ldg_stream_f4(p);
      p += a.ld;
      step(v, j);
    }
    *reinterpret_cast<float4*>(a.dist + i0) = acc;
    if (!a.no_prune) {
      const unsigned dead = ex & nib;
      if (dead & 1u) a.ev[i0] = e0;
      if (dead & 2u) a.ev[i0 + 1] = e1;
      if (dead & 4u) a.ev[i0 + 2] = e2;
      if (dead & 8u) a.ev[i0 + 3] = e3;
      nib &= ~dead;
      n_dead = __popc(dead);
    }
  }
  // eight lanes hold the 32 bits of one mask word
  unsigned w = nib << (4 * (lane & 7));
  w |= __shfl_xor_sync(FULL_MASK, w, 1);
  w |= __shfl_xor_sync(FULL_MASK, w, 2);
  w |= __shfl_xor_sync(FULL_MASK, w, 4);
  if (in && (lane & 7) == 0) a.mask_out[i0 >> 5] = w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_dead += __shfl_xor_sync(FULL_MASK, n_dead, o);
  if (lane == 0 && n_dead) atomicAdd(a.pruned, n_dead);
}

// thr[0] = (score of *key) * scale, thr[1] = thr[0] * 1.5  (src/batch.rs:486-487, :496, :546)
__global__ void adaptive_threshold_kernel(const uint64_t* key, float scale, float* thr) {
  const float t = __fmul_rn(__uint_as_float(order_bits_to_f32_bits((uint32_t)(*key >> 32))), scale);
  thr[0] = t;
  thr[1] = __fmul_rn(t, 1.5f);
}

// first pruning round (src/batch.rs:493-501): candidate i dies when dist[i] * ratio > threshold * 1.5
__global__ void __launch_bounds__(SCAN_THREADS) adaptive_mark_kernel(const float* __restrict__ dist, unsigned n, unsigned ld4,
                                                                    float ratio, const float* __restrict__ thr,
                                                                    uint32_t* __restrict__ mask_out, unsigned* pruned) {
  const int lane = threadIdx.x & 31;
  const unsigned i = blockIdx.x * SCAN_THREADS + threadIdx.x;
  const unsigned words32 = (ld4 + 31u) / 32u * 32u;
  bool alive = false, dead = false;
  if (i < n) {
    dead = __fmul_rn(dist[i], ratio) > thr[1];
    alive = !dead;
  }
  const unsigned w = __ballot_sync(FULL_MASK, alive);
  const unsigned nd = __popc(__ballot_sync(FULL_MASK, dead));
  if (lane == 0) {
    if (i < words32) mask_out[i >> 5] = w;
    if (nd) atomicAdd(pruned, nd);
  }
}

// vectors that were live before the epoch and are not after it, as keys ordered by DEcreasing (dimension, index)
__global__ void __launch_bounds__(SCAN_THREADS) adaptive_event_keys_kernel(const uint32_t* __restrict__ before,
                                                                          const uint32_t* __restrict__ after,
                                                                          const uint32_t* __restrict__ ev, unsigned n,
                                                                          uint64_t* __restrict__ keys) {
  const unsigned i = blockIdx.x * SCAN_THREADS + threadIdx.x;
  if (i >= n) return;
  const bool event = ((before[i >> 5] & ~after[i >> 5]) >> (i & 31)) & 1u;
  keys[i] = event ? ~(((uint64_t)ev[i] << 32) | i) : KEY_SENTINEL;
}

__global__ void adaptive_revive_kernel(const uint64_t* __restrict__ keys, unsigned m, uint32_t* mask) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const unsigned i = (unsigned)(~keys[t] & 0xFFFFFFFFull);
  atomicOr(mask + (i >> 5), 1u << (i & 31));
}

}  // namespace

cudaError_t launch_adaptive_threshold(const uint64_t* dev_key, float scale, float* dev_thr, cudaStream_t s,
                                      LaunchCounter* launches) {
  adaptive_threshold_kernel<<<1, 1, 0, s>>>(dev_key, scale, dev_thr);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_adaptive_mark(const PdxView& v, const float* dev_dist, float ratio, const float* dev_thr,
                                 uint32_t* dev_mask_out, unsigned* dev_pruned, cudaStream_t s, LaunchCounter* launches) {
  const unsigned ld4 = (unsigned)std::min<size_t>(v.ld, (v.n + 3) / 4 * 4);
  const unsigned span = (ld4 + 31u) / 32u * 32u;
  adaptive_mark_kernel<<<(span + SCAN_THREADS - 1) / SCAN_THREADS, SCAN_THREADS, 0, s>>>(dev_dist, (unsigned)v.n, ld4, ratio,
                                                                                      dev_thr, dev_mask_out, dev_pruned);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_adaptive_epoch(const PdxView& v, const float* dev_query, size_t d0, size_t d1, float* dev_dist,
                                  const uint32_t* dev_mask_in, uint32_t* dev_mask_out, const float* dev_thr, uint32_t* dev_ev,
                                  unsigned* dev_pruned, int no_prune, cudaStream_t s, LaunchCounter* launches) {
  AdaptArgs a{};
  a.data = v.data;
  a.ld = v.ld;
  a.ld4 = (unsigned)std::min<size_t>(v.ld, (v.n + 3) / 4 * 4);
  a.query = dev_query;
  a.d0 = (unsigned)d0;
  a.d1 = (unsigned)d1;
  a.dist = dev_dist;
  a.mask_in = dev_mask_in;
  a.mask_out = dev_mask_out;
  a.thr = dev_thr;
  a.ev = dev_ev;
  a.pruned = dev_pruned;
  a.no_prune = no_prune;
  const unsigned span = (a.ld4 + 31u) / 32u * 32u;  // whole mask words: every word of the output mask is written
  const unsigned grid = (span / VPT + SCAN_THREADS - 1) / SCAN_THREADS;
  adaptive_epoch_kernel<<<grid, SCAN_THREADS, (d1 - d0) * sizeof(float), s>>>(a);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_adaptive_event_keys(const uint32_t* dev_before, const uint32_t* dev_after, const uint32_t* dev_ev, size_t n,
                                       uint64_t* dev_keys, cudaStream_t s, LaunchCounter* launches) {
  adaptive_event_keys_kernel<<<(unsigned)((n + SCAN_THREADS - 1) / SCAN_THREADS), SCAN_THREADS, 0, s>>>(dev_before, dev_after,
                                                                                                     dev_ev, (unsigned)n, dev_keys);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_adaptive_revive(const uint64_t* dev_keys, size_t m, uint32_t* dev_mask, cudaStream_t s,
                                   LaunchCounter* launches) {
  if (m == 0) return cudaSuccess;
  adaptive_revive_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(dev_keys, (unsigned)m, dev_mask);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_topk_from_distances(const float* dev_dist, size_t n, size_t k, uint64_t* dev_keys,
                                       Workspace& ws, cudaStream_t s, LaunchCounter* launches) {
  return launch_topk_from_scores(dev_dist, 0, n, 0, k, dev_keys, ws, s, launches, nullptr, nullptr, 1, 1);
}

}  // namespace innr
