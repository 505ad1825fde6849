// K1, SIMT formulation: exact k-NN Hamming over 256-bit descriptors with XOR + POPC and a register top-k.
//
// Replaces `matcher_->knnMatch(descriptors, matches, 5)` at src/detection/DescriptorMatcher.cpp:211 of the reference
// (cv::BFMatcher(NORM_HAMMING) semantics, see oracle/hamming_knn.py).
//
// Work split: grid = (db chunks, query tiles).  A CTA owns 256*QPT queries (QPT per thread, held in registers as
// 8 x u32 each) and streams one contiguous chunk of DB rows through a 4-stage shared-memory ring filled by 1-D TMA
// bulk copies (cp.async.bulk -> UBLKCP) signalled on mbarriers.  Every lane of a warp reads the same DB row
// (broadcast LDS.128 x2) and scores it against its own QPT queries: 8 LOP3 + 8 POPC + adds per pair.  Each thread
// keeps its queries' best k candidates in registers as packed keys (distance << 23 | global_row); because a thread
// scans rows in ascending order a later row with an equal distance can never displace an earlier one, so the hot
// loop only tests `distance < threshold` and the sorted insert is a rare, divergent slow path.
// Output: per (chunk, query) k keys; the cross-chunk (and cross-GPU) merge is k1_merge.cu.
#include "ptx.cuh"
#include "tod_internal.h"

namespace tod {

namespace {

constexpr int kThreads = 256;
constexpr int kTileRows = 256;                 // DB rows per smem stage (8 KB)
constexpr int kStages = 4;
constexpr int kTileBytes = kTileRows * 32;

template <int K, int QPT>
__global__ void __launch_bounds__(kThreads, 2)
k1_popc_kernel(const uint4 *__restrict__ query, int nq, const uint4 *__restrict__ db, int shard_rows,
               uint32_t global_row_base, int rows_per_chunk, uint32_t thr_init, uint32_t *__restrict__ partial) {
  __shared__ __align__(128) uint4 tile[kStages][kTileRows * 2];
  __shared__ __align__(8) uint64_t full_bar[kStages];

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x;
  const int q_base = blockIdx.y * (kThreads * QPT);
  const int row0 = chunk * rows_per_chunk;
  const int row1 = min(shard_rows, row0 + rows_per_chunk);
  const int n_rows = row1 - row0;
  const int n_tiles = (n_rows + kTileRows - 1) / kTileRows;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(&full_bar[s], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int t) {  // called by thread 0 only
    const int s = t % kStages;
    const int r = row0 + t * kTileRows;
    const uint32_t bytes = static_cast<uint32_t>(min(kTileRows, row1 - r)) * 32u;
    ptx::mbar_arrive_expect_tx(&full_bar[s], bytes);
    ptx::bulk_g2s(&tile[s][0], db + size_t(r) * 2, bytes, &full_bar[s]);
  };
  if (tid == 0) {
    for (int t = 0; t < min(kStages, n_tiles); ++t) issue(t);
  }

  // this thread's queries: q_base + tid + 256*j
  uint32_t qv[QPT][8];
  uint32_t best[QPT][K];
  uint32_t thr[QPT];
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int qi = q_base + tid + kThreads * j;
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (qi < nq) {
      a = __ldg(query + size_t(qi) * 2);
      b = __ldg(query + size_t(qi) * 2 + 1);
    }
    qv[j][0] = a.x; qv[j][1] = a.y; qv[j][2] = a.z; qv[j][3] = a.w;
    qv[j][4] = b.x; qv[j][5] = b.y; qv[j][6] = b.z; qv[j][7] = b.w;
#pragma unroll
    for (int i = 0; i < K; ++i) best[j][i] = kKeyEmpty;
    thr[j] = thr_init;
  }

  for (int t = 0; t < n_tiles; ++t) {
    const int s = t % kStages;
    ptx::mbar_wait(&full_bar[s], (t / kStages) & 1);
    const uint4 *__restrict__ p = &tile[s][0];
    const int rows_here = min(kTileRows, n_rows - t * kTileRows);
    const uint32_t grow0 = global_row_base + uint32_t(row0 + t * kTileRows);
#pragma unroll 4
    for (int r = 0; r < rows_here; ++r) {
      const uint4 a = p[2 * r];
      const uint4 b = p[2 * r + 1];
#pragma unroll
      for (int j = 0; j < QPT; ++j) {
        const uint32_t d = __popc(qv[j][0] ^ a.x) + __popc(qv[j][1] ^ a.y) + __popc(qv[j][2] ^ a.z) +
                           __popc(qv[j][3] ^ a.w) + __popc(qv[j][4] ^ b.x) + __popc(qv[j][5] ^ b.y) +
                           __popc(qv[j][6] ^ b.z) + __popc(qv[j][7] ^ b.w);
        if (d < thr[j]) {
          // rare slow path: replace the worst entry, bubble it into place, tighten the threshold
          best[j][K - 1] = (d << kKeyRowBits) | (grow0 + uint32_t(r));
#pragma unroll
          for (int i = K - 1; i > 0; --i) {
            const uint32_t lo = min(best[j][i - 1], best[j][i]);
            const uint32_t hi = max(best[j][i - 1], best[j][i]);
            best[j][i - 1] = lo;
            best[j][i] = hi;
          }
          thr[j] = min(thr_init, best[j][K - 1] >> kKeyRowBits);
        }
      }
    }
    __syncthreads();  // every warp is done with stage s -> it can be refilled
    if (tid == 0 && t + kStages < n_tiles) issue(t + kStages);
  }

#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int qi = q_base + tid + kThreads * j;
    if (qi < nq) {
      uint32_t *o = partial + (size_t(chunk) * nq + qi) * K;
#pragma unroll
      for (int i = 0; i < K; ++i) o[i] = best[j][i];
    }
  }
}

template <int K, int QPT>
cudaError_t launch_kq(const K1Plan &plan, const void *q, int nq, const void *db, int64_t rows, uint32_t base,
                      uint32_t thr_init, uint32_t *partial, cudaStream_t stream) {
  dim3 grid(plan.n_chunks, plan.n_qtiles);
  k1_popc_kernel<K, QPT><<<grid, kThreads, 0, stream>>>(static_cast<const uint4 *>(q), nq,
                                                         static_cast<const uint4 *>(db), int(rows), base,
                                                         plan.rows_per_chunk, thr_init, partial);
  count_launch();
  return cudaGetLastError();
}

template <int K>
cudaError_t launch_k(const K1Plan &plan, const void *q, int nq, const void *db, int64_t rows, uint32_t base,
                     uint32_t thr_init, uint32_t *partial, cudaStream_t stream) {
  switch (plan.q_per_thread) {
    case 1: return launch_kq<K, 1>(plan, q, nq, db, rows, base, thr_init, partial, stream);
    case 2: return launch_kq<K, 2>(plan, q, nq, db, rows, base, thr_init, partial, stream);
    default: return launch_kq<K, 4>(plan, q, nq, db, rows, base, thr_init, partial, stream);
  }
}

}  // namespace

K1Plan k1_popc_plan(int nq, int64_t shard_rows, int sm_count) {
  K1Plan p{};
  p.q_per_thread = nq <= kThreads ? 1 : (nq <= 2 * kThreads ? 2 : 4);
  p.q_tile = kThreads * p.q_per_thread;
  p.n_qtiles = (nq + p.q_tile - 1) / p.q_tile;
  if (p.n_qtiles < 1) p.n_qtiles = 1;
  const int64_t max_chunks = (shard_rows + kTileRows - 1) / kTileRows;
  // 2 CTAs are resident per SM: size the grid to at most ONE full wave (floor), so that no tail wave runs at low
  // occupancy; with more query tiles than CTA slots the DB is not split at all
  int64_t target = (2LL * sm_count) / p.n_qtiles;
  if (target > max_chunks) target = max_chunks;
  if (target < 1) target = 1;
  int64_t rpc = (shard_rows + target - 1) / target;
  rpc = (rpc + kTileRows - 1) / kTileRows * kTileRows;
  if (rpc < kTileRows) rpc = kTileRows;
  p.rows_per_chunk = int(rpc);
  p.n_chunks = int((shard_rows + rpc - 1) / rpc);
  if (p.n_chunks < 1) p.n_chunks = 1;
  p.n_sources = p.n_chunks;
  return p;
}

cudaError_t launch_k1_popc(const K1Plan &plan, const void *d_query, int nq, const void *d_db, int64_t shard_rows,
                           uint32_t global_row_base, int k, uint32_t radius, uint32_t *d_partial,
                           cudaStream_t stream) {
  // distance <= radius  <=>  distance < radius + 1 ; radius 0 = no cut (any distance is < 511)
  const uint32_t thr_init = radius ? min(radius + 1u, 511u) : 511u;
  switch (k) {
    case 1: return launch_k<1>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 2: return launch_k<2>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 3: return launch_k<3>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 4: return launch_k<4>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 5: return launch_k<5>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 6: return launch_k<6>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 7: return launch_k<7>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 8: return launch_k<8>(plan, d_query, nq, d_db, shard_rows, global_row_base, thr_init, d_partial, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tod
