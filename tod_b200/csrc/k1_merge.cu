// Top-k merge for K1: reduces per-chunk (and, after the NCCL all-gather, per-GPU) candidate key lists to the final
// k per query, then applies the radius cut, decodes (imgIdx, trainIdx) and gathers the matched 3-D model points.
//
// Replaces the tail of DescriptorMatcher::process (src/detection/DescriptorMatcher.cpp:212-220 radius cut,
// :232-244 matches_3d gather) of the reference.  Keys are (distance << 23 | global_row); because the objects are
// concatenated in imgIdx order, ascending key order IS cv::BFMatcher's (distance, imgIdx, trainIdx) order, on one
// GPU or across shards.
#include "tod_internal.h"

namespace tod {
namespace {

template <int K, typename T>
__device__ __forceinline__ void topk_insert(T (&best)[K], T key) {
  if (key < best[K - 1]) {
    best[K - 1] = key;
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      const T lo = min(best[i - 1], best[i]);
      const T hi = max(best[i - 1], best[i]);
      best[i - 1] = lo;
      best[i] = hi;
    }
  }
}

// keys: n_src lists of nq x K keys, src_stride keys apart.  One thread per query; reads are coalesced across the warp
// for every (src, slot).  kPeerWritten: the lists were stored by other GPUs over NVLink during this launch's wait —
// read them through L2 (ld.global.cg), never through the non-coherent path.
template <int K, bool kPeerWritten = false>
__device__ __forceinline__ void reduce_query(const uint32_t *__restrict__ keys, int n_src, size_t src_stride, int q,
                                             uint32_t (&best)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) best[i] = kKeyEmpty;
  for (int s = 0; s < n_src; ++s) {
    const uint32_t *p = keys + size_t(s) * src_stride + size_t(q) * K;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const uint32_t key = kPeerWritten ? __ldcg(p + i) : __ldg(p + i);
      if (key >= best[K - 1]) break;  // source lists are ascending
      topk_insert<K, uint32_t>(best, key);
    }
  }
}

// Wide databases (more than 2^23 rows): the 23 row bits of a key count from the start of its source's row range
// (a segment of a shard), src_base[s] = first global row of source s.  Sources are ascending, disjoint row ranges, so
// (distance, global row) — cv::BFMatcher's order — is compared on 64-bit keys: distance << 32 | global row.
constexpr unsigned long long kKeyEmpty64 = ~0ull;
template <int K>
__device__ __forceinline__ void reduce_query_wide(const uint32_t *__restrict__ keys, int n_src, size_t src_stride,
                                                  const uint32_t *__restrict__ src_base, int q,
                                                  unsigned long long (&best)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) best[i] = kKeyEmpty64;
  for (int s = 0; s < n_src; ++s) {
    const uint32_t *p = keys + size_t(s) * src_stride + size_t(q) * K;
    const unsigned long long base = __ldg(src_base + s);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const uint32_t k32 = __ldg(p + i);
      if (k32 == kKeyEmpty) break;
      const unsigned long long key = ((unsigned long long)(k32 >> kKeyRowBits) << 32) | (base + (k32 & kKeyRowMask));
      if (key >= best[K - 1]) break;  // source lists are ascending
      topk_insert<K, unsigned long long>(best, key);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(128) reduce_keys_kernel(const uint32_t *__restrict__ keys, int n_src, int nq,
                                                          uint32_t *__restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t best[K];
  reduce_query<K>(keys, n_src, size_t(nq) * K, q, best);
#pragma unroll
  for (int i = 0; i < K; ++i) out[size_t(q) * K + i] = best[i];
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Fused top-k reduction + all-gather over peer memory: every rank reduces its chunk lists to K keys per query and
// stores them straight into slot `rank` of EVERY rank's exchange buffer (plain stores into CUDA-IPC-mapped peer memory:
// NVLink writes, no copy engine, no NCCL kernel); the block that finishes last raises this rank's flag on every peer
// with a system-scope release.  peers[r] = base of rank r's buffer for this step's parity: world slots of slot_stride
// keys, then (at flag_offset keys) world flags.
template <int K>
__global__ void __launch_bounds__(128)
reduce_push_kernel(const uint32_t *__restrict__ keys, int n_src, int nq, uint32_t *const *__restrict__ peers, int world,
                   int rank, size_t slot_stride, size_t flag_offset, uint32_t step, unsigned int *__restrict__ ticket) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq) {
    uint32_t best[K];
    reduce_query<K>(keys, n_src, size_t(nq) * K, q, best);
    for (int r = 0; r < world; ++r) {
      uint32_t *dst = peers[r] + size_t(rank) * slot_stride + size_t(q) * K;
#pragma unroll
      for (int i = 0; i < K; ++i) dst[i] = best[i];
    }
  }
  // the CTA barrier orders every thread's peer stores before thread 0's system-scope fence (cumulativity), which
  // publishes them before the block takes its ticket — one fence per block instead of one per thread
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(peers[r] + flag_offset + rank, step);
    }
  }
}

template <int K, bool kWide>
__global__ void __launch_bounds__(128)
finalize_matches_kernel(const uint32_t *__restrict__ keys, int n_src, int nq, uint32_t radius,
                        const uint32_t *__restrict__ obj_offsets, int n_objects, const float *__restrict__ points,
                        tod_match *__restrict__ matches, int32_t *__restrict__ counts,
                        float *__restrict__ points3d, int ratio_enabled, float ratio, uint32_t *__restrict__ rows_out,
                        size_t src_stride, const uint32_t *__restrict__ wait_flags, uint32_t wait_step,
                        uint32_t *__restrict__ wait_error, const uint32_t *__restrict__ src_base) {
  if (wait_flags) {
    // peer exchange: list s was pushed by rank s; its flag reaches wait_step once all of it is visible here.  Bounded
    // wait (~4 s): a rank that never arrives must not hang the GPU — the error word is reported by the next call.
    if (int(threadIdx.x) < n_src) {
      const long long t0 = clock64();
      while (ld_acquire_sys(wait_flags + threadIdx.x) != wait_step) {
        if (clock64() - t0 > (1ll << 33)) {
          atomicExch(wait_error, 1u);
          break;
        }
        __nanosleep(64);
      }
    }
    __syncthreads();
  }
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  // (distance, global row) of the K nearest, ascending; empty slots have distance 0xFFFFFFFF
  uint32_t dist_of[K], row_of[K];
  if (kWide) {
    unsigned long long best[K];
    reduce_query_wide<K>(keys, n_src, src_stride, src_base, q, best);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      dist_of[i] = uint32_t(best[i] >> 32);
      row_of[i] = uint32_t(best[i]);
    }
  } else {
    uint32_t best[K];
    if (wait_flags) reduce_query<K, true>(keys, n_src, src_stride, q, best);
    else reduce_query<K>(keys, n_src, src_stride, q, best);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      dist_of[i] = best[i] == kKeyEmpty ? 0xFFFFFFFFu : best[i] >> kKeyRowBits;
      row_of[i] = best[i] & kKeyRowMask;
    }
  }
  if (ratio_enabled && K >= 2) {
    // the ratio-test TODO of DescriptorMatcher.cpp:223-227 as Lowe's test on the two nearest neighbours (before the
    // radius cut): keep the best match only, and only if distance0 < ratio * distance1
    if (dist_of[0] != 0xFFFFFFFFu && dist_of[K >= 2 ? 1 : 0] != 0xFFFFFFFFu) {
      const float d0 = float(dist_of[0]), d1 = float(dist_of[K >= 2 ? 1 : 0]);
      if (!(d0 < __fmul_rn(ratio, d1))) dist_of[0] = 0xFFFFFFFFu;
    }
#pragma unroll
    for (int i = 1; i < K; ++i) dist_of[i] = 0xFFFFFFFFu;
  }
  int n = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const uint32_t dist = dist_of[i];
    // ascending list: the first empty slot or the first distance > radius ends it (DescriptorMatcher.cpp:215-219)
    const bool keep = (n == i) && dist != 0xFFFFFFFFu && (radius == 0 || dist <= radius);
    tod_match m;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (keep) {
      const uint32_t row = row_of[i];
      int lo = 0, hi = n_objects;  // largest o with obj_offsets[o] <= row
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(obj_offsets + mid) <= row) lo = mid; else hi = mid;
      }
      m.queryIdx = q;
      m.trainIdx = int(row - __ldg(obj_offsets + lo));
      m.imgIdx = lo;
      m.distance = float(dist);
      if (points3d) {
        px = __ldg(points + size_t(row) * 3);
        py = __ldg(points + size_t(row) * 3 + 1);
        pz = __ldg(points + size_t(row) * 3 + 2);
      }
      ++n;
    } else {
      m.queryIdx = -1; m.trainIdx = -1; m.imgIdx = -1; m.distance = 0.f;
    }
    matches[size_t(q) * K + i] = m;
    if (rows_out) rows_out[size_t(q) * K + i] = keep ? row_of[i] : 0xFFFFFFFFu;
    if (points3d) {
      float *o = points3d + (size_t(q) * K + i) * 3;
      o[0] = px; o[1] = py; o[2] = pz;
    }
  }
  counts[q] = n;
}

// ---- duplicate-match removal (the TODO at DescriptorMatcher.cpp:229) ------------------------------------------------
// A DB descriptor matched by several keypoints of one frame keeps only the match with the smallest (distance, queryIdx).
// Open-addressing hash table keyed by (frame, global row): pass 1 records the winner per key with atomicMin, pass 2
// drops the losers and compacts every query's list in place (order kept).
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
dedupe_insert_kernel(const tod_match *__restrict__ matches, const int32_t *__restrict__ counts,
                     const uint32_t *__restrict__ rows, int nq, int k, int frame_kp,
                     unsigned long long *__restrict__ hkeys, unsigned long long *__restrict__ hvals, uint64_t mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * k) return;
  const int q = i / k, j = i - q * k;
  if (j >= counts[q]) return;
  const uint64_t frame = frame_kp > 0 ? uint64_t(q / frame_kp) : 0ull;
  const unsigned long long key = ((frame << 32) | uint64_t(rows[i])) + 1ull;
  const unsigned long long val = (uint64_t(uint32_t(matches[i].distance)) << 32) | uint64_t(uint32_t(q));
  uint64_t h = mix64(key) & mask;
  for (;;) {
    const unsigned long long old = atomicCAS(hkeys + h, 0ull, key);
    if (old == 0ull || old == key) {
      atomicMin(hvals + h, val);
      return;
    }
    h = (h + 1) & mask;
  }
}

__global__ void __launch_bounds__(128)
dedupe_filter_kernel(tod_match *__restrict__ matches, int32_t *__restrict__ counts, float *__restrict__ points3d,
                     const uint32_t *__restrict__ rows, int nq, int k, int frame_kp,
                     const unsigned long long *__restrict__ hkeys, const unsigned long long *__restrict__ hvals,
                     uint64_t mask) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int n = counts[q];
  const uint64_t frame = frame_kp > 0 ? uint64_t(q / frame_kp) : 0ull;
  int out = 0;
  for (int j = 0; j < n; ++j) {
    const size_t i = size_t(q) * k + j;
    const tod_match m = matches[i];
    const unsigned long long key = ((frame << 32) | uint64_t(rows[i])) + 1ull;
    const unsigned long long val = (uint64_t(uint32_t(m.distance)) << 32) | uint64_t(uint32_t(q));
    uint64_t h = mix64(key) & mask;
    while (hkeys[h] != key) h = (h + 1) & mask;
    if (hvals[h] != val) continue;  // another keypoint of this frame owns the descriptor
    if (out != j) {
      matches[size_t(q) * k + out] = m;
      if (points3d)
        for (int d = 0; d < 3; ++d) points3d[(size_t(q) * k + out) * 3 + d] = points3d[i * 3 + d];
    }
    ++out;
  }
  for (int j = out; j < n; ++j) {
    tod_match e;
    e.queryIdx = -1; e.trainIdx = -1; e.imgIdx = -1; e.distance = 0.f;
    matches[size_t(q) * k + j] = e;
    if (points3d)
      for (int d = 0; d < 3; ++d) points3d[(size_t(q) * k + j) * 3 + d] = 0.f;
  }
  counts[q] = out;
}

}  // namespace

cudaError_t launch_remove_duplicates(tod_match *d_matches, int32_t *d_counts, float *d_points3d,
                                     const uint32_t *d_rows, int nq, int k, int frame_keypoints, void *d_hkeys,
                                     void *d_hvals, size_t table_slots, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_hkeys, 0, table_slots * 8, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_hvals, 0xFF, table_slots * 8, stream);
  if (e != cudaSuccess) return e;
  const int n = nq * k;
  dedupe_insert_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_matches, d_counts, d_rows, nq, k, frame_keypoints,
                                                            static_cast<unsigned long long *>(d_hkeys),
                                                            static_cast<unsigned long long *>(d_hvals),
                                                            uint64_t(table_slots - 1));
  dedupe_filter_kernel<<<(nq + 127) / 128, 128, 0, stream>>>(d_matches, d_counts, d_points3d, d_rows, nq, k,
                                                             frame_keypoints,
                                                             static_cast<const unsigned long long *>(d_hkeys),
                                                             static_cast<const unsigned long long *>(d_hvals),
                                                             uint64_t(table_slots - 1));
  count_launch(2);
  return cudaGetLastError();
}

#define TOD_DISPATCH_K(k, CALL)       \
  switch (k) {                        \
    case 1: { constexpr int K = 1; CALL; } break; \
    case 2: { constexpr int K = 2; CALL; } break; \
    case 3: { constexpr int K = 3; CALL; } break; \
    case 4: { constexpr int K = 4; CALL; } break; \
    case 5: { constexpr int K = 5; CALL; } break; \
    case 6: { constexpr int K = 6; CALL; } break; \
    case 7: { constexpr int K = 7; CALL; } break; \
    case 8: { constexpr int K = 8; CALL; } break; \
    default: return cudaErrorInvalidValue;        \
  }

cudaError_t launch_reduce_keys(const uint32_t *d_keys, int n_src, int nq, int k, uint32_t *d_out,
                               cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const int blocks = (nq + 127) / 128;
  TOD_DISPATCH_K(k, (reduce_keys_kernel<K><<<blocks, 128, 0, stream>>>(d_keys, n_src, nq, d_out)));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_finalize_matches(const uint32_t *d_keys, int n_src, int nq, int k, uint32_t radius,
                                    const uint32_t *d_obj_offsets, int n_objects, const float *d_points,
                                    tod_match *d_matches, int32_t *d_counts, float *d_points3d,
                                    cudaStream_t stream, int ratio_enabled, float ratio, uint32_t *d_rows_out,
                                    size_t src_stride, const uint32_t *d_wait_flags, uint32_t wait_step,
                                    uint32_t *d_wait_error, const uint32_t *d_src_base) {
  if (nq <= 0) return cudaSuccess;
  if (d_wait_flags && (n_src > 128 || d_src_base)) return cudaErrorInvalidValue;
  if (src_stride == 0) src_stride = size_t(nq) * size_t(k);
  const int blocks = (nq + 127) / 128;
  if (d_src_base) {
    TOD_DISPATCH_K(k, (finalize_matches_kernel<K, true><<<blocks, 128, 0, stream>>>(
                          d_keys, n_src, nq, radius, d_obj_offsets, n_objects, d_points, d_matches, d_counts,
                          d_points3d, ratio_enabled, ratio, d_rows_out, src_stride, nullptr, 0u, nullptr, d_src_base)));
  } else {
    TOD_DISPATCH_K(k, (finalize_matches_kernel<K, false><<<blocks, 128, 0, stream>>>(
                          d_keys, n_src, nq, radius, d_obj_offsets, n_objects, d_points, d_matches, d_counts,
                          d_points3d, ratio_enabled, ratio, d_rows_out, src_stride, d_wait_flags, wait_step,
                          d_wait_error, nullptr)));
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_reduce_push(const uint32_t *d_keys, int n_src, int nq, int k, uint32_t *const *d_peers, int world,
                               int rank, size_t slot_stride, size_t flag_offset, uint32_t step, unsigned int *d_ticket,
                               cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const int blocks = (nq + 127) / 128;
  TOD_DISPATCH_K(k, (reduce_push_kernel<K><<<blocks, 128, 0, stream>>>(d_keys, n_src, nq, d_peers, world, rank,
                                                                        slot_stride, flag_offset, step, d_ticket)));
  count_launch();
  return cudaGetLastError();
}

}  // namespace tod
