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

template <int K>
__device__ __forceinline__ void topk_insert(uint32_t (&best)[K], uint32_t key) {
  if (key < best[K - 1]) {
    best[K - 1] = key;
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      const uint32_t lo = min(best[i - 1], best[i]);
      const uint32_t hi = max(best[i - 1], best[i]);
      best[i - 1] = lo;
      best[i] = hi;
    }
  }
}

// keys: n_src x nq x K.  One thread per query; reads are coalesced across the warp for every (src, slot).
template <int K>
__device__ __forceinline__ void reduce_query(const uint32_t *__restrict__ keys, int n_src, int nq, int q,
                                             uint32_t (&best)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) best[i] = kKeyEmpty;
  for (int s = 0; s < n_src; ++s) {
    const uint32_t *p = keys + (size_t(s) * nq + q) * K;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const uint32_t key = __ldg(p + i);
      if (key >= best[K - 1]) break;  // source lists are ascending
      topk_insert<K>(best, key);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(128) reduce_keys_kernel(const uint32_t *__restrict__ keys, int n_src, int nq,
                                                          uint32_t *__restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t best[K];
  reduce_query<K>(keys, n_src, nq, q, best);
#pragma unroll
  for (int i = 0; i < K; ++i) out[size_t(q) * K + i] = best[i];
}

template <int K>
__global__ void __launch_bounds__(128)
finalize_matches_kernel(const uint32_t *__restrict__ keys, int n_src, int nq, uint32_t radius,
                        const uint32_t *__restrict__ obj_offsets, int n_objects, const float *__restrict__ points,
                        tod_match *__restrict__ matches, int32_t *__restrict__ counts,
                        float *__restrict__ points3d) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t best[K];
  reduce_query<K>(keys, n_src, nq, q, best);
  int n = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const uint32_t key = best[i];
    const uint32_t dist = key >> kKeyRowBits;
    // ascending list: the first empty slot or the first distance > radius ends it (DescriptorMatcher.cpp:215-219)
    const bool keep = (n == i) && key != kKeyEmpty && (radius == 0 || dist <= radius);
    tod_match m;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (keep) {
      const uint32_t row = key & kKeyRowMask;
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
    if (points3d) {
      float *o = points3d + (size_t(q) * K + i) * 3;
      o[0] = px; o[1] = py; o[2] = pz;
    }
  }
  counts[q] = n;
}

}  // namespace

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
                                    cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const int blocks = (nq + 127) / 128;
  TOD_DISPATCH_K(k, (finalize_matches_kernel<K><<<blocks, 128, 0, stream>>>(
                        d_keys, n_src, nq, radius, d_obj_offsets, n_objects, d_points, d_matches, d_counts,
                        d_points3d)));
  count_launch();
  return cudaGetLastError();
}

}  // namespace tod
