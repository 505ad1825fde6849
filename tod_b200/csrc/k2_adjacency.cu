// K2: pairwise 3-D consistency adjacency -> two bit-matrices per cluster.
//
// Replaces AdjacencyRansac::FillAdjacency (src/common/adjacency_ransac.cpp:127-172 of the reference).  For every pair
// (i, j) of correspondences of one (frame, object) cluster:
//   physical(i,j) <=> !(dq2 > (span+2e)^2) && !(|dt - dq| > 4e)                                   (:143-155)
//   sample(i,j)   <=> physical(i,j) && pixel_dist2(i,j) > 20*20 && |dt - dq| < 2e                 (:157-165)
// with dq2 = float distSq of the two query (camera-frame) points (sac_model_registration_graph.h:52-58),
// dq = sqrtf(dq2), dt = (float) sqrt( double sum of squares of the float difference of the two training points )
// (cv::norm(Vec3f) accumulates in double — SURVEY.md quirk Q8).  The arithmetic below reproduces those roundings
// one by one (this file is compiled with --fmad=false; products feeding the double sum are exact), so the
// bit-matrices are identical to the reference's neighbour lists: bit j of row i <=> neighbors(i) contains j.
// The comparisons keep the reference's polarity so NaN inputs behave the same.
//
// Layout: per cluster a full symmetric n x W u32 matrix, W = row_words(n) (multiple of 4 -> 16-byte rows).
// Mapping: the matrix is cut into 32 x 32 tiles and only the upper triangle of tiles is evaluated — the pair test is
// symmetric bit for bit (differences negate, squares are equal).  grid = (tile rows, clusters); a CTA owns tile row rb,
// stages its 32 rows' data in shared memory, and its 8 warps take the tile columns cb >= rb round-robin.  In a tile,
// lane L is column 32 cb + L and sweeps the 32 rows: a __ballot_sync per row yields that row's word (row-major
// store), while the lane's own results accumulate into the word of the transposed tile (column-major store).
#include "tod_internal.h"

namespace tod {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__global__ void __launch_bounds__(kThreads)
k2_adjacency_kernel(const int32_t *__restrict__ offsets, const int64_t *__restrict__ matrix_offsets,
                    const float *__restrict__ query, const float *__restrict__ train,
                    const float *__restrict__ pixels, const float *__restrict__ spans, float sensor_error,
                    uint32_t *__restrict__ physical, uint32_t *__restrict__ sample) {
  __shared__ float4 s_row[32][2];  // per row of this tile row: (qx qy qz tx) (ty tz px py)

  const int c = blockIdx.y;
  const int base = offsets[c];
  const int n = offsets[c + 1] - base;
  const int rb = blockIdx.x;
  if (rb * 32 >= n) return;
  const int W = ((n + 31) / 32 + 3) & ~3;
  uint32_t *P = physical + matrix_offsets[c];
  uint32_t *S = sample + matrix_offsets[c];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  const float span = spans[c];
  const float e2 = __fmul_rn(2.0f, sensor_error);
  const float e4 = __fmul_rn(4.0f, sensor_error);
  const float sp = __fadd_rn(span, e2);
  const float thr_span = __fmul_rn(sp, sp);

  if (threadIdx.x < 32) {
    const int i = min(rb * 32 + lane, n - 1) + base;
    s_row[lane][0] = make_float4(__ldg(query + size_t(i) * 3), __ldg(query + size_t(i) * 3 + 1),
                                 __ldg(query + size_t(i) * 3 + 2), __ldg(train + size_t(i) * 3));
    s_row[lane][1] = make_float4(__ldg(train + size_t(i) * 3 + 1), __ldg(train + size_t(i) * 3 + 2),
                                 __ldg(pixels + size_t(i) * 2), __ldg(pixels + size_t(i) * 2 + 1));
  }
  __syncthreads();

  // tile columns rb .. W-1: blocks past the last correspondence produce the zero padding words of the rows
  for (int cb = rb + warp; cb < W; cb += kWarps) {
    const int j = cb * 32 + lane;
    const bool col_ok = j < n;
    const size_t g = size_t(base) + (col_ok ? j : 0);
    const float qx = __ldg(query + g * 3), qy = __ldg(query + g * 3 + 1), qz = __ldg(query + g * 3 + 2);
    const float tx = __ldg(train + g * 3), ty = __ldg(train + g * 3 + 1), tz = __ldg(train + g * 3 + 2);
    const float px = __ldg(pixels + g * 2), py = __ldg(pixels + g * 2 + 1);
    uint32_t rowP = 0, rowS = 0;  // lane r keeps the word of row 32 rb + r
    uint32_t colP = 0, colS = 0;  // this lane's column: bit r = adjacent to row 32 rb + r
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const int i = rb * 32 + r;
      const float4 a = s_row[r][0], b = s_row[r][1];
      bool isP = false, isS = false;
      if (col_ok && i < n && j != i) {
        const float dx = __fsub_rn(a.x, qx), dy = __fsub_rn(a.y, qy), dz = __fsub_rn(a.z, qz);
        const float dq2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (!(dq2 > thr_span)) {
          const float ux = __fsub_rn(a.w, tx), uy = __fsub_rn(b.x, ty), uz = __fsub_rn(b.y, tz);
          const float dt2 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz));
          // Decide in cheap arithmetic first.  Against the reference's dq = sqrtf(dq2) and dt = (float) sqrt of the
          // double-accumulated sum, the approximate roots below are off by < 2^-22 relative each and dt2 by
          // < 3 * 2^-24, so |dt - dq| is known to within 1e-6 * max(dt, dq); `tol` is 4x that.  Only when the
          // difference lies that close to a threshold — about one pair in 10^5 — is the exact arithmetic replayed.
          float dq = sqrt_approx(dq2), dt = sqrt_approx(dt2);
          float diff = fabsf(__fsub_rn(dt, dq));
          const float tol = __fmul_rn(4e-6f, fmaxf(dt, dq));
          if (!(fabsf(__fsub_rn(diff, e4)) > tol && fabsf(__fsub_rn(diff, e2)) > tol)) {  // also taken for NaNs
            dq = __fsqrt_rn(dq2);
            const double vx = double(ux), vy = double(uy), vz = double(uz);
            const double s2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
            dt = __double2float_rn(__dsqrt_rn(s2));
            diff = fabsf(__fsub_rn(dt, dq));
          }
          if (!(diff > e4)) {
            isP = true;
            const float ax = __fsub_rn(b.z, px), ay = __fsub_rn(b.w, py);
            const float pd = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
            isS = (pd > 400.0f) && (diff < e2);
          }
        }
      }
      const uint32_t bp = __ballot_sync(0xffffffffu, isP);
      const uint32_t bs = __ballot_sync(0xffffffffu, isS);
      if (lane == r) {
        rowP = bp;
        rowS = bs;
      }
      colP |= isP ? (1u << r) : 0u;
      colS |= isS ? (1u << r) : 0u;
    }
    const int i = rb * 32 + lane;
    if (i < n) {  // row-major words of the tile (also the zero padding when cb is past the data)
      P[size_t(i) * W + cb] = rowP;
      S[size_t(i) * W + cb] = rowS;
    }
    if (cb != rb && col_ok) {  // the transposed tile
      P[size_t(j) * W + rb] = colP;
      S[size_t(j) * W + rb] = colS;
    }
  }
}

}  // namespace

cudaError_t launch_fill_adjacency(int n_clusters, const int32_t *d_offsets, const int64_t *d_matrix_offsets,
                                  const float *d_query, const float *d_train, const float *d_pixels,
                                  const float *d_spans, float sensor_error, uint32_t *d_physical,
                                  uint32_t *d_sample, int max_cluster, cudaStream_t stream) {
  if (n_clusters <= 0 || max_cluster <= 0) return cudaSuccess;
  for (int c0 = 0; c0 < n_clusters; c0 += 65535) {  // gridDim.y limit
    const int nc = min(65535, n_clusters - c0);
    dim3 grid((max_cluster + 31) / 32, nc);
    k2_adjacency_kernel<<<grid, kThreads, 0, stream>>>(d_offsets + c0, d_matrix_offsets + c0, d_query, d_train,
                                                       d_pixels, d_spans + c0, sensor_error, d_physical, d_sample);
    count_launch();
  }
  return cudaGetLastError();
}

}  // namespace tod
