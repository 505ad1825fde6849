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
// Mapping: grid = (row blocks of 32, clusters); a CTA has 8 warps, a warp owns 4 rows (row data in registers) and
// sweeps the columns in tiles of 1024 staged in shared memory as SoA; lane L of the warp tests column 32*w + L, a
// __ballot_sync turns 32 tests into one matrix word, lane w keeps it, and a tile ends with one coalesced 128-byte
// store per row and matrix.
#include "tod_internal.h"

namespace tod {
namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerWarp = 4;
constexpr int kRowsPerCta = (kThreads / 32) * kRowsPerWarp;  // 32
constexpr int kTileCols = 1024;

__global__ void __launch_bounds__(kThreads)
k2_adjacency_kernel(const int32_t *__restrict__ offsets, const int64_t *__restrict__ matrix_offsets,
                    const float *__restrict__ query, const float *__restrict__ train,
                    const float *__restrict__ pixels, const float *__restrict__ spans, float sensor_error,
                    uint32_t *__restrict__ physical, uint32_t *__restrict__ sample) {
  __shared__ float s_col[8][kTileCols];  // qx qy qz tx ty tz px py

  const int c = blockIdx.y;
  const int base = offsets[c];
  const int n = offsets[c + 1] - base;
  const int row_block = blockIdx.x * kRowsPerCta;
  if (row_block >= n) return;
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

  int row[kRowsPerWarp];
  float rq[kRowsPerWarp][3], rt[kRowsPerWarp][3], rp[kRowsPerWarp][2];
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    row[r] = row_block + warp * kRowsPerWarp + r;
    const int i = min(row[r], n - 1) + base;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      rq[r][d] = __ldg(query + size_t(i) * 3 + d);
      rt[r][d] = __ldg(train + size_t(i) * 3 + d);
    }
    rp[r][0] = __ldg(pixels + size_t(i) * 2);
    rp[r][1] = __ldg(pixels + size_t(i) * 2 + 1);
  }

  for (int col0 = 0; col0 < W * 32; col0 += kTileCols) {
    __syncthreads();
    for (int x = threadIdx.x; x < kTileCols; x += kThreads) {
      const int j = col0 + x;
      const bool ok = j < n;
      const size_t g = size_t(base) + (ok ? j : 0);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        s_col[d][x] = ok ? __ldg(query + g * 3 + d) : 0.f;
        s_col[3 + d][x] = ok ? __ldg(train + g * 3 + d) : 0.f;
      }
      s_col[6][x] = ok ? __ldg(pixels + g * 2) : 0.f;
      s_col[7][x] = ok ? __ldg(pixels + g * 2 + 1) : 0.f;
    }
    __syncthreads();

    uint32_t keepP[kRowsPerWarp], keepS[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) keepP[r] = keepS[r] = 0u;

    const int words_here = min(kTileCols / 32, W - col0 / 32);
    for (int w = 0; w < words_here; ++w) {
      const int x = w * 32 + lane;
      const int j = col0 + x;
      const float qx = s_col[0][x], qy = s_col[1][x], qz = s_col[2][x];
      const float tx = s_col[3][x], ty = s_col[4][x], tz = s_col[5][x];
      const float px = s_col[6][x], py = s_col[7][x];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        bool isP = false, isS = false;
        if (j < n && j != row[r]) {
          const float dx = __fsub_rn(rq[r][0], qx), dy = __fsub_rn(rq[r][1], qy), dz = __fsub_rn(rq[r][2], qz);
          const float dq2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          if (!(dq2 > thr_span)) {
            const float dq = __fsqrt_rn(dq2);
            const float ux = __fsub_rn(rt[r][0], tx), uy = __fsub_rn(rt[r][1], ty), uz = __fsub_rn(rt[r][2], tz);
            // dt in single precision first: it differs from the reference's double-accumulated norm by at most
            // 3.5 * 2^-24 * dt (three rounded products, two rounded adds, one rounded sqrt vs. one final rounding),
            // so the two comparisons below are already decided unless |dt - dq| is within `tol` (4x that bound) of
            // a threshold; only then — about one pair in 10^5 — is the reference's exact arithmetic replayed.
            float dt = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz)));
            float diff = fabsf(__fsub_rn(dt, dq));
            const float tol = __fmul_rn(1e-6f, fmaxf(dt, dq));
            if (!(fabsf(__fsub_rn(diff, e4)) > tol && fabsf(__fsub_rn(diff, e2)) > tol)) {  // also taken for NaNs
              const double vx = double(ux), vy = double(uy), vz = double(uz);
              const double s2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
              dt = __double2float_rn(__dsqrt_rn(s2));
              diff = fabsf(__fsub_rn(dt, dq));
            }
            if (!(diff > e4)) {
              isP = true;
              const float ax = __fsub_rn(rp[r][0], px), ay = __fsub_rn(rp[r][1], py);
              const float pd = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
              isS = (pd > 400.0f) && (diff < e2);
            }
          }
        }
        const uint32_t bp = __ballot_sync(0xffffffffu, isP);
        const uint32_t bs = __ballot_sync(0xffffffffu, isS);
        if (lane == w) {
          keepP[r] = bp;
          keepS[r] = bs;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      if (row[r] < n && lane < words_here) {
        const size_t o = size_t(row[r]) * W + col0 / 32 + lane;
        P[o] = keepP[r];
        S[o] = keepS[r];
      }
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
    dim3 grid((max_cluster + kRowsPerCta - 1) / kRowsPerCta, nc);
    k2_adjacency_kernel<<<grid, kThreads, 0, stream>>>(d_offsets + c0, d_matrix_offsets + c0, d_query, d_train,
                                                       d_pixels, d_spans + c0, sensor_error, d_physical, d_sample);
    count_launch();
  }
  return cudaGetLastError();
}

}  // namespace tod
