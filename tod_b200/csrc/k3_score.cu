// K3: batched RANSAC hypothesis scoring — rigid 3-point fit (Kabsch) in registers + warp-reduced inlier count.
//
// Replaces, per hypothesis, the body of pcl::RandomSampleConsensus::computeModel's loop (src/common/ransac.h:105-113
// of the reference): computeModelCoefficients -> estimateRigidTransformationSVD
// (sac_model_registration_graph.h:271-288, :304-347) and the candidate/inlier part of selectWithinDistance
// (:171-200).  One warp per 32 hypotheses (see the kernel); per hypothesis:
//   candidates = physical[s0] & physical[s1] & physical[s2] & valid            (:178-184, bit-rows instead of lists)
//   count      = #candidates (+ the 3 samples, :185-186) passing distSq(R q + T, t) < threshold^2   (:192-200)
// With the reference's never-set threshold (DBL_MAX, sac.h:70) threshold^2 is +inf and the count
// is popc(AND)+3 unless the distance is NaN/inf.  The clique gate (:203-268) stays on the host and is only evaluated
// for hypotheses that can still win.
//
// The rigid fit follows the reference's arithmetic where it matters (float centroids scaled by 1.f/3, float
// differences, double-accumulated correlation matrix rounded to float).  The 3x3 SVD is a Jacobi eigen-solve of
// H^T H in double; R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T, which equals the reference's
// "U * Vt with row 2 of Vt negated when det(U) det(Vt) < 0" (:337-343) for any sign convention of the SVD.
#include "tod_internal.h"

namespace tod {
namespace {

struct Sym3 {
  double a00, a01, a02, a11, a12, a22;
};

// Cyclic Jacobi on a symmetric 3x3 (double).  V columns = eigenvectors.
__device__ __forceinline__ void jacobi_rotate(double &app, double &aqq, double &apq, double &arp, double &arq,
                                              double (&V)[3][3], int p, int q) {
  if (apq == 0.0) return;
  const double theta = (aqq - app) / (2.0 * apq);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
  app -= t * apq;
  aqq += t * apq;
  apq = 0.0;
  const double rp = arp, rq = arq;
  arp = c * rp - s * rq;
  arq = s * rp + c * rq;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double vp = V[k][p], vq = V[k][q];
    V[k][p] = c * vp - s * vq;
    V[k][q] = s * vp + c * vq;
  }
}

__device__ void kabsch_from_H(const float (&Hf)[3][3], float (&R)[9]) {
  double H[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) H[i][j] = double(Hf[i][j]);
  // A = H^T H
  Sym3 A;
  A.a00 = H[0][0] * H[0][0] + H[1][0] * H[1][0] + H[2][0] * H[2][0];
  A.a01 = H[0][0] * H[0][1] + H[1][0] * H[1][1] + H[2][0] * H[2][1];
  A.a02 = H[0][0] * H[0][2] + H[1][0] * H[1][2] + H[2][0] * H[2][2];
  A.a11 = H[0][1] * H[0][1] + H[1][1] * H[1][1] + H[2][1] * H[2][1];
  A.a12 = H[0][1] * H[0][2] + H[1][1] * H[1][2] + H[2][1] * H[2][2];
  A.a22 = H[0][2] * H[0][2] + H[1][2] * H[1][2] + H[2][2] * H[2][2];
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A.a01) + fabs(A.a02) + fabs(A.a12);
    if (off <= 1e-300 || off <= 1e-22 * (fabs(A.a00) + fabs(A.a11) + fabs(A.a22))) break;
    jacobi_rotate(A.a00, A.a11, A.a01, A.a02, A.a12, V, 0, 1);
    jacobi_rotate(A.a00, A.a22, A.a02, A.a01, A.a12, V, 0, 2);
    jacobi_rotate(A.a11, A.a22, A.a12, A.a01, A.a02, V, 1, 2);
  }
  // two largest eigenvalues -> v1, v2
  double ev[3] = {A.a00, A.a11, A.a22};
  int i1 = 0;
  if (ev[1] > ev[i1]) i1 = 1;
  if (ev[2] > ev[i1]) i1 = 2;
  int i2 = (i1 == 0) ? 1 : 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k != i1 && ev[k] > ev[i2]) i2 = k;
  double v1[3] = {V[0][i1], V[1][i1], V[2][i1]};
  double v2[3] = {V[0][i2], V[1][i2], V[2][i2]};
  double u1[3], u2[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    u1[r] = H[r][0] * v1[0] + H[r][1] * v1[1] + H[r][2] * v1[2];
    u2[r] = H[r][0] * v2[0] + H[r][1] * v2[1] + H[r][2] * v2[2];
  }
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  if (n1 > 0.0) {
    u1[0] /= n1; u1[1] /= n1; u1[2] /= n1;
  } else {  // H == 0: any rotation is a solution; pick identity-compatible axes
    u1[0] = v1[0]; u1[1] = v1[1]; u1[2] = v1[2];
  }
  // Gram-Schmidt u2 against u1 (they are orthogonal in exact arithmetic)
  const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  u2[0] -= d12 * u1[0]; u2[1] -= d12 * u1[1]; u2[2] -= d12 * u1[2];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  if (n2 > 1e-12 * n1 && n2 > 0.0) {
    u2[0] /= n2; u2[1] /= n2; u2[2] /= n2;
  } else {
    // rank <= 1 (collinear sample): complete u2 with any unit vector orthogonal to u1 — stays finite like cv::SVD
    int m = 0;
    if (fabs(u1[1]) < fabs(u1[m])) m = 1;
    if (fabs(u1[2]) < fabs(u1[m])) m = 2;
    double e[3] = {0, 0, 0};
    e[m] = 1.0;
    const double d = u1[m];
    u2[0] = e[0] - d * u1[0]; u2[1] = e[1] - d * u1[1]; u2[2] = e[2] - d * u1[2];
    n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    u2[0] /= n2; u2[1] /= n2; u2[2] /= n2;
  }
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = float(u1[r] * v1[c] + u2[r] * v2[c] + u3[r] * v3[c]);
}

__device__ __forceinline__ bool within(const float (&R)[9], const float (&T)[3], const float *__restrict__ q,
                                       const float *__restrict__ t, double thr2) {
  // distSq(R * pt_src + T, pt_tgt) < threshold * threshold   (sac_model_registration_graph.h:198)
  const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
  const float px = (R[0] * qx + R[1] * qy + R[2] * qz) + T[0];
  const float py = (R[3] * qx + R[4] * qy + R[5] * qz) + T[1];
  const float pz = (R[6] * qx + R[7] * qy + R[8] * qz) + T[2];
  const float dx = px - __ldg(t), dy = py - __ldg(t + 1), dz = pz - __ldg(t + 2);
  const float d2 = dx * dx + dy * dy + dz * dz;
  return double(d2) < thr2;
}

// hyp[h] = (s0, s1, s2, cluster); clusters are described by K3Cluster records (tod_internal.h).

// One warp scores 32 consecutive hypotheses.
//   phase 1: lane l fits ITS OWN hypothesis h0 + l (32 different Kabsch solves per warp instruction instead of the same
//            one 32 times — the double-precision Jacobi is the expensive part of this kernel);
//   phase 2: for each of the 32 hypotheses in turn, the whole warp ANDs the three physical rows with the valid (and
//            finite) masks, 128 coalesced bytes per row and step, and reduces the popcounts with one REDUX; with a
//            finite threshold the owner's (R, T) is broadcast and the set bits are distance-tested, one per lane.
__global__ void __launch_bounds__(256)
k3_score_kernel(const K3Cluster *__restrict__ clusters, const float *__restrict__ query,
                const float *__restrict__ train, const uint32_t *__restrict__ physical,
                const uint32_t *__restrict__ valid, const uint32_t *__restrict__ finite, int n_hyp,
                const uint4 *__restrict__ hyps, double thr2,
                int exact_inf, int32_t *__restrict__ counts, float *__restrict__ Rout, float *__restrict__ Tout) {
  const int lane = threadIdx.x & 31;
  const int h0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (h0 >= n_hyp) return;
  const int h = h0 + lane;
  const bool mine = h < n_hyp;
  const uint4 hy = mine ? __ldg(hyps + h) : make_uint4(0, 0, 0, 0);
  const K3Cluster cl = clusters[mine ? hy.w : __ldg(hyps + h0).w];

  // ---- phase 1: rigid fit of this lane's own 3 samples ----
  float R[9], T[3];
  bool rt_finite = true;
  int add = 0;  // the 3 samples themselves (sac_model_registration_graph.h:185-186, :192-200)
  if (mine) {
    const float *q = query + cl.point_offset * 3;
    const float *t = train + cl.point_offset * 3;
    const uint32_t s[3] = {hy.x, hy.y, hy.z};
    float ct[3] = {0.f, 0.f, 0.f}, cq[3] = {0.f, 0.f, 0.f};
    float st[3][3], sq[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        st[i][d] = __ldg(t + size_t(s[i]) * 3 + d);
        sq[i][d] = __ldg(q + size_t(s[i]) * 3 + d);
        ct[d] += st[i][d];
        cq[d] += sq[i][d];
      }
    const float third = 1.f / 3.f;  // cv::Vec operator/= multiplies by 1.f/alpha
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      ct[d] *= third;
      cq[d] *= third;
    }
    // Reference-faithful mode (threshold^2 == +inf) without pose output: the count is popc(AND) + 3 unless (R, T) is
    // non-finite, and with finite samples of ordinary magnitude every step of the fit stays finite (the SVD of a
    // rank-deficient H does too, like cv::SVD) — so the double-precision Jacobi solve is skipped.  Samples with a
    // non-finite or astronomically large coordinate (where float products could overflow) take the full path.
    bool tame = true;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) tame = tame && fabsf(st[i][d]) < 1e15f && fabsf(sq[i][d]) < 1e15f;
    const bool need_fit = !exact_inf || Rout != nullptr || Tout != nullptr || !tame;
    if (need_fit) {
    float Hf[3][3];
    {
      double Hd[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float a[3], b[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          a[d] = st[i][d] - ct[d];
          b[d] = sq[i][d] - cq[d];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) Hd[r][c] += double(a[r]) * double(b[c]);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Hf[r][c] = float(Hd[r][c]);
    }
    kabsch_from_H(Hf, R);
#pragma unroll
    for (int d = 0; d < 3; ++d) T[d] = ct[d] - (R[d * 3] * cq[0] + R[d * 3 + 1] * cq[1] + R[d * 3 + 2] * cq[2]);
#pragma unroll
    for (int i = 0; i < 9; ++i) rt_finite = rt_finite && isfinite(R[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) rt_finite = rt_finite && isfinite(T[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) T[i] = 0.f;
    }
    // finite[i] = all six coordinates of correspondence i are finite.  With threshold^2 == +inf the reference's test
    // `distSq < inf` only fails for NaN/inf distances, i.e. for non-finite points or a non-finite (R, T).
    const uint32_t *F = finite ? finite + cl.valid_offset : nullptr;
    if (exact_inf) {
#pragma unroll
      for (int i = 0; i < 3; ++i) add += F ? int((__ldg(F + (s[i] >> 5)) >> (s[i] & 31)) & 1u) : 1;
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) add += within(R, T, q + size_t(s[i]) * 3, t + size_t(s[i]) * 3, thr2) ? 1 : 0;
    }
    if (Rout) {
#pragma unroll
      for (int i = 0; i < 9; ++i) Rout[size_t(h) * 9 + i] = R[i];
    }
    if (Tout) {
#pragma unroll
      for (int i = 0; i < 3; ++i) Tout[size_t(h) * 3 + i] = T[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) T[i] = 0.f;
  }

  // ---- phase 2: inlier count over the common physical neighbourhood, hypothesis by hypothesis ----
  int my_cnt = 0;
  const int n_here = min(32, n_hyp - h0);
  for (int j = 0; j < n_here; ++j) {
    const uint32_t s0 = __shfl_sync(0xffffffffu, hy.x, j), s1 = __shfl_sync(0xffffffffu, hy.y, j),
                   s2 = __shfl_sync(0xffffffffu, hy.z, j), ci = __shfl_sync(0xffffffffu, hy.w, j);
    const K3Cluster cj = clusters[ci];
    const uint32_t *P = physical + cj.matrix_offset;
    const uint32_t *V = valid + cj.valid_offset;
    const uint32_t *F = finite ? finite + cj.valid_offset : nullptr;
    const uint32_t *r0 = P + size_t(s0) * cj.W, *r1 = P + size_t(s1) * cj.W, *r2 = P + size_t(s2) * cj.W;
    int cnt = 0;
    if (exact_inf) {
      for (int w = lane; w < cj.W; w += 32) {
        uint32_t m = __ldg(r0 + w) & __ldg(r1 + w) & __ldg(r2 + w) & __ldg(V + w);
        if (F) m &= __ldg(F + w);
        cnt += __popc(m);
      }
    } else {
      float Rj[9], Tj[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rj[i] = __shfl_sync(0xffffffffu, R[i], j);
#pragma unroll
      for (int i = 0; i < 3; ++i) Tj[i] = __shfl_sync(0xffffffffu, T[i], j);
      const float *q = query + cj.point_offset * 3;
      const float *t = train + cj.point_offset * 3;
      for (int w = lane; w < cj.W; w += 32) {
        uint32_t m = __ldg(r0 + w) & __ldg(r1 + w) & __ldg(r2 + w) & __ldg(V + w);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int idx = w * 32 + b;
          cnt += within(Rj, Tj, q + size_t(idx) * 3, t + size_t(idx) * 3, thr2) ? 1 : 0;
        }
      }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == j) my_cnt = cnt;
  }
  if (mine) {
    if (exact_inf && !rt_finite) {
      my_cnt = 0;
      add = 0;
    }
    counts[h] = my_cnt + add;
  }
}

}  // namespace

cudaError_t launch_score_hypotheses_batched(const void *d_clusters, const float *d_query, const float *d_train,
                                            const uint32_t *d_physical, const uint32_t *d_valid,
                                            const uint32_t *d_finite, int n_hyp, const uint32_t *d_hyps, double threshold, int32_t *d_counts, float *d_R,
                                            float *d_T, cudaStream_t stream) {
  if (n_hyp <= 0) return cudaSuccess;
  const bool inf = !(threshold < 1e150);  // DBL_MAX * DBL_MAX == +inf in the reference
  const double thr2 = inf ? __builtin_inf() : threshold * threshold;
  const int warps_per_cta = 8;  // a warp scores 32 hypotheses
  const int blocks = (n_hyp + warps_per_cta * 32 - 1) / (warps_per_cta * 32);
  k3_score_kernel<<<blocks, warps_per_cta * 32, 0, stream>>>(
      static_cast<const K3Cluster *>(d_clusters), d_query, d_train, d_physical, d_valid, d_finite, n_hyp,
      reinterpret_cast<const uint4 *>(d_hyps), thr2, inf ? 1 : 0, d_counts, d_R, d_T);
  count_launch();
  return cudaGetLastError();
}

size_t k3_cluster_desc_size() { return sizeof(K3Cluster); }

void k3_fill_cluster_desc(void *dst, int32_t n, int32_t W, int64_t point_offset, int64_t matrix_offset,
                          int64_t valid_offset) {
  K3Cluster c{n, W, point_offset, matrix_offset, valid_offset};
  *static_cast<K3Cluster *>(dst) = c;
}

}  // namespace tod
