// Host-side geometry helpers of the guess generator: m-point rigid fit (used by the final refinement, which north_star
// leaves on the host) and the point-to-model distance tests, with the reference's operation order and precisions.
//
// Reference: SampleConsensusModelRegistrationGraph::estimateRigidTransformationSVD
// (src/common/sac_model_registration_graph.h:304-347) and the refinement loop of AdjacencyRansac::Ransac
// (src/common/adjacency_ransac.cpp:266-303).
#ifndef TOD_HOST_GEOMETRY_H_
#define TOD_HOST_GEOMETRY_H_

#include <algorithm>
#include <cmath>
#include <cstdint>

namespace tod {

// R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T from the two dominant singular pairs of H (3x3, row-major float).
// Equals the reference's U * Vt with row 2 of Vt negated when det(U) det(Vt) < 0 (:337-343), for every sign
// convention an SVD may choose.  One-sided (Hestenes) Jacobi in double.
inline void rotation_from_correlation(const float H[9], float R[9]) {
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A[i][j] = double(H[i * 3 + j]);
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, b = 0, g = 0;
        for (int k = 0; k < 3; ++k) {
          a += A[k][p] * A[k][p];
          b += A[k][q] * A[k][q];
          g += A[k][p] * A[k][q];
        }
        if (std::fabs(g) <= 1e-300 || std::fabs(g) <= 2.3e-16 * std::sqrt(a * b)) continue;
        rotated = true;
        const double zeta = (b - a) / (2.0 * g);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          const double x = A[k][p], y = A[k][q];
          A[k][p] = c * x - s * y;
          A[k][q] = s * x + c * y;
          const double vx = V[k][p], vy = V[k][q];
          V[k][p] = c * vx - s * vy;
          V[k][q] = s * vx + c * vy;
        }
      }
    if (!rotated) break;
  }
  double sv[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) sv[j] = std::sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  std::sort(ord, ord + 3, [&](int x, int y) { return sv[x] > sv[y]; });
  const int j1 = ord[0], j2 = ord[1];
  double u1[3], u2[3], v1[3], v2[3];
  for (int k = 0; k < 3; ++k) {
    v1[k] = V[k][j1];
    v2[k] = V[k][j2];
  }
  if (sv[j1] > 0.0) {
    for (int k = 0; k < 3; ++k) u1[k] = A[k][j1] / sv[j1];
  } else {
    for (int k = 0; k < 3; ++k) u1[k] = v1[k];
  }
  if (sv[j2] > 1e-12 * sv[j1] && sv[j2] > 0.0) {
    for (int k = 0; k < 3; ++k) u2[k] = A[k][j2] / sv[j2];
  } else {  // rank <= 1: any unit vector orthogonal to u1 keeps the result finite
    int m = 0;
    if (std::fabs(u1[1]) < std::fabs(u1[m])) m = 1;
    if (std::fabs(u1[2]) < std::fabs(u1[m])) m = 2;
    double e[3] = {0, 0, 0};
    e[m] = 1.0;
    const double d = u1[m];
    double nn = 0;
    for (int k = 0; k < 3; ++k) {
      u2[k] = e[k] - d * u1[k];
      nn += u2[k] * u2[k];
    }
    nn = std::sqrt(nn);
    for (int k = 0; k < 3; ++k) u2[k] /= nn;
  }
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = float(u1[r] * v1[c] + u2[r] * v2[c] + u3[r] * v3[c]);
}

// p = R * q + T with cv::Matx's float accumulation order.
inline void transform_point(const float R[9], const float T[3], const float *q, float p[3]) {
  for (int r = 0; r < 3; ++r) {
    float s = 0.f;
    s += R[r * 3 + 0] * q[0];
    s += R[r * 3 + 1] * q[1];
    s += R[r * 3 + 2] * q[2];
    p[r] = s + T[r];
  }
}

// estimateRigidTransformationSVD over `m` correspondences listed in idx (query -> training frame).
inline void rigid_fit(const float *query, const float *train, const uint32_t *idx, int m, float R[9], float T[3]) {
  float ct[3] = {0.f, 0.f, 0.f}, cq[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < m; ++i) {  // :312-315
    const float *t = train + size_t(idx[i]) * 3, *q = query + size_t(idx[i]) * 3;
    for (int d = 0; d < 3; ++d) {
      ct[d] += t[d];
      cq[d] += q[d];
    }
  }
  const float inv = 1.f / float(m);  // cv::Vec operator/=(float) multiplies by the reciprocal
  for (int d = 0; d < 3; ++d) {
    ct[d] *= inv;
    cq[d] *= inv;
  }
  double Hd[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // :330 — OpenCV's gemm accumulates this transposed product in double
  for (int i = 0; i < m; ++i) {
    const float *t = train + size_t(idx[i]) * 3, *q = query + size_t(idx[i]) * 3;
    float a[3], b[3];
    for (int d = 0; d < 3; ++d) {
      a[d] = t[d] - ct[d];
      b[d] = q[d] - cq[d];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Hd[r * 3 + c] += double(a[r]) * double(b[c]);
  }
  float H[9];
  for (int i = 0; i < 9; ++i) H[i] = float(Hd[i]);
  rotation_from_correlation(H, R);
  float rc[3];
  const float zero[3] = {0.f, 0.f, 0.f};
  transform_point(R, zero, cq, rc);
  for (int d = 0; d < 3; ++d) T[d] = ct[d] - rc[d];  // :344
}

}  // namespace tod
#endif
