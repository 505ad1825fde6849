// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// Minimal stand-in for the 12 cv:: symbols the reference's src/common/*.{h,cpp} use, so that those files compile
// UNMODIFIED from /root/reference into oracle/_ref/libtod_ref.so (OpenCV's C++ headers are not installed here).
// Only the arithmetic that influences results is modelled, following OpenCV's published behaviour:
//   * Vec /= float multiplies by 1.f/alpha (matx.hpp);  Matx * Vec accumulates left to right in float;
//   * cv::norm(Vec3f) = sqrt of a DOUBLE sum of squares (normL2Sqr<float,double>), SURVEY.md quirk Q8;
//   * Mat * Mat (gemm): the n x 3 transposed product accumulates in double and rounds to float; a 3x3 * 3x3 float
//     product takes the small-matrix float path (both verified against cv2 4.13);
//   * cv::SVD of a 3x3: singular values descending; computed here by one-sided Jacobi in double then rounded to float
//     (OpenCV runs the same algorithm in float) — cross-checked against cv2.SVDecomp in tests/test_oracle_geometry.py;
//   * rand() is routed to tod_oracle_rand() so the harness controls the sampler stream
//     (sac_model_registration_graph.h:111 calls plain rand()).
#pragma once
#include <algorithm>
#include <climits>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

extern "C" int tod_oracle_rand(void);
#define rand tod_oracle_rand

inline int cvIsNaN(double v) { return std::isnan(v) ? 1 : 0; }

namespace cv {

template <typename T, int cn>
struct Vec {
  T val[cn];
  Vec() { for (int i = 0; i < cn; ++i) val[i] = T(0); }
  Vec(T a, T b, T c) { static_assert(cn == 3, "3 channels"); val[0] = a; val[1] = b; val[2] = c; }
  T &operator[](int i) { return val[i]; }
  const T &operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;

template <typename T, int cn> inline Vec<T, cn> operator-(const Vec<T, cn> &a, const Vec<T, cn> &b) {
  Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = a.val[i] - b.val[i]; return r;
}
template <typename T, int cn> inline Vec<T, cn> operator+(const Vec<T, cn> &a, const Vec<T, cn> &b) {
  Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = a.val[i] + b.val[i]; return r;
}
template <typename T, int cn> inline Vec<T, cn> &operator+=(Vec<T, cn> &a, const Vec<T, cn> &b) {
  for (int i = 0; i < cn; ++i) a.val[i] = a.val[i] + b.val[i]; return a;
}
template <typename T, int cn> inline Vec<T, cn> &operator/=(Vec<T, cn> &a, float alpha) {
  float ialpha = 1.f / alpha; for (int i = 0; i < cn; ++i) a.val[i] = T(a.val[i] * ialpha); return a;
}
inline double norm(const Vec3f &v) {
  double s = 0; for (int i = 0; i < 3; ++i) { double x = v.val[i]; s += x * x; } return std::sqrt(s);
}

struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct KeyPoint {
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct DMatch {
  int queryIdx, trainIdx, imgIdx; float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
};
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };

class Mat;
template <typename T, int m, int n> struct Matx;

// Dense row-major float matrix with `ch` interleaved channels (enough for CV_32FC1 / CV_32FC3).
class Mat {
 public:
  int rows, cols, ch;
  bool texpr = false;  // produced by .t(): a product with it carries OpenCV's GEMM_1_T flag -> general (double) path
  std::vector<float> d;
  Mat() : rows(0), cols(0), ch(1) {}
  Mat(int r, int c, int channels = 1) : rows(r), cols(c), ch(channels), d(size_t(r) * c * channels, 0.f) {}
  template <int m, int n> Mat(const Matx<float, m, n> &M);
  Mat(const Vec3f &v) : rows(3), cols(1), ch(1), d(v.val, v.val + 3) {}
  bool empty() const { return d.empty(); }
  Mat clone() const { return *this; }
  void copyTo(Mat &o) const { o = *this; }
  template <typename V> V &at(int i, int j) { return *reinterpret_cast<V *>(&d[(size_t(i) * cols + j) * ch]); }
  template <typename V> const V &at(int i, int j) const { return *reinterpret_cast<const V *>(&d[(size_t(i) * cols + j) * ch]); }
  Mat reshape(int cn, int r) const {  // same data, new channel count / row count
    Mat o; o.ch = cn; o.rows = r; o.cols = int(d.size() / (size_t(cn) * r)); o.d = d; return o;
  }
  Mat t() const {
    Mat o(cols, rows, 1);
    for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) o.d[size_t(j) * rows + i] = d[size_t(i) * cols + j];
    o.texpr = true;
    return o;
  }
};

inline Mat operator*(const Mat &A, const Mat &B) {
  Mat C(A.rows, B.cols, 1);
  // small-matrix float path only for plain (flags == 0) products with 2 <= len <= 4 (OpenCV gemm)
  const bool small = !A.texpr && !B.texpr && A.cols >= 2 && A.cols <= 4 && (A.cols == B.cols || A.cols == A.rows);
  for (int i = 0; i < A.rows; ++i)
    for (int j = 0; j < B.cols; ++j) {
      if (small) {  // OpenCV gemm small-matrix path: float accumulation
        float s = 0.f;
        for (int k = 0; k < A.cols; ++k) s = (k == 0) ? A.d[size_t(i) * A.cols] * B.d[j] : s + A.d[size_t(i) * A.cols + k] * B.d[size_t(k) * B.cols + j];
        C.d[size_t(i) * B.cols + j] = s;
      } else {      // general path: double accumulation, rounded to float on store
        double s = 0;
        for (int k = 0; k < A.cols; ++k) s += double(A.d[size_t(i) * A.cols + k]) * double(B.d[size_t(k) * B.cols + j]);
        C.d[size_t(i) * B.cols + j] = float(s);
      }
    }
  return C;
}

template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c);
  Mat_(const Mat &m) : Mat(m) {}
  T &operator()(int i) { return *reinterpret_cast<T *>(&d[size_t(i) * ch]); }
  T &operator()(int i, int j) { return *reinterpret_cast<T *>(&d[(size_t(i) * cols + j) * ch]); }
};
template <> inline Mat_<Vec3f>::Mat_(int r, int c) : Mat(r, c, 3) {}
template <> inline Mat_<float>::Mat_(int r, int c) : Mat(r, c, 1) {}

template <typename T, int m, int n>
struct Matx {
  T val[m * n];
  Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
  Matx(const Mat &M) { for (int i = 0; i < m * n; ++i) val[i] = M.d[i]; }
  T &operator()(int i, int j) { return val[i * n + j]; }
  const T &operator()(int i, int j) const { return val[i * n + j]; }
  Matx<T, n, m> t() const { Matx<T, n, m> r; for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) r.val[j * m + i] = val[i * n + j]; return r; }
};
typedef Matx<float, 3, 3> Matx33f;
template <int m, int n> inline Mat::Mat(const Matx<float, m, n> &M) : rows(m), cols(n), ch(1), d(M.val, M.val + m * n) {}

inline Matx33f operator-(const Matx33f &a) { Matx33f r; for (int i = 0; i < 9; ++i) r.val[i] = a.val[i] * -1.f; return r; }
inline Vec3f operator*(const Matx33f &a, const Vec3f &b) {
  Vec3f r;
  for (int i = 0; i < 3; ++i) { float s = 0; for (int k = 0; k < 3; ++k) s += a(i, k) * b.val[k]; r.val[i] = s; }
  return r;
}

inline double determinant(const Mat &M) {
  const float *a = M.d.data();
  return a[0] * (double(a[4]) * a[8] - double(a[5]) * a[7]) - a[1] * (double(a[3]) * a[8] - double(a[5]) * a[6]) +
         a[2] * (double(a[3]) * a[7] - double(a[4]) * a[6]);
}

// One-sided Jacobi SVD (Hestenes), the algorithm of cv::JacobiSVD: A = U diag(w) Vt, w descending.
class SVD {
 public:
  Mat u, w, vt;
  explicit SVD(const Mat &A0) {
    const int n = 3;
    double A[3][3], V[3][3];
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { A[i][j] = A0.d[size_t(i) * 3 + j]; V[i][j] = (i == j); }
    // columns of A are rotated until mutually orthogonal
    for (int sweep = 0; sweep < 60; ++sweep) {
      bool changed = false;
      for (int p = 0; p < n - 1; ++p)
        for (int q = p + 1; q < n; ++q) {
          double a = 0, b = 0, g = 0;
          for (int k = 0; k < n; ++k) { a += A[k][p] * A[k][p]; b += A[k][q] * A[k][q]; g += A[k][p] * A[k][q]; }
          if (std::fabs(g) <= 1e-300 || std::fabs(g) <= 2.3e-16 * std::sqrt(a * b)) continue;
          changed = true;
          const double zeta = (b - a) / (2.0 * g);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
          for (int k = 0; k < n; ++k) {
            const double x = A[k][p], y = A[k][q];
            A[k][p] = c * x - s * y; A[k][q] = s * x + c * y;
            const double vx = V[k][p], vy = V[k][q];
            V[k][p] = c * vx - s * vy; V[k][q] = s * vx + c * vy;
          }
        }
      if (!changed) break;
    }
    double sv[3]; int order[3] = {0, 1, 2};
    for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += A[k][j] * A[k][j]; sv[j] = std::sqrt(s); }
    std::sort(order, order + 3, [&](int x, int y) { return sv[x] > sv[y]; });
    double U[3][3];
    for (int jj = 0; jj < n; ++jj) {
      const int j = order[jj];
      if (sv[j] > 1e-300 && sv[j] > 1e-14 * sv[order[0]]) { for (int k = 0; k < n; ++k) U[k][jj] = A[k][j] / sv[j]; }
      else { for (int k = 0; k < n; ++k) U[k][jj] = 0; }  // completed below
    }
    // complete U to an orthonormal basis for (numerically) zero singular values, as cv::JacobiSVD does
    for (int jj = 0; jj < n; ++jj) {
      double nn = 0; for (int k = 0; k < n; ++k) nn += U[k][jj] * U[k][jj];
      if (nn > 0.5) continue;
      for (int e = 0; e < n; ++e) {
        double v[3] = {0, 0, 0}; v[e] = 1;
        for (int c2 = 0; c2 < n; ++c2) {
          if (c2 == jj) continue;
          double n2 = 0, dp = 0; for (int k = 0; k < n; ++k) { n2 += U[k][c2] * U[k][c2]; dp += U[k][c2] * v[k]; }
          if (n2 > 0.5) for (int k = 0; k < n; ++k) v[k] -= dp * U[k][c2];
        }
        double vn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (vn > 0.3) { for (int k = 0; k < n; ++k) U[k][jj] = v[k] / vn; break; }
      }
    }
    u = Mat(3, 3); vt = Mat(3, 3); w = Mat(3, 1);
    for (int jj = 0; jj < n; ++jj) {
      w.d[jj] = float(sv[order[jj]]);
      for (int k = 0; k < n; ++k) { u.d[size_t(k) * 3 + jj] = float(U[k][jj]); vt.d[size_t(jj) * 3 + k] = float(V[k][order[jj]]); }
    }
  }
};

inline void drawKeypoints(const Mat &, const std::vector<KeyPoint> &, Mat &, const Scalar &) {}
inline void namedWindow(const std::string &, int = 0) {}
inline void imshow(const std::string &, const Mat &) {}

}  // namespace cv
