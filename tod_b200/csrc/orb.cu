// Feature stage in front of the hot path: ORB orientation + descriptors and depth -> 3-D on the
// GPU, so that the descriptors K1 consumes never leave HBM.
//
// Replaces, for keypoints that are already detected, the `FeatureDescriptor` cell (cv::ORB, n_features 5000, n_levels 3,
// scale_factor 1.2 — python/object_recognition_tod/detector.py:27,74; conf/detection.ork:23-31) and the `DepthTo3d`
// cell (detector.py:62-69).  OpenCV is un-vendored; the arithmetic follows oracle/orb.py, which restates what the
// cv2 4.13.0 binary does, stage by stage, and is pinned bit for bit against it (tests/test_orb_oracle.py):
//   pyramid      INTER_LINEAR_EXACT: two passes with 8-bit fixed-point weights (tables built on the host in double)
//   smoothing    7 x 7 Gaussian, sigma 2, BORDER_REFLECT_101, OpenCV's float separable filter: row pass an FMA chain
//                left to right, column pass centre first then fma(x[+j] + x[-j], k[j], s); round half to even
//   orientation  intensity centroid over the radius-15 disc, cv::fastAtan2's polynomial without contraction
//   descriptor   steered BRIEF, 256 comparisons at pattern points rotated by the angle (float products, cvRound)
// Every float operation that decides a bit is written with explicit rounding intrinsics (__fmul_rn / __fadd_rn /
// __fmaf_rn): the compiler may neither fuse nor split them.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "orb_pattern.h"
#include "tod_internal.h"

namespace tod {
namespace {

constexpr int kHalfPatch = 15;
constexpr int kMaxLevels = 8;
constexpr int kDescBorder = 22;  // the rotated pattern reaches ceil(15 * sqrt(2)) pixels from the centre

__constant__ signed char c_pattern[256][4];
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
// cv::getGaussianKernel(7, 2, CV_32F): centre, +-1, +-2, +-3
__constant__ float c_gauss[4] = {0x1.ba95cp-3f, 0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void __launch_bounds__(256)
resize_exact_kernel(const uint8_t *__restrict__ src, int sw, uint8_t *__restrict__ dst, int dw, int dh,
                    const int *__restrict__ x0, const int *__restrict__ x1, const int *__restrict__ fx,
                    const int *__restrict__ y0, const int *__restrict__ y1, const int *__restrict__ fy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const int a0 = x0[x], a1 = x1[x], wx = fx[x], b0 = y0[y], b1 = y1[y], wy = fy[y];
  const int h0 = int(src[size_t(b0) * sw + a0]) * (256 - wx) + int(src[size_t(b0) * sw + a1]) * wx;
  const int h1 = int(src[size_t(b1) * sw + a0]) * (256 - wx) + int(src[size_t(b1) * sw + a1]) * wx;
  dst[size_t(y) * dw + x] = uint8_t((h0 * (256 - wy) + h1 * wy + (1 << 15)) >> 16);
}

__global__ void __launch_bounds__(256)
smooth_rows_kernel(const uint8_t *__restrict__ src, int w, int h, float *__restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const uint8_t *row = src + size_t(y) * w;
  // s = k[0] x[0]; s = fma(x[j], k[j], s), taps left to right (k[0] = the +-3 coefficient)
  const float k3 = c_gauss[3], k2 = c_gauss[2], k1 = c_gauss[1], k0 = c_gauss[0];
  float s = __fmul_rn(float(row[reflect101(x - 3, w)]), k3);
  s = __fmaf_rn(float(row[reflect101(x - 2, w)]), k2, s);
  s = __fmaf_rn(float(row[reflect101(x - 1, w)]), k1, s);
  s = __fmaf_rn(float(row[x]), k0, s);
  s = __fmaf_rn(float(row[reflect101(x + 1, w)]), k1, s);
  s = __fmaf_rn(float(row[reflect101(x + 2, w)]), k2, s);
  s = __fmaf_rn(float(row[reflect101(x + 3, w)]), k3, s);
  out[size_t(y) * w + x] = s;
}

__global__ void __launch_bounds__(256)
smooth_cols_kernel(const float *__restrict__ src, int w, int h, uint8_t *__restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  float s = __fmul_rn(src[size_t(y) * w + x], c_gauss[0]);
#pragma unroll
  for (int j = 1; j <= 3; ++j) {
    const float a = src[size_t(reflect101(y + j, h)) * w + x], b = src[size_t(reflect101(y - j, h)) * w + x];
    s = __fmaf_rn(__fadd_rn(a, b), c_gauss[j], s);
  }
  out[size_t(y) * w + x] = uint8_t(min(255, max(0, __float2int_rn(s))));
}

struct LevelDesc {
  const uint8_t *img;     // unsmoothed level
  const uint8_t *smooth;  // smoothed level
  int w, h;
  float inv_scale;        // 1.f / layerScale
};
struct Levels {
  LevelDesc l[kMaxLevels];
};

// cv::fastAtan2 (degrees), operation by operation
__device__ float fast_atan2_deg(float y, float x) {
  const float p1 = 0x1.ca44dep+5f, p3 = -0x1.2aaddcp+4f, p5 = 0x1.1d3f7ep+3f, p7 = -0x1.4515b2p+1f;
  const float eps = 0x1p-52f;
  const float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2),
                                            p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

// One warp per keypoint: lane v sums row +-v of the disc (integers, exact), lane 0 turns the moments into the angle.
__global__ void __launch_bounds__(256)
orb_angle_kernel(Levels lv, const float *__restrict__ kx, const float *__restrict__ ky, const int *__restrict__ octave,
                 int n, float *__restrict__ angle) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const LevelDesc L = lv.l[octave[i]];
  const int cx = __float2int_rn(__fmul_rn(kx[i], L.inv_scale)), cy = __float2int_rn(__fmul_rn(ky[i], L.inv_scale));
  int m10 = 0, m01 = 0;
  if (lane <= kHalfPatch) {
    const int v = lane, d = c_umax[v];
    const uint8_t *rp = L.img + size_t(cy + v) * L.w + cx, *rm = L.img + size_t(cy - v) * L.w + cx;
    if (v == 0) {
      for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * int(rp[u]);
    } else {
      int vs = 0;
      for (int u = -d; u <= d; ++u) {
        const int p = rp[u], q = rm[u];
        vs += p - q;
        m10 += u * (p + q);
      }
      m01 = v * vs;
    }
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  if (lane == 0) angle[i] = fast_atan2_deg(float(m01), float(m10));
}

// One warp per keypoint, one descriptor byte per lane.
__global__ void __launch_bounds__(256)
orb_describe_kernel(Levels lv, const float *__restrict__ kx, const float *__restrict__ ky,
                    const int *__restrict__ octave, const float *__restrict__ angle, int n,
                    uint8_t *__restrict__ desc) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const LevelDesc L = lv.l[octave[i]];
  const int cx = __float2int_rn(__fmul_rn(kx[i], L.inv_scale)), cy = __float2int_rn(__fmul_rn(ky[i], L.inv_scale));
  // a = cosf(angle * pi/180), b = sinf(...): evaluated in double and rounded once, which is the correctly rounded
  // float — what the host libm returns
  const float rad = __fmul_rn(angle[i], 0x1.1df46ap-6f);
  const float a = float(cos(double(rad))), b = float(sin(double(rad)));
  const uint8_t *center = L.smooth + size_t(cy) * L.w + cx;
  unsigned byte = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const signed char *p = c_pattern[lane * 8 + t];
    const float xa = p[0], ya = p[1], xb = p[2], yb = p[3];
    const int ixa = __float2int_rn(__fsub_rn(__fmul_rn(xa, a), __fmul_rn(ya, b)));
    const int iya = __float2int_rn(__fadd_rn(__fmul_rn(xa, b), __fmul_rn(ya, a)));
    const int ixb = __float2int_rn(__fsub_rn(__fmul_rn(xb, a), __fmul_rn(yb, b)));
    const int iyb = __float2int_rn(__fadd_rn(__fmul_rn(xb, b), __fmul_rn(yb, a)));
    const int va = center[iya * L.w + ixa], vb = center[iyb * L.w + ixb];
    byte |= unsigned(va < vb) << t;
  }
  desc[size_t(i) * 32 + lane] = uint8_t(byte);
}

template <typename T>
__global__ void __launch_bounds__(256)
depth_to_3d_kernel(const T *__restrict__ depth, int w, int h, float fx, float fy, float cx, float cy,
                   float *__restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  float z;
  if (sizeof(T) == 2) {
    const unsigned d = unsigned(depth[size_t(y) * w + x]);
    z = d == 0 ? __int_as_float(0x7fc00000) : __fmul_rn(float(d), 0.001f);
  } else {
    z = float(depth[size_t(y) * w + x]);
  }
  float *o = out + (size_t(y) * w + x) * 3;
  if (isnan(z)) {
    o[0] = o[1] = o[2] = __int_as_float(0x7fc00000);
    return;
  }
  // cv::rgbd::depthTo3d caches (u - cx) * (1 / fx) per column, (v - cy) * (1 / fy) per row, then multiplies by z
  o[0] = __fmul_rn(__fmul_rn(__fsub_rn(float(x), cx), __fdiv_rn(1.f, fx)), z);
  o[1] = __fmul_rn(__fmul_rn(__fsub_rn(float(y), cy), __fdiv_rn(1.f, fy)), z);
  o[2] = z;
}

// ---- detection: FAST-9/16 score, 3 x 3 non-maximum suppression, Harris response ----------------------------------------
// cv::FAST(threshold 20, nonmaxSuppression) as cv::ORB runs it on every pyramid level.  score = the largest threshold
// for which the pixel is still a corner = max over the 16 arcs of 9 ring pixels of min(ring - p) (bright) or
// min(p - ring) (dark), minus 1; 0 when that does not exceed the threshold.
__constant__ int c_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

__global__ void __launch_bounds__(256)
fast_score_kernel(const uint8_t *__restrict__ img, int w, int h, int threshold, uint8_t *__restrict__ score) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  int out = 0;
  if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
    const uint8_t *pc = img + size_t(y) * w + x;
    const int p = pc[0];
    int d[16];
    // NOTE: the loops stay rolled on purpose.  Fully unrolled, nvcc 12.9 for sm_100a produced wrong scores for this
    // kernel (tools/fast_unroll_repro.cu reproduces it stand-alone); rolled, it equals cv2 and the host transliteration.
#pragma unroll 1
    for (int k = 0; k < 16; ++k) d[k] = int(pc[c_ring_dy[k] * w + c_ring_dx[k]]) - p;
    int best = -1000;
#pragma unroll 1
    for (int s0 = 0; s0 < 16; ++s0) {
      int mb = 1000, md = 1000;
#pragma unroll 1
      for (int j = 0; j < 9; ++j) {
        const int v = d[(s0 + j) & 15];
        mb = min(mb, v);
        md = min(md, -v);
      }
      best = max(best, max(mb, md));
    }
    if (best > threshold) out = best - 1;
  }
  score[size_t(y) * w + x] = uint8_t(out);
}

// keeps a corner whose score beats its 8 neighbours (strictly) and that lies at least `border` pixels inside the level
// (KeyPointsFilter::runByImageBorder with ORB's edgeThreshold); candidates are appended in arbitrary order.
__global__ void __launch_bounds__(256)
fast_nms_kernel(const uint8_t *__restrict__ score, int w, int h, int border, int level, int capacity,
                int4 *__restrict__ out, int *__restrict__ count, const uint8_t *__restrict__ mask) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < border || y < border || x >= w - border || y >= h - border) return;
  if (mask && mask[size_t(y) * w + x] == 0) return;  // KeyPointsFilter::runByPixelsMask on this level's mask
  const uint8_t *pc = score + size_t(y) * w + x;
  const int s = pc[0];
  if (s == 0) return;
  bool keep = true;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
      if (dx || dy) keep = keep && s > int(pc[dy * w + dx]);
  if (!keep) return;
  const int slot = atomicAdd(count, 1);
  if (slot < capacity) out[slot] = make_int4(x, y, s, level);
}

// cv::threshold(mask, mask, 254, 0, THRESH_TOZERO): the resized mask levels keep only fully covered pixels
__global__ void __launch_bounds__(256) threshold_tozero_kernel(uint8_t *__restrict__ m, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && m[i] <= 254) m[i] = 0;
}

// cv::cvtColor(COLOR_BGR2GRAY), 8-bit: (B 3735 + G 19235 + R 9798 + 2^14) >> 15
__global__ void __launch_bounds__(256)
bgr_to_gray_kernel(const uint8_t *__restrict__ bgr, uint8_t *__restrict__ gray, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
  gray[i] = uint8_t((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
}

// one pass of cv::erode with the default 3 x 3 rectangle; pixels outside the image do not erode
__global__ void __launch_bounds__(256)
erode3x3_kernel(const uint8_t *__restrict__ src, int w, int h, uint8_t *__restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  int m = 255;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx, yy = y + dy;
      if (xx >= 0 && yy >= 0 && xx < w && yy < h) m = min(m, int(src[size_t(yy) * w + xx]));
    }
  dst[size_t(y) * w + x] = uint8_t(m);
}

// rescale_depth (Trainer.cpp:63-81): depth -> float32 metres (uint16 millimetres, 0 -> NaN) on an image-sized canvas;
// when the depth image is smaller, nearest-neighbour resize into the top sub_h rows, NaN below
template <typename T>
__global__ void __launch_bounds__(256)
rescale_depth_kernel(const T *__restrict__ depth, int dw, int dh, int iw, int ih, int sub_h, float *__restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= iw || y >= ih) return;
  float z = __int_as_float(0x7fc00000);
  if (y < sub_h) {
    const int sy = (dh == ih && dw == iw) ? y : min(int(floor(double(y) * (double(dh) / double(sub_h)))), dh - 1);
    const int sx = (dh == ih && dw == iw) ? x : min(int(floor(double(x) * (double(dw) / double(iw)))), dw - 1);
    if (sizeof(T) == 2) {
      const unsigned d = unsigned(depth[size_t(sy) * dw + sx]);
      if (d != 0) z = __fmul_rn(float(d), 0.001f);
    } else {
      z = float(depth[size_t(sy) * dw + sx]);
    }
  }
  out[size_t(y) * iw + x] = z;
}

struct TrainView {
  float fx, fy, cx, cy;
  float R[9], T[3];
};

// validateKeyPoints (training.cpp:57-145) + depthTo3dSparse + cameraToWorld (training.cpp:175-195), one thread per
// keypoint: keep[i] = 1 and world[i] = ((p - T) * R) when the keypoint (or the nearest masked pixel of its 5 x 5
// neighbourhood) lies in the eroded mask and the depth there is valid.
__global__ void __launch_bounds__(128)
train_validate_kernel(const float *__restrict__ kx, const float *__restrict__ ky, int n,
                      const uint8_t *__restrict__ mask, const float *__restrict__ depth, int w, int h, TrainView v,
                      uint8_t *__restrict__ keep, float *__restrict__ world) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float fx = kx[i], fy = ky[i];
  int x = min(max(__float2int_rn(fx), 0), w), y = min(max(__float2int_rn(fy), 0), h);
  auto in_mask = [&](int yy, int xx) { return xx >= 0 && yy >= 0 && xx < w && yy < h && mask[size_t(yy) * w + xx] != 0; };
  bool good = in_mask(y, x);
  if (!good) {
    float best = __int_as_float(0x7f800000);
    const int x0 = x, y0 = y;
    for (int ii = max(x0 - 2, 0); ii <= min(x0 + 2, w); ++ii)
      for (int jj = max(y0 - 2, 0); jj <= min(y0 + 2, h); ++jj)
        if (in_mask(jj, ii)) {
          const float dx = __fsub_rn(float(ii), fx), dy = __fsub_rn(float(jj), fy);
          const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          if (d2 < best) {
            best = d2;
            x = ii;
            y = jj;
            good = true;
          }
        }
  }
  float z = 0.f;
  if (good) {
    z = depth[size_t(y) * w + x];
    good = !isnan(z);
  }
  keep[i] = good ? 1 : 0;
  if (!good) return;
  // depthTo3dSparse: ((u - cx) / fx) z ; cameraToWorld: (p - T) * R in float, products added left to right
  const float px = __fmul_rn(__fdiv_rn(__fsub_rn(float(x), v.cx), v.fx), z);
  const float py = __fmul_rn(__fdiv_rn(__fsub_rn(float(y), v.cy), v.fy), z);
  const float a0 = __fsub_rn(px, v.T[0]), a1 = __fsub_rn(py, v.T[1]), a2 = __fsub_rn(z, v.T[2]);
#pragma unroll
  for (int j = 0; j < 3; ++j)
    world[size_t(i) * 3 + j] =
        __fadd_rn(__fadd_rn(__fmul_rn(a0, v.R[j]), __fmul_rn(a1, v.R[3 + j])), __fmul_rn(a2, v.R[6 + j]));
}

// cv::ORB's HarrisResponses (blockSize 7, k = 0.04) on the unsmoothed level: integer sums of the Sobel products over the
// 7 x 7 block, then ((float) a b - (float) c c - k ((float) a + b)^2) scale^4 in float, operation by operation.
__global__ void __launch_bounds__(128)
harris_kernel(Levels lv, const int4 *__restrict__ cand, int n, float *__restrict__ response) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c4 = cand[i];
  const LevelDesc L = lv.l[c4.w];
  const uint8_t *base = L.img + size_t(c4.y) * L.w + c4.x;
  int a = 0, b = 0, c = 0;
  for (int dy = -3; dy <= 3; ++dy) {
    const uint8_t *r0 = base + (dy - 1) * L.w, *r1 = base + dy * L.w, *r2 = base + (dy + 1) * L.w;
    for (int dx = -3; dx <= 3; ++dx) {
      const int ix = (int(r1[dx + 1]) - int(r1[dx - 1])) * 2 + (int(r0[dx + 1]) - int(r0[dx - 1])) +
                     (int(r2[dx + 1]) - int(r2[dx - 1]));
      const int iy = (int(r2[dx]) - int(r0[dx])) * 2 + (int(r2[dx - 1]) - int(r0[dx - 1])) +
                     (int(r2[dx + 1]) - int(r0[dx + 1]));
      a += ix * ix;
      b += iy * iy;
      c += ix * iy;
    }
  }
  const float scale = __fdiv_rn(1.f, __fmul_rn(28.f, 255.f));
  const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
  const float fa = float(a), fb = float(b), fc = float(c);
  const float sum = __fadd_rn(fa, fb);
  const float t = __fmul_rn(__fmul_rn(0.04f, sum), sum);
  response[i] = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), t), s4);
}

void linear_exact_table(int src, int dst, std::vector<int> &i0, std::vector<int> &i1, std::vector<int> &f) {
  i0.resize(size_t(dst));
  i1.resize(size_t(dst));
  f.resize(size_t(dst));
  const double scale = 1.0 / (double(dst) / double(src));
  for (int i = 0; i < dst; ++i) {
    const double x = (double(i) + 0.5) * scale - 0.5;
    const double fl = std::floor(x);
    const int ix = int(fl);
    f[size_t(i)] = int(std::floor((x - fl) * 256.0 + 0.5));
    i0[size_t(i)] = std::min(std::max(ix, 0), src - 1);
    i1[size_t(i)] = std::min(std::max(ix + 1, 0), src - 1);
  }
}

}  // namespace
}  // namespace tod

using tod::DeviceBuffer;
using tod::fail;

struct tod_orb {
  tod_orb_params p{};
  cudaStream_t stream = nullptr;
  int height = 0, width = 0;  // geometry the pyramid buffers are built for
  int lw[tod::kMaxLevels] = {0}, lh[tod::kMaxLevels] = {0};
  float scale[tod::kMaxLevels] = {0};
  DeviceBuffer d_img[tod::kMaxLevels], d_smooth[tod::kMaxLevels], d_rowf, d_tables[tod::kMaxLevels];
  DeviceBuffer d_kx, d_ky, d_oct, d_angle, d_desc;
  DeviceBuffer d_score, d_cand, d_count, d_resp;  // detection: FAST score image, candidate list, counter, Harris
  DeviceBuffer d_mask[tod::kMaxLevels], d_bgr;    // optional detection mask pyramid; staging of a BGR frame
  bool pattern_uploaded = false;
};

extern "C" {

void tod_orb_default_params(tod_orb_params *p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->n_levels = 3;          // conf/detection.ork:27
  p->scale_factor = 1.2f;   // conf/detection.ork:28
  p->device = 0;
}

int tod_orb_create(const tod_orb_params *p, tod_orb **out) {
  TOD_REQUIRE(p && out, "null argument");
  TOD_REQUIRE(p->n_levels >= 1 && p->n_levels <= tod::kMaxLevels, "n_levels must be in 1..%d", tod::kMaxLevels);
  TOD_REQUIRE(p->scale_factor > 1.0f, "scale_factor must exceed 1");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return fail(TOD_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  TOD_REQUIRE(p->device >= 0 && p->device < n_dev, "device %d out of range (%d devices)", p->device, n_dev);
  TOD_CUDA(cudaSetDevice(p->device));
  tod_orb *o = new tod_orb();
  o->p = *p;
  if (cudaStreamCreate(&o->stream) != cudaSuccess) {
    delete o;
    return fail(TOD_ERR_CUDA, "creating the ORB stage's stream failed");
  }
  *out = o;
  return TOD_OK;
}

void tod_orb_destroy(tod_orb *o) {
  if (!o) return;
  cudaSetDevice(o->p.device);
  for (int l = 0; l < tod::kMaxLevels; ++l) {
    o->d_img[l].release();
    o->d_smooth[l].release();
    o->d_tables[l].release();
    o->d_mask[l].release();
  }
  o->d_bgr.release();
  for (DeviceBuffer *b : {&o->d_rowf, &o->d_kx, &o->d_ky, &o->d_oct, &o->d_angle, &o->d_desc, &o->d_score, &o->d_cand,
                          &o->d_count, &o->d_resp})
    b->release();
  if (o->stream) cudaStreamDestroy(o->stream);
  delete o;
}

}  // extern "C"

namespace {

// Level geometry and resize tables for a new image size.
int orb_prepare(tod_orb *o, int height, int width) {
  TOD_REQUIRE(height > 2 * tod::kDescBorder && width > 2 * tod::kDescBorder, "image too small");
  if (height == o->height && width == o->width) return TOD_OK;
  cudaStream_t st = o->stream;
  const int L = o->p.n_levels;
  // scale_l = (float) pow((double) scale_factor, l); size = cvRound(size0 / scale_l)
  for (int l = 0; l < L; ++l) {
    o->scale[l] = float(std::pow(double(o->p.scale_factor), double(l)));
    o->lw[l] = int(std::lrintf(float(width) / o->scale[l]));
    o->lh[l] = int(std::lrintf(float(height) / o->scale[l]));
    TOD_REQUIRE(o->lw[l] > 2 * tod::kDescBorder && o->lh[l] > 2 * tod::kDescBorder, "image too small for level %d", l);
    const size_t px = size_t(o->lw[l]) * size_t(o->lh[l]);
    TOD_CUDA(o->d_img[l].reserve(px));
    TOD_CUDA(o->d_smooth[l].reserve(px));
    if (l) {  // INTER_LINEAR_EXACT tables from level l-1: x0 x1 fx (lw) y0 y1 fy (lh), all int32
      std::vector<int> a0, a1, af, b0, b1, bf, all;
      tod::linear_exact_table(o->lw[l - 1], o->lw[l], a0, a1, af);
      tod::linear_exact_table(o->lh[l - 1], o->lh[l], b0, b1, bf);
      for (auto *v : {&a0, &a1, &af, &b0, &b1, &bf}) all.insert(all.end(), v->begin(), v->end());
      TOD_CUDA(o->d_tables[l].reserve(all.size() * sizeof(int)));
      TOD_CUDA(cudaMemcpyAsync(o->d_tables[l].ptr, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      TOD_CUDA(cudaStreamSynchronize(st));  // `all` dies at the end of this block
    }
  }
  TOD_CUDA(o->d_rowf.reserve(size_t(width) * size_t(height) * sizeof(float)));
  TOD_CUDA(o->d_score.reserve(size_t(width) * size_t(height)));
  o->height = height;
  o->width = width;
  return TOD_OK;
}

// Upload the frame, build the pyramid (unsmoothed + smoothed levels) on the handle's stream.
int orb_build_pyramid(tod_orb *o, const uint8_t *image, tod::Levels *lv, int channels = 1) {
  cudaStream_t st = o->stream;
  if (!o->pattern_uploaded) {
    TOD_CUDA(cudaMemcpyToSymbol(tod::c_pattern, tod::kOrbPattern, sizeof(tod::kOrbPattern)));
    o->pattern_uploaded = true;
  }
  const size_t px0 = size_t(o->width) * size_t(o->height);
  if (channels == 3) {  // cv::ORB converts a colour frame itself: cvtColor(BGR2GRAY)
    TOD_CUDA(o->d_bgr.reserve(px0 * 3));
    TOD_CUDA(cudaMemcpyAsync(o->d_bgr.ptr, image, px0 * 3, cudaMemcpyHostToDevice, st));
    tod::bgr_to_gray_kernel<<<unsigned((px0 + 255) / 256), 256, 0, st>>>(o->d_bgr.as<uint8_t>(),
                                                                          o->d_img[0].as<uint8_t>(), px0);
    tod::count_launch();
  } else {
    TOD_CUDA(cudaMemcpyAsync(o->d_img[0].ptr, image, px0, cudaMemcpyHostToDevice, st));
  }
  for (int l = 0; l < o->p.n_levels; ++l) {
    const int w = o->lw[l], h = o->lh[l];
    dim3 grid((w + 255) / 256, h);
    if (l) {
      const int *t = o->d_tables[l].as<int>();
      tod::resize_exact_kernel<<<grid, 256, 0, st>>>(o->d_img[l - 1].as<uint8_t>(), o->lw[l - 1],
                                                     o->d_img[l].as<uint8_t>(), w, h, t, t + w, t + 2 * w, t + 3 * w,
                                                     t + 3 * w + h, t + 3 * w + 2 * h);
      tod::count_launch();
    }
    tod::smooth_rows_kernel<<<grid, 256, 0, st>>>(o->d_img[l].as<uint8_t>(), w, h, o->d_rowf.as<float>());
    tod::smooth_cols_kernel<<<grid, 256, 0, st>>>(o->d_rowf.as<float>(), w, h, o->d_smooth[l].as<uint8_t>());
    tod::count_launch(2);
    lv->l[l].img = o->d_img[l].as<uint8_t>();
    lv->l[l].smooth = o->d_smooth[l].as<uint8_t>();
    lv->l[l].w = w;
    lv->l[l].h = h;
    lv->l[l].inv_scale = 1.f / o->scale[l];
  }
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

// Orientation (optional) + descriptors of n keypoints given as host arrays; angles come back in ha.
int orb_describe_points(tod_orb *o, const tod::Levels &lv, const std::vector<float> &hx, const std::vector<float> &hy,
                        const std::vector<int> &ho, std::vector<float> &ha, int n, bool compute_angles,
                        uint8_t *descriptors, const void **d_descriptors) {
  cudaStream_t st = o->stream;
  if (n == 0) {
    TOD_CUDA(cudaStreamSynchronize(st));
    if (d_descriptors) *d_descriptors = nullptr;
    return TOD_OK;
  }
  const size_t nn = size_t(n);
  TOD_CUDA(o->d_kx.reserve(nn * 4));
  TOD_CUDA(o->d_ky.reserve(nn * 4));
  TOD_CUDA(o->d_oct.reserve(nn * 4));
  TOD_CUDA(o->d_angle.reserve(nn * 4));
  TOD_CUDA(o->d_desc.reserve(nn * 32));
  TOD_CUDA(cudaMemcpyAsync(o->d_kx.ptr, hx.data(), nn * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(o->d_ky.ptr, hy.data(), nn * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(o->d_oct.ptr, ho.data(), nn * 4, cudaMemcpyHostToDevice, st));
  const int blocks = (n + 7) / 8;
  if (compute_angles) {
    tod::orb_angle_kernel<<<blocks, 256, 0, st>>>(lv, o->d_kx.as<float>(), o->d_ky.as<float>(), o->d_oct.as<int>(), n,
                                                  o->d_angle.as<float>());
    tod::count_launch();
  } else {
    TOD_CUDA(cudaMemcpyAsync(o->d_angle.ptr, ha.data(), nn * 4, cudaMemcpyHostToDevice, st));
  }
  tod::orb_describe_kernel<<<blocks, 256, 0, st>>>(lv, o->d_kx.as<float>(), o->d_ky.as<float>(), o->d_oct.as<int>(),
                                                   o->d_angle.as<float>(), n, o->d_desc.as<uint8_t>());
  tod::count_launch();
  TOD_CUDA(cudaGetLastError());
  if (compute_angles) TOD_CUDA(cudaMemcpyAsync(ha.data(), o->d_angle.ptr, nn * 4, cudaMemcpyDeviceToHost, st));
  if (descriptors) TOD_CUDA(cudaMemcpyAsync(descriptors, o->d_desc.ptr, nn * 32, cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  if (d_descriptors) *d_descriptors = o->d_desc.ptr;
  return TOD_OK;
}

// KeyPointsFilter::retainBest: keep everything whose response reaches the n-th best one (ties are all kept).
void retain_best(std::vector<int> &idx, const std::vector<float> &resp, size_t n) {
  if (idx.size() <= n) return;
  if (n == 0) {
    idx.clear();
    return;
  }
  std::vector<float> r;
  r.reserve(idx.size());
  for (int i : idx) r.push_back(resp[size_t(i)]);
  std::nth_element(r.begin(), r.begin() + (n - 1), r.end(), std::greater<float>());
  const float thr = r[n - 1];
  std::vector<int> keep;
  for (int i : idx)
    if (resp[size_t(i)] >= thr) keep.push_back(i);
  idx.swap(keep);
}

}  // namespace

extern "C" {

int tod_orb_describe(tod_orb *o, const uint8_t *image, int32_t height, int32_t width, tod_keypoint *keypoints,
                     int32_t n, int32_t compute_angles, uint8_t *descriptors, const void **d_descriptors) {
  TOD_REQUIRE(o && image && (keypoints || n == 0) && n >= 0, "bad argument");
  TOD_CUDA(cudaSetDevice(o->p.device));
  if (int rc = orb_prepare(o, height, width)) return rc;
  const int L = o->p.n_levels;
  // keypoints: level-0 coordinates and octave; every pattern / disc access must stay inside its level
  std::vector<float> hx(static_cast<size_t>(n)), hy(static_cast<size_t>(n)), ha(static_cast<size_t>(n));
  std::vector<int> ho(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    const tod_keypoint &k = keypoints[i];
    TOD_REQUIRE(k.octave >= 0 && k.octave < L, "keypoint %d: octave %d outside [0, %d)", i, k.octave, L);
    const float inv = 1.f / o->scale[k.octave];
    const long cx = std::lrintf(k.x * inv), cy = std::lrintf(k.y * inv);
    TOD_REQUIRE(cx >= tod::kDescBorder && cy >= tod::kDescBorder && cx < o->lw[k.octave] - tod::kDescBorder &&
                    cy < o->lh[k.octave] - tod::kDescBorder,
                "keypoint %d at (%g, %g) octave %d is closer than %d pixels to the border of its level (cv::ORB drops "
                "such keypoints: edgeThreshold)", i, k.x, k.y, k.octave, tod::kDescBorder);
    hx[size_t(i)] = k.x;
    hy[size_t(i)] = k.y;
    ho[size_t(i)] = k.octave;
    ha[size_t(i)] = k.angle;
  }
  tod::Levels lv{};
  if (int rc = orb_build_pyramid(o, image, &lv)) return rc;
  if (int rc = orb_describe_points(o, lv, hx, hy, ho, ha, n, compute_angles != 0, descriptors, d_descriptors)) return rc;
  if (compute_angles)
    for (int i = 0; i < n; ++i) keypoints[i].angle = ha[size_t(i)];
  return TOD_OK;
}

int tod_orb_detect_and_compute(tod_orb *o, const uint8_t *image, int32_t height, int32_t width, int32_t n_features,
                               tod_keypoint *keypoints, int32_t max_keypoints, int32_t *n_keypoints,
                               uint8_t *descriptors, const void **d_descriptors) {
  return tod_orb_detect_and_compute_masked(o, image, 1, height, width, nullptr, n_features, keypoints, max_keypoints,
                                           n_keypoints, descriptors, d_descriptors);
}

int tod_orb_detect_and_compute_masked(tod_orb *o, const uint8_t *image, int32_t channels, int32_t height,
                                      int32_t width, const uint8_t *mask, int32_t n_features, tod_keypoint *keypoints,
                                      int32_t max_keypoints, int32_t *n_keypoints, uint8_t *descriptors,
                                      const void **d_descriptors) {
  TOD_REQUIRE(o && image && keypoints && n_keypoints && n_features >= 0 && max_keypoints >= 0, "bad argument");
  TOD_REQUIRE(channels == 1 || channels == 3, "image must have 1 (grey) or 3 (BGR) channels");
  *n_keypoints = 0;
  TOD_CUDA(cudaSetDevice(o->p.device));
  if (int rc = orb_prepare(o, height, width)) return rc;
  cudaStream_t st = o->stream;
  const int L = o->p.n_levels;
  const int kEdge = 31, kFastThreshold = 20;  // cv::ORB defaults: edgeThreshold, fastThreshold
  tod::Levels lv{};
  if (int rc = orb_build_pyramid(o, image, &lv, channels)) return rc;
  if (mask) {
    // ORB's mask pyramid: level l = resize(level l-1, INTER_LINEAR_EXACT), then everything below 255 -> 0
    for (int l = 0; l < L; ++l) {
      const size_t px = size_t(o->lw[l]) * size_t(o->lh[l]);
      TOD_CUDA(o->d_mask[l].reserve(px));
      if (l == 0) {
        TOD_CUDA(cudaMemcpyAsync(o->d_mask[0].ptr, mask, px, cudaMemcpyHostToDevice, st));
      } else {
        const int w = o->lw[l], h = o->lh[l];
        const int *t = o->d_tables[l].as<int>();
        dim3 grid((w + 255) / 256, h);
        tod::resize_exact_kernel<<<grid, 256, 0, st>>>(o->d_mask[l - 1].as<uint8_t>(), o->lw[l - 1],
                                                       o->d_mask[l].as<uint8_t>(), w, h, t, t + w, t + 2 * w,
                                                       t + 3 * w, t + 3 * w + h, t + 3 * w + 2 * h);
        tod::threshold_tozero_kernel<<<unsigned((px + 255) / 256), 256, 0, st>>>(o->d_mask[l].as<uint8_t>(), px);
        tod::count_launch(2);
      }
    }
  }
  // features per level (ORB_Impl::computeKeyPoints): a geometric series over the levels, the last level takes the rest
  std::vector<int> per_level(static_cast<size_t>(L), 0);
  {
    const float factor = float(1.0 / double(o->p.scale_factor));
    float desired = float(n_features) * (1.f - factor) / (1.f - float(std::pow(double(factor), double(L))));  // float ops
    int sum = 0;
    for (int l = 0; l + 1 < L; ++l) {
      per_level[size_t(l)] = int(std::lrintf(desired));
      sum += per_level[size_t(l)];
      desired *= factor;
    }
    per_level[size_t(L - 1)] = std::max(n_features - sum, 0);
  }
  // FAST + non-maximum suppression + border filter on every level -> one candidate list
  const int capacity = std::max(1 << 16, (height * width) / 8);
  TOD_CUDA(o->d_cand.reserve(size_t(capacity) * sizeof(int4)));
  TOD_CUDA(o->d_count.reserve(sizeof(int)));
  TOD_CUDA(cudaMemsetAsync(o->d_count.ptr, 0, sizeof(int), st));
  for (int l = 0; l < L; ++l) {
    const int w = o->lw[l], h = o->lh[l];
    dim3 grid((w + 255) / 256, h);
    tod::fast_score_kernel<<<grid, 256, 0, st>>>(o->d_img[l].as<uint8_t>(), w, h, kFastThreshold,
                                                 o->d_score.as<uint8_t>());
    tod::fast_nms_kernel<<<grid, 256, 0, st>>>(o->d_score.as<uint8_t>(), w, h, kEdge, l, capacity,
                                               o->d_cand.as<int4>(), o->d_count.as<int>(),
                                               mask ? o->d_mask[l].as<uint8_t>() : nullptr);
    tod::count_launch(2);
  }
  TOD_CUDA(cudaGetLastError());
  int n_cand = 0;
  TOD_CUDA(cudaMemcpyAsync(&n_cand, o->d_count.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  if (n_cand > capacity) return fail(TOD_ERR_LIMIT, "%d FAST corners exceed the candidate capacity %d", n_cand, capacity);
  std::vector<int4> cand(static_cast<size_t>(n_cand));
  if (n_cand) TOD_CUDA(cudaMemcpy(cand.data(), o->d_cand.ptr, size_t(n_cand) * sizeof(int4), cudaMemcpyDeviceToHost));
  // deterministic order (the append order on the device is not): by level, row, column
  std::sort(cand.begin(), cand.end(), [](const int4 &a, const int4 &b) {
    if (a.w != b.w) return a.w < b.w;
    if (a.y != b.y) return a.y < b.y;
    return a.x < b.x;
  });
  // per level: keep the 2 N best FAST scores (ties kept), Harris response, keep the N best (ties kept)
  std::vector<int> first_cut;
  {
    std::vector<float> score(cand.size());
    for (size_t i = 0; i < cand.size(); ++i) score[i] = float(cand[i].z);
    size_t at = 0;
    for (int l = 0; l < L; ++l) {
      std::vector<int> idx;
      while (at < cand.size() && cand[at].w == l) idx.push_back(int(at++));
      retain_best(idx, score, size_t(2 * per_level[size_t(l)]));
      first_cut.insert(first_cut.end(), idx.begin(), idx.end());
    }
  }
  std::vector<int4> sel(first_cut.size());
  for (size_t i = 0; i < first_cut.size(); ++i) sel[i] = cand[size_t(first_cut[i])];
  std::vector<float> resp(sel.size());
  if (!sel.empty()) {
    TOD_CUDA(o->d_resp.reserve(sel.size() * sizeof(float)));
    TOD_CUDA(cudaMemcpyAsync(o->d_cand.ptr, sel.data(), sel.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    tod::harris_kernel<<<unsigned((sel.size() + 127) / 128), 128, 0, st>>>(lv, o->d_cand.as<int4>(), int(sel.size()),
                                                                             o->d_resp.as<float>());
    tod::count_launch();
    TOD_CUDA(cudaGetLastError());
    TOD_CUDA(cudaMemcpyAsync(resp.data(), o->d_resp.ptr, sel.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    TOD_CUDA(cudaStreamSynchronize(st));
  }
  std::vector<int> final_idx;
  {
    size_t at = 0;
    for (int l = 0; l < L; ++l) {
      std::vector<int> idx;
      while (at < sel.size() && sel[at].w == l) idx.push_back(int(at++));
      retain_best(idx, resp, size_t(per_level[size_t(l)]));
      final_idx.insert(final_idx.end(), idx.begin(), idx.end());
    }
  }
  const int n = int(final_idx.size());
  if (n > max_keypoints) return fail(TOD_ERR_LIMIT, "%d keypoints found but max_keypoints = %d", n, max_keypoints);
  // level coordinates -> level-0 coordinates (pt *= scale), size = patchSize * scale
  std::vector<float> hx(static_cast<size_t>(n)), hy(static_cast<size_t>(n)), ha(static_cast<size_t>(n), 0.f);
  std::vector<int> ho(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    const int4 c = sel[size_t(final_idx[size_t(i)])];
    const float sc = o->scale[c.w];
    hx[size_t(i)] = float(c.x) * sc;
    hy[size_t(i)] = float(c.y) * sc;
    ho[size_t(i)] = c.w;
  }
  if (int rc = orb_describe_points(o, lv, hx, hy, ho, ha, n, true, descriptors, d_descriptors)) return rc;
  for (int i = 0; i < n; ++i) {
    tod_keypoint &k = keypoints[i];
    k.x = hx[size_t(i)];
    k.y = hy[size_t(i)];
    k.size = 31.f * o->scale[ho[size_t(i)]];
    k.angle = ha[size_t(i)];
    k.response = resp[size_t(final_idx[size_t(i)])];
    k.octave = ho[size_t(i)];
    k.class_id = -1;
  }
  *n_keypoints = n;
  return TOD_OK;
}

}  // extern "C"

// ---- the offline training path (Trainer.cpp:121-187, training.cpp) ---------------------------------------------------
struct tod_trainer {
  tod_trainer_params p{};
  tod_orb *orb = nullptr;
  DeviceBuffer d_mask_a, d_mask_b, d_depth_in, d_depth, d_kx, d_ky, d_keep, d_world;
  std::vector<uint8_t> desc;   // merged model: n x 32
  std::vector<float> points;   // n x 3, object frame
};

extern "C" {

void tod_trainer_default_params(tod_trainer_params *p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->n_features = 500;     // cv::ORB's own defaults: Trainer.cpp:142-150 builds the extractor without parameters
  p->n_levels = 8;
  p->scale_factor = 1.2f;
  p->device = 0;
}

int tod_trainer_create(const tod_trainer_params *p, tod_trainer **out) {
  TOD_REQUIRE(p && out && p->n_features >= 0, "bad argument");
  tod_orb_params op;
  tod_orb_default_params(&op);
  op.n_levels = p->n_levels;
  op.scale_factor = p->scale_factor;
  op.device = p->device;
  tod_orb *orb = nullptr;
  if (int rc = tod_orb_create(&op, &orb)) return rc;
  tod_trainer *t = new tod_trainer();
  t->p = *p;
  t->orb = orb;
  *out = t;
  return TOD_OK;
}

void tod_trainer_destroy(tod_trainer *t) {
  if (!t) return;
  cudaSetDevice(t->p.device);
  for (DeviceBuffer *b : {&t->d_mask_a, &t->d_mask_b, &t->d_depth_in, &t->d_depth, &t->d_kx, &t->d_ky, &t->d_keep,
                          &t->d_world})
    b->release();
  tod_orb_destroy(t->orb);
  delete t;
}

int tod_trainer_clear(tod_trainer *t) {
  TOD_REQUIRE(t, "null argument");
  t->desc.clear();
  t->points.clear();
  return TOD_OK;
}

int64_t tod_trainer_num_points(const tod_trainer *t) { return t ? int64_t(t->desc.size() / 32) : 0; }

int tod_trainer_model(const tod_trainer *t, const uint8_t **descriptors, const float **points, int64_t *n) {
  TOD_REQUIRE(t, "null argument");
  if (descriptors) *descriptors = t->desc.data();
  if (points) *points = t->points.data();
  if (n) *n = int64_t(t->desc.size() / 32);
  return TOD_OK;
}

int tod_trainer_add_observation(tod_trainer *t, const uint8_t *image, int32_t channels, int32_t height, int32_t width,
                                const uint8_t *mask, const void *depth, int32_t depth_is_u16, int32_t depth_height,
                                int32_t depth_width, const float *K, const float *R, const float *T,
                                int32_t *n_added) {
  TOD_REQUIRE(t && image && mask && depth && K && R && T, "null argument");
  TOD_REQUIRE(depth_height > 0 && depth_width > 0, "bad depth size");
  if (n_added) *n_added = 0;
  TOD_CUDA(cudaSetDevice(t->p.device));
  // features on the masked image (Trainer.cpp:142-150)
  const int cap = 2 * t->p.n_features + 1024;
  std::vector<tod_keypoint> kp(static_cast<size_t>(cap));
  std::vector<uint8_t> desc(size_t(cap) * 32);
  int32_t n = 0;
  if (int rc = tod_orb_detect_and_compute_masked(t->orb, image, channels, height, width, mask, t->p.n_features,
                                                 kp.data(), cap, &n, desc.data(), nullptr))
    return rc;
  if (n == 0) return TOD_OK;
  cudaStream_t st = t->orb->stream;
  const size_t px = size_t(height) * size_t(width);
  // mask eroded 4 times (training.cpp:66-70)
  TOD_CUDA(t->d_mask_a.reserve(px));
  TOD_CUDA(t->d_mask_b.reserve(px));
  TOD_CUDA(cudaMemcpyAsync(t->d_mask_a.ptr, mask, px, cudaMemcpyHostToDevice, st));
  dim3 grid((width + 255) / 256, height);
  uint8_t *ma = t->d_mask_a.as<uint8_t>(), *mb = t->d_mask_b.as<uint8_t>();
  for (int it = 0; it < 4; ++it) {
    tod::erode3x3_kernel<<<grid, 256, 0, st>>>(ma, width, height, mb);
    std::swap(ma, mb);
  }
  tod::count_launch(4);
  // depth rescaled to the image (Trainer.cpp:63-81, :154)
  const size_t dpx = size_t(depth_height) * size_t(depth_width), dbytes = dpx * (depth_is_u16 ? 2 : 4);
  TOD_CUDA(t->d_depth_in.reserve(dbytes));
  TOD_CUDA(t->d_depth.reserve(px * 4));
  TOD_CUDA(cudaMemcpyAsync(t->d_depth_in.ptr, depth, dbytes, cudaMemcpyHostToDevice, st));
  const bool same = depth_height == height && depth_width == width;
  const int sub_h = same ? height : std::min(height, int(float(depth_height) * (float(width) / float(depth_width))));
  if (depth_is_u16)
    tod::rescale_depth_kernel<uint16_t><<<grid, 256, 0, st>>>(t->d_depth_in.as<uint16_t>(), depth_width, depth_height,
                                                              width, height, sub_h, t->d_depth.as<float>());
  else
    tod::rescale_depth_kernel<float><<<grid, 256, 0, st>>>(t->d_depth_in.as<float>(), depth_width, depth_height, width,
                                                           height, sub_h, t->d_depth.as<float>());
  tod::count_launch();
  // validateKeyPoints + depthTo3dSparse + cameraToWorld (Trainer.cpp:157-171)
  std::vector<float> hx(static_cast<size_t>(n)), hy(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    hx[size_t(i)] = kp[size_t(i)].x;
    hy[size_t(i)] = kp[size_t(i)].y;
  }
  TOD_CUDA(t->d_kx.reserve(size_t(n) * 4));
  TOD_CUDA(t->d_ky.reserve(size_t(n) * 4));
  TOD_CUDA(t->d_keep.reserve(size_t(n)));
  TOD_CUDA(t->d_world.reserve(size_t(n) * 12));
  TOD_CUDA(cudaMemcpyAsync(t->d_kx.ptr, hx.data(), size_t(n) * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(t->d_ky.ptr, hy.data(), size_t(n) * 4, cudaMemcpyHostToDevice, st));
  tod::TrainView v{};
  v.fx = K[0];
  v.fy = K[4];
  v.cx = K[2];
  v.cy = K[5];
  for (int i = 0; i < 9; ++i) v.R[i] = R[i];
  for (int i = 0; i < 3; ++i) v.T[i] = T[i];
  tod::train_validate_kernel<<<(n + 127) / 128, 128, 0, st>>>(t->d_kx.as<float>(), t->d_ky.as<float>(), n, ma,
                                                              t->d_depth.as<float>(), width, height, v,
                                                              t->d_keep.as<uint8_t>(), t->d_world.as<float>());
  tod::count_launch();
  TOD_CUDA(cudaGetLastError());
  std::vector<uint8_t> keep(static_cast<size_t>(n));
  std::vector<float> world(size_t(n) * 3);
  TOD_CUDA(cudaMemcpyAsync(keep.data(), t->d_keep.ptr, size_t(n), cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaMemcpyAsync(world.data(), t->d_world.ptr, size_t(n) * 12, cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  // mergePoints (training.cpp:147-173): kept keypoints appended in order
  int added = 0;
  for (int i = 0; i < n; ++i) {
    if (!keep[size_t(i)]) continue;
    t->desc.insert(t->desc.end(), desc.begin() + size_t(i) * 32, desc.begin() + size_t(i + 1) * 32);
    t->points.insert(t->points.end(), world.begin() + size_t(i) * 3, world.begin() + size_t(i + 1) * 3);
    ++added;
  }
  if (n_added) *n_added = added;
  return TOD_OK;
}

int tod_orb_read_level(tod_orb *o, int32_t level, int32_t kind, uint8_t *out, int32_t *height, int32_t *width) {
  TOD_REQUIRE(o && level >= 0 && level < o->p.n_levels && kind >= 0 && kind <= 2, "bad argument");
  if (o->height == 0) return fail(TOD_ERR_STATE, "no frame has been processed yet");
  TOD_CUDA(cudaSetDevice(o->p.device));
  const int w = o->lw[level], h = o->lh[level];
  if (height) *height = h;
  if (width) *width = w;
  if (!out) return TOD_OK;
  const uint8_t *src = kind == 0 ? o->d_img[level].as<uint8_t>() : o->d_smooth[level].as<uint8_t>();
  if (kind == 2) {  // FAST corner scores of the level, recomputed
    dim3 grid((w + 255) / 256, h);
    tod::fast_score_kernel<<<grid, 256, 0, o->stream>>>(o->d_img[level].as<uint8_t>(), w, h, 20,
                                                        o->d_score.as<uint8_t>());
    tod::count_launch();
    TOD_CUDA(cudaGetLastError());
    src = o->d_score.as<uint8_t>();
  }
  TOD_CUDA(cudaMemcpyAsync(out, src, size_t(w) * size_t(h), cudaMemcpyDeviceToHost, o->stream));
  TOD_CUDA(cudaStreamSynchronize(o->stream));
  return TOD_OK;
}

int tod_depth_to_3d(int32_t device, const void *depth, int32_t depth_is_u16, int32_t height, int32_t width,
                    const float *K, float *points3d) {
  TOD_REQUIRE(depth && K && points3d && height > 0 && width > 0, "bad argument");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return fail(TOD_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  TOD_REQUIRE(device >= 0 && device < n_dev, "device %d out of range", device);
  TOD_CUDA(cudaSetDevice(device));
  const size_t px = size_t(height) * size_t(width), in_bytes = px * (depth_is_u16 ? 2 : 4);
  // the device buffers of a calling thread are kept between calls (a frame loop calls this once per frame: no
  // cudaMalloc / cudaFree per frame); a pinned `points3d` makes the 12-byte-per-pixel read-back run at PCIe speed
  struct Cache {
    int device = -1;
    DeviceBuffer d_in, d_out;
    ~Cache() {
      if (device >= 0 && cudaSetDevice(device) == cudaSuccess) {
        d_in.release();
        d_out.release();
      }
    }
  };
  static thread_local Cache cache;
  if (cache.device != device) {
    if (cache.device >= 0 && cudaSetDevice(cache.device) == cudaSuccess) {
      cache.d_in.release();
      cache.d_out.release();
    }
    TOD_CUDA(cudaSetDevice(device));
    cache.device = device;
  }
  DeviceBuffer &d_in = cache.d_in, &d_out = cache.d_out;
  TOD_CUDA(d_in.reserve(in_bytes));
  TOD_CUDA(d_out.reserve(px * 12));
  TOD_CUDA(cudaMemcpy(d_in.ptr, depth, in_bytes, cudaMemcpyHostToDevice));
  dim3 grid((width + 255) / 256, height);
  if (depth_is_u16)
    tod::depth_to_3d_kernel<uint16_t><<<grid, 256>>>(d_in.as<uint16_t>(), width, height, K[0], K[4], K[2], K[5],
                                                     d_out.as<float>());
  else
    tod::depth_to_3d_kernel<float><<<grid, 256>>>(d_in.as<float>(), width, height, K[0], K[4], K[2], K[5],
                                                  d_out.as<float>());
  tod::count_launch();
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpy(points3d, d_out.ptr, px * 12, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return fail(TOD_ERR_CUDA, "depth -> 3-D failed: %s", cudaGetErrorString(ce));
  return TOD_OK;
}

}  // extern "C"
