// Stand-alone reproducer of a code-generation problem met while writing fast_score_kernel (tod_b200/csrc/orb.cu):
// with nvcc 12.9.86 for sm_100a, the FAST-9/16 score kernel gives WRONG scores when the 16-pixel ring gather (offsets
// from two __constant__ int arrays) is fully unrolled (variants 0 and 2: 80 474 "corners" on the 640 x 480 test frame),
// and the right ones (variant 1: 9 719, equal to cv2.FastFeatureDetector and to a host transliteration) with
// `#pragma unroll 1`.  The product kernel therefore keeps its loops rolled.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fast_unroll_repro tools/fast_unroll_repro.cu
// run:   ./fast_unroll_repro frame.bin      (frame.bin = 640 x 480 u8, e.g. tod_b200.synth.make_textured_image(480, 640, seed=11))
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
__constant__ int c_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
template <int VARIANT>
__global__ void k(const uint8_t *__restrict__ img, int w, int h, int threshold, uint8_t *__restrict__ score) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  int out = 0;
  if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
    const uint8_t *pc = img + size_t(y) * w + x;
    const int p = pc[0];
    int d[16];
    if (VARIANT == 0 || VARIANT == 2) {
#pragma unroll
      for (int k = 0; k < 16; ++k) d[k] = int(pc[c_ring_dy[k] * w + c_ring_dx[k]]) - p;
    } else {
#pragma unroll 1
      for (int k = 0; k < 16; ++k) d[k] = int(pc[c_ring_dy[k] * w + c_ring_dx[k]]) - p;
    }
    int best = -1000;
    if (VARIANT == 0) {
#pragma unroll
      for (int s0 = 0; s0 < 16; ++s0) {
        int mb = 1000, md = 1000;
#pragma unroll
        for (int j = 0; j < 9; ++j) { const int v = d[(s0 + j) & 15]; mb = min(mb, v); md = min(md, -v); }
        best = max(best, max(mb, md));
      }
    } else if (VARIANT == 2) {
#pragma unroll
      for (int s0 = 0; s0 < 16; ++s0) {
        int mn = d[s0], mx = d[s0];
#pragma unroll
        for (int j = 1; j < 9; ++j) { const int v = d[(s0 + j) & 15]; mn = min(mn, v); mx = max(mx, v); }
        best = max(best, max(mn, -mx));
      }
    } else {
#pragma unroll 1
      for (int s0 = 0; s0 < 16; ++s0) {
        int mb = 1000, md = 1000;
#pragma unroll 1
        for (int j = 0; j < 9; ++j) { const int v = d[(s0 + j) & 15]; mb = min(mb, v); md = min(md, -v); }
        best = max(best, max(mb, md));
      }
    }
    if (best > threshold) out = best - 1;
  }
  score[size_t(y) * w + x] = uint8_t(out);
}
int main(int argc, char **argv) {
  int w = 640, h = 480;
  std::vector<uint8_t> img(w * h), s0(w * h), s1(w * h);
  FILE *f = fopen(argv[1], "rb"); if (!f) return 1; if (fread(img.data(), 1, w * h, f) != size_t(w*h)) return 2; fclose(f);
  uint8_t *d_img, *d_s; cudaMalloc(&d_img, w * h); cudaMalloc(&d_s, w * h);
  cudaMemcpy(d_img, img.data(), w * h, cudaMemcpyHostToDevice);
  dim3 grid((w + 255) / 256, h);
  k<0><<<grid, 256>>>(d_img, w, h, 20, d_s); cudaMemcpy(s0.data(), d_s, w * h, cudaMemcpyDeviceToHost);
  k<1><<<grid, 256>>>(d_img, w, h, 20, d_s); cudaMemcpy(s1.data(), d_s, w * h, cudaMemcpyDeviceToHost);
  std::vector<uint8_t> s2(w*h); k<2><<<grid, 256>>>(d_img, w, h, 20, d_s); cudaMemcpy(s2.data(), d_s, w * h, cudaMemcpyDeviceToHost);
  long n2 = 0, diff = 0; for (int i = 0; i < w * h; ++i) { n2 += s2[i] > 0; diff += s2[i] != s1[i]; }
  printf("variant2 nonzero %ld, differs from variant1 in %ld pixels\n", n2, diff);
  long n0 = 0, n1 = 0; for (int i = 0; i < w * h; ++i) { n0 += s0[i] > 0; n1 += s1[i] > 0; }
  printf("variant0 nonzero %ld  variant1 nonzero %ld  err %s  (expected 9719)\n", n0, n1, cudaGetErrorString(cudaGetLastError()));
  printf("v0 (97,3)=%d (99,3)=%d (100,3)=%d (102,3)=%d | v1 %d %d %d %d\n", s0[3*w+97], s0[3*w+99], s0[3*w+100], s0[3*w+102], s1[3*w+97], s1[3*w+99], s1[3*w+100], s1[3*w+102]);
  return 0;
}
