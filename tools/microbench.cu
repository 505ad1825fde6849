// Micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json does not hold:
// INT POPC / LOP3 / IADD3 issue rates per SM, and the bare XOR+POPC "compare" rate of K1's inner loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints one JSON object.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITER = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(uint32_t *out, uint32_t seed) {
  uint32_t a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x * 31u + i * 977u;
  uint32_t b = seed * 3u + 1u, c = seed ^ 0x5bd1e995u;
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) a[i] = __popc(a[i]) + 0u;                                  // POPC chain (dependent per accumulator)
      if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));  // LOP3
      if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); // IADD
      if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));      // IMAD
      if (OP == 4) a[i] = __vimax3_s32(int(a[i]), int(b + it), int(c - i));                      // VIMNMX3
      if (OP == 5) a[i] = max(int(a[i]), int(b + it));                                           // VIMNMX
      if (OP == 6) a[i] = __vimax3_s16x2(a[i], b + it, c - i);                                   // VIMNMX3.S16x2
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= a[i];
  if (s == 0xdeadbeefu) out[0] = s;
}

// POPC with independent inputs each iteration (popc of xor with a changing operand) — closer to K1's use
__global__ void __launch_bounds__(1024) popc_xor_kernel(uint32_t *out, uint32_t seed) {
  uint32_t q[8], acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = seed * (i + 1) + threadIdx.x;
  uint32_t d = seed;
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      d = d * 1664525u + 1013904223u;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[r] += __popc(q[i] ^ (d + i));
    }
  }
  if ((acc[0] ^ acc[1] ^ acc[2] ^ acc[3]) == 0xdeadbeefu) out[0] = acc[0];
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
  uint32_t *d_out;
  CK(cudaMalloc(&d_out, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 2, threads = 1024;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_max_mhz\": %.0f", prop.name, sms, clock_khz / 1000.0);
  const char *names[7] = {"popc", "lop3", "iadd", "imad", "vimnmx3", "vimnmx", "vimnmx3_s16x2"};
  for (int op = 0; op < 7; ++op) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      if (op == 0) rate_kernel<0><<<blocks, threads>>>(d_out, 12345u + rep);
      if (op == 1) rate_kernel<1><<<blocks, threads>>>(d_out, 12345u + rep);
      if (op == 2) rate_kernel<2><<<blocks, threads>>>(d_out, 12345u + rep);
      if (op == 3) rate_kernel<3><<<blocks, threads>>>(d_out, 12345u + rep);
      if (op == 4) rate_kernel<4><<<blocks, threads>>>(d_out, 12345u + rep);
      if (op == 5) rate_kernel<5><<<blocks, threads>>>(d_out, 12345u + rep);
      if (op == 6) rate_kernel<6><<<blocks, threads>>>(d_out, 12345u + rep);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    const double ops = double(blocks) * threads * double(ITER) * ILP;
    printf(", \"%s_gops\": %.1f, \"%s_per_clk_per_sm_at_max_clock\": %.2f", names[op], ops / best / 1e6, names[op],
           ops / (best * 1e-3) / sms / (clock_khz * 1e3));
  }
  {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      popc_xor_kernel<<<blocks, threads>>>(d_out, 777u + rep);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    const double cmps = double(blocks) * threads * double(ITER) * 4;  // one cmp = 8 xor+popc
    printf(", \"xor_popc_gcmp\": %.1f", cmps / best / 1e6);
  }
  printf("}\n");
  return 0;
}
