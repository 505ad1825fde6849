// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM for the shapes / packing the K1 epilogue can use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_ARGS32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define REGS32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"

// MODE 0: 32x32b.x32 (32 columns -> 32 regs)   MODE 1: 32x32b.x32.pack::16b (64 columns -> 32 regs)
// MODE 2: 32x32b.x16 (16 regs)                  MODE 3: 16x256b.x8 (32 regs; 16 lanes x 256 bit x8)
template <int MODE>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&v)[32]) {
  if (MODE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " REGS32 ", [%32];" : LD_ARGS32(v) : "r"(taddr) : "memory");
  if (MODE == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " REGS32 ", [%32];" : LD_ARGS32(v) : "r"(taddr) : "memory");
  if (MODE == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
  if (MODE == 3) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " REGS32 ", [%32];" : LD_ARGS32(v) : "r"(taddr) : "memory");
}

template <int MODE, int DEPTH>
__global__ void __launch_bounds__(512, 1) ld_kernel(int iters, int n_warps, long long *cycles, uint32_t *sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < n_warps) {
    const uint32_t taddr = base + (uint32_t((warp & 3) * 32) << 16);
    uint32_t v[DEPTH][32];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) ld<MODE>(taddr + ((it * DEPTH + d) & 3) * 64, v[d]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc ^= v[d][it & 15];
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

template <int MODE, int DEPTH>
int run(const char *name, int n_warps, double regs_per_ld, double cols_per_ld) {
  long long *d_c; uint32_t *d_s;
  CK(cudaMalloc(&d_c, 148 * 8)); CK(cudaMalloc(&d_s, 64));
  const int iters = 4096;
  ld_kernel<MODE, DEPTH><<<148, 512>>>(iters, n_warps, d_c, d_s);
  CK(cudaDeviceSynchronize());
  ld_kernel<MODE, DEPTH><<<148, 512>>>(iters, n_warps, d_c, d_s);
  CK(cudaDeviceSynchronize());
  long long c[148];
  CK(cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
  const double lds = double(iters) * DEPTH * n_warps;            // warp-level ld instructions per SM
  printf("{\"variant\": \"%s\", \"warps\": %d, \"depth\": %d, \"clk_per_warp_ld\": %.1f, \"reg_bytes_per_clk_per_sm\": %.1f, \"tmem_cols_x_lanes_x4B_per_clk_per_sm\": %.1f}\n",
         name, n_warps, DEPTH, avg / (lds / n_warps), lds * regs_per_ld * 32 * 4 / avg, lds * cols_per_ld * 32 * 4 / avg);
  cudaFree(d_c); cudaFree(d_s);
  return 0;
}

int main() {
  for (int w : {1, 4, 8, 16}) {
    if (run<0, 1>("32x32b.x32", w, 32, 32)) return 1;
    if (run<0, 2>("32x32b.x32", w, 32, 32)) return 1;
    if (run<1, 2>("32x32b.x32.pack16", w, 32, 64)) return 1;
    if (run<2, 2>("32x32b.x16", w, 16, 16)) return 1;
    if (run<3, 2>("16x256b.x8", w, 32, 32)) return 1;
  }
  return 0;
}
