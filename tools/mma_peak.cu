// Micro-benchmark: dense int8 tcgen05.mma issue rate per SM (the denominator of the K1 tensor-form roofline).
// One CTA per SM; one thread issues M128 N256 K32 kind::i8 MMAs (cta_group::1) back to back from resident shared-
// memory operands (random +-1 bytes, so the datapath toggles like the real kernel) into two alternating TMEM stages.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_peak tools/mma_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= uint64_t((addr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// pattern: 0 = A and B random +-1 (the K1 encoding), 1 = A +-1, B random {0,1}, 2 = A and B {0,1}, 3 = all zero
template <int N>
__global__ void __launch_bounds__(128, 1) peak_kernel(int tiles, long long *cycles, int pattern) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a = smem;                 // 128 rows x 256 B  (2 k-halves of 128 B)
  uint8_t *b = smem + 32768;         // N rows x 256 B
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
  for (int i = threadIdx.x; i < (32768 + N * 256) / 4; i += 128) {
    uint32_t w = 0;
    for (int j = 0; j < 4; ++j) {
      x = x * 1664525u + 1013904223u;
      const bool bit = (x >> 16) & 1;
      const bool is_b = i >= 32768 / 4;
      uint32_t v = bit ? 0x01u : 0xFFu;
      if ((pattern == 1 && is_b) || pattern == 2) v = bit ? 0x01u : 0x00u;
      if (pattern == 3) v = 0u;
      w |= v << (8 * j);
    }
    reinterpret_cast<uint32_t *>(smem)[i] = w;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int t = 0; t < tiles; ++t) {
      const int s = t & 1;
      if (t >= 2) {  // the stage's previous tile must have completed (as an epilogue would require)
        const uint32_t parity = ((t >> 1) - 1) & 1;
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
      }
      const uint32_t d = tmem + s * (N <= 256 ? 256 : 0);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        const uint64_t ad = desc_sw128(smem_u32(a + kh * 128 * 128));
        const uint64_t bd = desc_sw128(smem_u32(b + kh * N * 128));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (kh | ks) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad + uint64_t(ks * 2)), "l"(bd + uint64_t(ks * 2)), "r"(idesc), "r"(acc) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[s])) : "memory");
    }
    for (int s = 0; s < 2; ++s) {  // drain: last tile of each stage
      const int n_s = (tiles + 1 - s) / 2;
      if (n_s == 0) continue;
      const uint32_t parity = (n_s - 1) & 1;
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N>
int run(int sms, int tiles, int reps, int pattern = 0) {
  long long *d_c;
  CK(cudaMalloc(&d_c, sizeof(long long) * sms));
  const int smem = 32768 + N * 256 + 1024;
  CK(cudaFuncSetAttribute(peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  peak_kernel<N><<<sms, 128, smem>>>(tiles, d_c, pattern);
  CK(cudaDeviceSynchronize());
  float best = 1e30f, sum = 0.f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    peak_kernel<N><<<sms, 128, smem>>>(tiles, d_c, pattern);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
    sum += ms;
  }
  long long *h = (long long *)malloc(sizeof(long long) * sms);
  CK(cudaMemcpy(h, d_c, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double cyc = 0;
  for (int i = 0; i < sms; ++i) cyc += double(h[i]);
  cyc /= sms;
  const double macs = double(sms) * tiles * 128.0 * N * 256.0;
  printf("{\"pattern\": %d, \"shape\": \"M128 N%d K256 tile, cta_group::1\", \"sms\": %d, \"tiles_per_cta\": %d, \"best_ms\": %.4f, \"mean_ms\": %.4f, "
         "\"tops_best\": %.1f, \"tops_mean\": %.1f, \"gcmp_best\": %.1f, \"gcmp_mean\": %.1f, \"mac_per_clk_per_sm\": %.1f}\n",
         pattern, N, sms, tiles, best, sum / reps, 2.0 * macs / (best * 1e-3) / 1e12, 2.0 * macs / (sum / reps * 1e-3) / 1e12,
         macs / 256.0 / (best * 1e-3) / 1e9, macs / 256.0 / (sum / reps * 1e-3) / 1e9, 128.0 * N * 256.0 * tiles / cyc);
  cudaFree(d_c);
  free(h);
  return 0;
}

int main(int argc, char **argv) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int tiles = argc > 1 ? atoi(argv[1]) : 20000;
  if (run<256>(sms, tiles, 5)) return 1;
  if (run<128>(sms, tiles, 5)) return 1;
  // a long run (seconds): the sustained figure under the power cap
  if (run<256>(sms, tiles * 20, 3)) return 1;
  // operand-value patterns (power under the cap depends on what the multipliers see)
  if (argc > 2)
    for (int pattern = 1; pattern <= 3; ++pattern)
      if (run<256>(sms, tiles * 20, 3, pattern)) return 1;
  return 0;
}
