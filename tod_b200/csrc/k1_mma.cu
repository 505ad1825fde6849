// K1, tensor-core formulation: exact Hamming k-NN as a dense int8 contraction on tcgen05 (sm_100a).
//
// Replaces `matcher_->knnMatch(descriptors, matches, 5)` at src/detection/DescriptorMatcher.cpp:211 of the reference
// (cv::BFMatcher(NORM_HAMMING) semantics, see oracle/hamming_knn.py) — same contract and same packed-key output as
// k1_popc.cu, so the merge kernels are shared.
//
// Identity: with every descriptor bit b mapped to the int8 value (1 - 2b), a . b = 256 - 2 * Hamming(a, b); the int32
// accumulation is exact, so distances are bit-identical to XOR+POPC.
//
// Structure (one CTA per SM, 192 threads, warp-specialised):
//   warp 0 / lane 0 : TMA producer.  Loads the CTA's QT resident query tiles (128 x 256 B each, once) and streams
//                     database tiles (256 rows x 256 B) through a 2-stage shared-memory ring, SWIZZLE_128B boxes of
//                     128 B x rows, completion on mbarriers (cp.async.bulk.tensor -> UTMALDG).
//   warp 1 / lane 0 : MMA issuer.  Per (db tile, query tile): 8 x tcgen05.mma.kind::i8 (M128 N256 K32) from shared-
//                     memory descriptors into one of two 256-column TMEM accumulator stages, then tcgen05.commit to
//                     the epilogue's mbarrier; a second commit frees the db ring slot.  Warp 1 also owns TMEM alloc.
//   warps 2-5       : epilogue.  Thread = one query row (TMEM lane).  tcgen05.ld 32 columns at a time, a max-reduce
//                     per group against the query's current threshold (fast path), and a rare divergent slow path
//                     that inserts candidates into a register top-k of packed keys (distance << 23 | global_row).
//                     Rows ascend within a thread, so ties never displace earlier rows (same argument as k1_popc.cu).
// Grid = (query groups, db chunks): CTAs that share a db chunk are adjacent in launch order, so the int8-expanded
// database (256 B / descriptor) is read from HBM about once per frame batch and served from L2 to the other groups.
#include <cuda.h>

#include <algorithm>

#include "ptx.cuh"
#include "tod_internal.h"

namespace tod {
namespace {

constexpr int kBlockM = 128;          // queries per tile (TMEM lanes)
constexpr int kBlockN = 256;          // db rows per tile (TMEM columns of one accumulator stage)
constexpr int kQT = 2;                // resident query tiles per CTA
constexpr int kBStages = 2;           // db ring depth
constexpr int kAccStages = 2;         // TMEM accumulator stages (2 x 256 = all 512 columns)
constexpr int kThreadsMma = 192;      // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kKBytes = 256;          // int8 elements (= bytes) per descriptor
constexpr int kSwizzleBytes = 128;    // inner TMA box / swizzle span
constexpr int kATileBytes = kBlockM * kKBytes;   // 32 KB
constexpr int kBTileBytes = kBlockN * kKBytes;   // 64 KB
constexpr int kSmemMma = kQT * kATileBytes + kBStages * kBTileBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t kSpinLimit = 1u << 26;  // bounded waits: a protocol bug traps instead of hanging the GPU

__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(map), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows packed at 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);        // start address        bits [0,14)
  d |= uint64_t(1) << 16;                            // leading byte offset  bits [16,30) (unused with swizzle)
  d |= uint64_t(1024 >> 4) << 32;                    // stride byte offset   bits [32,46)
  d |= uint64_t(1) << 46;                            // descriptor version 1 (sm_100)
  d |= uint64_t(2) << 61;                            // layout type SWIZZLE_128B
  return d;
}

// Instruction descriptor: kind::i8, A/B signed 8-bit K-major, D = S32, dense, M = 128, N = 256.
constexpr uint32_t kInstrDesc = (2u << 4) /*c S32*/ | (1u << 7) /*a S8*/ | (1u << 10) /*b S8*/ | (0u << 15) | (0u << 16) |
                                (uint32_t(kBlockN >> 3) << 17) | (uint32_t(kBlockM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int K>
__global__ void __launch_bounds__(kThreadsMma, 1)
k1_mma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db, int nq,
              int shard_rows, uint32_t global_row_base, int rows_per_chunk, uint32_t thr_init,
              uint32_t *__restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_smem = smem;                                   // [kQT][2 k-halves][128 rows][128 B]
  uint8_t *b_smem = smem + kQT * kATileBytes;               // [kBStages][2 k-halves][256 rows][128 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(b_smem + kBStages * kBTileBytes);
  uint64_t *a_full = bars;                                  // 1
  uint64_t *b_full = bars + 1;                              // kBStages
  uint64_t *b_empty = b_full + kBStages;                    // kBStages
  uint64_t *acc_full = b_empty + kBStages;                  // kAccStages
  uint64_t *acc_empty = acc_full + kAccStages;              // kAccStages
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_group = blockIdx.x;
  const int chunk = blockIdx.y;
  const int q_row0 = q_group * (kQT * kBlockM);
  const int row0 = chunk * rows_per_chunk;
  const int row1 = min(shard_rows, row0 + rows_per_chunk);
  const int n_tiles = (row1 - row0 + kBlockN - 1) / kBlockN;

  if (threadIdx.x == 0) {
    ptx::mbar_init(a_full, 1);
    for (int s = 0; s < kBStages; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], 4);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(a_full, kQT * kATileBytes);
      for (int j = 0; j < kQT; ++j)
        for (int kh = 0; kh < 2; ++kh)
          tma_load_2d(a_smem + j * kATileBytes + kh * (kBlockM * kSwizzleBytes), &map_q, kh * kSwizzleBytes,
                      q_row0 + j * kBlockM, a_full);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kBStages;
        mbar_wait_bounded(&b_empty[s], ((t / kBStages) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&b_full[s], kBTileBytes);
        for (int kh = 0; kh < 2; ++kh)
          tma_load_2d(b_smem + s * kBTileBytes + kh * (kBlockN * kSwizzleBytes), &map_db, kh * kSwizzleBytes,
                      row0 + t * kBlockN, &b_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      mbar_wait_bounded(a_full, 0);
      tc_fence_after();
      uint32_t acc_iter = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kBStages;
        mbar_wait_bounded(&b_full[s], (t / kBStages) & 1);
        tc_fence_after();
        for (int j = 0; j < kQT; ++j, ++acc_iter) {
          const uint32_t as = acc_iter % kAccStages;
          mbar_wait_bounded(&acc_empty[as], ((acc_iter / kAccStages) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kBlockN;
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            const uint64_t a_desc =
                make_kmajor_sw128_desc(ptx::smem_u32(a_smem + j * kATileBytes + kh * (kBlockM * kSwizzleBytes)));
            const uint64_t b_desc =
                make_kmajor_sw128_desc(ptx::smem_u32(b_smem + s * kBTileBytes + kh * (kBlockN * kSwizzleBytes)));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)  // 32-byte K steps inside the 128-byte swizzle span: +2 in 16-byte units
              umma_i8(d_tmem, a_desc + uint64_t(ks * 2), b_desc + uint64_t(ks * 2), (kh | ks) ? 1u : 0u);
          }
          umma_commit(&acc_full[as]);  // accumulator ready for the epilogue (implies fence::before_thread_sync)
        }
        umma_commit(&b_empty[s]);      // db ring slot may be refilled once these MMAs have read it
      }
    }
  } else {
    // ===================================== epilogue: warps 2..5 =====================================
    const int quarter = warp & 3;                         // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const int row_in_tile = quarter * 32 + lane;
    uint32_t best[kQT][K];
    int thr_dot[kQT];
#pragma unroll
    for (int j = 0; j < kQT; ++j) {
#pragma unroll
      for (int i = 0; i < K; ++i) best[j][i] = kKeyEmpty;
      thr_dot[j] = 256 - 2 * int(thr_init);               // distance < thr  <=>  dot > 256 - 2 thr
    }
    uint32_t acc_iter = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int cols_valid = min(kBlockN, (row1 - row0) - t * kBlockN);
      const uint32_t grow0 = global_row_base + uint32_t(row0 + t * kBlockN);
#pragma unroll
      for (int j = 0; j < kQT; ++j, ++acc_iter) {
        const uint32_t as = acc_iter % kAccStages;
        mbar_wait_bounded(&acc_full[as], (acc_iter / kAccStages) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + as * kBlockN;
#pragma unroll 1
        for (int c = 0; c < kBlockN; c += 32) {
          if (c >= cols_valid) break;
          uint32_t v[32];
          tmem_ld32(taddr + uint32_t(c), v);
          tmem_wait_ld();
          if (c + 32 > cols_valid) {  // ragged end of the shard: TMA zero-filled rows must never match
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c + i >= cols_valid) v[i] = 0x80000000u;
          }
          int m = int(v[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) m = max(m, int(v[i]));
          if (m > thr_dot[j]) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int dot = int(v[i]);
              if (dot > thr_dot[j]) {
                const uint32_t dist = uint32_t(256 - dot) >> 1;
                best[j][K - 1] = (dist << kKeyRowBits) | (grow0 + uint32_t(c + i));
#pragma unroll
                for (int x = K - 1; x > 0; --x) {
                  const uint32_t lo = min(best[j][x - 1], best[j][x]);
                  const uint32_t hi = max(best[j][x - 1], best[j][x]);
                  best[j][x - 1] = lo;
                  best[j][x] = hi;
                }
                thr_dot[j] = 256 - 2 * int(min(thr_init, best[j][K - 1] >> kKeyRowBits));
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
      }
    }
#pragma unroll
    for (int j = 0; j < kQT; ++j) {
      const int qi = q_row0 + j * kBlockM + row_in_tile;
      if (qi < nq) {
        uint32_t *o = partial + (size_t(chunk) * nq + qi) * K;
#pragma unroll
        for (int i = 0; i < K; ++i) o[i] = best[j][i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// bits -> +-1 int8:  out[row][8 * byte + b] = (in[row][byte] >> b) & 1 ? -1 : +1
__global__ void __launch_bounds__(256) expand_pm1_kernel(const uint8_t *__restrict__ in, uint2 *__restrict__ out,
                                                         size_t n_bytes) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_bytes) return;
  const uint32_t x = in[i];
  const uint32_t lo = ((x & 0xFu) * 0x00204081u) & 0x01010101u;
  const uint32_t hi = ((x >> 4) * 0x00204081u) & 0x01010101u;
  out[i] = make_uint2(lo * 0xFEu + 0x01010101u, hi * 0xFEu + 0x01010101u);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int K>
cudaError_t launch_mma_k(const K1Plan &plan, const CUtensorMap &map_q, const CUtensorMap &map_db, int nq, int64_t rows,
                         uint32_t base, uint32_t thr_init, uint32_t *partial, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k1_mma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMma);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(plan.n_qtiles, plan.n_chunks);
  k1_mma_kernel<K><<<grid, kThreadsMma, kSmemMma, stream>>>(map_q, map_db, nq, int(rows), base, plan.rows_per_chunk,
                                                            thr_init, partial);
  count_launch();
  return cudaGetLastError();
}

}  // namespace

K1Plan k1_mma_plan(int nq, int64_t shard_rows, int sm_count) {
  K1Plan p{};
  p.q_per_thread = 0;
  p.q_tile = kQT * kBlockM;
  p.n_qtiles = std::max(1, (nq + p.q_tile - 1) / p.q_tile);
  const int64_t max_chunks = std::max<int64_t>(1, (shard_rows + kBlockN - 1) / kBlockN);
  int64_t target = std::max<int64_t>(1, sm_count / p.n_qtiles);  // one CTA per SM, at most one wave
  target = std::min(target, max_chunks);
  int64_t rpc = (shard_rows + target - 1) / target;
  rpc = std::max<int64_t>(kBlockN, (rpc + kBlockN - 1) / kBlockN * kBlockN);
  p.rows_per_chunk = int(rpc);
  p.n_chunks = int(std::max<int64_t>(1, (shard_rows + rpc - 1) / rpc));
  return p;
}

cudaError_t launch_expand_pm1(const void *d_bits, void *d_int8, int64_t rows, cudaStream_t stream) {
  const size_t n_bytes = size_t(rows) * 32;
  if (n_bytes == 0) return cudaSuccess;
  const unsigned blocks = unsigned((n_bytes + 255) / 256);
  expand_pm1_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint8_t *>(d_bits), static_cast<uint2 *>(d_int8),
                                                n_bytes);
  count_launch();
  return cudaGetLastError();
}

// 2-D tensor map over an int8-expanded descriptor matrix [rows][256], box = 128 bytes x box_rows, SWIZZLE_128B.
bool make_desc_tensor_map(void *map_out, const void *d_int8, int64_t rows, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cuuint64_t(kKBytes), cuuint64_t(std::max<int64_t>(rows, 1))};
  const cuuint64_t strides[1] = {cuuint64_t(kKBytes)};
  const cuuint32_t box[2] = {cuuint32_t(kSwizzleBytes), cuuint32_t(box_rows)};
  const cuuint32_t elem[2] = {1, 1};
  return fn(static_cast<CUtensorMap *>(map_out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(d_int8), dims,
            strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int k1_mma_query_box_rows() { return kBlockM; }
int k1_mma_db_box_rows() { return kBlockN; }
size_t tensor_map_bytes() { return sizeof(CUtensorMap); }

cudaError_t launch_k1_mma(const K1Plan &plan, const void *map_q, const void *map_db, int nq, int64_t shard_rows,
                          uint32_t global_row_base, int k, uint32_t radius, uint32_t *d_partial, cudaStream_t stream) {
  const uint32_t thr_init = radius ? min(radius + 1u, 511u) : 511u;
  const CUtensorMap &mq = *static_cast<const CUtensorMap *>(map_q);
  const CUtensorMap &md = *static_cast<const CUtensorMap *>(map_db);
  switch (k) {
    case 1: return launch_mma_k<1>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 2: return launch_mma_k<2>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 3: return launch_mma_k<3>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 4: return launch_mma_k<4>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 5: return launch_mma_k<5>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 6: return launch_mma_k<6>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 7: return launch_mma_k<7>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    case 8: return launch_mma_k<8>(plan, mq, md, nq, shard_rows, global_row_base, thr_init, d_partial, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tod
