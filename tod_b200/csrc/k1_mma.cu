// K1, tensor-core formulation: exact Hamming k-NN as a dense int8 contraction on tcgen05 (sm_100a), CTA pairs.
//
// Replaces `matcher_->knnMatch(descriptors, matches, 5)` at src/detection/DescriptorMatcher.cpp:211 of the reference
// (cv::BFMatcher(NORM_HAMMING) semantics, see oracle/hamming_knn.py) — same contract and same packed-key output as
// k1_popc.cu, so the merge kernels are shared.
//
// Identity: query bits x are expanded to the int8 values 2x - 1 (+-1), database bits y to the int8 values y (0 / 1):
//   a . b = |x & y| - |~x & y|   and   Hamming(x, y) = popcount(x) - a . b ;
// the int32 accumulation is exact, so distances are bit-identical to XOR+POPC, and popcount(x) is a per-query
// constant folded into the thresholds.  Why 0/1 and not +-1 on the database side: the kernel is power-capped, and
// the multipliers burn measurably less on 0/1 operands — tools/mma_peak.cu sustains 4.59 POPS with a 0/1 B operand
// against 3.95 POPS with +-1 x +-1 under the same 1 kW cap (profiles/int_peaks.json).
//
// Structure — a cluster of 2 CTAs (one per SM of a TPC) runs tcgen05.mma.cta_group::2 tiles of M = 256 queries
// (128 per CTA) x N = 256 database rows (128 per CTA) x K = 256:
//   warp 0 / lane 0 : TMA producer (both CTAs).  Loads the CTA's own QT resident query tiles once and streams ITS
//                     128-row half of every database tile through a 4-stage shared-memory ring (SWIZZLE_128B boxes of
//                     128 B x 128 rows, cp.async.bulk.tensor.cta_group::2 -> UTMALDG, completion on the LEADER's
//                     mbarrier).  A fetched database row is thus shared by 512 queries: 0.5 byte of L2 traffic per
//                     comparison, which is what lets the tensor pipe run near its peak (DESIGN.md, K1 roofline).
//   warp 1 / lane 0 : MMA issuer (leader CTA only).  Per (db tile, query tile): 8 x tcgen05.mma.cta_group::2.kind::i8
//                     (M256 N256 K32) into one of two 256-column TMEM accumulator stages; tcgen05.commit multicast to
//                     both CTAs' mbarriers (accumulator ready / ring slot free).  Warp 1 of each CTA owns TMEM alloc.
//   warps 2-9       : epilogue (both CTAs), two warps per TMEM lane quarter (each takes 128 of the 256 columns).
//                     Thread = one query row.  Fast path (a few dozen instructions, instruction-cache resident):
//                     tcgen05.ld 32 columns at a time, software-pipelined over two register buffers, a VIMNMX3 tree
//                     per group and ONE compare against the query's threshold.  Slow path (rare, warp-uniform call of
//                     one out-of-line function): re-reads the 32 columns from TMEM and inserts candidates into the
//                     thread's top-k list of packed keys (distance << 23 | global_row) kept in shared memory.  Rows
//                     ascend within a thread, so ties never displace earlier rows (same argument as k1_popc.cu).
//                     The two column halves are separate merge sources.
//   Thresholds      : every list prunes with min(own k-th distance [strict], best k-th distance published by ANY
//                     list of the same query [non-strict]) — the latter through one u32 per query in global memory
//                     (atomicMin on publish, one relaxed load per tile).  Exact: a candidate farther than some list's
//                     k-th best can never be in the global top-k; equal distances are kept for the tie-break.
// Grid: 1-D, two phases (k1_mma_plan): first the query groups that fill whole rounds of the machine, one CTA each
// over the WHOLE shard (one cold start of the top-k lists per query, all CTAs sweeping the db in step so a row fetched
// from HBM serves every SM out of L2); then the left-over groups, split into db chunks (query group fastest) so that
// the last round is balanced too.  Cluster (2,1,1) in pair mode.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "ptx.cuh"
#include "tod_internal.h"

namespace tod {
namespace {

#ifndef TOD_MMA_PAIR
#define TOD_MMA_PAIR 1                // 1: CTA pairs (tcgen05 cta_group::2), 0: one CTA per tile (cta_group::1)
#endif
constexpr int kCtas = TOD_MMA_PAIR ? 2 : 1;   // CTAs cooperating on one MMA tile
constexpr int kBlockM = 128;          // queries per tile and CTA (TMEM lanes)
constexpr int kBlockN = 256;          // db rows per MMA tile (TMEM columns of one accumulator stage)
constexpr int kHalfN = kBlockN / kCtas;  // db rows each CTA stages per tile
constexpr int kQT = 2;                // resident query tiles per CTA
constexpr int kBStages = TOD_MMA_PAIR ? 4 : 2;  // db ring depth per CTA (4 x 32 KB or 2 x 64 KB)
constexpr int kAccStages = 2;         // TMEM accumulator stages (2 x 256 = all 512 columns)
constexpr int kEpiWarps = 8;
constexpr int kThreadsMma = 64 + 32 * kEpiWarps;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kKBytes = 256;          // int8 elements (= bytes) per descriptor
constexpr int kSwizzleBytes = 128;    // inner TMA box / swizzle span
constexpr int kATileBytes = kBlockM * kKBytes;   // 32 KB
constexpr int kBHalfBytes = kHalfN * kKBytes;    // 32 KB
constexpr int kListWords = kQT * TOD_MAX_K * kBlockM;  // top-k lists: [query tile][slot][row]
static_assert(kQT == 2 && kAccStages == 2 && kEpiWarps == 8,
              "epilogue warp set j (4 warps, one per TMEM lane quarter) serves query tile j = accumulator stage j");
constexpr int kSmemMma = kQT * kATileBytes + kBStages * kBHalfBytes + 1024 /*align*/ + 256 /*barriers*/ + kListWords * 4;
constexpr uint32_t kSpinLimit = 1u << 26;  // bounded waits: a protocol bug traps instead of hanging the GPU

#ifndef TOD_K1_STATS
#define TOD_K1_STATS 0                // 1: instrumented build (tools/k1_stats.py); never the production library
#endif
#if TOD_K1_STATS
// [0] slow calls (warp level) [1] cycles inside slow calls [2] MMA issuer cycles waiting acc_empty [3] ... waiting
// b_full [4] epilogue cycles waiting acc_full (sum over warps) [5] CTA lifetime cycles [6] epilogue groups (warp
// level) [7] epilogue busy cycles (sum over warps, excludes acc_full waits) [8] CTAs [9] producer cycles waiting b_empty
__device__ unsigned long long g_k1_stats[16];
// per tile index inside a CTA's sweep (63 = all later tiles): [0] MMA issuer cycles waiting for acc_empty, [1] slow
// calls (warp level), [2] cycles inside slow calls, [3] MMA issuer cycles waiting for b_full
__device__ unsigned long long g_k1_tile[4][64];
#define STAT_T0(var) const long long var = clock64()
#define STAT_ADD(acc, var) acc += (unsigned long long)(clock64() - var)
#else
#define STAT_T0(var)
#define STAT_ADD(acc, var)
#endif

__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
#if TOD_MMA_PAIR
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
#else
  return 0u;
#endif
}

// all threads of the CTA (single mode) or of both CTAs of the pair
__device__ __forceinline__ void cluster_sync_all() {
#if TOD_MMA_PAIR
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
  __syncthreads();
#endif
}

// shared::cluster address of `local` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
#if TOD_MMA_PAIR
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
#else
  return local;  // a shared::cta address is a valid shared::cluster address of the executing CTA
#endif
}

// Arrive on a barrier given by its shared::cluster address (own CTA or the pair's leader).  Default semantics
// (.release.cta) on purpose: a cluster-scope release costs hundreds of cycles on the accumulator hand-off, and no
// generic-proxy data is published through these barriers (TMEM reads are ordered by tcgen05.fence).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// plain TMA load (own smem, own mbarrier)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(map), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// pair TMA load: data lands in THIS CTA's smem, the transaction bytes complete on `bar_cluster_addr` (leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, int c0, int c1,
                                                 uint32_t bar_cluster_addr) {
#if TOD_MMA_PAIR
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
#else
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
#endif
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

// Instruction descriptor: kind::i8, A/B signed 8-bit K-major, D = S32, dense, M = 256 (pair), N = 256.
constexpr uint32_t kInstrDesc = (2u << 4) /*c S32*/ | (1u << 7) /*a S8*/ | (1u << 10) /*b S8*/ |
                                (uint32_t(kBlockN >> 3) << 17) | (uint32_t((kCtas * kBlockM) >> 4) << 24);

__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
#if TOD_MMA_PAIR
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
#endif
}

// commit all prior MMAs of the pair to the mbarrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
#if TOD_MMA_PAIR
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
#else
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar))
               : "memory");
#endif
}

// 64 accumulator columns -> 32 registers: register i = (low 16 bits of column 2i+1) << 16 | (low 16 bits of column 2i).
// The dot products lie in [-256, 256], so the low halves are the values as int16.
__device__ __forceinline__ void tmem_ld64_pack16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared::cta.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Slow path of the epilogue, COMPACT and fed from REGISTERS: the accumulator stage has already been handed back to the
// tensor core when it runs.  History (profiles/r1_k1_stats_*.jsonl): the first version, a 32-way unrolled scan with a
// divergent insert per column that re-read TMEM, cost ~2900 cycles per call; as one out-of-line copy of this routine
// ~1200; inlined at its 8 call sites (TOD_K1_INLINE_SLOW, default) another 7 % faster on small shards.
// Runs warp-uniformly for one 32-column sub-group in which some lane saw a
// candidate.  p0..p15 hold the sub-group's dot products as packed int16 pairs (register r = columns 2r, 2r + 1).
// Builds each lane's hit mask branch-free; every lane then walks its own hit columns (usually none or one) and inserts
// the candidates into its sorted top-k list in shared memory.
// Returns the new dot-product threshold.
//   list + i * 4 * kBlockM (i < k), a shared::cta address : ascending packed keys
//   n_valid : columns of the sub-group that are real db rows (the last tile of a chunk may be partial)
//   popx    : popcount of this thread's query descriptor (distance = popx - dot)
#ifndef TOD_K1_INLINE_SLOW
#define TOD_K1_INLINE_SLOW 1  // 1 (default): inlined at its 8 call sites — +7% on small shards vs one out-of-line copy
#endif
#if TOD_K1_INLINE_SLOW
__device__ __forceinline__
#else
__device__ __noinline__
#endif
int k1_mma_slow_scan(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4,
                                             uint32_t p5, uint32_t p6, uint32_t p7, uint32_t p8, uint32_t p9,
                                             uint32_t p10, uint32_t p11, uint32_t p12, uint32_t p13, uint32_t p14,
                                             uint32_t p15, uint32_t list, int k, uint32_t thr_init, uint32_t grow,
                                             int n_valid, int thr_dot, uint32_t *gthr, int popx) {
  constexpr uint32_t kStride = 4u * kBlockM;  // bytes between consecutive slots of one list
  const uint32_t v[16] = {p0, p1, p2, p3, p4, p5, p6, p7, p8, p9, p10, p11, p12, p13, p14, p15};
  // Hit mask, 3 instructions per register: d = v - (thr + 1) per int16 half (VIADD.16x2, no carry between halves;
  // |d| < 1100 so no wrap) is non-negative exactly where dot > thr.  Bit r of the mask <- low half of register r
  // (column 2r), bit 16 + r <- high half (column 2r + 1).
  const uint32_t neg = uint32_t(-(thr_dot + 1)) & 0xFFFFu;
  const uint32_t neg2 = neg | (neg << 16);
  uint32_t m0 = 0, m1 = 0;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const uint32_t d = __vadd2(v[r], neg2);
    const uint32_t sh = r < 15 ? (d >> (15 - r)) : d;          // sign bits of the halves -> bits r and 16 + r
    const uint32_t sel = 0x00010001u << r;
    if (r & 1) m1 |= ~sh & sel;
    else m0 |= ~sh & sel;
  }
  uint32_t mask = m0 | m1;
  if (n_valid < 32) {
    const int lo = (n_valid + 1) >> 1, hi = n_valid >> 1;     // valid even / odd columns
    mask &= ((1u << lo) - 1u) | (((1u << hi) - 1u) << 16);
  }
#if TOD_K1_STATS
  {
    const unsigned trips = __reduce_max_sync(0xffffffffu, unsigned(__popc(mask)));
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_k1_stats[13], (unsigned long long)trips);
  }
#endif
  // Every lane walks ITS OWN hit columns: the warp runs max-over-lanes trips, not one per distinct column — in the
  // cold first tiles of a list, where every lane has hits, that is up to 32x fewer trips.  The walk visits even
  // columns before odd ones, so inside the call candidates are compared by their full key (distance, row), which
  // makes the result independent of the visiting order; across calls rows only grow.
  bool inserted = false;
  uint32_t worst = lds_u32(list + uint32_t(k - 1) * kStride);   // k-th best key (0xFFFFFFFF while the list is not full)
#pragma unroll 1
  while (mask) {
    const int b = __ffs(int(mask)) - 1;
    mask &= mask - 1u;
    // v[b & 15] through a select tree (static register indices, per-lane predicates)
    uint32_t s8[8], s4[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s8[j] = (b & 8) ? v[8 + j] : v[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) s4[j] = (b & 4) ? s8[4 + j] : s8[j];
    const uint32_t s2a = (b & 2) ? s4[2] : s4[0], s2b = (b & 2) ? s4[3] : s4[1];
    const uint32_t x = (b & 1) ? s2b : s2a;
    const int dot = (b & 16) ? (int(x) >> 16) : (int(x << 16) >> 16);
    const uint32_t dist = uint32_t(popx - dot);               // Hamming = popcount(query) - dot
    const uint32_t key = (dist << kKeyRowBits) | (grow + uint32_t(2 * (b & 15) + (b >> 4)));
    if (key < worst) {
      // sorted insert: the new key displaces the worst entry and sinks to its place
      int pos = k - 1;
      while (pos > 0) {
        const uint32_t prev = lds_u32(list + uint32_t(pos - 1) * kStride);
        if (prev <= key) break;
        sts_u32(list + uint32_t(pos) * kStride, prev);
        --pos;
      }
      sts_u32(list + uint32_t(pos) * kStride, key);
      worst = lds_u32(list + uint32_t(k - 1) * kStride);
      inserted = true;
    }
  }
  if (inserted) {
    const uint32_t kth = min(thr_init, worst >> kKeyRowBits);  // strict bound of this list (511 while not full)
    thr_dot = max(thr_dot, popx - int(kth));                  // distance < kth  <=>  dot > popx - kth
  }
  if (inserted && gthr) {
    const uint32_t kth = worst >> kKeyRowBits;   // 511 while the list is not full
    // publish as a fire-and-forget reduction (RED.MIN): a returning atomic would stall the warp for a full L2
    // round trip on every insert.  Everyone, this list included, picks the bound up at its next per-tile refresh.
    if (kth < 511u) atomicMin(gthr, kth);
  }
  return thr_dot;
}

#if TOD_MMA_PAIR
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsMma, 1)
#else
__global__ void __launch_bounds__(kThreadsMma, 1)
#endif
k1_mma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db, int nq,
              int shard_rows, int seg_row0, uint32_t global_row_base, int rows_per_chunk, int n_chunks, int full_units,
              int n_sources, uint32_t thr_init, int K, uint32_t *__restrict__ partial, uint32_t *__restrict__ gthr, const uint32_t *__restrict__ popq
#if TOD_K1_STATS
              , int debug_mode   // ablation knobs of the instrumented build only (tools/k1_stats.py)
#endif
              ) {
#if !TOD_K1_STATS
  constexpr int debug_mode = 0;  // the production library has no run-time knobs: every ablation branch folds away
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_smem = smem;                                   // [kQT][2 k-halves][128 rows][128 B]   (own queries)
  uint8_t *b_smem = smem + kQT * kATileBytes;               // [kBStages][2 k-halves][128 rows][128 B] (own db half)
  uint64_t *bars = reinterpret_cast<uint64_t *>(b_smem + kBStages * kBHalfBytes);
  uint64_t *a_full = bars;                                  // local : own query tiles landed
  uint64_t *a_ready = bars + 1;                             // leader: both CTAs' query tiles landed (count 2)
  uint64_t *b_full = bars + 2;                              // leader: both halves of a db tile landed   [kBStages]
  uint64_t *b_empty = b_full + kBStages;                    // local : ring slot free (multicast commit) [kBStages]
  uint64_t *acc_full = b_empty + kBStages;                  // local : accumulator ready (multicast commit) [2]
  uint64_t *acc_empty = acc_full + kAccStages;              // leader: both CTAs' epilogues drained it (count 16) [2]
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccStages);
  uint32_t *lists = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(bars) + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  STAT_T0(t_cta);
#if TOD_K1_STATS
  unsigned long long st_a = 0, st_b = 0, st_c = 0, st_d = 0, st_e = 0;
#endif
  const uint32_t rank = cluster_ctarank();                  // 0 = leader
  // 1-D grid.  Units (CTA pairs) [0, full_units) take one query group each over the whole shard.  The remaining units
  // are (chunk, query group) items, query group fastest: the CTAs of db chunk c + 1 start after (most of) chunk c has
  // been scanned for the same queries, so all but the first chunk's CTAs begin with tight bounds in gthr.  A pair
  // handles query groups 2p and 2p + 1.  (debug_mode & 32: chunk fastest, for the ablation in profiles/r1_k1_stats.md.)
  const int unit = int(blockIdx.x) / kCtas;
  const bool full_sweep = unit < full_units;
  const int unit2 = unit - full_units;
  const int n_units_q = full_sweep ? 1 : (int(gridDim.x) / kCtas - full_units) / n_chunks;
  const int chunk = full_sweep ? 0 : ((debug_mode & 32) ? unit2 % n_chunks : unit2 / n_units_q);
  const int q_unit = full_sweep ? unit : full_units + ((debug_mode & 32) ? unit2 / n_chunks : unit2 % n_units_q);
  const int q_group = q_unit * kCtas + int(blockIdx.x) % kCtas;
  const int q_row0 = q_group * (kQT * kBlockM);
  const int row0 = full_sweep ? 0 : chunk * rows_per_chunk;
  const int row1 = full_sweep ? shard_rows : min(shard_rows, row0 + rows_per_chunk);
  const int n_tiles = (row1 - row0 + kBlockN - 1) / kBlockN;

  if (threadIdx.x == 0) {
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_ready, kCtas);
    for (int s = 0; s < kBStages; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], kCtas * 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns of both SMs, same warp id in both CTAs
#if TOD_MMA_PAIR
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
#else
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
#endif
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(a_full, kQT * kATileBytes);
      for (int j = 0; j < kQT; ++j)
        for (int kh = 0; kh < 2; ++kh)
          tma_load_2d(a_smem + j * kATileBytes + kh * (kBlockM * kSwizzleBytes), &map_q, kh * kSwizzleBytes,
                      q_row0 + j * kBlockM, a_full);
      mbar_wait_bounded(a_full, 0);
      mbar_arrive_cluster(mapa_u32(ptx::smem_u32(a_ready), 0));
      for (int t = 0; t < n_tiles; ++t) {
        if ((debug_mode & 2) && t >= kBStages) break;  // profiling only: no db streaming (results are garbage)
        const int s = t % kBStages;
        {
          STAT_T0(t0);
          mbar_wait_bounded(&b_empty[s], ((t / kBStages) & 1) ^ 1);
          STAT_ADD(st_a, t0);
        }
        if (rank == 0) ptx::mbar_arrive_expect_tx(&b_full[s], kCtas * kBHalfBytes);
        const uint32_t full_addr = mapa_u32(ptx::smem_u32(&b_full[s]), 0);
        for (int kh = 0; kh < 2; ++kh)
          tma_load_2d_pair(b_smem + s * kBHalfBytes + kh * (kHalfN * kSwizzleBytes), &map_db, kh * kSwizzleBytes,
                           seg_row0 + row0 + t * kBlockN + int(rank) * kHalfN, full_addr);
      }
#if TOD_K1_STATS
      atomicAdd(&g_k1_stats[9], st_a);
#endif
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================================
    if (lane == 0 && rank == 0) {
      mbar_wait_bounded(a_ready, 0);
      tc_fence_after();
      uint32_t acc_iter = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kBStages;
        {
          STAT_T0(t0);
          if (!((debug_mode & 2) && t >= kBStages)) mbar_wait_bounded(&b_full[s], (t / kBStages) & 1);
          STAT_ADD(st_b, t0);
#if TOD_K1_STATS
          atomicAdd(&g_k1_tile[3][min(t, 63)], (unsigned long long)(clock64() - t0));
#endif
        }
        tc_fence_after();
        for (int j = 0; j < kQT; ++j, ++acc_iter) {
          const uint32_t as = acc_iter % kAccStages;
          {
            STAT_T0(t0);
            mbar_wait_bounded(&acc_empty[as], ((acc_iter / kAccStages) & 1) ^ 1);
            STAT_ADD(st_a, t0);
#if TOD_K1_STATS
            atomicAdd(&g_k1_tile[0][min(t, 63)], (unsigned long long)(clock64() - t0));
#endif
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kBlockN;
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            const uint64_t a_desc =
                make_kmajor_sw128_desc(ptx::smem_u32(a_smem + j * kATileBytes + kh * (kBlockM * kSwizzleBytes)));
            const uint64_t b_desc =
                make_kmajor_sw128_desc(ptx::smem_u32(b_smem + s * kBHalfBytes + kh * (kHalfN * kSwizzleBytes)));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)  // 32-byte K steps inside the 128-byte swizzle span: +2 in 16-byte units
              umma_i8_pair(d_tmem, a_desc + uint64_t(ks * 2), b_desc + uint64_t(ks * 2), (kh | ks) ? 1u : 0u);
          }
          umma_commit_pair(&acc_full[as]);  // accumulator ready in both CTAs
        }
        umma_commit_pair(&b_empty[s]);      // ring slot of both CTAs may be refilled once these MMAs have read it
      }
#if TOD_K1_STATS
      atomicAdd(&g_k1_stats[2], st_a);
      atomicAdd(&g_k1_stats[3], st_b);
#endif
    }
  } else {
    // ===================================== epilogue: warps 2..9 (both CTAs) =====================================
    // Warp set j (warps 2 + 4j .. 5 + 4j, one warp per TMEM lane quarter) serves query tile j, whose accumulators
    // always land in TMEM stage j.  Thread = one query row, all 256 columns of the tile.
    const int quarter = warp & 3;                         // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const int j = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t my_list = ptx::smem_u32(lists + size_t(j * TOD_MAX_K) * kBlockM + row_in_tile);
    for (int i = 0; i < K; ++i) sts_u32(my_list + uint32_t(i) * 4u * kBlockM, kKeyEmpty);
    const int qi = q_row0 + j * kBlockM + row_in_tile;
    const int popx = qi < nq ? int(__ldg(popq + qi)) : 0;  // popcount of this thread's query: distance = popx - dot
    int thr_dot = popx - int(thr_init);                   // distance < thr  <=>  dot > popx - thr
    uint32_t *const my_gthr = (qi < nq && !(debug_mode & 16)) ? gthr + qi : nullptr;
    uint32_t g_next = 511u;                               // shared bound, loaded one tile ahead of its use
    const uint32_t acc_empty_leader = mapa_u32(ptx::smem_u32(&acc_empty[j]), 0);
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(j * kBlockN);

    // max over 32 accumulator columns held as 16 packed int16 pairs (VIMNMX3.S16x2 tree)
    auto max32p = [](const uint32_t (&v)[32], int o) -> int {
      const uint32_t r0 = __vimax3_s16x2(v[o + 0], v[o + 1], v[o + 2]), r1 = __vimax3_s16x2(v[o + 3], v[o + 4], v[o + 5]),
                     r2 = __vimax3_s16x2(v[o + 6], v[o + 7], v[o + 8]), r3 = __vimax3_s16x2(v[o + 9], v[o + 10], v[o + 11]),
                     r4 = __vimax3_s16x2(v[o + 12], v[o + 13], v[o + 14]);
      const uint32_t m = __vmaxs2(__vimax3_s16x2(r0, r1, r2), __vimax3_s16x2(r3, r4, v[o + 15]));
      return max(int(m << 16) >> 16, int(m) >> 16);
    };
    auto slow32 = [&](const uint32_t (&v)[32], int o, uint32_t grow, int n_valid) {
      thr_dot = k1_mma_slow_scan(v[o + 0], v[o + 1], v[o + 2], v[o + 3], v[o + 4], v[o + 5], v[o + 6], v[o + 7],
                                 v[o + 8], v[o + 9], v[o + 10], v[o + 11], v[o + 12], v[o + 13], v[o + 14], v[o + 15],
                                 my_list, K, thr_init, grow, n_valid, thr_dot, my_gthr, popx);
    };

    for (int t = 0; t < n_tiles; ++t) {
      const int cols_valid = min(kBlockN, (row1 - row0) - t * kBlockN);  // real db rows in this tile
      const uint32_t grow0 = global_row_base + uint32_t(row0 + t * kBlockN);
      {
        STAT_T0(t0);
        mbar_wait_bounded(&acc_full[j], t & 1);
        STAT_ADD(st_c, t0);
      }
      STAT_T0(t_busy);
      tc_fence_after();
      // the whole 128 x 256 accumulator as packed int16: four loads in flight, one wait, then the stage goes straight
      // back to the MMA warp — everything below works from registers
      uint32_t va[32], vb[32], vc[32], vd[32];
      tmem_ld64_pack16(taddr, va);
      tmem_ld64_pack16(taddr + 64u, vb);
      tmem_ld64_pack16(taddr + 128u, vc);
      tmem_ld64_pack16(taddr + 192u, vd);
      // shared bound of this query: use the value loaded during the previous tile, start the next load now
      thr_dot = max(thr_dot, popx - int(g_next) - 1);    // distance <= g  <=>  dot > popx - g - 1
      if (my_gthr) g_next = *reinterpret_cast<volatile uint32_t *>(my_gthr);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
      STAT_ADD(st_e, t_busy);

      uint32_t flags = 0;
      if (!(debug_mode & 4)) {
        const int m0 = max32p(va, 0), m1 = max32p(va, 16), m2 = max32p(vb, 0), m3 = max32p(vb, 16);
        const int m4 = max32p(vc, 0), m5 = max32p(vc, 16), m6 = max32p(vd, 0), m7 = max32p(vd, 16);
        if (t == 0 && cols_valid == kBlockN) {
          // Cold list: without a bound every one of the 256 columns is a candidate and the slow path walks them all
          // (24 000 cycles per CTA measured, tools/k1_stats.py).  The K-th largest of the eight sub-group maxima is
          // reached by K distinct rows of this tile, so no row below it can be among the K nearest: start from there.
          int a0 = m0, a1 = m1, a2 = m2, a3 = m3, a4 = m4, a5 = m5, a6 = m6, a7 = m7;
#define TOD_CSWAP(x, y) { const int lo_ = min(x, y); x = max(x, y); y = lo_; }   // descending
          TOD_CSWAP(a0, a1) TOD_CSWAP(a2, a3) TOD_CSWAP(a4, a5) TOD_CSWAP(a6, a7)
          TOD_CSWAP(a0, a2) TOD_CSWAP(a1, a3) TOD_CSWAP(a4, a6) TOD_CSWAP(a5, a7)
          TOD_CSWAP(a1, a2) TOD_CSWAP(a5, a6) TOD_CSWAP(a0, a4) TOD_CSWAP(a3, a7)
          TOD_CSWAP(a1, a5) TOD_CSWAP(a2, a6)
          TOD_CSWAP(a1, a4) TOD_CSWAP(a3, a6)
          TOD_CSWAP(a2, a4) TOD_CSWAP(a3, a5)
          TOD_CSWAP(a3, a4)
#undef TOD_CSWAP
          const int kth = K == 1 ? a0 : K == 2 ? a1 : K == 3 ? a2 : K == 4 ? a3 : K == 5 ? a4 : K == 6 ? a5 : K == 7 ? a6 : a7;
          thr_dot = max(thr_dot, kth - 1);               // keep dot >= kth
        }
        flags |= (m0 > thr_dot) ? 1u : 0u;
        flags |= (m1 > thr_dot) ? 2u : 0u;
        flags |= (m2 > thr_dot) ? 4u : 0u;
        flags |= (m3 > thr_dot) ? 8u : 0u;
        flags |= (m4 > thr_dot) ? 16u : 0u;
        flags |= (m5 > thr_dot) ? 32u : 0u;
        flags |= (m6 > thr_dot) ? 64u : 0u;
        flags |= (m7 > thr_dot) ? 128u : 0u;
        if (cols_valid < kBlockN) flags &= (1u << ((cols_valid + 31) >> 5)) - 1u;
      } else {
        thr_dot = max(thr_dot, int(va[0] ^ vb[7] ^ vc[3] ^ vd[9]) == 0x7fffffff ? 1 : 0);
      }
      flags = __reduce_or_sync(0xffffffffu, flags);
      if (debug_mode & 64) flags |= 1u;  // ablation: a (mostly empty) slow call on every tile, i.e. from a warm i-cache
      if (flags && !(debug_mode & 8)) {
        STAT_T0(t0);
        if (flags & 1u) slow32(va, 0, grow0, cols_valid);
        if (flags & 2u) slow32(va, 16, grow0 + 32u, cols_valid - 32);
        if (flags & 4u) slow32(vb, 0, grow0 + 64u, cols_valid - 64);
        if (flags & 8u) slow32(vb, 16, grow0 + 96u, cols_valid - 96);
        if (flags & 16u) slow32(vc, 0, grow0 + 128u, cols_valid - 128);
        if (flags & 32u) slow32(vc, 16, grow0 + 160u, cols_valid - 160);
        if (flags & 64u) slow32(vd, 0, grow0 + 192u, cols_valid - 192);
        if (flags & 128u) slow32(vd, 16, grow0 + 224u, cols_valid - 224);
        STAT_ADD(st_b, t0);
#if TOD_K1_STATS
        st_a += (unsigned long long)__popc(flags);
        if (lane == 0) {
          atomicAdd(&g_k1_tile[1][min(t, 63)], (unsigned long long)__popc(flags));
          atomicAdd(&g_k1_tile[2][min(t, 63)], (unsigned long long)(clock64() - t0));
        }
#endif
      }
      STAT_ADD(st_d, t_busy);
    }
#if TOD_K1_STATS
    if (lane == 0) {
      atomicAdd(&g_k1_stats[0], st_a);
      atomicAdd(&g_k1_stats[1], st_b);
      atomicAdd(&g_k1_stats[4], st_c);
      atomicAdd(&g_k1_stats[7], st_d);
      atomicAdd(&g_k1_stats[6], (unsigned long long)(n_tiles) * 8);  // 32-column sub-groups
      atomicAdd(&g_k1_stats[14], st_e);
    }
#endif
    if (qi < nq) {
      uint32_t *o = partial + (size_t(chunk) * nq + qi) * K;  // one candidate list per (chunk, query): a merge source
      for (int i = 0; i < K; ++i) o[i] = lds_u32(my_list + uint32_t(i) * 4u * kBlockM);
      if (full_sweep)  // the only list of this query: the other merge sources are empty
        for (int c = 1; c < n_sources; ++c)
          for (int i = 0; i < K; ++i) partial[(size_t(c) * nq + qi) * K + i] = kKeyEmpty;
    }
  }

#if TOD_K1_STATS
  if (threadIdx.x == 0) {
    atomicAdd(&g_k1_stats[5], (unsigned long long)(clock64() - t_cta));
    atomicAdd(&g_k1_stats[8], 1ull);
  }
#endif
  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal it or read its accumulators
  if (warp == 1) {
    tc_fence_after();
#if TOD_MMA_PAIR
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
#else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
#endif
  }
}

// database bits -> 0/1 int8:  out[row][8 * byte + b] = (in[row][byte] >> b) & 1
__global__ void __launch_bounds__(256) expand_db01_kernel(const uint8_t *__restrict__ in, uint2 *__restrict__ out,
                                                          size_t n_bytes) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_bytes) return;
  const uint32_t x = in[i];
  const uint32_t lo = ((x & 0xFu) * 0x00204081u) & 0x01010101u;
  const uint32_t hi = ((x >> 4) * 0x00204081u) & 0x01010101u;
  out[i] = make_uint2(lo, hi);
}

// query bits -> +-1 int8:  out[row][8 * byte + b] = (in[row][byte] >> b) & 1 ? +1 : -1.  A warp is one descriptor
// (32 bytes): it also writes the descriptor's popcount and resets the query's shared bound to "none" (511).
__global__ void __launch_bounds__(256) expand_query_kernel(const uint8_t *__restrict__ in, uint2 *__restrict__ out,
                                                           size_t n_bytes, uint32_t *__restrict__ popq,
                                                           uint32_t *__restrict__ gthr) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t x = i < n_bytes ? in[i] : 0u;
  const uint32_t pop = __reduce_add_sync(0xffffffffu, unsigned(__popc(x)));
  if (i >= n_bytes) return;
  const uint32_t lo = (((x & 0xFu) * 0x00204081u) & 0x01010101u) ^ 0x01010101u;   // 1 where the bit is 0
  const uint32_t hi = (((x >> 4) * 0x00204081u) & 0x01010101u) ^ 0x01010101u;
  out[i] = make_uint2(lo * 0xFEu + 0x01010101u, hi * 0xFEu + 0x01010101u);        // bit 0 -> 0xFF, bit 1 -> 0x01
  if ((threadIdx.x & 31) == 0) {
    popq[i >> 5] = pop;
    gthr[i >> 5] = 511u;
  }
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

}  // namespace

K1Plan k1_mma_plan(int nq, int64_t shard_rows, int sm_count) {
  K1Plan p{};
  p.q_per_thread = 0;
  p.q_tile = kQT * kBlockM;
  p.n_qtiles = std::max(1, (nq + p.q_tile - 1) / p.q_tile);
  if (kCtas == 2) p.n_qtiles = (p.n_qtiles + 1) & ~1;  // pairs: an odd last group gets an idle partner (queries OOB)
  const int64_t max_chunks = std::max<int64_t>(1, (shard_rows + kBlockN - 1) / kBlockN);
  // One CTA per SM, so the launch takes about waves(c) x (time of one CTA).  A CTA costs its db tiles plus a fixed
  // start-up: pipeline fill and, mostly, the cold top-k lists of its first tiles — ~28 tile periods measured with
  // tools/k1_stats.py (profiles/r1_k1_stats_*.jsonl).  More chunks fill the last wave better but pay the start-up
  // more often (and add merge sources); pick the chunk count that minimises the modelled time.
  constexpr double kStartupTiles = 28.0;
  const int64_t tiles_total = max_chunks;
  const int64_t c_hi = std::min<int64_t>(max_chunks, 64);
  // Whole rounds of the machine run unsplit (one start-up per round); only the left-over query groups are chunked.
  // full = 0 is the plain chunked plan; both are modelled and the cheaper one wins.
  const int64_t slots = std::max(kCtas, sm_count / kCtas * kCtas);
  int64_t best_c = 1, best_full = 0;
  double best_t = 1e300;
#ifndef TOD_K1_TWO_PHASE
#define TOD_K1_TWO_PHASE 1            // 0: variant build with the round-1 plan (every query group chunked), for A/B runs
#endif
  for (int64_t full : {int64_t(0), TOD_K1_TWO_PHASE ? int64_t(p.n_qtiles) / slots * slots : int64_t(0)}) {
    const int64_t rest = p.n_qtiles - full;
    const double t_full = double(full / slots) * (kStartupTiles + double(tiles_total));
    if (rest == 0) {
      if (t_full < best_t * (1.0 - 1e-9)) best_t = t_full, best_c = 1, best_full = full;
      continue;
    }
    for (int64_t c = 1; c <= std::max<int64_t>(1, c_hi); ++c) {
      const int64_t waves = (c * rest + sm_count - 1) / sm_count;
      const double t = t_full + double(waves) * (kStartupTiles + double((tiles_total + c - 1) / c));
      if (t < best_t * (1.0 - 1e-9)) best_t = t, best_c = c, best_full = full;
    }
  }
#if TOD_K1_STATS
  if (const char *e = getenv("TOD_K1_CHUNKS")) best_c = std::min<int64_t>(max_chunks, std::max(1, atoi(e)));  // experiments
  if (const char *e = getenv("TOD_K1_FULL")) best_full = std::min<int64_t>(p.n_qtiles, std::max(0, atoi(e)) / kCtas * kCtas);
#endif
  p.full_qtiles = int(best_full);
  int64_t rpc = (shard_rows + best_c - 1) / best_c;
  rpc = std::max<int64_t>(kBlockN, (rpc + kBlockN - 1) / kBlockN * kBlockN);
  p.rows_per_chunk = int(rpc);
  p.n_chunks = int(std::max<int64_t>(1, (shard_rows + rpc - 1) / rpc));
  if (p.full_qtiles == p.n_qtiles) p.n_chunks = 1;
  p.n_sources = p.n_chunks;
  return p;
}

cudaError_t launch_expand_db(const void *d_bits, void *d_int8, int64_t rows, cudaStream_t stream) {
  const size_t n_bytes = size_t(rows) * 32;
  if (n_bytes == 0) return cudaSuccess;
  const unsigned blocks = unsigned((n_bytes + 255) / 256);
  expand_db01_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint8_t *>(d_bits), static_cast<uint2 *>(d_int8),
                                                 n_bytes);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_expand_queries(const void *d_bits, void *d_int8, int64_t rows, uint32_t *d_popq, uint32_t *d_gthr,
                                  cudaStream_t stream) {
  const size_t n_bytes = size_t(rows) * 32;
  if (n_bytes == 0) return cudaSuccess;
  const unsigned blocks = unsigned((n_bytes + 255) / 256);
  expand_query_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint8_t *>(d_bits), static_cast<uint2 *>(d_int8),
                                                  n_bytes, d_popq, d_gthr);
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

#if TOD_K1_STATS
}  // namespace tod
extern "C" void tod_debug_k1_stats(unsigned long long *out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, tod::g_k1_stats, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(tod::g_k1_stats, z, sizeof(z));
  }
}
extern "C" void tod_debug_k1_tile_stats(unsigned long long *out256, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out256, tod::g_k1_tile, sizeof(unsigned long long) * 256);
  if (reset) {
    static unsigned long long z[256] = {0};
    cudaMemcpyToSymbol(tod::g_k1_tile, z, sizeof(z));
  }
}
namespace tod {
#endif

int k1_mma_query_box_rows() { return kBlockM; }
int k1_mma_db_box_rows() { return kHalfN; }
size_t tensor_map_bytes() { return sizeof(CUtensorMap); }

cudaError_t launch_k1_mma(const K1Plan &plan, const void *map_q, const void *map_db, int nq, int64_t shard_rows,
                          uint32_t global_row_base, int k, uint32_t radius, uint32_t *d_partial, uint32_t *d_gthr,
                          const uint32_t *d_popq, cudaStream_t stream, int64_t seg_row0) {
  if (k < 1 || k > TOD_MAX_K) return cudaErrorInvalidValue;
  const uint32_t thr_init = radius ? min(radius + 1u, 511u) : 511u;
  {  // per device and per context, so not cached in a static: a process may hold handles on several GPUs
    cudaError_t e = cudaFuncSetAttribute(k1_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMma);
    if (e != cudaSuccess) return e;
  }
#if TOD_K1_STATS
  static const int debug_mode = [] {  // TOD_K1_DEBUG_MODE: ablation knob of the instrumented build only
    const char *e = getenv("TOD_K1_DEBUG_MODE");
    return e ? atoi(e) : 0;
  }();
#endif
  // n_qtiles and full_qtiles are even in pair mode (cluster (2,1,1))
  dim3 grid(unsigned(plan.full_qtiles) + unsigned(plan.n_qtiles - plan.full_qtiles) * unsigned(plan.n_chunks));
  k1_mma_kernel<<<grid, kThreadsMma, kSmemMma, stream>>>(*static_cast<const CUtensorMap *>(map_q),
                                                         *static_cast<const CUtensorMap *>(map_db), nq, int(shard_rows),
                                                         int(seg_row0), global_row_base, plan.rows_per_chunk, plan.n_chunks,
                                                         plan.full_qtiles / kCtas, plan.n_sources, thr_init, k, d_partial,
                                                         d_gthr, d_popq
#if TOD_K1_STATS
                                                         , debug_mode
#endif
                                                         );
  count_launch();
  return cudaGetLastError();
}

}  // namespace tod
