// K4: the clique gate's exact pre-checks on the GPU, batched over hypotheses.
//
// Replaces, for the hypotheses that can still win a RANSAC round, the cheap part of selectWithinDistance
// (src/common/sac_model_registration_graph.h:203-265 of the reference): the degree filter (:209-218), the
// neighbourhood test (:222-238), and two exact proofs that the bounded clique search (:258-265 ->
// maximum_clique.cpp:286-369) cannot return a clique of more than 7 vertices — whatever its visiting order, early stop
// and step budget — because the induced sample sub-graph contains NO clique of 8:
//   * its 7-core (iterated removal of vertices with fewer than 7 live neighbours) has fewer than 8 vertices;
//   * a greedy colouring of the core needs fewer than 8 colours;
//   * (cores of at most 128 vertices) an exhaustive depth-first search over 128-bit candidate sets finds no 8-clique.
// Hypotheses whose sub-graph DOES hold an 8-clique — where the reference's exact stepping decides between "found 8" and
// "stopped at an exact 7"  — are settled by K5 (k5_search_kernel below) when the filtered graph
// has at most 128 vertices: the reference's bounded search itself, stepped exactly (clique_small.h, one source for the
// host and the device), one thread per hypothesis on the packed induced sub-graph K4 leaves in a job queue.  Larger
// graphs, and searches that exceed K5's step cap, go back to the host search.  The gate can only keep or zero a
// count, so "fails" / "passes" is all the host replay needs.
//
// One warp per hypothesis.  Bit-rows of the sample graph are read straight from K2's output; masks live in shared
// memory (W words per warp).  Reference-faithful (+inf threshold) mode only: the candidate set is
// P[s0] & P[s1] & P[s2] & valid & finite, plus the three samples.
#include "clique_small.h"
#include "k5_warp.cuh"
#include "tod_internal.h"

namespace tod {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kSmallCore = 128;   // exhaustive search limit (two 64-bit words per set)
constexpr int kProofMax = tod::kGateProofMax;  // larger filtered graphs skip the proofs (host search)
#ifndef TOD_K4_SEARCH_EARLY
#define TOD_K4_SEARCH_EARLY 1      // n > 0: graphs K5 takes leave the proofs after n peeling sweeps (0: full proofs).
                                   // Measured at C5: K4 + K5 63 -> 57 ms of device time with 1 (K5 is bound by its slowest
                                   // search, not by the number of searches: 933 k jobs take the 24 ms that 768 k took)
#endif
constexpr int kSearchMax = 256;   // largest filtered graph K5 takes (four 64-bit words per row)
constexpr int kMaxSweeps = 64;
constexpr int kDfsBudget = 6000;  // node expansions per lane before giving the hypothesis back to the host
constexpr int kK5Threads = 64;
constexpr int kK5StepCap = 512;   // searches of the gate take tens of steps; longer ones are left to the host

__global__ void __launch_bounds__(256)
sample_degree_mask_kernel(const K3Cluster *__restrict__ clusters, const int32_t *__restrict__ active,
                          const uint32_t *__restrict__ sample, const uint32_t *__restrict__ valid,
                          uint32_t *__restrict__ deg_mask, int min_degree) {
  const K3Cluster cl = clusters[active[blockIdx.y]];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t *S = sample + cl.matrix_offset;
  const uint32_t *V = valid + cl.valid_offset;
  uint32_t *D = deg_mask + cl.valid_offset;
  // a warp owns 32 consecutive rows = one word of the mask: no atomics
  for (int w0 = blockIdx.x * kWarpsPerCta + warp; w0 < cl.W; w0 += gridDim.x * kWarpsPerCta) {
    const uint32_t vw = __ldg(V + w0);
    uint32_t out = 0;
    uint32_t m = vw;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t *row = S + size_t(w0 * 32 + b) * cl.W;
      int d = 0;
      for (int w = lane; w < cl.W; w += 32) d += __popc(__ldg(row + w) & __ldg(V + w));
      d = __reduce_add_sync(0xffffffffu, d);
      if (d >= min_degree) out |= 1u << b;
    }
    if (lane == 0) D[w0] = out;
  }
}

struct U128 {
  unsigned long long lo, hi;
};
__device__ __forceinline__ int popc128(U128 a) { return __popcll(a.lo) + __popcll(a.hi); }

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k4_gate_kernel(const K3Cluster *__restrict__ clusters, const uint32_t *__restrict__ physical,
               const uint32_t *__restrict__ sample, const uint32_t *__restrict__ valid,
               const uint32_t *__restrict__ finite, const uint32_t *__restrict__ deg_mask, int n_hyp,
               const uint4 *__restrict__ hyps, const int32_t *__restrict__ counts, const int32_t *__restrict__ floor_,
               int max_words, uint8_t *__restrict__ verdict, int4 *__restrict__ job_hdr,
               unsigned long long *__restrict__ job_pool, unsigned long long *__restrict__ job_ctl,
               unsigned long long pool_words) {
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = blockIdx.x * kWarpsPerCta + warp;
  if (h >= n_hyp) return;
  // per warp: adj[128] (U128) | alive[max_words] | work[max_words] | filt[max_words] | ids[256] (u16)
  const int per_warp = kSmallCore * 4 + 3 * max_words + kSearchMax / 2;
  U128 *adj = reinterpret_cast<U128 *>(smem + size_t(warp) * per_warp);
  uint32_t *alive = smem + size_t(warp) * per_warp + kSmallCore * 4;
  uint32_t *work = alive + max_words;
  uint32_t *filt = work + max_words;   // the filtered set before the core peeling: the graph the reference searches
  uint16_t *ids = reinterpret_cast<uint16_t *>(filt + max_words);

  const int cnt = __ldg(counts + h);
  const uint4 hy = __ldg(hyps + h);
  auto finish = [&](int v) {
    if (lane == 0) verdict[h] = uint8_t(v);
  };
  if (cnt <= 7 || cnt <= __ldg(floor_ + hy.w)) return finish(kGateNotEvaluated);
  const K3Cluster cl = clusters[hy.w];
  const int W = cl.W;
  if (W > max_words) return finish(kGateNeedsHost);
  const uint32_t *P = physical + cl.matrix_offset;
  const uint32_t *S = sample + cl.matrix_offset;
  const uint32_t *V = valid + cl.valid_offset;
  const uint32_t *F = finite + cl.valid_offset;
  const uint32_t *D = deg_mask + cl.valid_offset;

  // ---- filtered = (common valid physical neighbours of the samples + the samples) with sample-degree >= 7 ----------
  const uint32_t *r0 = P + size_t(hy.x) * W, *r1 = P + size_t(hy.y) * W, *r2 = P + size_t(hy.z) * W;
  int nf = 0;
  for (int w = lane; w < W; w += 32) {
    uint32_t m = __ldg(r0 + w) & __ldg(r1 + w) & __ldg(r2 + w) & __ldg(V + w) & __ldg(F + w);
    if (int(hy.x >> 5) == w) m |= 1u << (hy.x & 31);  // the samples pass the +inf test themselves (count > 7 implies
    if (int(hy.y >> 5) == w) m |= 1u << (hy.y & 31);  // finite samples and a finite fit)
    if (int(hy.z >> 5) == w) m |= 1u << (hy.z & 31);
    m &= __ldg(D + w);
    alive[w] = m;
    filt[w] = m;
    nf += __popc(m);
  }
  nf = __reduce_add_sync(0xffffffffu, nf);
  if (nf <= 7) return finish(kGateFails);  // :214-218
  // A filtered graph of thousands of vertices (one object filling the frame) is too large for K5 and its proofs cost
  // milliseconds per warp (2 ms of C1's 4.8 ms frame) while they rarely succeed there: straight to the host search.
  if (nf > kProofMax) return finish(kGateNeedsHost);
  __syncwarp();

  // ---- neighbourhood test (:222-238) and 7-core: Jacobi sweeps of "degree inside the live set" ------------------------
  int n_alive = nf;
  bool search_early = false;
  for (int sweep = 0;; ++sweep) {
    if (sweep >= kMaxSweeps) return finish(kGateNeedsHost);
    int max_d = 0, removed = 0;
    for (int w = lane; w < W; w += 32) work[w] = 0u;  // vertices to drop after this sweep
    __syncwarp();
    // The live vertices are visited kBatch at a time: their rows are loaded together (kBatch independent loads in
    // flight per lane and step) before the popcounts are reduced — the sweep is latency-bound on row loads otherwise.
    constexpr int kBatch = 8;
    int w0 = 0;
    uint32_t m = alive[0];
    for (;;) {
      int vid[kBatch];
      int nb = 0;
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        while (m == 0 && w0 + 1 < W) m = alive[++w0];
        if (m) {
          vid[j] = w0 * 32 + __ffs(m) - 1;
          m &= m - 1;
          nb = j + 1;
        } else {
          vid[j] = -1;
        }
      }
      if (nb == 0) break;
      int d[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) d[j] = 0;
      for (int w = lane; w < W; w += 32) {
        const uint32_t a = alive[w];
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
          if (vid[j] >= 0) d[j] += __popc(__ldg(S + size_t(vid[j]) * W + w) & a);
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        if (vid[j] >= 0) {
          const int dj = __reduce_add_sync(0xffffffffu, d[j]);
          max_d = max(max_d, dj);
          if (dj < 7) {
            ++removed;
            if (lane == 0) work[vid[j] >> 5] |= 1u << (vid[j] & 31);
          }
        }
      }
      if (nb < kBatch) break;
    }
    // the reference scans for ONE filtered vertex with more than 7 sample-neighbours inside filtered
    if (sweep == 0 && max_d <= 7) return finish(kGateFails);
#if TOD_K4_SEARCH_EARLY
    // graphs K5 takes anyway: one peeling sweep (it settles the neighbourhood test and most of the small cores), then
    // straight to the exact search instead of iterating to the 7-core and colouring it
    if (job_hdr != nullptr && nf <= kSearchMax && sweep >= TOD_K4_SEARCH_EARLY - 1 && removed != 0) {
      __syncwarp();
      for (int w = lane; w < W; w += 32) alive[w] &= ~work[w];
      n_alive -= removed;
      __syncwarp();
      if (n_alive < 8) return finish(kGateFails);
      search_early = true;
      break;
    }
#endif
    if (removed == 0) break;
    __syncwarp();
    for (int w = lane; w < W; w += 32) alive[w] &= ~work[w];
    n_alive -= removed;
    __syncwarp();
    if (n_alive < 8) return finish(kGateFails);  // an 8-clique lives inside the 7-core
  }

  // ---- greedy colouring of the core: fewer than 8 colours bound the clique number below 8 ---------------------------
  __syncwarp();
  for (int w = lane; w < W; w += 32) work[w] = alive[w];  // uncoloured
  __syncwarp();
  int colours = 0;
  bool undecided = search_early;
  for (int left = search_early ? 0 : n_alive; left > 0;) {
    if (++colours >= 8) {
      undecided = true;
      break;
    }
    // candidates of this colour class: uncoloured vertices without a neighbour in the class so far (<= 4 words/lane)
    uint32_t q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = (lane + 32 * i < W) ? work[lane + 32 * i] : 0u;
    for (;;) {
      int mine = 0x7fffffff;
#pragma unroll
      for (int i = 3; i >= 0; --i)
        if (q[i]) mine = (lane + 32 * i) * 32 + __ffs(q[i]) - 1;
      const int v = __reduce_min_sync(0xffffffffu, mine);
      if (v == 0x7fffffff) break;
      --left;
      const uint32_t *row = S + size_t(v) * W;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int w = lane + 32 * i;
        if (w < W) {
          if (w == (v >> 5)) {
            q[i] &= ~(1u << (v & 31));
            work[w] &= ~(1u << (v & 31));
          }
          q[i] &= ~__ldg(row + w);
        }
      }
    }
    __syncwarp();
  }
  if (!undecided) return finish(kGateFails);
  // K5 job: the filtered graph (NOT its core — the reference's search order depends on every vertex) when it is small
  const bool to_search = job_hdr != nullptr && nf <= kSearchMax;
  if (!to_search && n_alive > kSmallCore) return finish(kGateNeedsHost);
  const uint32_t *members = to_search ? filt : alive;
  const int nn = to_search ? nf : n_alive;

  // ---- the induced sub-graph's vertices: vertex a = a-th smallest member (the renumbering of :241-255) -------------
  __syncwarp();
  if (lane == 0) {
    int k = 0;
    for (int w = 0; w < W; ++w) {
      uint32_t m = members[w];
      while (m) {
        ids[k++] = uint16_t(w * 32 + __ffs(m) - 1);
        m &= m - 1;
      }
    }
  }
  __syncwarp();
  if (to_search) {
    // queue: rows of 1, 2 or 4 words (<= 64, 128, 256 vertices), bump-allocated in the pool; the headers of each width
    // are kept together so that the warps of K5 run one instantiation of the search each
    const int nw = nn <= 64 ? 1 : (nn <= 128 ? 2 : 4);
    const unsigned long long words = (unsigned long long)nw * (unsigned long long)nn;
    unsigned long long slot = 0, off = 0;
    if (lane == 0) {
      slot = atomicAdd(job_ctl + (nw == 1 ? 0 : (nw == 2 ? 1 : 3)), 1ull);
      off = atomicAdd(job_ctl + 2, words);
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    off = __shfl_sync(0xffffffffu, off, 0);
    const bool fits = off + words <= pool_words;
    if (fits) {
      unsigned long long *dst = job_pool + off;
      for (int i = lane; i < nn; i += 32) {
        const uint32_t *row = S + size_t(ids[i]) * W;
        unsigned long long a[4] = {0ull, 0ull, 0ull, 0ull};
        for (int j = 0; j < nn; ++j) {
          const uint32_t u = ids[j];
          const unsigned long long bit = (__ldg(row + (u >> 5)) >> (u & 31)) & 1u;
          a[j >> 6] |= bit << (j & 63);
        }
        for (int x = 0; x < nw; ++x) dst[size_t(i) * nw + x] = a[x];
      }
    }
    // header slots: [0, n1) one-word jobs | [n_hyp - n2, n_hyp) two-word jobs | four-word jobs in a second array of
    // n_hyp slots behind the first (a hypothesis takes at most one slot, so none of them can overflow)
    if (lane == 0) {
      const unsigned long long at = nw == 1 ? slot : (nw == 2 ? (unsigned long long)(n_hyp - 1) - slot
                                                              : (unsigned long long)n_hyp + slot);
      job_hdr[at] = make_int4(h, fits ? nn : 0, int(off & 0xffffffffull), int(off >> 32));
    }
    return finish(kGateNeedsHost);  // K5 overwrites the verdict unless the pool was full or its step cap is reached
  }

  // ---- pack the core into shared memory ------------------------------------------------------------------------------
  for (int i = lane; i < nn; i += 32) {
    const uint32_t *row = S + size_t(ids[i]) * W;
    U128 a{0ull, 0ull};
    for (int j = 0; j < nn; ++j) {
      const uint32_t u = ids[j];
      const unsigned long long bit = (__ldg(row + (u >> 5)) >> (u & 31)) & 1u;
      if (j < 64) a.lo |= bit << j;
      else a.hi |= bit << (j - 64);
    }
    adj[i] = a;
  }
  __syncwarp();

  // ---- small core of a larger graph: exhaustive search for an 8-clique over 128-bit sets -----------------------------
  bool found = false, overflow = false;
  for (int root = lane; root < n_alive && !found && !overflow; root += 32) {
    // cliques are enumerated with increasing core indices: candidates of level 1 = neighbours of root above root
    U128 st[8];
    U128 above{~0ull, ~0ull};
    if (root < 63) above.lo = ~0ull << (root + 1);
    else {
      above.lo = 0ull;
      above.hi = root >= 127 ? 0ull : ~0ull << (root - 63);
    }
    st[1].lo = adj[root].lo & above.lo;
    st[1].hi = adj[root].hi & above.hi;
    int size = 1, steps = 0;
    while (size >= 1) {
      U128 p = st[size];
      if (popc128(p) < 8 - size) {
        --size;
        continue;
      }
      int v;
      if (p.lo) {
        v = __ffsll(p.lo) - 1;
        p.lo &= p.lo - 1;
      } else {
        v = 64 + __ffsll(p.hi) - 1;
        p.hi &= p.hi - 1;
      }
      st[size] = p;  // v consumed at this level; what is left of p lies above v
      if (size + 1 == 8) {
        found = true;
        break;
      }
      U128 nx;
      nx.lo = p.lo & adj[v].lo;
      nx.hi = p.hi & adj[v].hi;
      if (popc128(nx) >= 8 - (size + 1)) {
        ++size;
        st[size] = nx;
      }
      if (++steps > kDfsBudget) {
        overflow = true;
        break;
      }
    }
  }
  const bool any_found = __any_sync(0xffffffffu, found);
  const bool any_overflow = __any_sync(0xffffffffu, overflow);
  finish((any_found || any_overflow) ? kGateNeedsHost : kGateFails);
}

// K5, thread form: the reference's bounded clique search, stepped exactly, one THREAD per queued hypothesis whose
// graph has at most 64 vertices (one-word rows) — the bulk of the queue, tens of steps each.
__global__ void __launch_bounds__(kK5Threads)
k5_search_kernel(const int4 *__restrict__ job_hdr, const unsigned long long *__restrict__ job_pool,
                 const unsigned long long *__restrict__ job_ctl, int n_hyp, int step_cap,
                 uint8_t *__restrict__ verdict) {
  const long long j = (long long)blockIdx.x * kK5Threads + threadIdx.x;  // header slot
  if (j >= (long long)job_ctl[0]) return;
  const int4 hd = job_hdr[j];
  if (hd.y <= 0) return;  // the pool was full: the verdict stays "host"
  const unsigned long long *rows = job_pool + ((unsigned long long)(unsigned)hd.z | ((unsigned long long)(unsigned)hd.w << 32));
  const int r = small_gate_search(reinterpret_cast<const Bits64 *>(rows), hd.y, step_cap, nullptr);
  if (r == 1) verdict[hd.x] = uint8_t(kGatePasses);
  else if (r == 0) verdict[hd.x] = uint8_t(kGateFailsSearch);
}

// K5, warp form (k5_warp.cuh): one WARP per queued hypothesis of 65..256 vertices.  A fixed grid of warps pulls jobs
// from a counter, widest graphs first (they are the slowest).
constexpr int kK5WarpsPerCta = 4;
constexpr int kK5WarpCtas = 296;   // two CTAs per SM
constexpr size_t kK5WarpStateBytes = (sizeof(k5w::WarpState<4>) + 15) & ~size_t(15);
__global__ void __launch_bounds__(kK5WarpsPerCta * 32)
k5_warp_kernel(const int4 *__restrict__ job_hdr, const unsigned long long *__restrict__ job_pool,
               unsigned long long *__restrict__ job_ctl, int n_hyp, int step_cap, uint8_t *__restrict__ verdict) {
  extern __shared__ __align__(16) unsigned char k5w_smem[];
  const int lane = threadIdx.x & 31;
  void *state = k5w_smem + size_t(threadIdx.x >> 5) * kK5WarpStateBytes;
  const unsigned long long n2 = job_ctl[1], n4 = job_ctl[3];
  for (;;) {
    unsigned long long idx = 0;
    if (lane == 0) idx = atomicAdd(job_ctl + 4, 1ull);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    long long slot;
    if (idx < n4) slot = (long long)n_hyp + (long long)idx;
    else if (idx - n4 < n2) slot = (long long)n_hyp - 1 - (long long)(idx - n4);
    else break;
    const int4 hd = job_hdr[slot];
    if (hd.y <= 0) continue;
    const unsigned long long *rows = job_pool + ((unsigned long long)(unsigned)hd.z | ((unsigned long long)(unsigned)hd.w << 32));
    const int r = hd.y <= 128
                      ? k5w::warp_gate_search<2>(*static_cast<k5w::WarpState<2> *>(state), rows, hd.y, step_cap, lane)
                      : k5w::warp_gate_search<4>(*static_cast<k5w::WarpState<4> *>(state), rows, hd.y, step_cap, lane);
    if (lane == 0) {
      if (r == 1) verdict[hd.x] = uint8_t(kGatePasses);
      else if (r == 0) verdict[hd.x] = uint8_t(kGateFailsSearch);
    }
    __syncwarp();
  }
}

// both forms of K5 over a filled job queue
cudaError_t launch_k5(int4 *job_hdr, unsigned long long *job_pool, unsigned long long *job_ctl, int n_hyp,
                      uint8_t *d_verdict, cudaStream_t stream) {
  const size_t smem = size_t(kK5WarpsPerCta) * kK5WarpStateBytes;
  cudaError_t e = cudaFuncSetAttribute(k5_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  // the warp form first: its jobs are the long ones, the thread form fills the machine around them
  k5_warp_kernel<<<kK5WarpCtas, kK5WarpsPerCta * 32, smem, stream>>>(job_hdr, job_pool, job_ctl, n_hyp, kK5StepCap,
                                                                    d_verdict);
  k5_search_kernel<<<(n_hyp + kK5Threads - 1) / kK5Threads, kK5Threads, 0, stream>>>(job_hdr, job_pool, job_ctl, n_hyp,
                                                                                    kK5StepCap, d_verdict);
  count_launch(2);
  return cudaGetLastError();
}

}  // namespace

// job queue layout: [u64 control: 1-word jobs, 2-word jobs, pool words used, 4-word jobs, warp-form job cursor | pad
// to 256 B]
// [2 * n_hyp headers] [pool]
size_t gate_job_bytes(int n_hyp, size_t pool_bytes) {
  return 256 + ((2 * size_t(n_hyp) * sizeof(int4) + 255) & ~size_t(255)) + pool_bytes;
}

// K5 alone on a prepared job queue (layout of gate_job_bytes): for the stage-level entry point tod_gate_search_device.
cudaError_t launch_gate_search_jobs(int n_hyp, void *d_jobs, uint8_t *d_verdict, cudaStream_t stream) {
  if (n_hyp <= 0) return cudaSuccess;
  unsigned long long *job_ctl = static_cast<unsigned long long *>(d_jobs);
  int4 *job_hdr = reinterpret_cast<int4 *>(static_cast<char *>(d_jobs) + 256);
  unsigned long long *job_pool = reinterpret_cast<unsigned long long *>(
      static_cast<char *>(d_jobs) + 256 + ((2 * size_t(n_hyp) * sizeof(int4) + 255) & ~size_t(255)));
  return launch_k5(job_hdr, job_pool, job_ctl, n_hyp, d_verdict, stream);
}

cudaError_t launch_sample_degree_mask(const void *d_clusters, const int32_t *d_active, int n_active, int max_n,
                                      const uint32_t *d_sample, const uint32_t *d_valid, uint32_t *d_deg_mask,
                                      int min_degree, cudaStream_t stream) {
  if (n_active <= 0 || max_n <= 0) return cudaSuccess;
  const int max_w = adjacency_row_words(max_n);
  const int bx = std::max(1, std::min(64, (max_w + kWarpsPerCta - 1) / kWarpsPerCta));
  for (int c0 = 0; c0 < n_active; c0 += 65535) {
    dim3 grid(bx, std::min(65535, n_active - c0));
    sample_degree_mask_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(static_cast<const K3Cluster *>(d_clusters),
                                                                      d_active + c0, d_sample, d_valid, d_deg_mask,
                                                                      min_degree);
    count_launch();
  }
  return cudaGetLastError();
}

cudaError_t launch_gate_prechecks(const void *d_clusters, const uint32_t *d_physical, const uint32_t *d_sample,
                                  const uint32_t *d_valid, const uint32_t *d_finite, const uint32_t *d_deg_mask,
                                  int n_hyp, const uint32_t *d_hyps, const int32_t *d_counts, const int32_t *d_floor,
                                  int max_words, uint8_t *d_verdict, void *d_jobs, size_t pool_bytes,
                                  cudaStream_t stream, cudaEvent_t ev_between) {
  if (n_hyp <= 0) return cudaSuccess;
  max_words = std::min(max_words, 128);  // clusters of more than 4096 correspondences go to the host search
  max_words = std::max(max_words, 4);
  const size_t smem = size_t(kWarpsPerCta) * (3 * size_t(max_words) + kSearchMax / 2 + kSmallCore * 4) * sizeof(uint32_t);
  cudaError_t e = cudaFuncSetAttribute(k4_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  unsigned long long *job_ctl = nullptr, *job_pool = nullptr;
  int4 *job_hdr = nullptr;
  if (d_jobs && pool_bytes >= 2048) {
    job_ctl = static_cast<unsigned long long *>(d_jobs);
    job_hdr = reinterpret_cast<int4 *>(static_cast<char *>(d_jobs) + 256);
    job_pool = reinterpret_cast<unsigned long long *>(static_cast<char *>(d_jobs) + 256 +
                                                      ((2 * size_t(n_hyp) * sizeof(int4) + 255) & ~size_t(255)));
    e = cudaMemsetAsync(job_ctl, 0, 8 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
  }
  const int blocks = (n_hyp + kWarpsPerCta - 1) / kWarpsPerCta;
  k4_gate_kernel<<<blocks, kWarpsPerCta * 32, smem, stream>>>(
      static_cast<const K3Cluster *>(d_clusters), d_physical, d_sample, d_valid, d_finite, d_deg_mask, n_hyp,
      reinterpret_cast<const uint4 *>(d_hyps), d_counts, d_floor, max_words, d_verdict, job_hdr, job_pool, job_ctl,
      (unsigned long long)(pool_bytes / 8));
  count_launch();
  if (ev_between) {
    e = cudaEventRecord(ev_between, stream);
    if (e != cudaSuccess) return e;
  }
  if (job_hdr) {
    e = launch_k5(job_hdr, job_pool, job_ctl, n_hyp, d_verdict, stream);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace tod
