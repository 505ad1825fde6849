// C-ABI of the GuessGenerator half of the hot path (include/tod_b200.h: tod_guess_*).
//
// Host driver mirroring tod::GuessGenerator::process (src/detection/GuessGenerator.cpp:127-250 of the reference) and
// tod::AdjacencyRansac (src/common/adjacency_ransac.cpp), restructured around two GPU kernels:
//
//   ClusterPerObject (adjacency_ransac.cpp:176-205)            host, one pass over the matches
//   FillAdjacency    (:127-172)                                K2, one launch for all (object) clusters of the frame
//   per RANSAC round, for all still-active objects together:
//     sampler        (sac_model_registration_graph.h:102-168)  host, on bit-rows, seeded stream (replaces libc rand())
//     model + inlier count (:171-200, :271-347)                K3, one launch for every hypothesis of every object
//     best / adaptive-k scan (ransac.h:95-135)                 host replay over the K3 counts, in hypothesis order
//     clique gate    (:203-268)                                host, only for hypotheses that beat the current best
//     refinement + pose inversion (adjacency_ransac.cpp:255-308)   host
//     InvalidateQueryIndices (:93-123) + cascade (:63-89)      host, as a valid-bit mask ANDed in at use
//
// Equivalences used: sorted neighbour lists <-> bit-rows; InvalidateCluster <-> AND with the valid
// mask; the gate can only keep or zero a count, so it is evaluated lazily; the early stop of computeModel is a prefix
// of the hypothesis list, so hypotheses are generated and scored in growing batches.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "clique.h"
#include "host_geometry.h"
#include "tod_internal.h"

using tod::DeviceBuffer;
using tod::fail;

namespace {

inline int popc32(uint32_t x) { return __builtin_popcount(x); }

struct Cluster {
  int frame = 0;
  int object = 0;
  int n = 0, W = 0;
  std::vector<float> q, t, px;     // n x 3, n x 3, n x 2
  std::vector<uint32_t> qidx;      // query_indices_ (non-decreasing)
  std::vector<uint32_t> valid;     // W words: valid_indices_ as a mask
  std::vector<uint32_t> deg7;      // W words: valid vertices with >= 7 valid sample-neighbours this round (gate :209-213);
                                   // computed by sample_degree_mask_kernel on the GPU, or by host_degree_mask
  std::vector<uint32_t> finite;    // W words: all six coordinates finite
  std::vector<uint32_t> vpre;      // sampler: prefix popcounts of `valid` per 64-bit word, rebuilt when a round starts
  int n_valid = 0;
  int64_t point_offset = 0, matrix_offset = 0, valid_offset = 0;
  const uint32_t *P = nullptr, *S = nullptr;  // host copies of the bit-matrices (n x W)
  bool active = true;
  unsigned round = 0;

  // --- per-round RANSAC replay state (pcl::RandomSampleConsensus::computeModel) ---
  uint64_t rng = 0;
  int iterations = 0;
  int n_best = -INT_MAX;
  double k = 1.0;
  bool stopped = false;
  std::vector<uint32_t> best_inliers;
  float best_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, best_T[3] = {0, 0, 0};  // R_, T_ of the best hypothesis (finite mode)
  std::vector<uint32_t> hyps;  // triples of the current batch
  std::vector<uint32_t> hyps_next;  // triples drawn ahead for the next batch while the GPU scores this one
  bool drawn_ahead = false;
  int batch_begin = 0;         // offset of this cluster's hypotheses in the frame-wide batch
};

inline int mask_count(const uint32_t *m, int W) {
  int c = 0;
  for (int w = 0; w < W; ++w) c += popc32(m[w]);
  return c;
}

#if defined(__x86_64__)
inline bool have_bmi2() {
  static const bool yes = [] {
    __builtin_cpu_init();
    return __builtin_cpu_supports("bmi2") != 0;
  }();
  return yes;
}
// position of the r-th set bit: deposit a single bit at the r-th set position of x
__attribute__((target("bmi2"))) inline int bit_select_bmi2(uint64_t x, uint32_t r) {
  return __builtin_ctzll(_pdep_u64(1ull << r, x));
}
#endif

// Per-thread scratch of the sampler: one candidate mask per level (64-bit words, no allocation per hypothesis).
struct SamplerScratch {
  std::vector<uint64_t> level[3];
  void fit(int W64) {
    for (auto &v : level)
      if (int(v.size()) < W64) v.resize(size_t(W64));
  }
};

inline int popc64(uint64_t x) { return __builtin_popcountll(x); }

// position of the r-th (0-based) set bit of x; x has more than r bits set
inline int select_in_word(uint64_t x, uint32_t r) {
#if defined(__x86_64__)
  if (have_bmi2()) return bit_select_bmi2(x, r);
#endif
  for (uint32_t i = 0; i < r; ++i) x &= x - 1;
  return __builtin_ctzll(x);
}

// index of the r-th set bit (ascending) of a W64-word mask
inline uint32_t select_bit64(const uint64_t *m, int W64, uint32_t r) {
  for (int w = 0; w < W64; ++w) {
    const uint32_t c = uint32_t(popc64(m[w]));
    if (r < c) return uint32_t(w) * 64u + uint32_t(select_in_word(m[w], r));
    r -= c;
  }
  return 0xFFFFFFFFu;
}

// the same on the round's valid mask through its prefix popcounts (vpre[w] = set bits in words < w)
inline uint32_t select_valid(const Cluster &c, int W64, uint32_t r) {
  const uint64_t *m = reinterpret_cast<const uint64_t *>(c.valid.data());
  int lo = 0, hi = W64;  // largest w with vpre[w] <= r
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (c.vpre[size_t(mid)] <= r) lo = mid; else hi = mid;
  }
  return uint32_t(lo) * 64u + uint32_t(select_in_word(m[lo], r - c.vpre[size_t(lo)]));
}

// Called when a round starts (the valid mask is fixed during a round).
void prepare_sampler(Cluster &c) {
  const int W64 = c.W / 2;
  const uint64_t *m = reinterpret_cast<const uint64_t *>(c.valid.data());
  c.vpre.resize(size_t(W64) + 1);
  uint32_t acc = 0;
  for (int w = 0; w < W64; ++w) {
    c.vpre[size_t(w)] = acc;
    acc += uint32_t(popc64(m[w]));
  }
  c.vpre[size_t(W64)] = acc;
}

// getSamples (sac_model_registration_graph.h:141-168) with drawIndexSampleHelper (:102-132) unrolled for 3 samples, on
// 64-bit mask words: s = valid[rand() % size]; valid' = valid & sample_adj.neighbors(s); recurse; on failure remove s
// and retry.  The samples come out deepest-first (s3, s2, s1) like samples_.push_back in the reference.  Returns false
// when the valid set holds no triangle of the sample graph (every one of the reference's 1000 retries would then fail
// identically, so one exhaustive attempt decides).  The level-0 mask is the round's valid mask itself until a retry
// has to remove a vertex from it.
bool get_samples(const Cluster &c, SamplerScratch &sc, uint64_t &rng, uint32_t triple[3]) {
  if (c.n_valid < 3) return false;
  const int W64 = c.W / 2;
  sc.fit(W64);
  const uint64_t *S = reinterpret_cast<const uint64_t *>(c.S);
  const uint64_t *cur0 = reinterpret_cast<const uint64_t *>(c.valid.data());
  uint64_t *own0 = sc.level[0].data(), *l1 = sc.level[1].data(), *l2 = sc.level[2].data();
  bool copied = false;
  int count0 = c.n_valid;
  for (;;) {
    const uint32_t r0 = uint32_t(uint64_t(tod_rng_next(&rng)) % uint64_t(count0));
    const uint32_t s0 = copied ? select_bit64(cur0, W64, r0) : select_valid(c, W64, r0);
    const uint64_t *row0 = S + size_t(s0) * W64;
    int count1 = tod::bitops::and_store_popcount(l1, cur0, row0, W64);
    while (count1 > 0) {
      const uint32_t r1 = uint32_t(uint64_t(tod_rng_next(&rng)) % uint64_t(count1));
      const uint32_t s1 = select_bit64(l1, W64, r1);
      const uint64_t *row1 = S + size_t(s1) * W64;
      const int count2 = tod::bitops::and_store_popcount(l2, l1, row1, W64);
      if (count2 > 0) {
        const uint32_t r2 = uint32_t(uint64_t(tod_rng_next(&rng)) % uint64_t(count2));
        triple[0] = select_bit64(l2, W64, r2);
        triple[1] = s1;
        triple[2] = s0;
        return true;
      }
      l1[s1 >> 6] &= ~(1ull << (s1 & 63));
      --count1;
    }
    if (!copied) {
      std::copy(cur0, cur0 + W64, own0);
      cur0 = own0;
      copied = true;
    }
    own0[s0 >> 6] &= ~(1ull << (s0 & 63));
    if (--count0 == 0) return false;
  }
}

// Small persistent pool: the per-cluster host work of a round (sampler, replay + gate, refinement + invalidation) is
// independent across objects, so it is spread over the host cores.  Results do not depend on the thread count: every
// cluster owns its sampler stream (tod_rng_seed(seed, object, round)) and its state.
class HostPool {
 public:
  explicit HostPool(int n_threads) {
    for (int t = 1; t < n_threads; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto &w : workers_) w.join();
  }
  int size() const { return int(workers_.size()) + 1; }
  // fn(item, thread) for item in [0, n); the caller is thread 0
  void run(int n, const std::function<void(int, int)> &fn) {
    if (n <= 0) return;
    if (workers_.empty() || n == 1) {
      for (int i = 0; i < n; ++i) fn(i, 0);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      n_ = n;
      next_.store(0);
      pending_ = int(workers_.size());
      ++epoch_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work(int t) {
    for (;;) {
      const int i = next_.fetch_add(1);
      if (i >= n_) break;
      (*fn_)(i, t);
    }
  }
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (quit_) return;
      }
      work(t);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int, int)> *fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool quit_ = false;
};

using Clock = std::chrono::steady_clock;
inline double ms_since(Clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

struct GateScratch {
  std::vector<uint32_t> inliers, filtered, mask;
  std::vector<uint32_t> adj, alive, uncoloured, cls;  // induced sub-graph as a dense bit-matrix + work masks
  std::vector<int> deg, rank;
  long calls = 0, proved_empty = 0;  // gate evaluations / those settled by the no-8-clique proof
  long core_rejects = 0;             // settled even earlier: fewer than 8 candidates inside the round's 7-core
  double ms_setup = 0, ms_proof = 0, ms_search = 0;
  // shape of the gate's work (tod_guess_last_gate_stats): [0..7] gates by induced-graph size, [8..15] by the size of
  // its 7-core (buckets <=16, 32, 64, 128, 256, 512, 1024, more), [16] core smaller than 8, [17] settled by the
  // colouring bound, [18] bounded searches run, [19] of those, passes, [20] search steps summed
  long hist[24] = {0};
  int last_core = 0;
};

inline int size_bucket(int n) {
  int b = 0;
  for (int lim = 16; b < 7 && n > lim; lim <<= 1) ++b;
  return b;
}

// Exact early "no": true iff the induced graph provably has NO clique of `size` vertices, so the bounded search of the
// reference (whatever its order, early stop and step budget) cannot return one and the gate fails
// (sac_model_registration_graph.h:260-265 needs clique.size() > minimal_size).  Proof = (size-1)-core peeling, then a
// greedy colouring of the core: fewer than `size` colours bound the clique number.  adj: nv x words bit-matrix.
bool proves_no_clique(GateScratch &g, int nv, int words, int size) {
  g.alive.assign(size_t(words), 0u);
  for (int v = 0; v < nv; ++v) g.alive[size_t(v) >> 5] |= 1u << (v & 31);
  g.deg.assign(size_t(nv), 0);
  int n_alive = nv;
  bool changed = true;
  while (changed) {  // peel vertices that have fewer than size-1 live neighbours
    changed = false;
    for (int v = 0; v < nv; ++v) {
      if (!((g.alive[size_t(v) >> 5] >> (v & 31)) & 1u)) continue;
      const uint32_t *row = g.adj.data() + size_t(v) * words;
      int d = 0;
      for (int w = 0; w < words; ++w) d += popc32(row[w] & g.alive[size_t(w)]);
      if (d < size - 1) {
        g.alive[size_t(v) >> 5] &= ~(1u << (v & 31));
        --n_alive;
        changed = true;
      }
    }
  }
  g.last_core = n_alive;
  if (n_alive < size) return true;
  // greedy colouring of the core with independent sets built on bit masks
  g.uncoloured = g.alive;
  int colours = 0, left = n_alive;
  while (left > 0) {
    if (++colours >= size) return false;  // bound not good enough: run the real search
    g.cls = g.uncoloured;                 // candidates for this colour class
    for (int w = 0; w < words; ++w) {
      while (g.cls[size_t(w)]) {
        const int v = w * 32 + __builtin_ctz(g.cls[size_t(w)]);
        g.uncoloured[size_t(v) >> 5] &= ~(1u << (v & 31));
        --left;
        const uint32_t *row = g.adj.data() + size_t(v) * words;
        for (int x = w; x < words; ++x) g.cls[size_t(x)] &= ~row[x];  // neighbours cannot share the colour
        g.cls[size_t(w)] &= ~(1u << (v & 31));
      }
    }
  }
  return colours < size;
}

// Inlier list of one hypothesis exactly as selectWithinDistance builds it before the gate (:178-200): common valid
// physical neighbours in ascending order, then the samples; each kept if it passes the distance test.
void hypothesis_inliers(const Cluster &c, const uint32_t s[3], bool inf_threshold, double thr2, const float *R,
                        const float *T, std::vector<uint32_t> &out) {
  out.clear();
  const int W = c.W;
  if (inf_threshold) {
    // a sample with a non-finite coordinate makes (R, T) NaN and every `distSq < inf` test false (K3 returns 0 too)
    for (int k = 0; k < 3; ++k)
      if (!((c.finite[s[k] >> 5] >> (s[k] & 31)) & 1u)) return;
  }
  const uint32_t *r0 = c.P + size_t(s[0]) * W, *r1 = c.P + size_t(s[1]) * W, *r2 = c.P + size_t(s[2]) * W;
  auto passes = [&](uint32_t i) -> bool {
    if (inf_threshold) return (c.finite[i >> 5] >> (i & 31)) & 1u;
    float p[3];
    tod::transform_point(R, T, c.q.data() + size_t(i) * 3, p);
    const float *t = c.t.data() + size_t(i) * 3;
    const float dx = p[0] - t[0], dy = p[1] - t[1], dz = p[2] - t[2];
    const float d2 = dx * dx + dy * dy + dz * dz;
    return double(d2) < thr2;
  };
  for (int w = 0; w < W; ++w) {
    uint32_t m = r0[w] & r1[w] & r2[w] & c.valid[size_t(w)];
    while (m) {
      const uint32_t i = uint32_t(w) * 32u + uint32_t(__builtin_ctz(m));
      m &= m - 1;
      if (passes(i)) out.push_back(i);
    }
  }
  for (int k = 0; k < 3; ++k)
    if (passes(s[k])) out.push_back(s[k]);
}

// Row of the induced sub-graph: the bits of `row` at the positions set in `mask`, packed in rank order into out
// (`words` u32, zeroed by the caller).  Portable version: walk the set bits.
inline void compress_row_generic(const uint32_t *row, const uint32_t *mask, const int *rank, int W, uint32_t *out) {
  for (int w = 0; w < W; ++w) {
    uint32_t m = row[w] & mask[w];
    while (m) {
      const int bit = __builtin_ctz(m);
      m &= m - 1;
      const int b = rank[w] + popc32(mask[w] & ((1u << bit) - 1u));
      out[b >> 5] |= 1u << (b & 31);
    }
  }
}

#if defined(__x86_64__)
// BMI2 version: one PEXT per word, appended to the output bit stream.
__attribute__((target("bmi2"))) inline void compress_row_bmi2(const uint32_t *row, const uint32_t *mask,
                                                              const int *rank, int W, uint32_t *out) {
  for (int w = 0; w < W; ++w) {
    const uint32_t mk = mask[w];
    if (!mk) continue;
    const uint64_t packed = _pext_u32(row[w], mk);
    const int b = rank[w];
    const int sh = b & 31;
    const uint64_t v = packed << sh;
    out[b >> 5] |= uint32_t(v);
    if (v >> 32) out[(b >> 5) + 1] |= uint32_t(v >> 32);
  }
}
#endif

// The clique gate of selectWithinDistance (:203-268) on an inlier list of size > 7.  Returns true if the list stands.
bool clique_gate(const Cluster &c, const std::vector<uint32_t> &inliers, GateScratch &g, bool proofs_done = false) {
  const size_t minimal = 7;  // std::min(best_inlier_number_, 7) with best_inlier_number_ >= 8 always (:85, :203)
  const int W = c.W;
  const Clock::time_point t0 = Clock::now();
  g.filtered.clear();
  for (uint32_t v : inliers)  // :209-213 — sample-degree inside the current valid set (one mask per round)
    if ((c.deg7[v >> 5] >> (v & 31)) & 1u) g.filtered.push_back(v);
  if (g.filtered.size() <= minimal) return false;
  std::sort(g.filtered.begin(), g.filtered.end());
  g.mask.assign(size_t(W), 0u);
  for (uint32_t v : g.filtered) g.mask[v >> 5] |= 1u << (v & 31);
  size_t reach = 0;  // :222-238
  for (uint32_t v : g.filtered) {
    const uint32_t *row = c.S + size_t(v) * W;
    int m = 0;
    for (int w = 0; w < W; ++w) m += popc32(row[w] & g.mask[size_t(w)]);
    reach = size_t(m);
    if (reach > minimal) break;
  }
  if (reach <= minimal) return false;
  const int nv = int(g.filtered.size());
  ++g.calls;
  if (2 * nv > c.n) {
    // most of the cluster survives (one object fills the frame: thousands of vertices): search the cluster's own
    // bit-rows through a view instead of compressing an induced copy.  No proofs here — K4 ran them where it could.
    g.ms_setup += ms_since(t0);
    const Clock::time_point ts = Clock::now();
    tod::CliqueFinder finder(c.n, W / 2, reinterpret_cast<const uint64_t *>(c.S), g.filtered.data(), nv);
    const bool ok = finder.finds_more_than(unsigned(minimal));
    g.ms_search += ms_since(ts);
    ++g.hist[size_bucket(nv)];
    ++g.hist[18];
    g.hist[19] += ok ? 1 : 0;
    g.hist[20] += finder.steps();
    return ok;
  }
  // induced sample sub-graph on `filtered` (:241-255) as a dense nv x words bit-matrix: vertex a = a-th smallest
  // member of `filtered`, i.e. its rank inside the mask
  const int words = (nv + 31) / 32;
  g.rank.resize(size_t(W) + 1);
  g.rank[0] = 0;
  for (int w = 0; w < W; ++w) g.rank[size_t(w) + 1] = g.rank[size_t(w)] + popc32(g.mask[size_t(w)]);
  g.adj.assign(size_t(nv) * words, 0u);
  for (int a = 0; a < nv; ++a) {
    const uint32_t *row = c.S + size_t(g.filtered[size_t(a)]) * W;
    uint32_t *out = g.adj.data() + size_t(a) * words;
#if defined(__x86_64__)
    if (have_bmi2()) {
      compress_row_bmi2(row, g.mask.data(), g.rank.data(), W, out);
      continue;
    }
#endif
    compress_row_generic(row, g.mask.data(), g.rank.data(), W, out);
  }
  const Clock::time_point t1 = Clock::now();
  g.ms_setup += std::chrono::duration<double, std::milli>(t1 - t0).count();
  // (when K4 has already run the proofs on the GPU and could not settle the hypothesis, they are not repeated)
  const bool none = (proofs_done && nv <= tod::kGateProofMax) ? false : proves_no_clique(g, nv, words, int(minimal) + 1);
  const Clock::time_point t2 = Clock::now();
  g.ms_proof += std::chrono::duration<double, std::milli>(t2 - t1).count();
  ++g.hist[size_bucket(nv)];
  if (!(proofs_done && nv <= tod::kGateProofMax)) ++g.hist[8 + size_bucket(g.last_core)];
  if (none) {
    ++g.proved_empty;
    ++g.hist[g.last_core < int(minimal) + 1 ? 16 : 17];
    return false;
  }
  // the bounded clique search (:258-265)
  tod::CliqueFinder finder(nv, g.adj.data());
  const bool ok = finder.finds_more_than(unsigned(minimal));
  g.ms_search += ms_since(t2);
  ++g.hist[18];
  g.hist[19] += ok ? 1 : 0;
  g.hist[20] += finder.steps();
  return ok;
}

// Host version of sample_degree_mask_kernel (k4_gate.cu), for the host-only entry points: valid vertices with at
// least 7 valid neighbours in the sample graph (:209-213).
void host_degree_mask(Cluster &c) {
  c.deg7.assign(size_t(c.W), 0u);
  for (int v = 0; v < c.n; ++v) {
    if (!((c.valid[size_t(v) >> 5] >> (v & 31)) & 1u)) continue;
    const uint32_t *row = c.S + size_t(v) * c.W;
    int d = 0;
    for (int w = 0; w < c.W; ++w) d += popc32(row[w] & c.valid[size_t(w)]);
    if (d >= 7) c.deg7[size_t(v) >> 5] |= 1u << (v & 31);
  }
}

}  // namespace

struct tod_guess {
  tod_guess_params p{};
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr, ev_k2 = nullptr, ev_S = nullptr,
              ev_P = nullptr;
  DeviceBuffer d_off, d_mo, d_q, d_t, d_px, d_sp, d_P, d_S, d_desc, d_valid, d_finite, d_hyps, d_counts, d_R, d_T;
  DeviceBuffer d_deg, d_active, d_floor, d_verdict;  // K4: degree masks, active-cluster list, per-cluster best, verdicts
  DeviceBuffer d_jobs;                               // K5: queue of packed induced sub-graphs (<= 128 vertices each)
  int64_t k5_stats[4] = {0, 0, 0, 0};
  tod::PinnedBuffer h_P, h_S;  // host copies of the bit-matrices (read by the sampler and the gate)
  tod::PinnedBuffer h_pts;     // staging of the clusters' query / training points and pixels for the upload
  tod::PinnedBuffer h_counts, h_verdict, h_hyps;  // per-batch K3 counts / gate verdicts / triples: pinned, so that the
                                                  // copies are truly asynchronous and the host can draw ahead
  float k2_ms = 0, k3_ms = 0;
  double k2_bytes = 0, k3_bytes = 0;  // algorithmic bytes of the last call's K2 / K3 launches (SURVEY.md §8d units)
  int64_t n_clusters = 0, n_correspondences = 0;
  int64_t n_hyp_total = 0;
  int32_t n_rounds = 0;
  // host wall-clock profile of the last process call (ms): see tod_guess_last_profile
  double prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int64_t gate_hist[24] = {0};
  HostPool *pool = nullptr;
};

extern "C" {

// Host-only: getSamples (sac_model_registration_graph.h:141-168) as this library runs it — n_hyp triples drawn from the
// sample graph given as an n x row_words(n) bit-matrix and a valid mask, consuming the stream *rng_state.  Returns the
// number of triples produced (the draw stops when the valid set holds no triangle).  For CPU tests against the reference.
int32_t tod_sample_triples(int32_t n, const uint32_t *sample_bits, const uint32_t *valid_bits, uint64_t *rng_state,
                           int32_t n_hyp, uint32_t *triples) {
  if (n < 0 || n_hyp < 0 || !sample_bits || !valid_bits || !rng_state || (n_hyp > 0 && !triples)) {
    tod::set_error("bad argument");
    return -1;
  }
  Cluster c;
  c.n = n;
  c.W = tod::adjacency_row_words(n);
  c.S = sample_bits;
  c.valid.assign(valid_bits, valid_bits + c.W);
  c.n_valid = mask_count(c.valid.data(), c.W);
  prepare_sampler(c);
  SamplerScratch sc;
  int32_t made = 0;
  for (; made < n_hyp; ++made)
    if (!get_samples(c, sc, *rng_state, triples + size_t(made) * 3)) break;
  return made;
}

// Host-only: selectWithinDistance (sac_model_registration_graph.h:171-269) with the reference's never-set threshold, as
// this library evaluates it — candidate list from the physical bit-rows, then the clique gate on the sample graph
// (degree filter, 7-core test, no-8-clique proofs, bounded search).  Returns the number of inliers written to
// `inliers` (capacity n + 3): the sorted list when the gate passes or at most 7 candidates exist, 0 when it fails.
int32_t tod_select_inliers(int32_t n, const uint32_t *physical_bits, const uint32_t *sample_bits,
                           const uint32_t *valid_bits, const uint32_t *triple, uint32_t *inliers) {
  if (n < 3 || !physical_bits || !sample_bits || !valid_bits || !triple || !inliers || triple[0] >= uint32_t(n) ||
      triple[1] >= uint32_t(n) || triple[2] >= uint32_t(n)) {
    tod::set_error("bad argument");
    return -1;
  }
  Cluster c;
  c.n = n;
  c.W = tod::adjacency_row_words(n);
  c.P = physical_bits;
  c.S = sample_bits;
  c.valid.assign(valid_bits, valid_bits + c.W);
  c.n_valid = mask_count(c.valid.data(), c.W);
  c.finite.assign(size_t(c.W), 0xFFFFFFFFu);
  host_degree_mask(c);
  std::vector<uint32_t> list;
  hypothesis_inliers(c, triple, true, std::numeric_limits<double>::infinity(), nullptr, nullptr, list);
  GateScratch gs;
  if (list.size() > 7) {  // :204-205 returns shorter lists as they are (candidates ascending, then the samples)
    if (!clique_gate(c, list, gs)) list.clear();
    std::sort(list.begin(), list.end());  // :267
  }
  for (size_t i = 0; i < list.size(); ++i) inliers[i] = list[i];
  return int32_t(list.size());
}

void tod_guess_default_params(tod_guess_params *p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->min_inliers = 15;
  p->n_ransac_iterations = 1000;
  p->sensor_error = 0.01f;
  p->device = 0;
  p->ransac_threshold = std::numeric_limits<double>::infinity();
  p->seed = 0;
}

int tod_guess_create(const tod_guess_params *p, tod_guess **out) {
  TOD_REQUIRE(p && out, "null argument");
  TOD_REQUIRE(p->sensor_error >= 0.f, "negative sensor_error");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return fail(TOD_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  TOD_REQUIRE(p->device >= 0 && p->device < n_dev, "device %d out of range (%d devices)", p->device, n_dev);
  TOD_CUDA(cudaSetDevice(p->device));
  tod_guess *g = new tod_guess();
  g->p = *p;
  // highest stream priority: the guess generator's kernels are short and the host waits for each of them; when a
  // matcher call of the next batch occupies the GPU (K1: thousands of long CTAs) they must not queue behind it
  int prio_lo = 0, prio_hi = 0;
  cudaError_t ce = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&g->stream, cudaStreamDefault, prio_hi);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&g->copy_stream, cudaStreamNonBlocking, prio_hi);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&g->ev_k2, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&g->ev_S, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&g->ev_P, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreate(&g->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&g->ev1);
  if (ce == cudaSuccess) ce = cudaEventCreate(&g->ev2);
  if (ce == cudaSuccess) ce = cudaEventCreate(&g->ev3);
  if (ce == cudaSuccess) ce = cudaEventCreate(&g->ev4);
  if (ce != cudaSuccess) {
    tod_guess_destroy(g);
    return fail(TOD_ERR_CUDA, "creating the guess generator's stream/events failed: %s", cudaGetErrorString(ce));
  }
  *out = g;
  return TOD_OK;
}

void tod_guess_destroy(tod_guess *g) {
  if (!g) return;
  cudaSetDevice(g->p.device);
  for (DeviceBuffer *b : {&g->d_off, &g->d_mo, &g->d_q, &g->d_t, &g->d_px, &g->d_sp, &g->d_P, &g->d_S, &g->d_desc,
                          &g->d_valid, &g->d_finite, &g->d_hyps, &g->d_counts, &g->d_R, &g->d_T, &g->d_deg,
                          &g->d_active, &g->d_floor, &g->d_verdict, &g->d_jobs})
    b->release();
  if (g->ev0) cudaEventDestroy(g->ev0);
  if (g->ev1) cudaEventDestroy(g->ev1);
  if (g->ev2) cudaEventDestroy(g->ev2);
  if (g->ev3) cudaEventDestroy(g->ev3);
  if (g->ev4) cudaEventDestroy(g->ev4);
  if (g->ev_k2) cudaEventDestroy(g->ev_k2);
  if (g->ev_S) cudaEventDestroy(g->ev_S);
  if (g->ev_P) cudaEventDestroy(g->ev_P);
  if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
  if (g->stream) cudaStreamDestroy(g->stream);
  g->h_P.release();
  g->h_S.release();
  g->h_counts.release();
  g->h_pts.release();
  g->h_verdict.release();
  g->h_hyps.release();
  delete g->pool;
  delete g;
}

void tod_guess_last_profile(const tod_guess *g, double *ms12) {
  if (!ms12) return;
  for (int i = 0; i < 12; ++i) ms12[i] = g ? g->prof[i] : 0.0;
}

void tod_guess_last_gate_stats(const tod_guess *g, int64_t *out24) {
  if (!out24) return;
  for (int i = 0; i < 24; ++i) out24[i] = g ? g->gate_hist[i] : 0;
}

void tod_guess_last_k5_stats(const tod_guess *g, int64_t *out4) {
  if (!out4) return;
  for (int i = 0; i < 4; ++i) out4[i] = g ? g->k5_stats[i] : 0;
}

void tod_guess_last_traffic(const tod_guess *g, double *k2_bytes, double *k3_bytes, int64_t *n_clusters,
                            int64_t *n_correspondences) {
  if (k2_bytes) *k2_bytes = g ? g->k2_bytes : 0.0;
  if (k3_bytes) *k3_bytes = g ? g->k3_bytes : 0.0;
  if (n_clusters) *n_clusters = g ? g->n_clusters : 0;
  if (n_correspondences) *n_correspondences = g ? g->n_correspondences : 0;
}

void tod_guess_last_stats(const tod_guess *g, float *k2_ms, float *k3_ms, int64_t *n_hyp, int32_t *n_rounds) {
  if (k2_ms) *k2_ms = g ? g->k2_ms : 0.f;
  if (k3_ms) *k3_ms = g ? g->k3_ms : 0.f;
  if (n_hyp) *n_hyp = g ? g->n_hyp_total : 0;
  if (n_rounds) *n_rounds = g ? g->n_rounds : 0;
}

int tod_guess_process_batch(tod_guess *g, int32_t n_frames, const int32_t *kp_offsets, const tod_keypoint *keypoints,
                            const float *clouds, int32_t height, int32_t width, const tod_match *matches,
                            const int32_t *counts, int32_t k, const float *points3d, const float *spans,
                            int32_t n_objects, tod_pose *poses, int32_t *pose_frames, int32_t max_poses,
                            int32_t *n_poses, int32_t *inlier_keypoints, int32_t max_inlier_total) {
  TOD_REQUIRE(g && n_poses, "null argument");
  *n_poses = 0;
  g->k2_ms = g->k3_ms = 0.f;
  g->k2_bytes = g->k3_bytes = 0.0;
  g->n_clusters = g->n_correspondences = 0;
  g->n_hyp_total = 0;
  g->n_rounds = 0;
  for (double &v : g->prof) v = 0.0;
  for (int64_t &v : g->gate_hist) v = 0;
  const Clock::time_point t_total = Clock::now();
  Clock::time_point t_phase = t_total;
  TOD_REQUIRE(n_frames >= 0 && k >= 1 && n_objects >= 0 && max_poses >= 0, "bad sizes");
  if (n_frames == 0) return TOD_OK;
  TOD_REQUIRE(kp_offsets && kp_offsets[0] == 0, "kp_offsets must start at 0");
  for (int f = 0; f < n_frames; ++f) TOD_REQUIRE(kp_offsets[f + 1] >= kp_offsets[f], "kp_offsets must not decrease");
  const int32_t n_kp_total = kp_offsets[n_frames];
  if (n_kp_total == 0) return TOD_OK;
  TOD_REQUIRE(keypoints && matches && counts && points3d && spans && (poses || max_poses == 0), "null input buffer");
  // "if (point_cloud.empty()) { TODO 2d-3d }" — no cloud, no poses (GuessGenerator.cpp:147-152)
  if (!clouds || height <= 0 || width <= 0) return TOD_OK;
  TOD_CUDA(cudaSetDevice(g->p.device));
  cudaStream_t st = g->stream;
  const float err = g->p.sensor_error;
  const bool inf_thr = !(g->p.ransac_threshold < 1e150);
  const double thr2 = inf_thr ? std::numeric_limits<double>::infinity() : g->p.ransac_threshold * g->p.ransac_threshold;

  // host threads: the per-frame and per-cluster work is independent (see HostPool); small inputs stay on the caller
  if (!g->pool) {
    int want = g->p.host_threads > 0 ? g->p.host_threads : int(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
    if (const char *e = getenv("TOD_HOST_THREADS")) want = std::max(1, atoi(e));
    g->pool = new HostPool(want);
  }
  HostPool &pool = *g->pool;

  // ---- ClusterPerObject (adjacency_ransac.cpp:176-205): one cluster per (frame, object) -------------------------------
  // Frames in parallel; when there are fewer frames than host threads (a single frame with 200 000 correspondences at
  // C5) a frame's keypoints are cut into consecutive chunks that are clustered in parallel and then appended in chunk
  // order — the same order of correspondences inside every cluster as one pass over the keypoints.
  std::vector<std::map<int, Cluster> > per_frame(static_cast<size_t>(n_frames));  // per frame: object -> cluster
  const int chunks = n_frames >= pool.size() ? 1 : (pool.size() + n_frames - 1) / n_frames;
  std::vector<std::map<int, Cluster> > partial(chunks > 1 ? size_t(n_frames) * size_t(chunks) : size_t(0));
  std::vector<std::string> frame_error(size_t(n_frames) * size_t(chunks));
  const size_t cloud_stride = size_t(height) * size_t(width) * 3;
  pool.run(n_frames * chunks, [&](int item, int) {
    const int f = item / chunks, ch = item % chunks;
    const float *cloud = clouds + size_t(f) * cloud_stride;
    std::map<int, Cluster> &mine = chunks > 1 ? partial[size_t(item)] : per_frame[size_t(f)];
    const int64_t n_f = kp_offsets[f + 1] - kp_offsets[f];
    const int32_t g_lo = kp_offsets[f] + int32_t(n_f * ch / chunks), g_hi = kp_offsets[f] + int32_t(n_f * (ch + 1) / chunks);
    char buf[200];
    for (int32_t gi = g_lo; gi < g_hi; ++gi) {
      const int32_t qi = gi - kp_offsets[f];  // keypoint index inside its frame
      const int cnt = counts[gi];
      if (cnt < 0 || cnt > k) {
        snprintf(buf, sizeof(buf), "counts[%d] = %d outside [0, k=%d]", gi, cnt, k);
        frame_error[size_t(item)] = buf;
        return;
      }
      // point_cloud.at<Vec3f>(pt.y, pt.x): float -> int conversion truncates (quirk Q9)
      const int y = int(keypoints[gi].y), x = int(keypoints[gi].x);
      if (!(y >= 0 && y < height && x >= 0 && x < width)) {
        snprintf(buf, sizeof(buf), "keypoint %d at (%g, %g) outside the %dx%d cloud", gi, keypoints[gi].x,
                 keypoints[gi].y, width, height);
        frame_error[size_t(item)] = buf;
        return;
      }
      const float *qp = cloud + (size_t(y) * width + x) * 3;
      if (std::isnan(qp[0])) continue;  // x only, like cvIsNaN(query_point[0]) (:189)
      for (int j = 0; j < cnt; ++j) {
        const tod_match &m = matches[size_t(gi) * k + j];
        if (m.imgIdx < 0 || m.imgIdx >= n_objects) {
          snprintf(buf, sizeof(buf), "match imgIdx %d outside [0, %d)", m.imgIdx, n_objects);
          frame_error[size_t(item)] = buf;
          return;
        }
        Cluster &c = mine[m.imgIdx];
        c.frame = f;
        c.object = m.imgIdx;
        const float *tp = points3d + (size_t(gi) * k + j) * 3;
        c.t.insert(c.t.end(), tp, tp + 3);
        c.q.insert(c.q.end(), qp, qp + 3);
        c.px.push_back(keypoints[gi].x);
        c.px.push_back(keypoints[gi].y);
        c.qidx.push_back(uint32_t(qi));
      }
    }
  });
  if (chunks > 1) {
    pool.run(n_frames, [&](int f, int) {
      std::map<int, Cluster> &dst = per_frame[size_t(f)];
      for (int ch = 0; ch < chunks; ++ch)
        for (auto &kv : partial[size_t(f) * size_t(chunks) + size_t(ch)]) {
          Cluster &part = kv.second;
          auto it = dst.find(kv.first);
          if (it == dst.end()) {
            dst.emplace(kv.first, std::move(part));
            continue;
          }
          Cluster &c = it->second;
          c.t.insert(c.t.end(), part.t.begin(), part.t.end());
          c.q.insert(c.q.end(), part.q.begin(), part.q.end());
          c.px.insert(c.px.end(), part.px.begin(), part.px.end());
          c.qidx.insert(c.qidx.end(), part.qidx.begin(), part.qidx.end());
        }
    });
  }
  for (const std::string &e : frame_error)
    if (!e.empty()) return fail(TOD_ERR_INVALID, "%s", e.c_str());
  std::vector<Cluster *> by_object;  // frames, then objects, ascending: the reference's std::map order per frame
  for (auto &mp : per_frame)
    for (auto &kv : mp) by_object.push_back(&kv.second);
  if (by_object.empty()) return TOD_OK;

  std::vector<Cluster *> clusters;
  std::vector<int32_t> offsets(1, 0);
  std::vector<int64_t> mo(1, 0);
  std::vector<float> spans_c;
  int64_t vo = 0;
  int max_n = 0;
  for (Cluster *cp : by_object) {
    Cluster &c = *cp;
    c.n = int(c.qidx.size());
    c.W = tod::adjacency_row_words(c.n);
    c.point_offset = offsets.back();
    c.matrix_offset = mo.back();
    c.valid_offset = vo;
    c.n_valid = c.n;
    offsets.push_back(offsets.back() + c.n);
    mo.push_back(mo.back() + int64_t(c.n) * c.W);
    vo += c.W;
    spans_c.push_back(spans[c.object]);
    max_n = std::max(max_n, c.n);
    clusters.push_back(&c);
  }
  const int nc = int(clusters.size());
  const int64_t N = offsets.back();
  const size_t mat_words = size_t(mo.back());
  g->n_clusters = nc;
  g->n_correspondences = N;
  g->k2_bytes = 32.0 * double(N) + 2.0 * 4.0 * double(mat_words);  // 32 n in + two n x W bit-matrices out

  // ---- FillAdjacency for every cluster: K2 --------------------------------------------------------------------------
  // staged in pinned memory (the uploads run at PCIe speed and do not park the host), clusters packed in parallel
  TOD_CUDA(g->h_pts.reserve(size_t(N) * 8 * sizeof(float)));
  float *all_q = g->h_pts.as<float>(), *all_t = all_q + size_t(N) * 3, *all_px = all_t + size_t(N) * 3;
  std::vector<uint32_t> all_valid(static_cast<size_t>(vo)), all_finite(static_cast<size_t>(vo));
  std::vector<unsigned char> desc(size_t(nc) * tod::k3_cluster_desc_size());
  pool.run(nc, [&](int ci, int) {
    Cluster &c = *clusters[size_t(ci)];
    c.valid.assign(size_t(c.W), 0u);
    c.finite.assign(size_t(c.W), 0u);
    for (int i = 0; i < c.n; ++i) {
      c.valid[size_t(i) >> 5] |= 1u << (i & 31);
      bool fin = true;
      for (int d = 0; d < 3; ++d) fin = fin && std::isfinite(c.q[size_t(i) * 3 + d]) && std::isfinite(c.t[size_t(i) * 3 + d]);
      if (fin) c.finite[size_t(i) >> 5] |= 1u << (i & 31);
    }
    std::copy(c.q.begin(), c.q.end(), all_q + c.point_offset * 3);
    std::copy(c.t.begin(), c.t.end(), all_t + c.point_offset * 3);
    std::copy(c.px.begin(), c.px.end(), all_px + c.point_offset * 2);
    std::copy(c.finite.begin(), c.finite.end(), all_finite.begin() + c.valid_offset);
    tod::k3_fill_cluster_desc(desc.data() + size_t(ci) * tod::k3_cluster_desc_size(), c.n, c.W, c.point_offset,
                              c.matrix_offset, c.valid_offset);
  });
  TOD_CUDA(g->d_off.reserve(offsets.size() * 4));
  TOD_CUDA(g->d_mo.reserve(mo.size() * 8));
  TOD_CUDA(g->d_q.reserve(size_t(N) * 3 * 4));
  TOD_CUDA(g->d_t.reserve(size_t(N) * 3 * 4));
  TOD_CUDA(g->d_px.reserve(size_t(N) * 2 * 4));
  TOD_CUDA(g->d_sp.reserve(spans_c.size() * 4));
  TOD_CUDA(g->d_P.reserve(mat_words * 4));
  TOD_CUDA(g->d_S.reserve(mat_words * 4));
  TOD_CUDA(g->d_desc.reserve(desc.size()));
  TOD_CUDA(g->d_valid.reserve(all_valid.size() * 4));
  TOD_CUDA(g->d_finite.reserve(all_finite.size() * 4));
  TOD_CUDA(g->d_deg.reserve(all_valid.size() * 4));
  TOD_CUDA(g->d_active.reserve(size_t(nc) * 4));
  TOD_CUDA(g->d_floor.reserve(size_t(nc) * 4));
  std::vector<uint32_t> all_deg(all_valid.size(), 0u);
  std::vector<int32_t> floor_by_cluster(static_cast<size_t>(nc), 0);
  const uint8_t *batch_verdict = nullptr;
  int max_W = 0;
  for (const Cluster *c : clusters) max_W = std::max(max_W, c->W);
  long k4_fails_used = 0, k4_host_used = 0, k5_pass_used = 0, k5_fail_used = 0;
  float k4_ms = 0.f, k5_ms = 0.f;  // K4 + K5 together / K5 alone
  TOD_CUDA(cudaMemcpyAsync(g->d_off.ptr, offsets.data(), offsets.size() * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(g->d_mo.ptr, mo.data(), mo.size() * 8, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(g->d_q.ptr, all_q, size_t(N) * 3 * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(g->d_t.ptr, all_t, size_t(N) * 3 * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(g->d_px.ptr, all_px, size_t(N) * 2 * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(g->d_sp.ptr, spans_c.data(), spans_c.size() * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(g->d_desc.ptr, desc.data(), desc.size(), cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(g->d_finite.ptr, all_finite.data(), all_finite.size() * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaEventRecord(g->ev0, st));
  TOD_CUDA(tod::launch_fill_adjacency(nc, g->d_off.as<int32_t>(), g->d_mo.as<int64_t>(), g->d_q.as<float>(),
                                      g->d_t.as<float>(), g->d_px.as<float>(), g->d_sp.as<float>(), err,
                                      g->d_P.as<uint32_t>(), g->d_S.as<uint32_t>(), max_n, st));
  TOD_CUDA(cudaEventRecord(g->ev1, st));
  TOD_CUDA(g->h_P.reserve(mat_words * 4));
  TOD_CUDA(g->h_S.reserve(mat_words * 4));
  // The sample matrix comes back first (the sampler of round 0 needs it at once); the physical matrix follows on the
  // copy stream while the first hypotheses are already being drawn and scored — the host reads it only when it builds
  // the inlier list of a hypothesis that reaches the host gate.
  TOD_CUDA(cudaEventRecord(g->ev_k2, st));
  TOD_CUDA(cudaStreamWaitEvent(g->copy_stream, g->ev_k2, 0));
  TOD_CUDA(cudaMemcpyAsync(g->h_S.ptr, g->d_S.ptr, mat_words * 4, cudaMemcpyDeviceToHost, g->copy_stream));
  TOD_CUDA(cudaEventRecord(g->ev_S, g->copy_stream));
  TOD_CUDA(cudaMemcpyAsync(g->h_P.ptr, g->d_P.ptr, mat_words * 4, cudaMemcpyDeviceToHost, g->copy_stream));
  TOD_CUDA(cudaEventRecord(g->ev_P, g->copy_stream));
  bool physical_on_host = false;
  TOD_CUDA(cudaEventSynchronize(g->ev_S));
  TOD_CUDA(cudaEventElapsedTime(&g->k2_ms, g->ev0, g->ev1));
  for (Cluster *c : clusters) {
    c->P = g->h_P.as<uint32_t>() + c->matrix_offset;
    c->S = g->h_S.as<uint32_t>() + c->matrix_offset;
  }
  // "InvalidateIndices({})" at the end of FillAdjacency is a no-op (quirk Q4): no pruning before the first round.
  g->prof[0] = ms_since(t_phase);  // ClusterPerObject + upload + K2 + bit-matrix download

  // ---- RANSAC rounds, all active objects in lock-step ------------------------------------------------------------------
  struct Found {
    int frame;
    int object;
    unsigned round;
    tod_pose pose;
    std::vector<uint32_t> kp;
  };
  const int max_iter = int(g->p.n_ransac_iterations);
  bool long_rounds_seen = false;  // a round of this call went past its first batch of hypotheses
  const int n_thr = pool.size();
  struct ThreadScratch {
    GateScratch gate;
    SamplerScratch sampler;
    std::vector<uint32_t> inliers;
    std::string error;
    long k4_fails = 0, k4_host = 0, k5_pass = 0, k5_fail = 0;
  };
  std::vector<ThreadScratch> ts(static_cast<size_t>(n_thr));
  std::vector<std::vector<Found>> found_by_cluster(clusters.size());
  std::vector<uint32_t> batch_hyps;
  const int32_t *batch_counts = nullptr;
  std::vector<float> batch_R, batch_T;
  std::vector<int> active_idx;

  while (true) {
    // start a round on every active cluster (AdjacencyRansac::Ransac, adjacency_ransac.cpp:234-253)
    active_idx.clear();
    for (size_t ci = 0; ci < clusters.size(); ++ci) {
      Cluster *c = clusters[ci];
      if (!c->active) continue;
      if (c->n_valid < 3) {  // :238-241 -> no inliers -> below min_inliers -> object finished
        c->active = false;
        continue;
      }
      c->rng = tod_rng_seed(g->p.seed, uint32_t(c->object), c->round);
      prepare_sampler(*c);
      c->iterations = 0;
      c->n_best = -INT_MAX;
      c->k = 1.0;
      c->stopped = false;
      c->best_inliers.clear();
      std::copy(c->valid.begin(), c->valid.end(), all_valid.begin() + c->valid_offset);
      active_idx.push_back(int(ci));
    }
    if (active_idx.empty()) break;
    ++g->n_rounds;
    t_phase = Clock::now();
    TOD_CUDA(cudaMemcpyAsync(g->d_valid.ptr, all_valid.data(), all_valid.size() * 4, cudaMemcpyHostToDevice, st));
    // the gate's degree filter (:209-213) for this round: one mask per active cluster, computed on the GPU (K4 reads
    // it there) and copied back for the few gates that end on the host
    {
      int max_active_n = 0;
      for (int ci : active_idx) max_active_n = std::max(max_active_n, clusters[size_t(ci)]->n);
      TOD_CUDA(cudaMemcpyAsync(g->d_active.ptr, active_idx.data(), active_idx.size() * 4, cudaMemcpyHostToDevice, st));
      TOD_CUDA(tod::launch_sample_degree_mask(g->d_desc.ptr, g->d_active.as<int32_t>(), int(active_idx.size()),
                                              max_active_n, g->d_S.as<uint32_t>(), g->d_valid.as<uint32_t>(),
                                              g->d_deg.as<uint32_t>(), 7, st));
      TOD_CUDA(cudaMemcpyAsync(all_deg.data(), g->d_deg.ptr, all_deg.size() * 4, cudaMemcpyDeviceToHost, st));
    }
    bool deg_on_host = false;
    g->prof[4] += ms_since(t_phase);

    // computeModel (ransac.h:80-143): hypotheses are drawn and scored in growing batches; the replay below consumes
    // them in order and stops exactly where the reference's loop would.
    // batches of 64, 256, 1024, 1024, ... hypotheses per cluster: while the GPU scores one the host draws the next, so
    // only the last batch of a round is waited for in full — hence the cap (4096 left 3 ms per round exposed at C5)
    constexpr int kMaxBatch = 1024;
    int batch_size = 64;
    int batches_this_round = 0;
    while (true) {
      // -- sampler (getSamples, sac_model_registration_graph.h:141-168), clusters in parallel ---------------------------
      t_phase = Clock::now();
      pool.run(int(active_idx.size()), [&](int ai, int t) {
        Cluster *c = clusters[size_t(active_idx[size_t(ai)])];
        c->hyps.clear();
        if (c->stopped) return;
        if (c->drawn_ahead) {  // this batch was drawn while the GPU scored the previous one
          c->hyps.swap(c->hyps_next);
          c->drawn_ahead = false;
          return;
        }
        // the loop can run at most until iterations_ exceeds max_iterations_ (ransac.h:132-134)
        const int room = std::min(batch_size, max_iter + 1 - c->iterations);
        for (int h = 0; h < room; ++h) {
          uint32_t tr[3];
          if (!get_samples(*c, ts[size_t(t)].sampler, c->rng, tr)) break;  // empty selection -> break (ransac.h:100-101)
          c->hyps.insert(c->hyps.end(), tr, tr + 3);
        }
      });
      batch_hyps.clear();
      for (int ci : active_idx) {
        Cluster *c = clusters[size_t(ci)];
        c->batch_begin = int(batch_hyps.size() / 4);
        for (size_t h = 0; h + 2 < c->hyps.size(); h += 3) {
          batch_hyps.push_back(c->hyps[h]);
          batch_hyps.push_back(c->hyps[h + 1]);
          batch_hyps.push_back(c->hyps[h + 2]);
          batch_hyps.push_back(uint32_t(ci));
        }
      }
      g->prof[1] += ms_since(t_phase);

      // -- K3: one launch for every hypothesis of every object ----------------------------------------------------------
      t_phase = Clock::now();
      const int H = int(batch_hyps.size() / 4);
      if (H > 0) {
        TOD_CUDA(g->h_counts.reserve(size_t(H) * 4));
        batch_counts = g->h_counts.as<int32_t>();
        TOD_CUDA(g->d_hyps.reserve(batch_hyps.size() * 4));
        TOD_CUDA(g->d_counts.reserve(size_t(H) * 4));
        if (!inf_thr) {
          TOD_CUDA(g->d_R.reserve(size_t(H) * 36));
          TOD_CUDA(g->d_T.reserve(size_t(H) * 12));
          batch_R.resize(size_t(H) * 9);
          batch_T.resize(size_t(H) * 3);
        }
        TOD_CUDA(g->h_hyps.reserve(batch_hyps.size() * 4));
        std::memcpy(g->h_hyps.ptr, batch_hyps.data(), batch_hyps.size() * 4);
        TOD_CUDA(cudaMemcpyAsync(g->d_hyps.ptr, g->h_hyps.ptr, batch_hyps.size() * 4, cudaMemcpyHostToDevice, st));
        TOD_CUDA(cudaEventRecord(g->ev0, st));
        TOD_CUDA(tod::launch_score_hypotheses_batched(
            g->d_desc.ptr, g->d_q.as<float>(), g->d_t.as<float>(), g->d_P.as<uint32_t>(), g->d_valid.as<uint32_t>(),
            g->d_finite.as<uint32_t>(), H, g->d_hyps.as<uint32_t>(), g->p.ransac_threshold, g->d_counts.as<int32_t>(),
            inf_thr ? nullptr : g->d_R.as<float>(), inf_thr ? nullptr : g->d_T.as<float>(), st));
        TOD_CUDA(cudaEventRecord(g->ev1, st));
        TOD_CUDA(cudaMemcpyAsync(g->h_counts.ptr, g->d_counts.ptr, size_t(H) * 4, cudaMemcpyDeviceToHost, st));
        if (inf_thr) {
          // K4: the gate's exact pre-checks for every hypothesis of the batch that beats its cluster's best so far
          for (int ci : active_idx) {
            const Cluster *c = clusters[size_t(ci)];
            floor_by_cluster[size_t(ci)] = c->n_best < 0 ? 0 : c->n_best;
          }
          TOD_CUDA(g->h_verdict.reserve(size_t(H)));
          batch_verdict = g->h_verdict.as<uint8_t>();
          TOD_CUDA(g->d_verdict.reserve(size_t(H)));
          // K5 queue: room for every hypothesis of the launch at 2 KB (128 vertices; the rare graphs of up to 256
          // take 8 KB, most take 0.5 KB), capped at 1 GiB; a hypothesis that finds the pool full goes to the host
          const size_t pool_bytes = std::min<size_t>(std::max<size_t>(size_t(H) * 2048, 65536), size_t(1) << 30);
          TOD_CUDA(g->d_jobs.reserve(tod::gate_job_bytes(H, pool_bytes)));
          TOD_CUDA(cudaMemcpyAsync(g->d_floor.ptr, floor_by_cluster.data(), floor_by_cluster.size() * 4,
                                   cudaMemcpyHostToDevice, st));
          TOD_CUDA(cudaEventRecord(g->ev2, st));
          TOD_CUDA(tod::launch_gate_prechecks(g->d_desc.ptr, g->d_P.as<uint32_t>(), g->d_S.as<uint32_t>(),
                                              g->d_valid.as<uint32_t>(), g->d_finite.as<uint32_t>(),
                                              g->d_deg.as<uint32_t>(), H, g->d_hyps.as<uint32_t>(),
                                              g->d_counts.as<int32_t>(), g->d_floor.as<int32_t>(), max_W,
                                              g->d_verdict.as<uint8_t>(), g->d_jobs.ptr, pool_bytes, st, g->ev4));
          TOD_CUDA(cudaEventRecord(g->ev3, st));
          TOD_CUDA(cudaMemcpyAsync(g->h_verdict.ptr, g->d_verdict.ptr, size_t(H), cudaMemcpyDeviceToHost, st));
        }
        if (!inf_thr) {
          TOD_CUDA(cudaMemcpyAsync(batch_R.data(), g->d_R.ptr, size_t(H) * 36, cudaMemcpyDeviceToHost, st));
          TOD_CUDA(cudaMemcpyAsync(batch_T.data(), g->d_T.ptr, size_t(H) * 12, cudaMemcpyDeviceToHost, st));
        }
        // While the GPU scores this batch: draw the NEXT batch of every cluster that can still need one.  The sampler
        // stream of a round does not depend on the replay (the valid set is fixed during a round), so the triples are
        // the ones the next batch would draw anyway; if the replay stops the cluster they are dropped with the round.
        // Only in rounds that are already known to be long (this is their second batch, or an earlier round of this
        // call needed one): short rounds — a clean object found in the first 64 hypotheses — pay nothing.
        if (inf_thr && (batches_this_round >= 1 || long_rounds_seen)) {
          const Clock::time_point t_ahead = Clock::now();
          const int next_size = std::min(batch_size * 4, kMaxBatch);
          pool.run(int(active_idx.size()), [&](int ai, int t) {
            Cluster *c = clusters[size_t(active_idx[size_t(ai)])];
            c->drawn_ahead = false;
            if (c->stopped) return;
            const int nh = int(c->hyps.size() / 3);
            const int room_now = std::min(batch_size, max_iter + 1 - c->iterations);
            if (nh < room_now) return;                        // the sampler ran dry: the round ends with this batch
            const int room = std::min(next_size, max_iter + 1 - (c->iterations + nh));
            if (room <= 0) return;
            c->hyps_next.clear();
            for (int h = 0; h < room; ++h) {
              uint32_t tr[3];
              if (!get_samples(*c, ts[size_t(t)].sampler, c->rng, tr)) break;
              c->hyps_next.insert(c->hyps_next.end(), tr, tr + 3);
            }
            c->drawn_ahead = true;
          });
          g->prof[1] += ms_since(t_ahead);
          t_phase += Clock::now() - t_ahead;                  // accounted under "sampler", not under "K3 launch + sync"
        }
        TOD_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        TOD_CUDA(cudaEventElapsedTime(&ms, g->ev0, g->ev1));
        g->k3_ms += ms;
        if (inf_thr) {
          TOD_CUDA(cudaEventElapsedTime(&ms, g->ev2, g->ev3));
          k4_ms += ms;
          TOD_CUDA(cudaEventElapsedTime(&ms, g->ev4, g->ev3));
          k5_ms += ms;
        }
        if (!deg_on_host) {  // the round's degree masks arrived with this synchronisation
          for (int ci : active_idx) {
            Cluster *c = clusters[size_t(ci)];
            c->deg7.assign(all_deg.begin() + c->valid_offset, all_deg.begin() + c->valid_offset + c->W);
          }
          deg_on_host = true;
        }
        g->n_hyp_total += H;
        for (int ci : active_idx) {  // per hypothesis: 3 physical rows + valid + finite masks, samples, triple, result
          const Cluster *c = clusters[size_t(ci)];
          g->k3_bytes += double(c->hyps.size() / 3) * (5.0 * 4.0 * double(c->W) + 72.0 + 16.0 + 52.0);
        }
      }
      g->prof[2] += ms_since(t_phase);

      // -- replay of computeModel's scan (ransac.h:95-135) + lazy clique gate ---------------------------------------------
      t_phase = Clock::now();
      if (!physical_on_host) {
        TOD_CUDA(cudaEventSynchronize(g->ev_P));
        physical_on_host = true;
      }
      std::atomic<int> more{0};
      // exact post-gate count of hypothesis h of cluster c (the gate keeps or zeroes the pre-gate count); the inlier
      // list is left in sc.inliers (cleared when the gate fails).  Depends only on h and the round's state, never on
      // the best-so-far, so it may be evaluated ahead of the sequential scan.  Returns -1 on an internal error.
      auto evaluate = [&](Cluster *c, int h, ThreadScratch &sc) -> int {
        const int pre = batch_counts[size_t(c->batch_begin + h)];
        bool proofs_done = false;
        if (inf_thr && pre > 7) {
          const uint8_t v = batch_verdict[size_t(c->batch_begin + h)];
          proofs_done = c->W <= 128;  // K4 evaluates clusters of up to 4096 correspondences
          if (v == tod::kGateFails || v == tod::kGateFailsSearch) {  // K4 proved / K5 found that the gate clears this list
            sc.inliers.clear();
            ++(v == tod::kGateFails ? sc.k4_fails : sc.k5_fail);
            return 0;
          }
          if (v == tod::kGatePasses) {  // K5 ran the reference's search to its end: the candidate list stands
            ++sc.k5_pass;
            hypothesis_inliers(*c, c->hyps.data() + size_t(h) * 3, inf_thr, thr2, nullptr, nullptr, sc.inliers);
            if (int(sc.inliers.size()) != pre) {
              sc.error = "K3 count disagrees with the host candidate list";
              return -1;
            }
            return pre;
          }
          if (v == tod::kGateNeedsHost) ++sc.k4_host;
          else {
            sc.error = "K4 left a hypothesis that beats its cluster's best unevaluated";
            return -1;
          }
        }
        const uint32_t *s = c->hyps.data() + size_t(h) * 3;
        const float *R = inf_thr ? nullptr : batch_R.data() + size_t(c->batch_begin + h) * 9;
        const float *T = inf_thr ? nullptr : batch_T.data() + size_t(c->batch_begin + h) * 3;
        hypothesis_inliers(*c, s, inf_thr, thr2, R, T, sc.inliers);
        if (int(sc.inliers.size()) != pre) {
          char buf[160];
          snprintf(buf, sizeof(buf), "K3 count %d disagrees with the host candidate list %zu (object %d)", pre,
                   sc.inliers.size(), c->object);
          sc.error = buf;
          return -1;
        }
        if (sc.inliers.size() > 7 && !clique_gate(*c, sc.inliers, sc.gate, proofs_done)) {
          sc.inliers.clear();
          return 0;
        }
        return pre;
      };
      // the sequential scan of one cluster's batch; result(h, inliers_out) yields the post-gate count of h
      auto scan = [&](Cluster *c, const std::function<int(int, std::vector<uint32_t> &)> &result) {
        const int nh = int(c->hyps.size() / 3);
        const int room = std::min(batch_size, max_iter + 1 - c->iterations);
        std::vector<uint32_t> kept;
        for (int h = 0; h < nh; ++h) {
          if (!(double(c->iterations) < c->k)) {  // while (iterations_ < k)
            c->stopped = true;
            break;
          }
          const int pre = batch_counts[size_t(c->batch_begin + h)];
          if (pre > c->n_best) {
            // only now does the exact post-gate count matter
            const int final_count = result(h, kept);
            if (final_count < 0) {
              c->stopped = true;
              return;
            }
            if (final_count > c->n_best) {  // ransac.h:115-130
              c->n_best = final_count;
              c->best_inliers = kept;
              if (!inf_thr) {
                std::memcpy(c->best_R, batch_R.data() + size_t(c->batch_begin + h) * 9, sizeof(c->best_R));
                std::memcpy(c->best_T, batch_T.data() + size_t(c->batch_begin + h) * 3, sizeof(c->best_T));
              }
              const double w = double(c->n_best) / double(c->n_valid);
              double p_no = 1.0 - std::pow(w, 3.0);
              p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
              p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
              c->k = std::log(1.0 - 0.99) / std::log(p_no);
            }
          }
          ++c->iterations;
          if (c->iterations > max_iter) {
            c->stopped = true;
            break;
          }
        }
        if (!c->stopped) {
          if (nh < room) c->stopped = true;                        // sampler ran dry: selection.empty() -> break
          else if (!(double(c->iterations) < c->k)) c->stopped = true;
          else more.store(1, std::memory_order_relaxed);
        }
      };
      int n_scanning = 0;
      for (int ci : active_idx) n_scanning += clusters[size_t(ci)]->stopped ? 0 : 1;
      if (n_thr == 1 || n_scanning != 1) {
        // one cluster per work item, gates evaluated inside the scan (with 3 objects in a frame this already beats the
        // wave scheme below: 3.5 ms against 7 ms on the 3 x 400 case, because passes keep the waves short)
        pool.run(int(active_idx.size()), [&](int ai, int t) {
          Cluster *c = clusters[size_t(active_idx[size_t(ai)])];
          if (c->stopped) return;
          ThreadScratch &sc = ts[size_t(t)];
          scan(c, [&](int h, std::vector<uint32_t> &kept) {
            const int r = evaluate(c, h, sc);
            kept.swap(sc.inliers);  // sc.inliers is refilled from scratch by the next evaluation
            return r;
          });
        });
      } else {
        // a single cluster left (one object in the frame, or the last object still producing poses): the gates of the
        // NEXT hypotheses that still matter are evaluated ahead of the scan, in parallel, in waves that double while
        // every gate fails (the common case in outlier-heavy rounds) and fall back to 1 after a pass — a pass raises
        // the best-so-far and usually makes the following evaluations unnecessary.
        std::vector<int> cached;
        std::vector<std::vector<uint32_t> > cached_inliers;
        std::vector<int> wave_items;
        for (int ci : active_idx) {
          Cluster *c = clusters[size_t(ci)];
          if (c->stopped) continue;
          const int nh = int(c->hyps.size() / 3);
          cached.assign(size_t(nh), -2);  // -2 = not evaluated
          cached_inliers.assign(size_t(nh), std::vector<uint32_t>());
          int wave = 1;
          scan(c, [&](int h, std::vector<uint32_t> &kept) {
            if (cached[size_t(h)] == -2) {
              wave_items.clear();
              for (int x = h; x < nh && int(wave_items.size()) < wave; ++x)
                if (cached[size_t(x)] == -2 && batch_counts[size_t(c->batch_begin + x)] > c->n_best)
                  wave_items.push_back(x);
              pool.run(int(wave_items.size()), [&](int wi, int t) {
                const int x = wave_items[size_t(wi)];
                ThreadScratch &sc = ts[size_t(t)];
                cached[size_t(x)] = evaluate(c, x, sc);
                if (cached[size_t(x)] > 0) cached_inliers[size_t(x)] = sc.inliers;
              });
              bool any_pass = false;
              for (int x : wave_items) any_pass = any_pass || cached[size_t(x)] > 0;
              wave = any_pass ? 1 : std::min(wave * 2, 4 * n_thr);
            }
            kept = cached_inliers[size_t(h)];
            return cached[size_t(h)];
          });
        }
      }
      g->prof[3] += ms_since(t_phase);
      for (const ThreadScratch &sc : ts)
        if (!sc.error.empty()) return fail(TOD_ERR_STATE, "%s", sc.error.c_str());
      if (!more.load()) break;
      batch_size = std::min(batch_size * 4, kMaxBatch);
      ++batches_this_round;
      long_rounds_seen = true;
    }
    for (int ci : active_idx) clusters[size_t(ci)]->drawn_ahead = false;  // what was drawn ahead ends with the round

    // finish the round per cluster: refinement, pose, invalidation (adjacency_ransac.cpp:255-308,
    // GuessGenerator.cpp:205-230), clusters in parallel
    t_phase = Clock::now();
    pool.run(int(active_idx.size()), [&](int ai, int) {
      const int ci = active_idx[size_t(ai)];
      Cluster *c = clusters[size_t(ci)];
      if (c->best_inliers.empty()) {  // computeModel() returned false -> inliers_in stays empty -> below min_inliers
        c->active = false;
        return;
      }
      std::vector<uint32_t> inliers(c->best_inliers);
      std::sort(inliers.begin(), inliers.end());
      std::vector<uint32_t> in_mask(size_t(c->W), 0u);
      for (uint32_t i : inliers) in_mask[i >> 5] |= 1u << (i & 31);
      std::vector<uint32_t> rest;  // valid_indices_ \ inliers, ascending
      for (int w = 0; w < c->W; ++w) {
        uint32_t m = c->valid[size_t(w)] & ~in_mask[size_t(w)];
        while (m) {
          rest.push_back(uint32_t(w) * 32u + uint32_t(__builtin_ctz(m)));
          m &= m - 1;
        }
      }
      bool do_final = false;
      double thresh = double(err * err);  // :267 (float product widened)
      float R[9], T[3];
      std::memcpy(R, c->best_R, sizeof(R));
      std::memcpy(T, c->best_T, sizeof(T));
      while (true) {
        // optimizeModelCoefficients leaves (R, T) untouched with fewer than 3 inliers (sac_model...h:307-308)
        if (inliers.size() >= 3)
          tod::rigid_fit(c->q.data(), c->t.data(), inliers.data(), int(inliers.size()), R, T);  // :272
        std::vector<uint32_t> extra, keep;
        for (uint32_t i : rest) {  // :276-283
          float p[3];
          tod::transform_point(R, T, c->q.data() + size_t(i) * 3, p);
          const float *t = c->t.data() + size_t(i) * 3;
          const double dx = double(p[0] - t[0]), dy = double(p[1] - t[1]), dz = double(p[2] - t[2]);
          const double nrm = std::sqrt(dx * dx + dy * dy + dz * dz);  // cv::norm(Vec3f): double accumulation
          if (nrm * nrm < thresh) extra.push_back(i);
          else keep.push_back(i);
        }
        if (!extra.empty()) {
          std::vector<uint32_t> merged(inliers.size() + extra.size());
          std::merge(inliers.begin(), inliers.end(), extra.begin(), extra.end(), merged.begin());
          inliers.swap(merged);
          rest.swap(keep);
        }
        if (do_final) break;
        if (extra.empty()) {
          do_final = true;
          thresh *= 4;
        }
      }
      // R = R^T ; T = -R * T   (object -> camera, :304-305)
      float Rt[9];
      for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) Rt[r * 3 + cc] = R[cc * 3 + r];
      float nR[9], Tn[3];
      for (int i = 0; i < 9; ++i) nR[i] = Rt[i] * -1.f;
      const float zero[3] = {0.f, 0.f, 0.f};
      tod::transform_point(nR, zero, T, Tn);
      std::vector<uint32_t> kp;
      kp.reserve(inliers.size());
      for (uint32_t i : inliers) kp.push_back(c->qidx[i]);
      std::sort(kp.begin(), kp.end());
      kp.erase(std::unique(kp.begin(), kp.end()), kp.end());
      if (kp.size() < g->p.min_inliers) {  // GuessGenerator.cpp:205-206
        c->active = false;
        return;
      }
      // InvalidateQueryIndices (:93-123): drop every still-valid match whose keypoint is an inlier, then cascade
      std::vector<uint32_t> todo;
      for (int w = 0; w < c->W; ++w) {
        uint32_t m = c->valid[size_t(w)];
        while (m) {
          const uint32_t i = uint32_t(w) * 32u + uint32_t(__builtin_ctz(m));
          m &= m - 1;
          if (std::binary_search(kp.begin(), kp.end(), c->qidx[i])) todo.push_back(i);
        }
      }
      while (!todo.empty()) {  // InvalidateIndices (:63-89)
        for (uint32_t i : todo) c->valid[i >> 5] &= ~(1u << (i & 31));
        c->n_valid -= int(todo.size());
        todo.clear();
        for (int w = 0; w < c->W; ++w) {
          uint32_t m = c->valid[size_t(w)];
          while (m) {
            const uint32_t i = uint32_t(w) * 32u + uint32_t(__builtin_ctz(m));
            m &= m - 1;
            const uint32_t *row = c->S + size_t(i) * c->W;
            int deg = 0;
            for (int ww = 0; ww < c->W; ++ww) deg += popc32(row[ww] & c->valid[size_t(ww)]);
            if (deg < 3) todo.push_back(i);  // min_sample_size_ (adjacency_ransac.h:58)
          }
        }
      }
      Found f;
      f.frame = c->frame;
      f.object = c->object;
      f.round = c->round;
      std::memcpy(f.pose.R, Rt, sizeof(Rt));
      std::memcpy(f.pose.T, Tn, sizeof(Tn));
      f.pose.object_index = c->object;
      f.pose.n_inliers = int32_t(kp.size());
      f.kp.swap(kp);
      found_by_cluster[size_t(ci)].push_back(std::move(f));
      ++c->round;
    });
    g->prof[4] += ms_since(t_phase);
  }

  if (!physical_on_host) TOD_CUDA(cudaEventSynchronize(g->ev_P));  // nothing may still be in flight when we return
  // emission order of the reference: objects ascending (std::map), rounds in order (GuessGenerator.cpp:170-235);
  // `clusters` is already in ascending (frame, object) order and every cluster appended its rounds in order
  std::vector<Found> found;
  for (auto &v : found_by_cluster)
    for (auto &f : v) found.push_back(std::move(f));
  for (const ThreadScratch &sc : ts) {
    g->prof[5] += double(sc.gate.calls);
    g->prof[6] += double(sc.gate.proved_empty);
    g->prof[11] += double(sc.gate.core_rejects);
    g->prof[8] += sc.gate.ms_setup;
    g->prof[9] += sc.gate.ms_proof;
    g->prof[10] += sc.gate.ms_search;
    for (int i = 0; i < 24; ++i) g->gate_hist[i] += sc.gate.hist[i];
    k4_fails_used += sc.k4_fails;
    k4_host_used += sc.k4_host;
    k5_pass_used += sc.k5_pass;
    k5_fail_used += sc.k5_fail;
  }
  g->k5_stats[0] = k5_pass_used;
  g->k5_stats[1] = k5_fail_used;
  g->k5_stats[2] = k4_host_used;
  g->k5_stats[3] = int64_t(k5_ms * 1000.f);
  g->gate_hist[21] = k4_fails_used;
  g->gate_hist[22] = k4_host_used;
  g->gate_hist[23] = int64_t(k4_ms * 1000.f);
  g->prof[7] = ms_since(t_total);
  if (int64_t(found.size()) > max_poses)
    return fail(TOD_ERR_LIMIT, "%zu poses found but max_poses = %d", found.size(), max_poses);
  int64_t n_inl = 0;
  for (size_t i = 0; i < found.size(); ++i) {
    poses[i] = found[i].pose;
    if (pose_frames) pose_frames[i] = found[i].frame;
    if (inlier_keypoints) {
      if (n_inl + int64_t(found[i].kp.size()) > max_inlier_total)
        return fail(TOD_ERR_LIMIT, "inlier_keypoints capacity %d too small", max_inlier_total);
      for (uint32_t v : found[i].kp) inlier_keypoints[n_inl++] = int32_t(v);
    }
  }
  *n_poses = int32_t(found.size());
  return TOD_OK;
}

int tod_guess_process(tod_guess *g, const tod_keypoint *keypoints, int32_t n_kp, const float *cloud, int32_t height,
                      int32_t width, const tod_match *matches, const int32_t *counts, int32_t k,
                      const float *points3d, const float *spans, int32_t n_objects, tod_pose *poses,
                      int32_t max_poses, int32_t *n_poses, int32_t *inlier_keypoints, int32_t max_inlier_total) {
  TOD_REQUIRE(n_kp >= 0, "bad sizes");
  const int32_t off[2] = {0, n_kp};
  return tod_guess_process_batch(g, 1, off, keypoints, cloud, height, width, matches, counts, k, points3d, spans,
                                 n_objects, poses, nullptr, max_poses, n_poses, inlier_keypoints, max_inlier_total);
}

}  // extern "C"
