// The clique gate's bounded search for graphs of at most 256 vertices, written once for the host and the device.
//
// Same contract as CliqueFinder::finds_more_than(7) (clique.h), i.e. "would tod::maximum_clique::Graph::FindClique
// (src/common/maximum_clique.cpp:343-369 of the reference) called with minimal_size = 7 return MORE than 7 vertices?"
// (sac_model_registration_graph.h:258-265), stepping exactly like the reference: DegreeSort (:263-284), ColorSort
// (:219-261) with its shared colour vector, the 0.025 re-sort rule (:313), early exit at the first clique of 7
// (:290, :325).  Only the SIZES of the incumbent and of the current clique are tracked — that is all the gate asks.
// Fixed-capacity state (no allocation): vertex sets are masks of 64, 128 or 256 bits (one instantiation each), the
// per-level candidate lists are byte arrays.
// The 100000-step budget of the reference is replaced by a caller-given cap far below it: a search that reaches the
// cap returns -1 and is decided by the host's CliqueFinder instead.
#ifndef TOD_CLIQUE_SMALL_H_
#define TOD_CLIQUE_SMALL_H_

#include <cstdint>

#if defined(__CUDACC__)
#define TOD_HD __host__ __device__ __forceinline__
#else
#define TOD_HD inline
#endif

namespace tod {

// A vertex set of up to 64 * NW vertices.
template <int NW>
struct BitsN {
  unsigned long long w[NW];
};
typedef BitsN<1> Bits64;
typedef BitsN<2> Bits128;
typedef BitsN<4> Bits256;

constexpr int kSmallGraphMax = 256;
constexpr int kSmallGateMinimal = 7;  // std::min(best_inlier_number_, 7) with best_inlier_number_ >= 8 (:85, :203)

namespace small_clique {

TOD_HD int popc64(unsigned long long x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
TOD_HD int clz64(unsigned long long x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}

template <int NW>
TOD_HD BitsN<NW> empty_set() {
  BitsN<NW> s;
#pragma unroll
  for (int i = 0; i < NW; ++i) s.w[i] = 0ull;
  return s;
}
template <int NW>
TOD_HD bool test_bit(const BitsN<NW> &s, int v) {
  if (NW == 1) return ((s.w[0] >> v) & 1ull) != 0;
  return ((s.w[v >> 6] >> (v & 63)) & 1ull) != 0;
}
template <int NW>
TOD_HD void set_bit(BitsN<NW> &s, int v) {
  if (NW == 1) s.w[0] |= 1ull << v;
  else s.w[v >> 6] |= 1ull << (v & 63);
}
template <int NW>
TOD_HD int and_popc(const BitsN<NW> &a, const BitsN<NW> &b) {
  int c = 0;
#pragma unroll
  for (int i = 0; i < NW; ++i) c += popc64(a.w[i] & b.w[i]);
  return c;
}
template <int NW>
TOD_HD bool intersects(const BitsN<NW> &a, const BitsN<NW> &b) {
  unsigned long long x = 0ull;
#pragma unroll
  for (int i = 0; i < NW; ++i) x |= a.w[i] & b.w[i];
  return x != 0ull;
}
// highest member, removed from the set (-1 when empty)
template <int NW>
TOD_HD int pop_highest(BitsN<NW> &s) {
#pragma unroll
  for (int i = NW - 1; i >= 0; --i)
    if (s.w[i]) {
      const int b = 63 - clz64(s.w[i]);
      s.w[i] &= ~(1ull << b);
      return 64 * i + b;
    }
  return -1;
}

// Idx: unsigned char up to 128 vertices, unsigned short above (positions and class numbers reach the vertex count)
template <int NW, typename Idx>
struct State {
  static constexpr int kMaxN = 64 * NW;
  unsigned char lists[kSmallGateMinimal + 2][kMaxN];  // candidate list of every recursion level (1-based); ids < 256
  Idx colour[kMaxN];                                  // the colour vector shared by all levels
  unsigned char snapshot[kMaxN];
  Idx cls[kMaxN];
  Idx count[kMaxN + 2];
  BitsN<NW> class_mask[kMaxN + 5];   // classes are 1-based; four are tested per trip
  int size[kSmallGateMinimal + 2];
  unsigned level_steps[kSmallGateMinimal + 3], level_steps_old[kSmallGateMinimal + 3];
};

// DegreeSort: descending by (degree inside the list, vertex id) — std::sort of (degree, vertex) pairs read backwards.
// Returns the largest degree.
template <int NW, typename Idx>
TOD_HD int sort_by_degree(const BitsN<NW> *adj, unsigned char *r, int m, Idx *count) {
  BitsN<NW> mask = empty_set<NW>();
  for (int i = 0; i < m; ++i) set_bit(mask, r[i]);
  for (int d = 0; d <= m; ++d) count[d] = 0;
  int top = 0;
  for (int i = 0; i < m; ++i) {
    const int d = and_popc(adj[r[i]], mask);
    ++count[d];
    top = d > top ? d : top;
  }
  // count[d] <- first position of degree d in the descending order
  int pos = 0;
  for (int d = top; d >= 0; --d) {
    const int c = count[d];
    count[d] = (Idx)pos;
    pos += c;
  }
  // vertices by descending id; equal degrees keep that order
  BitsN<NW> walk = mask;
  for (int v = pop_highest(walk); v >= 0; v = pop_highest(walk)) r[count[and_popc(adj[v], mask)]++] = (unsigned char)v;
  return top;
}

// ColorSort with min_k == 1: every vertex, in list order, joins the first class that holds none of its neighbours;
// the list is rewritten class by class and colour[position] = class number.
template <int NW, typename Idx>
TOD_HD void colour_sort(const BitsN<NW> *adj, State<NW, Idx> &s, unsigned char *r, int m) {
  int n_classes = 0;
  for (int i = 0; i < m; ++i) {
    const int p = r[i];
    s.snapshot[i] = (unsigned char)p;
    const BitsN<NW> a = adj[p];
    // first class without a neighbour of p, four classes per trip: the four tests are independent loads, so a thread
    // that walks a 200-vertex list through ~25 classes is not serialised on one branch per class
    int k = 1;
    for (;; k += 4) {
      // (masks past n_classes hold stale words: read, then masked out — no short-circuit, the loads stay independent)
      const bool h0 = intersects(a, s.class_mask[k]) & (k <= n_classes);
      const bool h1 = intersects(a, s.class_mask[k + 1]) & (k + 1 <= n_classes);
      const bool h2 = intersects(a, s.class_mask[k + 2]) & (k + 2 <= n_classes);
      const bool h3 = intersects(a, s.class_mask[k + 3]) & (k + 3 <= n_classes);
      if (!h0) break;
      if (!h1) { k += 1; break; }
      if (!h2) { k += 2; break; }
      if (!h3) { k += 3; break; }
    }
    if (k > n_classes) {
      n_classes = k;
      s.class_mask[k] = empty_set<NW>();
      s.class_mask[k + 4] = empty_set<NW>();   // keeps every mask the four-wide test can read defined
      s.count[k] = 0;
    }
    set_bit(s.class_mask[k], p);
    s.cls[i] = (Idx)k;
    ++s.count[k];
  }
  int pos = 0;
  for (int k = 1; k <= n_classes; ++k) {
    const int c = s.count[k];
    s.count[k] = (Idx)pos;
    pos += c;
  }
  for (int i = 0; i < m; ++i) {
    const int k = s.cls[i];
    const int at = s.count[k]++;
    r[at] = s.snapshot[i];
    s.colour[at] = (Idx)k;
  }
}

}  // namespace small_clique

// adj: n rows of NW 64-bit words, bit j of row i <=> vertices i and j are adjacent (symmetric, no self-loops),
// n <= 64 * NW.
// Returns 1 (the search returns more than 7 vertices: the gate passes), 0 (it does not), -1 (step cap reached).
template <int NW, typename Idx>
TOD_HD int small_gate_search_t(const BitsN<NW> *adj, int n, int step_cap, int *steps_out) {
  using namespace small_clique;
  constexpr int kMinimal = kSmallGateMinimal;
  if (steps_out) *steps_out = 0;
  if (n <= 0) return 0;
  State<NW, Idx> s;
  for (int i = 0; i < kMinimal + 3; ++i) s.level_steps[i] = s.level_steps_old[i] = 0u;
  for (int i = 0; i < 5; ++i) s.class_mask[i] = small_clique::empty_set<NW>();   // read (and masked) before first use
  int steps = 1;
  int best = 0, cur = 0;
  unsigned char *order = s.lists[1];
  for (int i = 0; i < n; ++i) order[i] = (unsigned char)i;
  const int top = sort_by_degree(adj, order, n, s.count);
  for (int i = 0; i < n; ++i) s.colour[i] = (Idx)(i < top ? i + 1 : top + 1);
  long colour_size = n;
  // word in front of the colour array on a glibc heap (read by the reference when the shared vector underflows)
  unsigned long chunk = (4ul * (unsigned long)n + 8ul + 15ul) & ~15ul;
  if (chunk < 32ul) chunk = 32ul;
  int level = 1;
  s.size[1] = n;
  s.level_steps[1] = s.level_steps[1] + s.level_steps[0] - s.level_steps_old[1];
  s.level_steps_old[1] = s.level_steps[0];
  int result = 0;
  for (;;) {
    bool returned = true;   // leave the level unless the loop below descends
    while (s.size[level] > 0) {
      unsigned char *r = s.lists[level];
      const int m = s.size[level];
      const int p = r[m - 1];
      const unsigned long c = colour_size > 0 ? s.colour[colour_size - 1] : (colour_size == -1 ? (chunk | 1ul) : 0ul);
      if ((unsigned long)cur + c > (unsigned long)best) {
        ++cur;
        if (cur > kMinimal) {  // a clique of 8: from here the reference only descends to a leaf of at least this size
          result = 1;
          goto done;
        }
        unsigned char *next = s.lists[level + 1];
        const BitsN<NW> a = adj[p];
        int mn = 0;
        for (int i = 0; i < m; ++i) {
          const int v = r[i];
          if (test_bit(a, v)) next[mn++] = (unsigned char)v;
        }
        if (mn > 0) {
          if (double(s.level_steps[level]) / double(steps) < 0.025) sort_by_degree(adj, next, mn, s.count);
          if (best - cur + 1 > 1) s.colour[mn - 1] = 0;  // ColorSort with min_k > 1 keeps the order (class 1 stays empty)
          else colour_sort(adj, s, next, mn);
          ++s.level_steps[level];
          ++steps;
          if (steps > step_cap) {
            result = -1;
            goto done;
          }
          ++level;
          s.size[level] = mn;
          s.level_steps[level] = s.level_steps[level] + s.level_steps[level - 1] - s.level_steps_old[level];
          s.level_steps_old[level] = s.level_steps[level - 1];
          returned = false;
          break;
        }
        if (cur > best) {
          best = cur;
          if (best >= kMinimal) {  // the search stops at its first clique of 7: exactly 7 here, the gate fails
            result = 0;
            goto done;
          }
        }
        --cur;
      } else {
        break;
      }
      --s.size[level];
      --colour_size;
    }
    if (!returned) continue;
    if (level == 1) break;
    // back in the parent, behind its call of expand(): current_.pop_back(); r.pop_back(); --colour_size_
    --level;
    --cur;
    --s.size[level];
    --colour_size;
  }
done:
  if (steps_out) *steps_out = steps;
  return result;
}

TOD_HD int small_gate_search(const Bits64 *adj, int n, int step_cap, int *steps_out) {
  return small_gate_search_t<1, unsigned char>(adj, n, step_cap, steps_out);
}
TOD_HD int small_gate_search(const Bits128 *adj, int n, int step_cap, int *steps_out) {
  return small_gate_search_t<2, unsigned char>(adj, n, step_cap, steps_out);
}
TOD_HD int small_gate_search(const Bits256 *adj, int n, int step_cap, int *steps_out) {
  return small_gate_search_t<4, unsigned short>(adj, n, step_cap, steps_out);
}

}  // namespace tod
#endif
