// K5, warp-cooperative form: the same search as clique_small.h (small_gate_search_t — the reference's MaxCliqueDyn
// stepped exactly, maximum_clique.cpp:219-369), one WARP per hypothesis, for induced graphs of 65..256 vertices.
// A single thread takes 1-2 ms for a 200-vertex graph (dependent local-memory loads), and a launch waits for its slowest
// search; here the lanes share the work of every step:
//   filter (Intersection)   lanes test 32 list entries at a time, ballot-compacted in list order
//   DegreeSort              lanes count degrees of different vertices; counting sort by degree with the vertices of
//                           equal degree placed in descending id order (match_any groups, leader bumps the counter)
//   ColorSort               vertices stay sequential (the greedy rule is), but the search for the first class without
//                           a neighbour tests 32 classes at a time (one per lane, ballot + ffs)
// State lives in shared memory (one WarpState per warp).  Control flow is warp-uniform; every decision is taken from
// shared memory or from warp-wide reductions, so all lanes agree.  Equality with the thread form / the host
// compilation / the compiled reference: tests/test_k5_gpu.py.
#ifndef TOD_K5_WARP_CUH_
#define TOD_K5_WARP_CUH_

#include "clique_small.h"

namespace tod {
namespace k5w {

constexpr unsigned kFull = 0xffffffffu;

template <int NW>
struct WarpState {
  static constexpr int kMaxN = 64 * NW;
  BitsN<NW> adj[kMaxN];
  BitsN<NW> class_mask[kMaxN + 2];   // 1-based
  unsigned int count[kMaxN + 2];     // histogram, then running positions
  unsigned short colour[kMaxN];      // the colour vector shared by all levels
  unsigned short cls[kMaxN];
  unsigned char lists[kSmallGateMinimal + 2][kMaxN];
  unsigned char snapshot[kMaxN];
  unsigned char deg_of[kMaxN];       // by vertex id
  int size[kSmallGateMinimal + 2];
  unsigned level_steps[kSmallGateMinimal + 3], level_steps_old[kSmallGateMinimal + 3];
};

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long x) {
  const unsigned lo = __reduce_or_sync(kFull, unsigned(x));
  const unsigned hi = __reduce_or_sync(kFull, unsigned(x >> 32));
  return (static_cast<unsigned long long>(hi) << 32) | lo;
}

// count[lo..hi] (a histogram) -> first output position of every bin, bins taken from `hi` DOWN to `lo` when
// descending, from `lo` up otherwise.  Warp-wide; count[] must be visible (caller syncs before), synced after.
__device__ __forceinline__ void warp_positions(unsigned int *count, int lo, int hi, bool descending, int lane) {
  const int n = hi - lo + 1;
  const int per = (n + 31) / 32;
  unsigned local = 0;
  for (int j = 0; j < per; ++j) {
    const int i = lane * per + j;
    if (i < n) local += count[descending ? hi - i : lo + i];
  }
  unsigned incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += t;
  }
  unsigned run = incl - local;
  __syncwarp();
  for (int j = 0; j < per; ++j) {
    const int i = lane * per + j;
    if (i < n) {
      const int b = descending ? hi - i : lo + i;
      const unsigned c = count[b];
      count[b] = run;
      run += c;
    }
  }
  __syncwarp();
}

// DegreeSort (descending by (degree inside the list, vertex id)); returns the largest degree.
template <int NW>
__device__ __forceinline__ int sort_by_degree(WarpState<NW> &S, unsigned char *r, int m, int lane) {
  using namespace small_clique;
  BitsN<NW> mine = empty_set<NW>();
  for (int i = lane; i < m; i += 32) set_bit(mine, r[i]);
  BitsN<NW> mask;
#pragma unroll
  for (int w = 0; w < NW; ++w) mask.w[w] = warp_or64(mine.w[w]);
  for (int d = lane; d <= m; d += 32) S.count[d] = 0u;
  __syncwarp();
  int top = 0;
  for (int i = lane; i < m; i += 32) {
    const int v = r[i];
    const int d = and_popc(S.adj[v], mask);
    S.deg_of[v] = static_cast<unsigned char>(d);
    atomicAdd(&S.count[d], 1u);
    top = max(top, d);
  }
  top = __reduce_max_sync(kFull, top);
  __syncwarp();
  warp_positions(S.count, 0, top, true, lane);
  // vertices by descending id, 32 ids at a time (lane 0 = the highest id of the group); equal degrees keep that order
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int w = NW - 1; w >= 0; --w) {
    for (int half = 1; half >= 0; --half) {
      const unsigned word = unsigned(mask.w[w] >> (32 * half));
      if (word == 0u) continue;
      const int id = w * 64 + half * 32 + 31 - lane;
      const bool active = (word >> (31 - lane)) & 1u;
      const unsigned key = active ? unsigned(S.deg_of[id]) : 0x1000u + unsigned(lane);
      const unsigned peers = __match_any_sync(kFull, key);
      unsigned base = 0;
      if (active) base = S.count[key];
      __syncwarp();
      if (active) {
        r[base + __popc(peers & lt)] = static_cast<unsigned char>(id);
        if ((peers & lt) == 0u) S.count[key] = base + unsigned(__popc(peers));
      }
      __syncwarp();
    }
  }
  return top;
}

// ColorSort with min_k == 1.
template <int NW>
__device__ __forceinline__ void colour_sort(WarpState<NW> &S, unsigned char *r, int m, int lane) {
  using namespace small_clique;
  int n_classes = 0;
  for (int i = 0; i < m; ++i) {
    const int p = r[i];
    const BitsN<NW> a = S.adj[p];
    int k = 0;
    for (int j = 0; j * 32 < n_classes; ++j) {
      const int kk = j * 32 + lane + 1;
      const bool is_free = kk <= n_classes && !intersects(a, S.class_mask[kk]);
      const unsigned fm = __ballot_sync(kFull, is_free);
      if (fm) {
        k = j * 32 + __ffs(int(fm));
        break;
      }
    }
    if (k == 0) {
      k = ++n_classes;
      if (lane == 0) {
        S.class_mask[k] = empty_set<NW>();
        S.count[k] = 0u;
      }
      __syncwarp();
    }
    if (lane == 0) {
      S.snapshot[i] = static_cast<unsigned char>(p);
      set_bit(S.class_mask[k], p);
      S.cls[i] = static_cast<unsigned short>(k);
      S.count[k] += 1u;
    }
    __syncwarp();
  }
  warp_positions(S.count, 1, n_classes, false, lane);
  const unsigned lt = (1u << lane) - 1u;
  for (int b0 = 0; b0 < m; b0 += 32) {
    const int i = b0 + lane;
    const bool active = i < m;
    const unsigned key = active ? unsigned(S.cls[i]) : 0x1000u + unsigned(lane);
    const unsigned peers = __match_any_sync(kFull, key);
    unsigned base = 0;
    if (active) base = S.count[key];
    __syncwarp();
    if (active) {
      const unsigned at = base + unsigned(__popc(peers & lt));
      r[at] = S.snapshot[i];
      S.colour[at] = static_cast<unsigned short>(key);
      if ((peers & lt) == 0u) S.count[key] = base + unsigned(__popc(peers));
    }
    __syncwarp();
  }
}

// rows: n rows of NW words in global memory (the job pool).  Returns 1 / 0 / -1 like small_gate_search_t.
template <int NW>
__device__ int warp_gate_search(WarpState<NW> &S, const unsigned long long *__restrict__ rows, int n, int step_cap,
                                int lane) {
  using namespace small_clique;
  constexpr int kMinimal = kSmallGateMinimal;
  if (n <= 0) return 0;
  for (int i = lane; i < n * NW; i += 32) S.adj[i / NW].w[i % NW] = rows[i];
  if (lane < kMinimal + 3) S.level_steps[lane] = S.level_steps_old[lane] = 0u;
  unsigned char *order = S.lists[1];
  for (int i = lane; i < n; i += 32) order[i] = static_cast<unsigned char>(i);
  __syncwarp();
  const int top = sort_by_degree(S, order, n, lane);
  for (int i = lane; i < n; i += 32) S.colour[i] = static_cast<unsigned short>(i < top ? i + 1 : top + 1);
  long colour_size = n;
  unsigned long chunk = (4ul * static_cast<unsigned long>(n) + 8ul + 15ul) & ~15ul;
  if (chunk < 32ul) chunk = 32ul;
  int steps = 1, best = 0, cur = 0, level = 1;
  if (lane == 0) {
    S.size[1] = n;
    S.level_steps[1] = S.level_steps[1] + S.level_steps[0] - S.level_steps_old[1];
    S.level_steps_old[1] = S.level_steps[0];
  }
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  for (;;) {
    bool returned = true;
    while (S.size[level] > 0) {
      unsigned char *r = S.lists[level];
      const int m = S.size[level];
      const int p = r[m - 1];
      const unsigned long c = colour_size > 0 ? S.colour[colour_size - 1] : (colour_size == -1 ? (chunk | 1ul) : 0ul);
      if (static_cast<unsigned long>(cur) + c > static_cast<unsigned long>(best)) {
        ++cur;
        if (cur > kMinimal) return 1;
        unsigned char *next = S.lists[level + 1];
        const BitsN<NW> a = S.adj[p];
        int mn = 0;
        for (int b0 = 0; b0 < m; b0 += 32) {
          const int i = b0 + lane;
          const int v = i < m ? r[i] : 0;
          const bool keep = i < m && test_bit(a, v);
          const unsigned bal = __ballot_sync(kFull, keep);
          if (keep) next[mn + __popc(bal & lt)] = static_cast<unsigned char>(v);
          mn += __popc(bal);
        }
        __syncwarp();
        if (mn > 0) {
          if (double(S.level_steps[level]) / double(steps) < 0.025) sort_by_degree(S, next, mn, lane);
          if (best - cur + 1 > 1) {
            if (lane == 0) S.colour[mn - 1] = 0;
          } else {
            colour_sort(S, next, mn, lane);
          }
          ++steps;
          if (steps > step_cap) return -1;
          __syncwarp();
          if (lane == 0) {
            ++S.level_steps[level];
            S.size[level + 1] = mn;
            S.level_steps[level + 1] = S.level_steps[level + 1] + S.level_steps[level] - S.level_steps_old[level + 1];
            S.level_steps_old[level + 1] = S.level_steps[level];
          }
          ++level;
          __syncwarp();
          returned = false;
          break;
        }
        if (cur > best) {
          best = cur;
          if (best >= kMinimal) return 0;
        }
        --cur;
      } else {
        break;
      }
      __syncwarp();
      if (lane == 0) --S.size[level];
      --colour_size;
      __syncwarp();
    }
    if (!returned) continue;
    if (level == 1) break;
    --level;
    --cur;
    __syncwarp();
    if (lane == 0) --S.size[level];
    --colour_size;
    __syncwarp();
  }
  return 0;
}

}  // namespace k5w
}  // namespace tod
#endif
