// Host-side clique gate: bounded Konc–Janezic MaxCliqueDyn search with the reference's exact stopping rules.
//
// Stays on the host by design (north_star: "only the branchy clique search and the final refinement left on the host").
// Behavioural contract = tod::maximum_clique::Graph::FindClique (src/common/maximum_clique.cpp:343-369 of the
// reference) with its helpers DegreeSort (:263-284), ColorSort (:219-261), Intersection (:209-217) and MaxCliqueDyn
// (:286-336): early exit as soon as a clique of `minimal_size` exists (:290, :325), step budget of 100000 (:318),
// re-sort by degree while the per-level step ratio is under 0.025 (:313).  Data structure is a dense bit-matrix
// instead of sorted neighbour lists.
//
// One reference quirk is reproduced on purpose: the colour vector is shared by all recursion levels (it is passed by
// reference and popped by every level).  When it runs empty the reference reads past the front of the array; this
// implementation returns the values a glibc heap yields there (0, then the chunk size word) — see DESIGN.md.
#ifndef TOD_CLIQUE_H_
#define TOD_CLIQUE_H_

#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

#include "bitops.h"

namespace tod {

class CliqueFinder {
 public:
  // adjacency is kept as n rows of 64-bit words
  explicit CliqueFinder(int n_vertices)
      : n_(n_vertices), id_space_(n_vertices), words_((n_vertices + 63) / 64),
        bits_(size_t(n_vertices) * size_t((n_vertices + 63) / 64), 0ull) {
    rows_ = bits_.data();
  }

  // from a dense symmetric n x ceil(n/32) bit-matrix of 32-bit words (no self-loops)
  CliqueFinder(int n_vertices, const uint32_t *adjacency) : CliqueFinder(n_vertices) {
    const int w32 = (n_ + 31) / 32;
    for (int v = 0; v < n_; ++v) {
      const uint32_t *src = adjacency + size_t(v) * w32;
      uint64_t *dst = bits_.data() + size_t(v) * words_;
      for (int w = 0; w < w32; ++w) dst[w >> 1] |= uint64_t(src[w]) << (32 * (w & 1));
    }
  }

  // VIEW of a larger graph: the sub-graph induced by `vertices` (ascending ids) of a graph whose rows (`words64`
  // 64-bit words each, id_space rows) stay where they are — no induced copy is built.  Every set operation below is
  // masked by the current vertex set, and ids keep the order of their ranks, so the search steps exactly as it would
  // on the renumbered induced graph (the reference builds that one, sac_model_registration_graph.h:241-255).
  CliqueFinder(int id_space, int words64, const uint64_t *rows, const uint32_t *vertices, int n_vertices)
      : n_(n_vertices), id_space_(id_space), words_(words64), rows_(rows) {
    view_.assign(vertices, vertices + n_vertices);
  }

  void add_edge(int a, int b) {  // owned storage only
    if (a == b || connected(a, b)) return;
    bits_[size_t(a) * words_ + (b >> 6)] |= 1ull << (b & 63);
    bits_[size_t(b) * words_ + (a >> 6)] |= 1ull << (a & 63);
  }

  int steps() const { return steps_; }  // search steps of the last find()

  bool connected(int a, int b) const { return (rows_[size_t(a) * words_ + (b >> 6)] >> (b & 63)) & 1ull; }

  // The gate's question (sac_model_registration_graph.h:260-265): would find(minimal_size) return MORE than
  // minimal_size vertices?  Same search, same answer, but it stops as soon as the answer is certain.
  bool finds_more_than(unsigned minimal_size) {
    decide_only_ = true;
    decided_ = false;
    const bool yes = find(minimal_size).size() > minimal_size;
    decide_only_ = false;
    return yes;
  }

  // Returns the clique found: the first one reaching `minimal_size`, else the largest seen within the step budget.
  std::vector<int> find(unsigned minimal_size) {
    best_.clear();
    if (n_ == 0) return best_;
    minimal_ = minimal_size;
    steps_ = 1;
    ratio_limit_ = 0.025;
    std::vector<int> order(static_cast<size_t>(n_));
    if (view_.empty())
      for (int i = 0; i < n_; ++i) order[size_t(i)] = i;
    else
      for (int i = 0; i < n_; ++i) order[size_t(i)] = int(view_[size_t(i)]);
    sort_by_degree(order);  // leaves the largest degree and the degree sum behind
    dense_ = 2ull * last_degree_sum_ > (unsigned long long)(n_) * (unsigned long long)(n_ - 1);
    const unsigned top = last_top_;
    colour_.assign(size_t(n_), 0u);
    for (unsigned i = 0; i < top && i < unsigned(n_); ++i) colour_[i] = i + 1;
    for (unsigned i = top; i < unsigned(n_); ++i) colour_[i] = top + 1;
    colour_size_ = n_;
    level_steps_.assign(size_t(n_) + 1, 0u);
    level_steps_old_.assign(size_t(n_) + 1, 0u);
    current_.clear();
    expand(order, 1);
    return best_;
  }

 private:
  // descending by (degree inside `r`, vertex id): what std::sort of (degree, vertex) pairs read backwards gives the
  // reference (DegreeSort, maximum_clique.cpp:263-284) — here a counting sort: degrees are below |r|, and walking the
  // set's bit mask from the top yields the vertices of equal degree in descending id order.
  // degree inside r = popcount(row & mask of r) — the same numbers the reference gets from pairwise tests.
  // Leaves the largest degree in last_top_ and the degree sum in last_degree_sum_.
  void sort_by_degree(std::vector<int> &r) {
    const size_t m = r.size();
    set_mask_.assign(size_t(words_), 0ull);
    for (int v : r) set_mask_[size_t(v) >> 6] |= 1ull << (v & 63);
    if (deg_of_.size() < size_t(id_space_)) deg_of_.resize(size_t(id_space_));
    if (deg_count_.size() < m + 2) deg_count_.resize(m + 2);
    std::fill(deg_count_.begin(), deg_count_.begin() + long(m) + 1, 0u);
    unsigned top = 0;
    unsigned long long sum = 0;
    for (size_t i = 0; i < m; ++i) {
      const uint64_t *row = rows_ + size_t(r[i]) * words_;
      const unsigned d = unsigned(bitops::and_popcount(row, set_mask_.data(), words_));
      deg_of_[size_t(r[i])] = d;
      ++deg_count_[d];
      top = std::max(top, d);
      sum += d;
    }
    last_top_ = top;
    last_degree_sum_ = sum;
    // deg_count_[d] <- first position of degree d in the descending order
    unsigned pos = 0;
    for (long d = long(top); d >= 0; --d) {
      const unsigned c = deg_count_[size_t(d)];
      deg_count_[size_t(d)] = pos;
      pos += c;
    }
    for (int w = words_ - 1; w >= 0; --w) {
      uint64_t x = set_mask_[size_t(w)];
      while (x) {
        const int b = 63 - __builtin_clzll(x);
        x &= ~(1ull << b);
        const int v = w * 64 + b;
        r[deg_count_[deg_of_[size_t(v)]]++] = v;
      }
    }
  }

  // greedy sequential colouring; vertices whose colour cannot extend the incumbent go first with colour 0.
  // "first class without a neighbour of p" is found from p's side — by walking p's already-coloured neighbours (sparse
  // graphs) or non-neighbours (dense graphs: the inlier graphs of the gate are ~85 % dense) — instead of scanning
  // class after class; it is the same class the reference's sequential scan picks.
  void colour_sort(std::vector<int> &r) {
    const int gap = int(best_.size()) - int(current_.size()) + 1;
    const unsigned min_k = unsigned(std::max(1, gap));
    if (dense_) {
      colour_sort_dense(r, min_k);
      return;
    }
    size_t n_classes = 2;
    if (classes_.size() < 2) classes_.resize(2);
    classes_[0].clear();
    classes_[1].clear();
    set_mask_.assign(size_t(words_), 0ull);          // vertices already pushed into a class
    if (class_of_.size() < size_t(id_space_)) class_of_.resize(size_t(id_space_));
    if (used_.size() < r.size() + 3) used_.resize(r.size() + 3, 0u);
    size_t keep = 0;
    snapshot_.assign(r.begin(), r.end());
    for (int p : snapshot_) {
      const uint64_t *row = rows_ + size_t(p) * words_;
      if (++stamp_ == 0u) {  // wrapped: start over with clean stamps
        std::fill(used_.begin(), used_.end(), 0u);
        stamp_ = 1u;
      }
      // walk p's already-coloured NEIGHBOURS and stamp their classes; first unstamped class wins
      for (int w = 0; w < words_; ++w) {
        uint64_t m = row[w] & set_mask_[size_t(w)];
        while (m) {
          const int v = w * 64 + __builtin_ctzll(m);
          m &= m - 1;
          used_[size_t(class_of_[size_t(v)])] = stamp_;
        }
      }
      size_t k = 1;
      while (used_[k] == stamp_) {
        ++k;
        if (k >= n_classes) break;
      }
      if (k >= n_classes) {  // every existing class holds a neighbour of p: open a new one
        k = n_classes;
        ++n_classes;
        if (classes_.size() < n_classes) classes_.resize(n_classes);
        classes_[n_classes - 1].clear();
      }
      if (k < min_k) {
        r[keep++] = p;
      } else {
        classes_[k].push_back(p);
        class_of_[size_t(p)] = int(k);
        set_mask_[size_t(p) >> 6] |= 1ull << (p & 63);
      }
    }
    if (keep > 0) colour_[keep - 1] = 0;
    size_t pos = keep;
    for (size_t k = min_k; k < n_classes; ++k)
      for (int v : classes_[k]) {
        r[pos] = v;
        colour_[pos] = unsigned(k);
        ++pos;
      }
  }

  // The same colouring for dense graphs (the inlier graphs of the gate are ~85 % dense), built CLASS BY CLASS on bit
  // sets: class k = the uncoloured vertices, taken in the order of r, that have no neighbour among the members chosen
  // before them.  That is exactly the class the sequential rule ("smallest class without a neighbour of p") gives
  // every vertex: p lands in class k iff each earlier class already held a neighbour of p when p was reached, and
  // class k did not.  One AND-NOT over the row per vertex instead of a walk over its coloured non-neighbours.
  void colour_sort_dense(std::vector<int> &r, unsigned min_k) {
    const size_t m = r.size();
    if (m == 0) return;
    if (min_k > 1) {
      // classes below min_k never receive members (their vertices are only kept in front, ColorSort :245-248), so
      // class 1 stays empty, every vertex gets k = 1 < min_k and R keeps its order
      colour_[m - 1] = 0;
      return;
    }
    if (rank_.size() < size_t(id_space_)) rank_.resize(size_t(id_space_));
    snapshot_.assign(r.begin(), r.end());
    set_mask_.assign(size_t(words_), 0ull);  // uncoloured vertices
    for (size_t i = 0; i < m; ++i) {
      const int v = snapshot_[i];
      rank_[size_t(v)] = int(i);
      set_mask_[size_t(v) >> 6] |= 1ull << (v & 63);
    }
    cand_.resize(size_t(words_));
    size_t first = 0, pos = 0;
    unsigned k = 0;
    while (pos < m) {
      ++k;
      while (!((set_mask_[size_t(snapshot_[first]) >> 6] >> (snapshot_[first] & 63)) & 1ull)) ++first;
      int v = snapshot_[first];
      size_t at = first;  // position of the last member in the order
      bool more = true;
      for (bool first_member = true; more; first_member = false) {
        // take v: out of the uncoloured set, into the class; the candidates lose v's neighbours
        set_mask_[size_t(v) >> 6] &= ~(1ull << (v & 63));
        r[pos] = v;
        colour_[pos] = k;
        ++pos;
        const uint64_t *row = rows_ + size_t(v) * words_;
        uint64_t any = 0;
        if (first_member) {
          any = bitops::andnot_store_any(cand_.data(), set_mask_.data(), row, words_);
        } else {
          cand_[size_t(v) >> 6] &= ~(1ull << (v & 63));
          any = bitops::andnot_store_any(cand_.data(), cand_.data(), row, words_);
        }
        if (!any) break;
        // next member = the candidate that comes first in the order of r; it lies behind the last member.  Look a few
        // positions ahead (a hit is likely while the candidate set is large), else take the minimum rank directly.
        // (Choosing between the two by the size of the candidate set was measured and is slower: 0.145 against 0.095 ms
        // on a 630-vertex graph.)
        more = false;
        const size_t lim = std::min(m, at + 1 + 12);
        for (size_t i = at + 1; i < lim; ++i) {
          const int u = snapshot_[i];
          if ((cand_[size_t(u) >> 6] >> (u & 63)) & 1ull) {
            v = u;
            at = i;
            more = true;
            break;
          }
        }
        if (!more) {
          int best_rank = int(m);
          for (int w = 0; w < words_; ++w) {
            uint64_t x = cand_[size_t(w)];
            while (x) {
              const int u = w * 64 + __builtin_ctzll(x);
              x &= x - 1;
              if (rank_[size_t(u)] < best_rank) best_rank = rank_[size_t(u)];
            }
          }
          v = snapshot_[size_t(best_rank)];
          at = size_t(best_rank);
          more = true;
        }
      }
    }
  }

  unsigned colour_back() const {
    if (colour_size_ > 0) return colour_[size_t(colour_size_ - 1)];
    if (colour_size_ == -1) {  // word in front of the array on a glibc heap: chunk size | PREV_INUSE
      const unsigned long chunk = std::max(32ul, (4ul * unsigned(n_) + 8ul + 15ul) & ~15ul);
      return unsigned(chunk | 1ul);
    }
    return 0u;
  }

  void expand(std::vector<int> &r, size_t level) {
    if (best_.size() >= minimal_) return;
    if (level >= level_steps_.size()) {
      level_steps_.push_back(0u);
      level_steps_old_.push_back(0u);
    }
    level_steps_[level] = level_steps_[level] + level_steps_[level - 1] - level_steps_old_[level];
    level_steps_old_[level] = level_steps_[level - 1];
    while (!r.empty()) {
      const int p = r.back();
      const unsigned c = colour_back();
      if (current_.size() + c > best_.size()) {
        current_.push_back(p);
        // Decision-only mode: current_ is a clique of minimal_ + 1 vertices and the search is still alive (best_ <
        // minimal_), so from here it can only descend — size + colour > best_ holds at every deeper level — to a leaf
        // of at least this size, unless the 100000-step budget ran out first; the descent is at most |r| levels deep.
        if (decide_only_ && current_.size() > minimal_ && steps_ + int(r.size()) + 1 <= 100000) {
          best_ = current_;
          decided_ = true;
          return;
        }
        std::vector<int> next;
        for (int v : r)
          if (connected(p, v)) next.push_back(v);
        if (!next.empty()) {
          if (double(level_steps_[level]) / steps_ < ratio_limit_) sort_by_degree(next);
          colour_sort(next);
          ++level_steps_[level];
          ++steps_;
          if (steps_ > 100000) return;
          expand(next, level + 1);
          if (decided_) return;
        } else if (current_.size() > best_.size()) {
          best_ = current_;
          if (best_.size() >= minimal_) return;
        }
        current_.pop_back();
      } else {
        return;
      }
      r.pop_back();
      --colour_size_;
    }
  }

  bool decide_only_ = false, decided_ = false;
  int n_;         // vertices of the graph being searched
  int id_space_;  // vertex ids are < id_space_ (== n_ unless this is a view)
  int words_;
  const uint64_t *rows_ = nullptr;
  std::vector<uint32_t> view_;
  std::vector<uint64_t> set_mask_;
  std::vector<uint32_t> used_;
  std::vector<int> rank_;
  std::vector<uint64_t> cand_;
  bool dense_ = false;  // more than half of all vertex pairs are edges: colour_sort walks non-neighbours
  std::vector<int> class_of_;
  uint32_t stamp_ = 0u;
  std::vector<unsigned> deg_of_, deg_count_;
  unsigned last_top_ = 0;
  unsigned long long last_degree_sum_ = 0;
  std::vector<std::vector<int> > classes_;
  std::vector<int> snapshot_;
  std::vector<uint64_t> bits_;
  std::vector<unsigned> colour_;
  long colour_size_ = 0;
  std::vector<unsigned> level_steps_, level_steps_old_;
  std::vector<int> current_, best_;
  unsigned minimal_ = 0;
  int steps_ = 1;
  double ratio_limit_ = 0.025;
};

}  // namespace tod
#endif
