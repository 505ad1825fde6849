// Stage-level C-ABI entry points of the geometry half (include/tod_b200.h: tod_fill_adjacency,
// tod_score_hypotheses): host buffers in, host buffers out, compute in K2 / K3 on the GPU.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "clique.h"
#include "clique_small.h"
#include "host_geometry.h"
#include "tod_internal.h"

using tod::DeviceBuffer;
using tod::fail;

namespace {

struct Scratch {  // freed on scope exit
  std::vector<DeviceBuffer *> bufs;
  ~Scratch() {
    for (DeviceBuffer *b : bufs) b->release();
  }
  void own(DeviceBuffer &b) { bufs.push_back(&b); }
};

thread_local float g_last_stage_ms = -1.f;

// CUDA events around one kernel launch on `st`; the elapsed time lands in g_last_stage_ms once the stream is synced.
struct StageTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st;
  explicit StageTimer(cudaStream_t s) : st(s) {
    g_last_stage_ms = -1.f;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) e0 = e1 = nullptr;
  }
  void begin() { if (e0) cudaEventRecord(e0, st); }
  void end() { if (e1) cudaEventRecord(e1, st); }
  void finish() {
    if (e0 && e1 && cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&g_last_stage_ms, e0, e1);
  }
  ~StageTimer() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

int check_device(int device) {
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return fail(TOD_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  TOD_REQUIRE(device >= 0 && device < n_dev, "device %d out of range (%d devices)", device, n_dev);
  TOD_CUDA(cudaSetDevice(device));
  return TOD_OK;
}

}  // namespace

extern "C" {

float tod_last_stage_ms(void) { return g_last_stage_ms; }

// ---- host-only pieces of the guess generator, exposed so that they can be checked without a GPU -------------------
int32_t tod_clique_find(int32_t n_vertices, const int32_t *edges, int32_t n_edges, uint32_t minimal_size,
                        int32_t *out_vertices, int32_t *finds_more) {
  if (n_vertices < 0 || n_edges < 0 || (n_edges > 0 && !edges)) {
    tod::set_error("bad graph");
    return -1;
  }
  for (int32_t e = 0; e < n_edges * 2; ++e)
    if (edges[e] < 0 || edges[e] >= n_vertices) {
      tod::set_error("edge endpoint %d out of range", edges[e]);
      return -1;
    }
  tod::CliqueFinder finder(n_vertices);
  for (int32_t e = 0; e < n_edges; ++e) finder.add_edge(edges[2 * e], edges[2 * e + 1]);
  const std::vector<int> best = finder.find(minimal_size);
  if (out_vertices)
    for (size_t i = 0; i < best.size(); ++i) out_vertices[i] = best[i];
  if (finds_more) {
    tod::CliqueFinder again(n_vertices);
    for (int32_t e = 0; e < n_edges; ++e) again.add_edge(edges[2 * e], edges[2 * e + 1]);
    *finds_more = again.finds_more_than(minimal_size) ? 1 : 0;
  }
  return int32_t(best.size());
}

int32_t tod_clique_gate_small(int32_t n_vertices, const int32_t *edges, int32_t n_edges, int32_t step_cap,
                              int32_t *steps) {
  if (n_vertices < 0 || n_vertices > tod::kSmallGraphMax || n_edges < 0 || (n_edges > 0 && !edges) || step_cap < 1) {
    tod::set_error("bad graph (at most %d vertices)", tod::kSmallGraphMax);
    return -2;
  }
  // widest form first; the narrower forms K5 uses for small graphs must agree with it
  std::vector<tod::Bits256> adj(size_t(std::max(n_vertices, 1)), tod::small_clique::empty_set<4>());
  for (int32_t e = 0; e < n_edges; ++e) {
    const int32_t a = edges[2 * e], b = edges[2 * e + 1];
    if (a < 0 || a >= n_vertices || b < 0 || b >= n_vertices) {
      tod::set_error("edge endpoint out of range");
      return -2;
    }
    if (a == b) continue;
    tod::small_clique::set_bit(adj[size_t(a)], b);
    tod::small_clique::set_bit(adj[size_t(b)], a);
  }
  int st = 0;
  const int r = tod::small_gate_search(adj.data(), n_vertices, step_cap, &st);
  if (n_vertices <= 128) {
    std::vector<tod::Bits128> a2(adj.size());
    for (size_t i = 0; i < adj.size(); ++i) a2[i].w[0] = adj[i].w[0], a2[i].w[1] = adj[i].w[1];
    int st2 = 0;
    if (tod::small_gate_search(a2.data(), n_vertices, step_cap, &st2) != r || st2 != st) {
      tod::set_error("the 128-bit and 256-bit forms of the search disagree");
      return -2;
    }
  }
  if (n_vertices <= 64) {
    std::vector<tod::Bits64> a1(adj.size());
    for (size_t i = 0; i < adj.size(); ++i) a1[i].w[0] = adj[i].w[0];
    int st1 = 0;
    if (tod::small_gate_search(a1.data(), n_vertices, step_cap, &st1) != r || st1 != st) {
      tod::set_error("the 64-bit and 256-bit forms of the search disagree");
      return -2;
    }
  }
  if (steps) *steps = st;
  return r;
}

// K5 on the device for a batch of graphs given as edge lists: the same kernel, job queue layout and step cap as inside
// tod_guess_process (where K4 fills the queue).  results[g] = 1 passes, 0 fails, -1 left to the host (step cap).
int tod_gate_search_device(int32_t device, int32_t n_graphs, const int32_t *n_vertices, const int32_t *edge_offsets,
                           const int32_t *edges, int32_t *results) {
  TOD_REQUIRE(n_graphs >= 0 && (n_graphs == 0 || (n_vertices && edge_offsets && results)), "bad argument");
  if (n_graphs == 0) return TOD_OK;
  if (int rc = check_device(device)) return rc;
  size_t pool_words = 0;
  for (int32_t g = 0; g < n_graphs; ++g) {
    TOD_REQUIRE(n_vertices[g] >= 1 && n_vertices[g] <= tod::kSmallGraphMax, "graph %d has %d vertices (1..%d)", g,
                n_vertices[g], tod::kSmallGraphMax);
    const int nw = n_vertices[g] <= 64 ? 1 : (n_vertices[g] <= 128 ? 2 : 4);
    pool_words += size_t(nw) * size_t(n_vertices[g]);
  }
  const size_t hdr_bytes = (2 * size_t(n_graphs) * 16 + 255) & ~size_t(255);
  std::vector<unsigned char> host(256 + hdr_bytes + pool_words * 8, 0);
  unsigned long long *ctl = reinterpret_cast<unsigned long long *>(host.data());
  int32_t *hdr = reinterpret_cast<int32_t *>(host.data() + 256);
  unsigned long long *pool = reinterpret_cast<unsigned long long *>(host.data() + 256 + hdr_bytes);
  size_t off = 0;
  for (int32_t g = 0; g < n_graphs; ++g) {
    const int n = n_vertices[g];
    const int nw = n <= 64 ? 1 : (n <= 128 ? 2 : 4);
    unsigned long long *rows = pool + off;
    for (int32_t e = edge_offsets[g]; e < edge_offsets[g + 1]; ++e) {
      const int32_t a = edges[2 * e], b = edges[2 * e + 1];
      TOD_REQUIRE(a >= 0 && a < n && b >= 0 && b < n, "edge endpoint out of range in graph %d", g);
      if (a == b) continue;
      rows[size_t(a) * nw + (b >> 6)] |= 1ull << (b & 63);
      rows[size_t(b) * nw + (a >> 6)] |= 1ull << (a & 63);
    }
    const size_t slot = nw == 1 ? size_t(ctl[0]++) : (nw == 2 ? size_t(n_graphs) - 1 - size_t(ctl[1]++)
                                                               : size_t(n_graphs) + size_t(ctl[3]++));
    int32_t *h = hdr + slot * 4;
    h[0] = g;
    h[1] = n;
    h[2] = int32_t(off & 0xffffffffull);
    h[3] = int32_t(off >> 32);
    off += size_t(nw) * size_t(n);
  }
  ctl[2] = off;
  void *d_jobs = nullptr;
  uint8_t *d_verdict = nullptr;
  TOD_CUDA(cudaMalloc(&d_jobs, host.size()));
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&d_verdict), size_t(n_graphs));
  if (e == cudaSuccess) e = cudaMemcpy(d_jobs, host.data(), host.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(d_verdict, tod::kGateNeedsHost, size_t(n_graphs));
  if (e == cudaSuccess) e = tod::launch_gate_search_jobs(n_graphs, d_jobs, d_verdict, nullptr);
  std::vector<uint8_t> v(static_cast<size_t>(n_graphs));
  if (e == cudaSuccess) e = cudaMemcpy(v.data(), d_verdict, size_t(n_graphs), cudaMemcpyDeviceToHost);
  cudaFree(d_jobs);
  if (d_verdict) cudaFree(d_verdict);
  if (e != cudaSuccess) return tod::fail(TOD_ERR_CUDA, "K5 launch failed: %s", cudaGetErrorString(e));
  for (int32_t g = 0; g < n_graphs; ++g)
    results[g] = v[size_t(g)] == tod::kGatePasses ? 1 : (v[size_t(g)] == tod::kGateFailsSearch ? 0 : -1);
  return TOD_OK;
}

int tod_rigid_fit(const float *query_pts, const float *train_pts, const uint32_t *indices, int32_t m, float *R, float *T) {
  TOD_REQUIRE(query_pts && train_pts && indices && R && T && m >= 1, "bad argument");
  tod::rigid_fit(query_pts, train_pts, indices, m, R, T);
  return TOD_OK;
}

int32_t tod_adjacency_row_words(int32_t n) { return n <= 0 ? 0 : tod::adjacency_row_words(n); }

int tod_fill_adjacency(int32_t device, int32_t n_clusters, const int32_t *offsets, const float *query_pts,
                       const float *train_pts, const float *pixels, const float *spans, float sensor_error,
                       uint32_t *physical, uint32_t *sample, int64_t *matrix_offsets) {
  TOD_REQUIRE(n_clusters >= 0 && offsets && spans, "bad cluster table");
  if (n_clusters == 0) return TOD_OK;
  const int64_t N = offsets[n_clusters];
  TOD_REQUIRE(offsets[0] == 0 && N >= 0, "offsets must start at 0 and be non-decreasing");
  TOD_REQUIRE(N == 0 || (query_pts && train_pts && pixels && physical && sample), "null point/matrix buffer");
  std::vector<int64_t> mo(size_t(n_clusters) + 1, 0);
  int max_cluster = 0;
  for (int c = 0; c < n_clusters; ++c) {
    const int n = offsets[c + 1] - offsets[c];
    TOD_REQUIRE(n >= 0, "offsets must be non-decreasing");
    max_cluster = std::max(max_cluster, n);
    mo[size_t(c) + 1] = mo[size_t(c)] + int64_t(n) * tod_adjacency_row_words(n);
  }
  if (matrix_offsets) std::memcpy(matrix_offsets, mo.data(), mo.size() * sizeof(int64_t));
  if (N == 0 || mo.back() == 0) return TOD_OK;
  if (int rc = check_device(device)) return rc;

  DeviceBuffer d_off, d_mo, d_q, d_t, d_px, d_sp, d_P, d_S;
  Scratch scratch;
  for (DeviceBuffer *b : {&d_off, &d_mo, &d_q, &d_t, &d_px, &d_sp, &d_P, &d_S}) scratch.own(*b);
  const size_t mat_bytes = size_t(mo.back()) * sizeof(uint32_t);
  TOD_CUDA(d_off.reserve((size_t(n_clusters) + 1) * sizeof(int32_t)));
  TOD_CUDA(d_mo.reserve(mo.size() * sizeof(int64_t)));
  TOD_CUDA(d_q.reserve(size_t(N) * 12));
  TOD_CUDA(d_t.reserve(size_t(N) * 12));
  TOD_CUDA(d_px.reserve(size_t(N) * 8));
  TOD_CUDA(d_sp.reserve(size_t(n_clusters) * 4));
  TOD_CUDA(d_P.reserve(mat_bytes));
  TOD_CUDA(d_S.reserve(mat_bytes));
  cudaStream_t st = nullptr;
  TOD_CUDA(cudaMemcpyAsync(d_off.ptr, offsets, (size_t(n_clusters) + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_mo.ptr, mo.data(), mo.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_q.ptr, query_pts, size_t(N) * 12, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_t.ptr, train_pts, size_t(N) * 12, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_px.ptr, pixels, size_t(N) * 8, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_sp.ptr, spans, size_t(n_clusters) * 4, cudaMemcpyHostToDevice, st));
  StageTimer timer(st);
  timer.begin();
  TOD_CUDA(tod::launch_fill_adjacency(n_clusters, d_off.as<int32_t>(), d_mo.as<int64_t>(), d_q.as<float>(),
                                      d_t.as<float>(), d_px.as<float>(), d_sp.as<float>(), sensor_error,
                                      d_P.as<uint32_t>(), d_S.as<uint32_t>(), max_cluster, st));
  timer.end();
  TOD_CUDA(cudaMemcpyAsync(physical, d_P.ptr, mat_bytes, cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaMemcpyAsync(sample, d_S.ptr, mat_bytes, cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  timer.finish();
  return TOD_OK;
}

int tod_score_hypotheses(int32_t device, int32_t n, const float *query_pts, const float *train_pts,
                         const uint32_t *physical, const uint32_t *valid, int32_t n_hyp, const uint32_t *triples,
                         double threshold, int32_t *counts, float *R, float *T) {
  TOD_REQUIRE(n >= 0 && n_hyp >= 0, "negative size");
  if (n_hyp == 0) return TOD_OK;
  TOD_REQUIRE(n >= 3 && query_pts && train_pts && physical && valid && triples && counts, "null/too-small input");
  for (int64_t i = 0; i < int64_t(n_hyp) * 3; ++i)
    TOD_REQUIRE(triples[i] < uint32_t(n), "sample index %u out of range (n=%d)", triples[i], n);
  if (int rc = check_device(device)) return rc;
  const int W = tod::adjacency_row_words(n);

  // finite mask (see k3_score.cu): correspondence i can only pass `distSq < inf` if all its coordinates are finite
  std::vector<uint32_t> finite(size_t(W), 0u);
  for (int i = 0; i < n; ++i) {
    bool ok = true;
    for (int d = 0; d < 3; ++d) ok = ok && std::isfinite(query_pts[size_t(i) * 3 + d]) && std::isfinite(train_pts[size_t(i) * 3 + d]);
    if (ok) finite[size_t(i) >> 5] |= 1u << (i & 31);
  }
  std::vector<uint32_t> hyps(size_t(n_hyp) * 4);
  for (int h = 0; h < n_hyp; ++h) {
    hyps[size_t(h) * 4 + 0] = triples[size_t(h) * 3 + 0];
    hyps[size_t(h) * 4 + 1] = triples[size_t(h) * 3 + 1];
    hyps[size_t(h) * 4 + 2] = triples[size_t(h) * 3 + 2];
    hyps[size_t(h) * 4 + 3] = 0;
  }
  std::vector<unsigned char> desc(tod::k3_cluster_desc_size());
  tod::k3_fill_cluster_desc(desc.data(), n, W, 0, 0, 0);

  DeviceBuffer d_desc, d_q, d_t, d_P, d_V, d_F, d_h, d_c, d_R, d_T;
  Scratch scratch;
  for (DeviceBuffer *b : {&d_desc, &d_q, &d_t, &d_P, &d_V, &d_F, &d_h, &d_c, &d_R, &d_T}) scratch.own(*b);
  const size_t mat_bytes = size_t(n) * W * sizeof(uint32_t);
  TOD_CUDA(d_desc.reserve(desc.size()));
  TOD_CUDA(d_q.reserve(size_t(n) * 12));
  TOD_CUDA(d_t.reserve(size_t(n) * 12));
  TOD_CUDA(d_P.reserve(mat_bytes));
  TOD_CUDA(d_V.reserve(size_t(W) * 4));
  TOD_CUDA(d_F.reserve(size_t(W) * 4));
  TOD_CUDA(d_h.reserve(hyps.size() * 4));
  TOD_CUDA(d_c.reserve(size_t(n_hyp) * 4));
  if (R) TOD_CUDA(d_R.reserve(size_t(n_hyp) * 36));
  if (T) TOD_CUDA(d_T.reserve(size_t(n_hyp) * 12));
  cudaStream_t st = nullptr;
  TOD_CUDA(cudaMemcpyAsync(d_desc.ptr, desc.data(), desc.size(), cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_q.ptr, query_pts, size_t(n) * 12, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_t.ptr, train_pts, size_t(n) * 12, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_P.ptr, physical, mat_bytes, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_V.ptr, valid, size_t(W) * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_F.ptr, finite.data(), size_t(W) * 4, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaMemcpyAsync(d_h.ptr, hyps.data(), hyps.size() * 4, cudaMemcpyHostToDevice, st));
  StageTimer timer(st);
  timer.begin();
  TOD_CUDA(tod::launch_score_hypotheses_batched(d_desc.ptr, d_q.as<float>(), d_t.as<float>(), d_P.as<uint32_t>(),
                                                d_V.as<uint32_t>(), d_F.as<uint32_t>(), n_hyp, d_h.as<uint32_t>(),
                                                threshold, d_c.as<int32_t>(), R ? d_R.as<float>() : nullptr,
                                                T ? d_T.as<float>() : nullptr, st));
  timer.end();
  TOD_CUDA(cudaMemcpyAsync(counts, d_c.ptr, size_t(n_hyp) * 4, cudaMemcpyDeviceToHost, st));
  if (R) TOD_CUDA(cudaMemcpyAsync(R, d_R.ptr, size_t(n_hyp) * 36, cudaMemcpyDeviceToHost, st));
  if (T) TOD_CUDA(cudaMemcpyAsync(T, d_T.ptr, size_t(n_hyp) * 12, cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  timer.finish();
  return TOD_OK;
}

}  // extern "C"
