// Internal helpers shared by the C-ABI translation units of libtod_b200.so (not installed).
#ifndef TOD_INTERNAL_H_
#define TOD_INTERNAL_H_

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "tod_b200.h"

namespace tod {

// thread-local error message behind tod_last_error()
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(uint64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

#define TOD_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return ::tod::fail(TOD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                                      \
  } while (0)

#define TOD_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::tod::fail(TOD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Growable device buffer (cudaMalloc'd, never shrinks).
struct DeviceBuffer {
  void *ptr = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    size_t want = n + n / 4;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <typename T>
  T *as() const { return static_cast<T *>(ptr); }
};

// Growable pinned host buffer (cudaHostAlloc'd, never shrinks): device -> host copies into it run at PCIe speed
// instead of through the driver's pageable staging path.
struct PinnedBuffer {
  void *ptr = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    bytes = 0;
    size_t want = n + n / 4;
    cudaError_t e = cudaHostAlloc(&ptr, want, cudaHostAllocDefault);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <typename T>
  T *as() const { return static_cast<T *>(ptr); }
};

// ---- kernel launchers (defined in the .cu files) ---------------------------------------------------------------

// Packed key of one candidate: distance (9 bits) << 23 | global DB row (23 bits).  min() over keys is exactly the
// (distance, imgIdx, trainIdx) order of cv::BFMatcher because objects are concatenated in imgIdx order.
constexpr int kKeyRowBits = 23;
constexpr uint32_t kKeyRowMask = (1u << kKeyRowBits) - 1;
constexpr uint32_t kKeyEmpty = 0xFFFFFFFFu;
constexpr int64_t kMaxGlobalRows = int64_t(1) << kKeyRowBits;  // rows one packed key can address (a "segment")
constexpr int64_t kMaxDbRows = (int64_t(1) << 31) - 1;        // whole database (larger ones are cut into segments)

struct K1Plan {
  int q_per_thread;    // 1, 2 or 4
  int q_tile;          // queries per CTA = 256 * q_per_thread
  int n_qtiles;
  int rows_per_chunk;  // multiple of the smem tile
  int n_chunks;
  int n_sources;       // candidate lists per query written to `partial` (merge fan-in)
  int full_qtiles;     // k1_mma: the first full_qtiles query groups sweep the whole shard in one CTA (one cold start);
                       // only the remaining n_qtiles - full_qtiles groups are split into n_chunks db chunks
};
K1Plan k1_popc_plan(int nq, int64_t shard_rows, int sm_count);

// K1 (SIMT popc formulation).  partial: n_chunks x nq x k keys.
cudaError_t launch_k1_popc(const K1Plan &plan, const void *d_query, int nq, const void *d_db, int64_t shard_rows,
                           uint32_t global_row_base, int k, uint32_t radius, uint32_t *d_partial,
                           cudaStream_t stream);

// K1 (tcgen05 int8 formulation, k1_mma.cu).  The database is expanded to 0/1 int8 and the queries to +-1 int8 (256 B
// per descriptor), both described by TMA tensor maps (opaque CUtensorMap blobs of tensor_map_bytes() bytes, host
// memory).  launch_expand_queries also writes the queries' popcounts and resets their shared bounds to 511.
K1Plan k1_mma_plan(int nq, int64_t shard_rows, int sm_count);
cudaError_t launch_expand_db(const void *d_bits, void *d_int8, int64_t rows, cudaStream_t stream);
cudaError_t launch_expand_queries(const void *d_bits, void *d_int8, int64_t rows, uint32_t *d_popq, uint32_t *d_gthr,
                                  cudaStream_t stream);
bool make_desc_tensor_map(void *map_out, const void *d_int8, int64_t rows, int box_rows);
int k1_mma_query_box_rows();
int k1_mma_db_box_rows();
size_t tensor_map_bytes();
// d_gthr: nq u32 shared per-query bounds, must hold 511 (no bound) before the launch; d_popq: nq query popcounts
// (both written by launch_expand_queries).
// seg_row0: the launch covers rows [seg_row0, seg_row0 + shard_rows) of the matrix behind map_db (a segment of a
// wide database); keys count rows from global_row_base + 0 at seg_row0.
cudaError_t launch_k1_mma(const K1Plan &plan, const void *map_q, const void *map_db, int nq, int64_t shard_rows,
                          uint32_t global_row_base, int k, uint32_t radius, uint32_t *d_partial, uint32_t *d_gthr,
                          const uint32_t *d_popq, cudaStream_t stream, int64_t seg_row0 = 0);

// Reduce n_src x nq x k key lists to nq x k keys (ascending).
cudaError_t launch_reduce_keys(const uint32_t *d_keys, int n_src, int nq, int k, uint32_t *d_out,
                               cudaStream_t stream);
// Reduce + radius cut + decode (imgIdx, trainIdx) + matches_3d gather.  d_src_base (wide databases, more than 2^23
// rows): first global row of every source list, whose keys then count rows from there; the merge compares
// (distance, global row) on 64 bits.
cudaError_t launch_finalize_matches(const uint32_t *d_keys, int n_src, int nq, int k, uint32_t radius,
                                    const uint32_t *d_obj_offsets, int n_objects, const float *d_points,
                                    tod_match *d_matches, int32_t *d_counts, float *d_points3d,
                                    cudaStream_t stream, int ratio_enabled = 0, float ratio = 0.f,
                                    uint32_t *d_rows_out = nullptr, size_t src_stride = 0,
                                    const uint32_t *d_wait_flags = nullptr, uint32_t wait_step = 0,
                                    uint32_t *d_wait_error = nullptr, const uint32_t *d_src_base = nullptr);
// Fused top-k reduction + all-gather over peer memory (k1_merge.cu: reduce_push_kernel): the reduced keys of this rank
// go straight into slot `rank` of every rank's exchange buffer (d_peers[r] = that buffer's base for this step's parity,
// CUDA-IPC-mapped), then this rank's flag (at d_peers[r] + flag_offset + rank) is raised to `step` on every peer.
// The matching launch_finalize_matches call passes d_wait_flags (= own buffer + flag_offset) and wait_step = step:
// it starts merging once every rank's flag has arrived.
cudaError_t launch_reduce_push(const uint32_t *d_keys, int n_src, int nq, int k, uint32_t *const *d_peers, int world,
                               int rank, size_t slot_stride, size_t flag_offset, uint32_t step, unsigned int *d_ticket,
                               cudaStream_t stream);
// Duplicate-match removal inside each frame (DescriptorMatcher.cpp:229 TODO): d_rows = the global DB row of every
// match slot (written by launch_finalize_matches); the hash tables hold table_slots (a power of two) u64 each.
cudaError_t launch_remove_duplicates(tod_match *d_matches, int32_t *d_counts, float *d_points3d,
                                     const uint32_t *d_rows, int nq, int k, int frame_keypoints, void *d_hkeys,
                                     void *d_hvals, size_t table_slots, cudaStream_t stream);

// K2: one launch over all clusters. Device pointers.
cudaError_t launch_fill_adjacency(int n_clusters, const int32_t *d_offsets, const int64_t *d_matrix_offsets,
                                  const float *d_query, const float *d_train, const float *d_pixels,
                                  const float *d_spans, float sensor_error, uint32_t *d_physical,
                                  uint32_t *d_sample, int max_cluster, cudaStream_t stream);

// One (frame, object) cluster of correspondences as the geometry kernels see it: n, row words W, and the offsets of
// its points / bit-matrices / W-word bit-vectors (valid, finite, degree mask) inside the batch-wide buffers.
struct K3Cluster {
  int32_t n;
  int32_t W;
  int64_t point_offset;   // in points
  int64_t matrix_offset;  // in u32 words
  int64_t valid_offset;   // in u32 words (offset into the `valid`, `finite` and `deg_mask` bit-vectors)
};

// K3: hypotheses (s0, s1, s2, cluster) of a batch of clusters described by K3Cluster records.
// d_finite may be null (all points finite).
cudaError_t launch_score_hypotheses_batched(const void *d_clusters, const float *d_query, const float *d_train,
                                            const uint32_t *d_physical, const uint32_t *d_valid,
                                            const uint32_t *d_finite, int n_hyp, const uint32_t *d_hyps,
                                            double threshold, int32_t *d_counts, float *d_R, float *d_T,
                                            cudaStream_t stream);
size_t k3_cluster_desc_size();
void k3_fill_cluster_desc(void *dst, int32_t n, int32_t W, int64_t point_offset, int64_t matrix_offset,
                          int64_t valid_offset);

// Per round, for the clusters listed in d_active: deg_mask bit v <=> v is valid and has at least `min_degree` valid
// neighbours in the sample graph (the degree filter of the clique gate, sac_model_registration_graph.h:209-213).
// deg_mask words of the listed clusters are overwritten.
cudaError_t launch_sample_degree_mask(const void *d_clusters, const int32_t *d_active, int n_active, int max_n,
                                      const uint32_t *d_sample, const uint32_t *d_valid, uint32_t *d_deg_mask,
                                      int min_degree, cudaStream_t stream);

// K4: the clique gate's exact pre-checks for every hypothesis whose pre-gate count beats both 7 and its cluster's
// current best (d_floor, per cluster).  verdict: 0 = not evaluated, 1 = the gate FAILS for certain (no clique of 8 can
// be returned by the reference's search), 2 = undecided: run the exact host search.  Reference-faithful (+inf
// threshold) mode only.
// K5 (same launch call): hypotheses K4 cannot settle whose filtered graph has at most 128 vertices are queued in
// d_jobs (gate_job_bytes(n_hyp, pool_bytes) bytes, may be null = no K5) and decided by the reference's bounded search
// stepped exactly on the GPU: verdict 3 = the gate PASSES, 4 = it fails (decided by the search, not by a proof).
constexpr int kGateProofMax = 1024;  // K4 runs its proofs on filtered graphs of at most this many vertices
enum { kGateNotEvaluated = 0, kGateFails = 1, kGateNeedsHost = 2, kGatePasses = 3, kGateFailsSearch = 4 };
size_t gate_job_bytes(int n_hyp, size_t pool_bytes);
cudaError_t launch_gate_search_jobs(int n_hyp, void *d_jobs, uint8_t *d_verdict, cudaStream_t stream);
cudaError_t launch_gate_prechecks(const void *d_clusters, const uint32_t *d_physical, const uint32_t *d_sample,
                                  const uint32_t *d_valid, const uint32_t *d_finite, const uint32_t *d_deg_mask,
                                  int n_hyp, const uint32_t *d_hyps, const int32_t *d_counts, const int32_t *d_floor,
                                  int max_words, uint8_t *d_verdict, void *d_jobs, size_t pool_bytes,
                                  cudaStream_t stream, cudaEvent_t ev_between = nullptr);

inline int adjacency_row_words(int n) { return ((n + 31) / 32 + 3) & ~3; }

}  // namespace tod
#endif
