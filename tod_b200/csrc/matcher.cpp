// C-ABI of the DescriptorMatcher half of the hot path (include/tod_b200.h: tod_matcher_*).
// Host logic mirrors src/detection/DescriptorMatcher.cpp of the reference: parameter_callback (:60-129) ->
// add_object/train, configure (:154-188) -> params_from_json, process (:195-252) -> knn.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "json_min.h"
#include "nccl_dl.h"
#include "tod_internal.h"

using tod::DeviceBuffer;
using tod::fail;

struct tod_matcher {
  tod_matcher_params p{};
  // host copy of the DB (concatenated in imgIdx order)
  std::vector<std::string> ids;
  std::vector<float> spans;
  std::vector<uint32_t> offsets{0};  // n_objects + 1 global row offsets
  std::vector<uint8_t> h_desc;
  std::vector<float> h_pts;
  bool trained = false;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  // K1 is bracketed by a ring of event pairs, one pair per call: a caller that streams many steps can read every
  // step's kernel time afterwards (tod_matcher_k1_ms_ago) without synchronising inside its loop
  static constexpr int kEvRing = 64;
  cudaEvent_t ev0[kEvRing] = {nullptr}, ev1[kEvRing] = {nullptr};
  uint64_t k1_calls = 0;
  int64_t shard_begin = 0, shard_rows = 0;
  DeviceBuffer d_db, d_pts, d_offsets, d_query, d_partial, d_matches, d_counts, d_pts3d;
  // tensor-core formulation: int8 copies (db 0/1, queries +-1; 256 B / descriptor) and their TMA tensor maps
  DeviceBuffer d_db8, d_q8, d_gthr, d_popq;
  alignas(64) unsigned char map_db[128];
  alignas(64) unsigned char map_q[128];
  bool have_db8 = false;
  const void *map_q_ptr = nullptr;
  int64_t map_q_rows = -1;
  const char *last_kernel = "none";
  // opt-in post-filters (ratio test / duplicate removal): global row per match slot + the (frame, row) hash table
  DeviceBuffer d_rows, d_hkeys, d_hvals;
  // sharded handle with a communicator (tod_matcher_set_comm)
  tod::ncclComm_t comm = nullptr;
  int comm_mode = 0;                 // 0 none, 1 NCCL
  DeviceBuffer d_keys_local, d_keys_all;
  cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;  // around the all-gather of the last sharded call
  bool ev_x_valid = false;
  bool stage_timing = false;         // tod_matcher_set_stage_timing: also bracket the exchange with events
  int32_t reserved_nq = 0;
  // wide database (more than 2^23 rows in all): the shard is scanned in segments of 2^23 rows, keys count rows from
  // their segment's start, and the merge orders (distance, global row) on 64 bits with the table d_src_base
  bool wide = false;
  int n_seg = 1;                     // segments per shard (the same on every rank)
  DeviceBuffer d_src_base;           // shard_count x n_seg first global rows
  // peer-memory exchange (comm_mode 2): every rank owns one buffer of 2 parities x (world key slots + world flags),
  // mapped into all the others with CUDA IPC; reduce_push_kernel stores into them over NVLink
  bool peer_enabled = true;          // tod_matcher_set_exchange: 0 keeps the ncclAllGather path (A/B runs)
  bool peer_tried = false;           // the mapping was attempted for the current communicator
  size_t peer_cap = 0;               // keys per slot
  size_t peer_half = 0;              // u32 words per parity half: world * peer_cap keys + flags
  uint32_t *peer_local = nullptr;    // own buffer (cudaMalloc)
  std::vector<void *> peer_mapped;   // bases of the other ranks' buffers (cudaIpcOpenMemHandle), own rank = peer_local
  DeviceBuffer d_peer_tbl;           // 2 x world device pointers (one table per parity) | ticket | error word
  uint32_t peer_step = 0;
};

namespace {

int use_device(const tod_matcher *m) {
  TOD_CUDA(cudaSetDevice(m->p.device));
  return TOD_OK;
}

// TOD_KERNEL_AUTO picks the tensor-core formulation (5-13x the popc kernel on B200) whenever this
// shard holds rows; the popc kernel stays selectable and is the cross-check in tests/.
bool use_mma(const tod_matcher *m) {
  return m->p.kernel == TOD_KERNEL_MMA || (m->p.kernel == TOD_KERNEL_AUTO && m->shard_rows > 0);
}

// K1 on rows [seg_row0, seg_row0 + seg_rows) of this shard; keys = distance << 23 | (key_base + row - seg_row0).
// The whole shard with key_base = shard_begin unless the database is wide.  first / last: the first launch of a call
// expands the queries (and resets their shared bounds, which later segments then start from), first and last bracket
// the call's K1 time.
int run_k1(tod_matcher *m, const void *d_query, int nq, cudaStream_t st, tod::K1Plan *plan_out, int64_t seg_row0,
           int64_t seg_rows, uint32_t key_base, bool first, bool last) {
  if (use_mma(m)) {
    if (!m->have_db8) return fail(TOD_ERR_STATE, "tensor-core K1 requested but the int8 database was not built");
    tod::K1Plan plan = tod::k1_mma_plan(nq, seg_rows, m->sm_count);
    TOD_CUDA(m->d_partial.reserve(size_t(plan.n_sources) * size_t(std::max(nq, 1)) * m->p.k * sizeof(uint32_t)));
    TOD_CUDA(m->d_q8.reserve(size_t(nq) * 256));
    if (m->map_q_ptr != m->d_q8.ptr || m->map_q_rows != nq) {
      if (!tod::make_desc_tensor_map(m->map_q, m->d_q8.ptr, nq, tod::k1_mma_query_box_rows()))
        return fail(TOD_ERR_CUDA, "cuTensorMapEncodeTiled failed for the query matrix");
      m->map_q_ptr = m->d_q8.ptr;
      m->map_q_rows = nq;
    }
    TOD_CUDA(m->d_gthr.reserve(size_t(nq) * sizeof(uint32_t)));
    TOD_CUDA(m->d_popq.reserve(size_t(nq) * sizeof(uint32_t)));
    if (first)
      TOD_CUDA(tod::launch_expand_queries(d_query, m->d_q8.ptr, nq, m->d_popq.as<uint32_t>(),
                                          m->d_gthr.as<uint32_t>(), st));
    const int er = int(m->k1_calls % tod_matcher::kEvRing);
    if (first) TOD_CUDA(cudaEventRecord(m->ev0[er], st));
    TOD_CUDA(tod::launch_k1_mma(plan, m->map_q, m->map_db, nq, seg_rows, key_base, m->p.k, m->p.radius,
                                m->d_partial.as<uint32_t>(), m->d_gthr.as<uint32_t>(), m->d_popq.as<uint32_t>(), st,
                                seg_row0));
    if (last) {
      TOD_CUDA(cudaEventRecord(m->ev1[er], st));
      ++m->k1_calls;
    }
    m->last_kernel = "mma";
    *plan_out = plan;
    return TOD_OK;
  }
  tod::K1Plan plan = tod::k1_popc_plan(nq, seg_rows, m->sm_count);
  TOD_CUDA(m->d_partial.reserve(size_t(plan.n_sources) * size_t(std::max(nq, 1)) * m->p.k * sizeof(uint32_t)));
  const int er = int(m->k1_calls % tod_matcher::kEvRing);
  if (first) TOD_CUDA(cudaEventRecord(m->ev0[er], st));
  TOD_CUDA(tod::launch_k1_popc(plan, d_query, nq, m->d_db.as<uint8_t>() + size_t(seg_row0) * 32, seg_rows, key_base,
                               m->p.k, m->p.radius, m->d_partial.as<uint32_t>(), st));
  if (last) {
    TOD_CUDA(cudaEventRecord(m->ev1[er], st));
    ++m->k1_calls;
  }
  m->last_kernel = "popc";
  *plan_out = plan;
  return TOD_OK;
}

int run_k1(tod_matcher *m, const void *d_query, int nq, cudaStream_t st, tod::K1Plan *plan_out) {
  return run_k1(m, d_query, nq, st, plan_out, 0, m->shard_rows, uint32_t(m->shard_begin), true, true);
}

#define TOD_NCCL(expr)                                                                                \
  do {                                                                                                \
    int _r = (expr);                                                                                  \
    if (_r != tod::kNcclSuccess)                                                                      \
      return fail(TOD_ERR_CUDA, "%s failed: %s", #expr, tod::nccl_api().GetErrorString(_r));          \
  } while (0)

// Merge (+ radius cut, decode, 3-D gather) and the opt-in post-filters, from n_src key lists per query.
int finalize(tod_matcher *m, const uint32_t *d_keys, int n_src, int nq, tod_match *d_matches, int32_t *d_counts,
             float *d_points3d, cudaStream_t st, size_t src_stride = 0, const uint32_t *wait_flags = nullptr,
             uint32_t wait_step = 0, uint32_t *wait_error = nullptr, const uint32_t *src_base = nullptr) {
  const int k = m->p.k;
  const bool ratio = m->p.ratio_enabled != 0 && k >= 2;
  const bool dedupe = m->p.remove_duplicates != 0;
  uint32_t *rows = nullptr;
  size_t slots = 0;
  if (dedupe) {
    TOD_CUDA(m->d_rows.reserve(size_t(nq) * k * sizeof(uint32_t)));
    slots = 1024;
    while (slots < 2 * size_t(nq) * size_t(k)) slots <<= 1;
    TOD_CUDA(m->d_hkeys.reserve(slots * 8));
    TOD_CUDA(m->d_hvals.reserve(slots * 8));
    rows = m->d_rows.as<uint32_t>();
  }
  TOD_CUDA(tod::launch_finalize_matches(d_keys, n_src, nq, k, m->p.radius, m->d_offsets.as<uint32_t>(),
                                        int(m->ids.size()), m->d_pts.as<float>(), d_matches, d_counts, d_points3d, st,
                                        ratio ? 1 : 0, m->p.ratio, rows, src_stride, wait_flags, wait_step, wait_error,
                                        src_base));
  if (dedupe)
    TOD_CUDA(tod::launch_remove_duplicates(d_matches, d_counts, d_points3d, rows, nq, k, m->p.frame_keypoints,
                                           m->d_hkeys.ptr, m->d_hvals.ptr, slots, st));
  return TOD_OK;
}

int setup_peer_exchange(tod_matcher *m, size_t need_keys, cudaStream_t st);

// The whole DescriptorMatcher.process on device buffers, enqueued on `st`: K1 on this shard, then (sharded) top-k
// reduction -> ncclAllGather of the packed keys -> merge; or (one shard) merge of the chunk lists directly.
int run_process(tod_matcher *m, const void *d_query, int nq, tod_match *d_matches, int32_t *d_counts,
                float *d_points3d, cudaStream_t st) {
  const int k = m->p.k;
  tod::K1Plan plan;
  if (m->wide) {
    // more than 2^23 rows in all: one K1 launch + top-k reduction per segment of this shard (keys local to the
    // segment; the shared per-query bounds carry over, so later segments start warm), then the merge over every
    // (rank, segment) list on 64-bit (distance, global row) keys.  The exchange is the ncclAllGather one.
    const int world = m->p.shard_count;
    if (world > 1 && !m->comm)
      return fail(TOD_ERR_STATE, "this handle holds shard %d of %d: call tod_matcher_set_comm first", m->p.shard_rank,
                  world);
    const size_t nk = size_t(nq) * k;
    TOD_CUDA(m->d_keys_local.reserve(nk * sizeof(uint32_t) * size_t(m->n_seg)));
    int last_seg = 0;
    for (int sg = 0; sg < m->n_seg; ++sg)
      if (int64_t(sg) * tod::kMaxGlobalRows < m->shard_rows) last_seg = sg;
    for (int sg = 0; sg < m->n_seg; ++sg) {
      const int64_t r0 = int64_t(sg) * tod::kMaxGlobalRows;
      const int64_t rows = std::min<int64_t>(tod::kMaxGlobalRows, m->shard_rows - r0);
      uint32_t *dst = m->d_keys_local.as<uint32_t>() + size_t(sg) * nk;
      if (rows <= 0) {  // the last rank's shard may end before the last segment
        TOD_CUDA(cudaMemsetAsync(dst, 0xFF, nk * sizeof(uint32_t), st));
        continue;
      }
      if (int rc = run_k1(m, d_query, nq, st, &plan, r0, rows, 0u, sg == 0, sg == last_seg)) return rc;
      TOD_CUDA(tod::launch_reduce_keys(m->d_partial.as<uint32_t>(), plan.n_sources, nq, k, dst, st));
    }
    if (world == 1)
      return finalize(m, m->d_keys_local.as<uint32_t>(), m->n_seg, nq, d_matches, d_counts, d_points3d, st, 0, nullptr,
                      0, nullptr, m->d_src_base.as<uint32_t>());
    TOD_CUDA(m->d_keys_all.reserve(nk * sizeof(uint32_t) * size_t(m->n_seg) * size_t(world)));
    if (m->stage_timing) TOD_CUDA(cudaEventRecord(m->ev_x0, st));
    TOD_NCCL(tod::nccl_api().AllGather(m->d_keys_local.ptr, m->d_keys_all.ptr, nk * size_t(m->n_seg), tod::kNcclUint32,
                                       m->comm, st));
    if (m->stage_timing) TOD_CUDA(cudaEventRecord(m->ev_x1, st));
    m->ev_x_valid = m->stage_timing;
    return finalize(m, m->d_keys_all.as<uint32_t>(), world * m->n_seg, nq, d_matches, d_counts, d_points3d, st, 0,
                    nullptr, 0, nullptr, m->d_src_base.as<uint32_t>());
  }
  if (m->p.shard_count == 1) {
    if (int rc = run_k1(m, d_query, nq, st, &plan)) return rc;
    return finalize(m, m->d_partial.as<uint32_t>(), plan.n_sources, nq, d_matches, d_counts, d_points3d, st);
  }
  if (!m->comm)
    return fail(TOD_ERR_STATE, "this handle holds shard %d of %d: call tod_matcher_set_comm first (or use the "
                               "*_device stage calls with your own exchange)", m->p.shard_rank, m->p.shard_count);
  const size_t nk = size_t(nq) * k;
  if (m->peer_enabled && !m->wide && (!m->peer_tried || (m->comm_mode == 2 && nk > m->peer_cap)))
    if (int rc = setup_peer_exchange(m, nk, st)) return rc;
  if (m->comm_mode == 2) {
    // K1 -> [top-k reduction fused with the all-gather: stores into every rank's buffer over NVLink] -> merge, which
    // starts as soon as every rank's flag has arrived
    const int world = m->p.shard_count;
    uint32_t *const *tbl = m->d_peer_tbl.as<uint32_t *>();
    unsigned int *ticket = reinterpret_cast<unsigned int *>(m->d_peer_tbl.as<uint64_t>() + 2 * world);
    uint32_t *err = reinterpret_cast<uint32_t *>(m->d_peer_tbl.as<uint64_t>() + 2 * world + 1);
    if (int rc = run_k1(m, d_query, nq, st, &plan)) return rc;
    if (++m->peer_step == 0u) m->peer_step = 2u;  // 0 is what freshly zeroed flags hold; 2 keeps the parity sequence
    const uint32_t step = m->peer_step;
    const int par = int(step & 1u);
    const size_t flag_off = size_t(world) * m->peer_cap;
    if (m->stage_timing) TOD_CUDA(cudaEventRecord(m->ev_x0, st));
    TOD_CUDA(tod::launch_reduce_push(m->d_partial.as<uint32_t>(), plan.n_sources, nq, k, tbl + size_t(par) * world, world,
                                     m->p.shard_rank, m->peer_cap, flag_off, step, ticket, st));
    if (m->stage_timing) TOD_CUDA(cudaEventRecord(m->ev_x1, st));
    m->ev_x_valid = m->stage_timing;
    const uint32_t *mine = m->peer_local + size_t(par) * m->peer_half;
    return finalize(m, mine, world, nq, d_matches, d_counts, d_points3d, st, m->peer_cap, mine + flag_off, step, err);
  }
  TOD_CUDA(m->d_keys_local.reserve(nk * sizeof(uint32_t)));
  TOD_CUDA(m->d_keys_all.reserve(nk * sizeof(uint32_t) * size_t(m->p.shard_count)));
  if (int rc = run_k1(m, d_query, nq, st, &plan)) return rc;
  TOD_CUDA(tod::launch_reduce_keys(m->d_partial.as<uint32_t>(), plan.n_sources, nq, k, m->d_keys_local.as<uint32_t>(), st));
  if (m->stage_timing) TOD_CUDA(cudaEventRecord(m->ev_x0, st));
  TOD_NCCL(tod::nccl_api().AllGather(m->d_keys_local.ptr, m->d_keys_all.ptr, nk, tod::kNcclUint32, m->comm, st));
  if (m->stage_timing) TOD_CUDA(cudaEventRecord(m->ev_x1, st));
  m->ev_x_valid = m->stage_timing;
  return finalize(m, m->d_keys_all.as<uint32_t>(), m->p.shard_count, nq, d_matches, d_counts, d_points3d, st);
}

void close_peer_exchange(tod_matcher *m) {
  for (size_t r = 0; r < m->peer_mapped.size(); ++r)
    if (m->peer_mapped[r] && m->peer_mapped[r] != m->peer_local) cudaIpcCloseMemHandle(m->peer_mapped[r]);
  m->peer_mapped.clear();
  if (m->peer_local) cudaFree(m->peer_local);
  m->peer_local = nullptr;
  m->peer_cap = m->peer_half = 0;
  m->peer_step = 0;
  if (m->comm_mode == 2) m->comm_mode = 1;
}

// Collective over the ranks of the communicator (every rank reaches it in the same call, with the same nq): allocate
// the exchange buffer, trade CUDA IPC handles through NCCL, map the peers' buffers.  Any rank that cannot map a peer
// (no peer access, handles opened inside one process, ...) makes ALL ranks stay on the ncclAllGather path.
// (worker: `local` / `mapped` / `d_rec` belong to the caller, which frees them when the mapping is not adopted)
int setup_peer_exchange_impl(tod_matcher *m, size_t need_keys, cudaStream_t st, void *&local, std::vector<void *> &mapped,
                             DeviceBuffer &d_rec, bool &adopted) {
  const tod::NcclApi &api = tod::nccl_api();
  const int world = m->p.shard_count, rank = m->p.shard_rank;
  TOD_CUDA(cudaStreamSynchronize(st));
  close_peer_exchange(m);
  m->peer_tried = true;
  const size_t cap = (std::max(need_keys, size_t(std::max(m->reserved_nq, 0)) * size_t(m->p.k)) + 63) & ~size_t(63);
  const size_t half = size_t(world) * cap + 64 * ((size_t(world) + 63) / 64);
  bool ok = world <= 128;
  if (ok && cudaMalloc(&local, 2 * half * sizeof(uint32_t)) != cudaSuccess) ok = false, local = nullptr, cudaGetLastError();
  if (local) ok = cudaMemset(local, 0, 2 * half * sizeof(uint32_t)) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, local) != cudaSuccess) ok = false, cudaGetLastError();
  // trade {handle, ok} records
  struct Rec {
    cudaIpcMemHandle_t h;
    int32_t ok;
    int32_t pad[15];
  };
  static_assert(sizeof(Rec) == 128, "IPC record");
  TOD_CUDA(d_rec.reserve(sizeof(Rec) * size_t(world + 1)));
  Rec my{};
  my.h = mine;
  my.ok = ok ? 1 : 0;
  std::vector<Rec> all(static_cast<size_t>(world));
  Rec *d_my = d_rec.as<Rec>() + world;
  TOD_CUDA(cudaMemcpyAsync(d_my, &my, sizeof(Rec), cudaMemcpyHostToDevice, st));
  TOD_NCCL(api.AllGather(d_my, d_rec.ptr, sizeof(Rec), tod::kNcclUint8, m->comm, st));
  TOD_CUDA(cudaMemcpyAsync(all.data(), d_rec.ptr, sizeof(Rec) * size_t(world), cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  for (const Rec &r : all) ok = ok && r.ok == 1;
  mapped.assign(static_cast<size_t>(world), nullptr);
  if (ok) {
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) {
        mapped[size_t(r)] = local;
        continue;
      }
      if (cudaIpcOpenMemHandle(&mapped[size_t(r)], all[size_t(r)].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        mapped[size_t(r)] = nullptr;
        ok = false;
        cudaGetLastError();
      }
    }
  }
  // second round: did everybody map everybody?  (also the barrier behind which every buffer is zeroed)
  my.ok = ok ? 1 : 0;
  TOD_CUDA(cudaMemcpyAsync(d_my, &my, sizeof(Rec), cudaMemcpyHostToDevice, st));
  TOD_NCCL(api.AllGather(d_my, d_rec.ptr, sizeof(Rec), tod::kNcclUint8, m->comm, st));
  TOD_CUDA(cudaMemcpyAsync(all.data(), d_rec.ptr, sizeof(Rec) * size_t(world), cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  for (const Rec &r : all) ok = ok && r.ok == 1;
  if (!ok) return TOD_OK;  // stay on NCCL (the caller releases what was allocated and mapped)
  m->peer_local = static_cast<uint32_t *>(local);
  m->peer_mapped = mapped;
  m->peer_cap = cap;
  m->peer_half = half;
  m->peer_step = 0;
  // device table: parity 0 bases, parity 1 bases, then the ticket and the error word
  std::vector<uint64_t> tbl(size_t(2 * world) + 2, 0ull);
  for (int par = 0; par < 2; ++par)
    for (int r = 0; r < world; ++r)
      tbl[size_t(par * world + r)] = uint64_t(reinterpret_cast<uintptr_t>(static_cast<uint32_t *>(mapped[size_t(r)]) + size_t(par) * half));
  TOD_CUDA(m->d_peer_tbl.reserve(tbl.size() * 8));
  TOD_CUDA(cudaMemcpyAsync(m->d_peer_tbl.ptr, tbl.data(), tbl.size() * 8, cudaMemcpyHostToDevice, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  m->comm_mode = 2;
  adopted = true;
  return TOD_OK;
}

int setup_peer_exchange(tod_matcher *m, size_t need_keys, cudaStream_t st) {
  void *local = nullptr;
  std::vector<void *> mapped;
  DeviceBuffer d_rec;
  bool adopted = false;
  const int rc = setup_peer_exchange_impl(m, need_keys, st, local, mapped, d_rec, adopted);
  d_rec.release();
  if (!adopted) {  // error or "stay on NCCL": nothing of this attempt survives
    for (size_t r = 0; r < mapped.size(); ++r)
      if (mapped[r] && mapped[r] != local) cudaIpcCloseMemHandle(mapped[r]);
    if (local) cudaFree(local);
    m->peer_local = nullptr;
    m->peer_mapped.clear();
    m->peer_cap = m->peer_half = 0;
    if (m->comm_mode == 2) m->comm_mode = 1;
  }
  return rc;
}

void close_comm(tod_matcher *m) {
  close_peer_exchange(m);
  m->peer_tried = false;
  // ncclCommDestroy waits for the peer ranks (it hung a run whose ranks closed their handles at different times);
  // the handle's stream is idle here, so the communicator is torn down locally with ncclCommAbort instead.
  if (m->comm) {
    if (tod::nccl_api().CommAbort) tod::nccl_api().CommAbort(m->comm);
    else tod::nccl_api().CommDestroy(m->comm);
  }
  m->comm = nullptr;
  m->comm_mode = 0;
}

}  // namespace

extern "C" {

void tod_matcher_default_params(tod_matcher_params *p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->k = 5;
  p->radius = 0;
  p->search_type = TOD_SEARCH_EXACT;
  p->device = 0;
  p->shard_rank = 0;
  p->shard_count = 1;
  p->kernel = TOD_KERNEL_AUTO;
  p->ratio_enabled = 0;
  p->ratio = 0.f;
  p->remove_duplicates = 0;
  p->frame_keypoints = 0;
}

int tod_matcher_params_from_json(const char *search_json_params, tod_matcher_params *p) {
  TOD_REQUIRE(search_json_params && p, "null argument");
  tod::json::Value v;
  if (!tod::json::parse(search_json_params, v) || v.type != tod::json::Value::Object)
    return fail(TOD_ERR_PARSE, "search_json_params is not a JSON object");
  // radius_ and ratio_ are `unsigned int` members assigned from get_real() (DescriptorMatcher.cpp:170-171,257-259):
  // the value is truncated toward zero; ratio 0.8 becomes 0 and the ratio block is empty anyway (:223-227).
  if (v.has("radius")) {
    const double r = v.at("radius").num;
    TOD_REQUIRE(v.at("radius").type == tod::json::Value::Number && r >= 0 && r < 4294967296.0, "bad radius");
    p->radius = static_cast<uint32_t>(r);
  }
  // `ratio` is kept as the real number the .ork file holds, but the test itself is an opt-in extension: the
  // reference's block is empty (:223-227).  "ratio_enabled" / "remove_duplicates" are this library's own keys.
  if (v.has("ratio") && v.at("ratio").type == tod::json::Value::Number) p->ratio = float(v.at("ratio").num);
  if (v.has("ratio_enabled")) p->ratio_enabled = tod::json::truthy(v.at("ratio_enabled")) ? 1 : 0;
  if (v.has("remove_duplicates")) p->remove_duplicates = tod::json::truthy(v.at("remove_duplicates")) ? 1 : 0;
  const tod::json::Value &type = v.at("type");
  TOD_REQUIRE(type.type == tod::json::Value::String, "search type missing");
  if (type.str == "LSH") {
    p->search_type = TOD_SEARCH_LSH;  // n_tables / key_size / multi_probe_level are accepted and ignored: exact search
  } else if (type.str == "BruteForce" || type.str == "exact" || type.str == "BruteForce-Hamming") {
    p->search_type = TOD_SEARCH_EXACT;
  } else {
    return fail(TOD_ERR_INVALID, "Search not implemented for that type: %s", type.str.c_str());
  }
  return TOD_OK;
}

int tod_shard_range(int64_t total_rows, int32_t shard_rank, int32_t shard_count, int64_t *begin, int64_t *rows) {
  TOD_REQUIRE(begin && rows, "null argument");
  TOD_REQUIRE(total_rows >= 0 && shard_count >= 1 && shard_rank >= 0 && shard_rank < shard_count, "bad shard %d/%d",
              shard_rank, shard_count);
  const int64_t per = (total_rows + shard_count - 1) / shard_count;
  *begin = std::min<int64_t>(total_rows, per * shard_rank);
  *rows = std::min<int64_t>(total_rows, *begin + per) - *begin;
  return TOD_OK;
}

uint32_t tod_pack_key(uint32_t distance, uint32_t global_row) {
  return (std::min<uint32_t>(distance, 511u) << tod::kKeyRowBits) | (global_row & tod::kKeyRowMask);
}

int tod_matcher_create(const tod_matcher_params *p, tod_matcher **out) {
  TOD_REQUIRE(p && out, "null argument");
  TOD_REQUIRE(p->k >= 1 && p->k <= TOD_MAX_K, "k must be in 1..%d (got %d)", TOD_MAX_K, p->k);
  TOD_REQUIRE(p->shard_count >= 1 && p->shard_rank >= 0 && p->shard_rank < p->shard_count, "bad shard %d/%d",
              p->shard_rank, p->shard_count);
  TOD_REQUIRE(p->search_type == TOD_SEARCH_EXACT || p->search_type == TOD_SEARCH_LSH, "bad search_type");
  TOD_REQUIRE(!p->ratio_enabled || (p->k >= 2 && p->ratio > 0.f), "the ratio test needs k >= 2 and ratio > 0");
  TOD_REQUIRE(p->frame_keypoints >= 0, "negative frame_keypoints");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return fail(TOD_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  TOD_REQUIRE(p->device >= 0 && p->device < n_dev, "device %d out of range (%d devices)", p->device, n_dev);
  TOD_CUDA(cudaSetDevice(p->device));
  int major = 0;
  TOD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, p->device));
  if (major != 10) return fail(TOD_ERR_CUDA, "device %d is sm_%dx; this library is built for sm_100a only", p->device, major);
  tod_matcher *m = new tod_matcher();
  m->p = *p;
  cudaError_t ce = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, p->device);
  if (ce == cudaSuccess) ce = cudaStreamCreate(&m->stream);
  for (int i = 0; i < tod_matcher::kEvRing && ce == cudaSuccess; ++i) {
    ce = cudaEventCreate(&m->ev0[i]);
    if (ce == cudaSuccess) ce = cudaEventCreate(&m->ev1[i]);
  }
  if (ce == cudaSuccess) ce = cudaEventCreate(&m->ev_x0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&m->ev_x1);
  if (ce != cudaSuccess) {
    tod_matcher_destroy(m);
    return fail(TOD_ERR_CUDA, "creating the matcher's stream/events failed: %s", cudaGetErrorString(ce));
  }
  *out = m;
  return TOD_OK;
}

void tod_matcher_destroy(tod_matcher *m) {
  if (!m) return;
  cudaSetDevice(m->p.device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  close_comm(m);
  for (DeviceBuffer *b : {&m->d_db, &m->d_pts, &m->d_offsets, &m->d_query, &m->d_partial, &m->d_matches,
                          &m->d_counts, &m->d_pts3d, &m->d_db8, &m->d_q8, &m->d_gthr, &m->d_popq, &m->d_rows,
                          &m->d_hkeys, &m->d_hvals, &m->d_keys_local, &m->d_keys_all, &m->d_peer_tbl, &m->d_src_base})
    b->release();
  for (int i = 0; i < tod_matcher::kEvRing; ++i) {
    if (m->ev0[i]) cudaEventDestroy(m->ev0[i]);
    if (m->ev1[i]) cudaEventDestroy(m->ev1[i]);
  }
  if (m->ev_x0) cudaEventDestroy(m->ev_x0);
  if (m->ev_x1) cudaEventDestroy(m->ev_x1);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int tod_matcher_add_object(tod_matcher *m, const char *object_id, const uint8_t *descriptors, const float *points,
                           int32_t n) {
  TOD_REQUIRE(m && object_id, "null argument");
  TOD_REQUIRE(n >= 0 && (n == 0 || (descriptors && points)), "bad descriptor/point buffers");
  const int64_t total = int64_t(m->offsets.back()) + n;
  if (total > tod::kMaxDbRows)
    return fail(TOD_ERR_LIMIT, "DB would hold %lld descriptors; at most %lld are supported", (long long)total,
                (long long)tod::kMaxDbRows);
  m->ids.emplace_back(object_id);
  m->h_desc.insert(m->h_desc.end(), descriptors, descriptors + size_t(n) * 32);
  m->h_pts.insert(m->h_pts.end(), points, points + size_t(n) * 3);
  m->offsets.push_back(uint32_t(total));
  // span of the object: DescriptorMatcher.cpp:106-121 (bounding-box diagonal, float arithmetic)
  float min_x = std::numeric_limits<float>::max(), max_x = -std::numeric_limits<float>::max(), min_y = min_x,
        max_y = max_x, min_z = min_x, max_z = max_x;
  for (int32_t i = 0; i < n; ++i) {
    const float *v = points + size_t(i) * 3;
    min_x = std::min(min_x, v[0]); max_x = std::max(max_x, v[0]);
    min_y = std::min(min_y, v[1]); max_y = std::max(max_y, v[1]);
    min_z = std::min(min_z, v[2]); max_z = std::max(max_z, v[2]);
  }
  const float max_span_sq =
      (max_x - min_x) * (max_x - min_x) + (max_y - min_y) * (max_y - min_y) + (max_z - min_z) * (max_z - min_z);
  m->spans.push_back(std::sqrt(max_span_sq));
  m->trained = false;
  return TOD_OK;
}

int tod_matcher_clear(tod_matcher *m) {
  TOD_REQUIRE(m, "null argument");
  m->ids.clear();
  m->spans.clear();
  m->offsets.assign(1, 0u);
  m->h_desc.clear();
  m->h_pts.clear();
  m->trained = false;
  m->shard_begin = m->shard_rows = 0;
  return TOD_OK;
}

int tod_matcher_train(tod_matcher *m) {
  TOD_REQUIRE(m, "null argument");
  if (int rc = use_device(m)) return rc;
  const int64_t total = m->offsets.back();
  if (int rc = tod_shard_range(total, m->p.shard_rank, m->p.shard_count, &m->shard_begin, &m->shard_rows)) return rc;
  TOD_CUDA(m->d_db.reserve(std::max<size_t>(size_t(m->shard_rows) * 32, 256)));
  TOD_CUDA(m->d_pts.reserve(std::max<size_t>(size_t(total) * 3 * sizeof(float), 256)));
  TOD_CUDA(m->d_offsets.reserve(m->offsets.size() * sizeof(uint32_t)));
  if (m->shard_rows)
    TOD_CUDA(cudaMemcpyAsync(m->d_db.ptr, m->h_desc.data() + size_t(m->shard_begin) * 32, size_t(m->shard_rows) * 32,
                             cudaMemcpyHostToDevice, m->stream));
  if (total)
    TOD_CUDA(cudaMemcpyAsync(m->d_pts.ptr, m->h_pts.data(), size_t(total) * 3 * sizeof(float),
                             cudaMemcpyHostToDevice, m->stream));
  TOD_CUDA(cudaMemcpyAsync(m->d_offsets.ptr, m->offsets.data(), m->offsets.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, m->stream));
  // wide database: segments of 2^23 rows per shard, the same count on every rank (shards are ceil(total / world) rows)
  m->wide = total > tod::kMaxGlobalRows;
  m->n_seg = 1;
  if (m->wide) {
    const int64_t per = (total + m->p.shard_count - 1) / m->p.shard_count;
    m->n_seg = int((per + tod::kMaxGlobalRows - 1) / tod::kMaxGlobalRows);
    std::vector<uint32_t> base(size_t(m->p.shard_count) * size_t(m->n_seg));
    for (int r = 0; r < m->p.shard_count; ++r)
      for (int sg = 0; sg < m->n_seg; ++sg)
        base[size_t(r) * m->n_seg + sg] =
            uint32_t(std::min<int64_t>(total, per * r) + int64_t(sg) * tod::kMaxGlobalRows);
    TOD_CUDA(m->d_src_base.reserve(base.size() * sizeof(uint32_t)));
    TOD_CUDA(cudaMemcpy(m->d_src_base.ptr, base.data(), base.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  m->have_db8 = false;
  if (use_mma(m)) {
    static_assert(sizeof(m->map_db) >= 128, "CUtensorMap is 128 bytes");
    if (tod::tensor_map_bytes() > sizeof(m->map_db)) return fail(TOD_ERR_CUDA, "unexpected CUtensorMap size");
    TOD_CUDA(m->d_db8.reserve(std::max<size_t>(size_t(m->shard_rows) * 256, 1024)));
    TOD_CUDA(tod::launch_expand_db(m->d_db.ptr, m->d_db8.ptr, m->shard_rows, m->stream));
    if (!tod::make_desc_tensor_map(m->map_db, m->d_db8.ptr, m->shard_rows, tod::k1_mma_db_box_rows()))
      return fail(TOD_ERR_CUDA, "cuTensorMapEncodeTiled failed for the database");
    m->have_db8 = true;
  }
  TOD_CUDA(cudaStreamSynchronize(m->stream));
  m->trained = true;
  return TOD_OK;
}

int32_t tod_matcher_num_objects(const tod_matcher *m) { return m ? int32_t(m->ids.size()) : 0; }
int64_t tod_matcher_num_descriptors(const tod_matcher *m) { return m ? int64_t(m->offsets.back()) : 0; }
int64_t tod_matcher_shard_rows(const tod_matcher *m) { return m ? m->shard_rows : 0; }
const char *tod_matcher_object_id(const tod_matcher *m, int32_t i) {
  return (m && i >= 0 && size_t(i) < m->ids.size()) ? m->ids[size_t(i)].c_str() : "";
}
float tod_matcher_span(const tod_matcher *m, int32_t i) {
  return (m && i >= 0 && size_t(i) < m->spans.size()) ? m->spans[size_t(i)] : 0.f;
}
int32_t tod_matcher_k(const tod_matcher *m) { return m ? m->p.k : 0; }

int tod_matcher_knn(tod_matcher *m, const uint8_t *descriptors, int32_t nq, tod_match *matches, int32_t *counts,
                    float *points3d) {
  TOD_REQUIRE(m && matches && counts, "null argument");
  TOD_REQUIRE(nq >= 0 && (nq == 0 || descriptors), "bad query buffer");
  if (!m->trained) return fail(TOD_ERR_STATE, "tod_matcher_knn called before tod_matcher_train");
  if (nq == 0) return TOD_OK;
  if (int rc = use_device(m)) return rc;
  const int k = m->p.k;
  const size_t nk = size_t(nq) * k;
  TOD_CUDA(m->d_query.reserve(size_t(nq) * 32));
  TOD_CUDA(m->d_matches.reserve(nk * sizeof(tod_match)));
  TOD_CUDA(m->d_counts.reserve(size_t(nq) * sizeof(int32_t)));
  TOD_CUDA(m->d_pts3d.reserve(nk * 3 * sizeof(float)));
  TOD_CUDA(cudaMemcpyAsync(m->d_query.ptr, descriptors, size_t(nq) * 32, cudaMemcpyHostToDevice, m->stream));
  if (int rc = run_process(m, m->d_query.ptr, nq, m->d_matches.as<tod_match>(), m->d_counts.as<int32_t>(),
                           points3d ? m->d_pts3d.as<float>() : nullptr, m->stream))
    return rc;
  TOD_CUDA(cudaMemcpyAsync(matches, m->d_matches.ptr, nk * sizeof(tod_match), cudaMemcpyDeviceToHost, m->stream));
  TOD_CUDA(cudaMemcpyAsync(counts, m->d_counts.ptr, size_t(nq) * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
  if (points3d)
    TOD_CUDA(cudaMemcpyAsync(points3d, m->d_pts3d.ptr, nk * 3 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  TOD_CUDA(cudaStreamSynchronize(m->stream));
  if (m->comm_mode == 2) {
    uint32_t err = 0;
    TOD_CUDA(cudaMemcpy(&err, m->d_peer_tbl.as<uint64_t>() + 2 * m->p.shard_count + 1, 4, cudaMemcpyDeviceToHost));
    if (err) return fail(TOD_ERR_STATE, "peer exchange timed out: a rank of the communicator did not deliver its keys");
  }
  return TOD_OK;
}

int tod_matcher_knn_device(tod_matcher *m, const void *d_descriptors, int32_t nq, tod_match *d_matches,
                           int32_t *d_counts, float *d_points3d, void *stream) {
  TOD_REQUIRE(m && d_matches && d_counts, "null argument");
  TOD_REQUIRE(nq >= 0 && (nq == 0 || d_descriptors), "bad query buffer");
  if (!m->trained) return fail(TOD_ERR_STATE, "tod_matcher_knn_device called before tod_matcher_train");
  if (nq == 0) return TOD_OK;
  if (int rc = use_device(m)) return rc;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : m->stream;
  return run_process(m, d_descriptors, nq, d_matches, d_counts, d_points3d, st);
}

int tod_matcher_reserve(tod_matcher *m, int32_t max_nq) {
  TOD_REQUIRE(m && max_nq >= 0, "bad argument");
  if (!m->trained) return fail(TOD_ERR_STATE, "tod_matcher_reserve called before tod_matcher_train");
  if (int rc = use_device(m)) return rc;
  const size_t nq = size_t(std::max(max_nq, 1)), k = size_t(m->p.k), nk = nq * k;
  // the candidate buffer depends on the chunk plan, which depends on nq: take the largest over every query-tile count
  size_t partial = 0;
  const int step = 256;
  for (int64_t q = step; q < int64_t(nq) + step; q += step) {
    const int qq = int(std::min<int64_t>(q, int64_t(nq)));
    for (int64_t rows : {std::min<int64_t>(m->shard_rows, tod::kMaxGlobalRows),
                         m->shard_rows % tod::kMaxGlobalRows ? m->shard_rows % tod::kMaxGlobalRows : m->shard_rows}) {
      const tod::K1Plan pl = use_mma(m) ? tod::k1_mma_plan(qq, rows, m->sm_count)
                                        : tod::k1_popc_plan(qq, rows, m->sm_count);
      partial = std::max(partial, size_t(pl.n_sources) * size_t(qq) * k * sizeof(uint32_t));
    }
  }
  TOD_CUDA(m->d_partial.reserve(partial));
  TOD_CUDA(m->d_query.reserve(nq * 32));
  TOD_CUDA(m->d_matches.reserve(nk * sizeof(tod_match)));
  TOD_CUDA(m->d_counts.reserve(nq * sizeof(int32_t)));
  TOD_CUDA(m->d_pts3d.reserve(nk * 3 * sizeof(float)));
  if (use_mma(m)) {
    TOD_CUDA(m->d_q8.reserve(nq * 256));
    TOD_CUDA(m->d_gthr.reserve(nq * sizeof(uint32_t)));
    TOD_CUDA(m->d_popq.reserve(nq * sizeof(uint32_t)));
  }
  if (m->p.shard_count > 1 || m->wide) {
    TOD_CUDA(m->d_keys_local.reserve(nk * sizeof(uint32_t) * size_t(m->n_seg)));
    if (m->p.shard_count > 1)
      TOD_CUDA(m->d_keys_all.reserve(nk * sizeof(uint32_t) * size_t(m->n_seg) * size_t(m->p.shard_count)));
  }
  if (m->p.remove_duplicates) {
    size_t slots = 1024;
    while (slots < 2 * nk) slots <<= 1;
    TOD_CUDA(m->d_rows.reserve(nk * sizeof(uint32_t)));
    TOD_CUDA(m->d_hkeys.reserve(slots * 8));
    TOD_CUDA(m->d_hvals.reserve(slots * 8));
  }
  m->reserved_nq = std::max(m->reserved_nq, max_nq);
  return TOD_OK;
}

int tod_comm_unique_id(void *id_out) {
  TOD_REQUIRE(id_out, "null argument");
  const tod::NcclApi &api = tod::nccl_api();
  if (!api.ok) return fail(TOD_ERR_STATE, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbols");
  tod::ncclUniqueId id;
  TOD_NCCL(api.GetUniqueId(&id));
  static_assert(sizeof(id) == TOD_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  std::memcpy(id_out, &id, sizeof(id));
  return TOD_OK;
}

int32_t tod_matcher_comm_mode(const tod_matcher *m) { return m ? m->comm_mode : 0; }

int tod_matcher_set_comm(tod_matcher *m, const void *unique_id) {
  TOD_REQUIRE(m && unique_id, "null argument");
  const tod::NcclApi &api = tod::nccl_api();
  if (!api.ok) return fail(TOD_ERR_STATE, "libnccl.so.2 could not be loaded");
  if (int rc = use_device(m)) return rc;
  TOD_CUDA(cudaStreamSynchronize(m->stream));
  close_comm(m);
  tod::ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  const int world = m->p.shard_count, rank = m->p.shard_rank;
  TOD_NCCL(api.CommInitRank(&m->comm, world, id, rank));
  m->comm_mode = 1;
  return TOD_OK;
}

int tod_matcher_set_exchange(tod_matcher *m, int32_t peer_memory) {
  TOD_REQUIRE(m, "null argument");
  if (int rc = use_device(m)) return rc;
  TOD_CUDA(cudaStreamSynchronize(m->stream));
  m->peer_enabled = peer_memory != 0;
  if (!m->peer_enabled) close_peer_exchange(m);
  m->peer_tried = false;
  return TOD_OK;
}

int32_t tod_matcher_exchange_error(const tod_matcher *m) {
  if (!m || m->comm_mode != 2) return 0;
  uint32_t err = 0;
  if (cudaMemcpy(&err, m->d_peer_tbl.as<uint64_t>() + 2 * m->p.shard_count + 1, 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return int32_t(err);
}

void tod_matcher_set_stage_timing(tod_matcher *m, int32_t on) {
  if (m) m->stage_timing = on != 0;
}

float tod_matcher_last_exchange_ms(const tod_matcher *m) {
  if (!m || !m->ev_x_valid) return -1.f;
  if (cudaEventSynchronize(m->ev_x1) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, m->ev_x0, m->ev_x1) != cudaSuccess) return -1.f;
  return ms;
}

int tod_matcher_knn_keys_device(tod_matcher *m, const void *d_descriptors, int32_t nq, uint32_t *d_keys,
                                void *stream) {
  TOD_REQUIRE(m && d_keys, "null argument");
  TOD_REQUIRE(nq >= 0 && (nq == 0 || d_descriptors), "bad query buffer");
  if (!m->trained) return fail(TOD_ERR_STATE, "tod_matcher_knn_keys_device called before tod_matcher_train");
  if (m->wide)
    return fail(TOD_ERR_LIMIT, "the stage calls carry 32-bit global keys (2^23 rows); this database is wider: use "
                               "tod_matcher_knn / tod_matcher_knn_device");
  if (nq == 0) return TOD_OK;
  if (int rc = use_device(m)) return rc;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : m->stream;
  tod::K1Plan plan;
  if (int rc = run_k1(m, d_descriptors, nq, st, &plan)) return rc;
  TOD_CUDA(tod::launch_reduce_keys(m->d_partial.as<uint32_t>(), plan.n_sources, nq, m->p.k, d_keys, st));
  return TOD_OK;
}

int tod_matcher_merge_device(tod_matcher *m, const uint32_t *d_keys_all, int32_t n_src, int32_t nq,
                             tod_match *d_matches, int32_t *d_counts, float *d_points3d, void *stream) {
  TOD_REQUIRE(m && d_keys_all && d_matches && d_counts, "null argument");
  TOD_REQUIRE(n_src >= 1 && nq >= 0, "bad sizes");
  if (!m->trained) return fail(TOD_ERR_STATE, "tod_matcher_merge_device called before tod_matcher_train");
  if (m->wide)
    return fail(TOD_ERR_LIMIT, "the stage calls carry 32-bit global keys (2^23 rows); this database is wider: use "
                               "tod_matcher_knn / tod_matcher_knn_device");
  if (nq == 0) return TOD_OK;
  if (int rc = use_device(m)) return rc;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : m->stream;
  return finalize(m, d_keys_all, n_src, nq, d_matches, d_counts, d_points3d, st);
}

float tod_matcher_k1_ms_ago(const tod_matcher *m, int32_t calls_ago) {
  if (!m || calls_ago < 0 || calls_ago >= tod_matcher::kEvRing || uint64_t(calls_ago) >= m->k1_calls) return -1.f;
  const int er = int((m->k1_calls - 1 - uint64_t(calls_ago)) % tod_matcher::kEvRing);
  if (cudaEventSynchronize(m->ev1[er]) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, m->ev0[er], m->ev1[er]) != cudaSuccess) return -1.f;
  return ms;
}

float tod_matcher_last_k1_ms(const tod_matcher *m) { return tod_matcher_k1_ms_ago(m, 0); }

const char *tod_matcher_last_kernel(const tod_matcher *m) { return m ? m->last_kernel : "none"; }

}  // extern "C"
