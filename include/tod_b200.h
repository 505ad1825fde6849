/*
 * tod_b200.h — C-ABI of libtod_b200.so: the B200 (sm_100a) implementation of TOD's detection hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one piece of the reference's two ecto cells
 * (reference = wg-perception/tod 0.5.6; file:line below are relative to the reference tree):
 *
 *   tod::DescriptorMatcher   src/detection/DescriptorMatcher.cpp:58-270   -> tod_matcher_*
 *   tod::GuessGenerator      src/detection/GuessGenerator.cpp:69-276      -> tod_guess_*
 *   tod::AdjacencyRansac     src/common/adjacency_ransac.{h,cpp}          -> tod_fill_adjacency, tod_score_hypotheses,
 *                                                                            tod_guess_process
 *
 * Conventions: plain pointers and sizes only, caller-owned buffers, int status returns (TOD_OK == 0),
 * thread-local tod_last_error(), no exceptions cross the ABI.  A handle is NOT thread-safe — the reference's
 * scheduler calls the cells serially on one thread.  There is no CPU fallback: every compute entry
 * point fails with TOD_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef TOD_B200_H_
#define TOD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOD_B200_ABI_VERSION 2

/* ---- status codes -------------------------------------------------------------------------------------------- */
enum {
  TOD_OK = 0,
  TOD_ERR_INVALID = 1,   /* bad argument (null pointer, negative size, k out of range, ...) */
  TOD_ERR_STATE = 2,     /* call order violated (knn before train, ...) */
  TOD_ERR_CUDA = 3,      /* CUDA runtime error or no usable device */
  TOD_ERR_LIMIT = 4,     /* a documented capacity limit was exceeded */
  TOD_ERR_PARSE = 5      /* malformed JSON parameter string */
};

/* Message of the last failing call on this thread ("" if none). Never NULL. */
const char *tod_last_error(void);
int tod_abi_version(void);
/* Number of kernels launched by this library in this process so far (bench.py reports it as gpu_launches). */
uint64_t tod_kernel_launch_count(void);

/* ---- POD mirrors of the OpenCV / ORK types that cross the cell boundary -------------------- */

/* cv::DMatch {int queryIdx; int trainIdx; int imgIdx; float distance;}  — 16 bytes, same field order. */
typedef struct tod_match {
  int32_t queryIdx;
  int32_t trainIdx; /* row inside the object's descriptor matrix */
  int32_t imgIdx;   /* object index (order of tod_matcher_add_object calls) */
  float distance;   /* Hamming distance, an exact integer in [0,256] */
} tod_match;

/* cv::KeyPoint {Point2f pt; float size; float angle; float response; int octave; int class_id;} — 28 bytes. */
typedef struct tod_keypoint {
  float x, y;
  float size, angle, response;
  int32_t octave, class_id;
} tod_keypoint;

/* object_recognition_core::common::PoseResult as filled at GuessGenerator.cpp:224-228 (R, T, object id). */
typedef struct tod_pose {
  float R[9];            /* row-major 3x3, object -> camera (adjacency_ransac.cpp:304) */
  float T[3];            /* metres (adjacency_ransac.cpp:305) */
  int32_t object_index;  /* index into the matcher's object_ids (GuessGenerator.cpp:175-176) */
  int32_t n_inliers;     /* number of distinct inlier keypoints (query_inliers.size(), GuessGenerator.cpp:202) */
} tod_pose;

/* ================================================================================================================
 * DescriptorMatcher  (DescriptorMatcher.cpp)
 * ============================================================================================================== */

typedef struct tod_matcher tod_matcher;

enum { TOD_SEARCH_EXACT = 0, TOD_SEARCH_LSH = 1 /* accepted for .ork compatibility; runs the exact search */ };
enum { TOD_KERNEL_AUTO = 0, TOD_KERNEL_POPC = 1, TOD_KERNEL_MMA = 2 };
#define TOD_MAX_K 8

typedef struct tod_matcher_params {
  int32_t k;            /* neighbours per query, 1..TOD_MAX_K. The reference hard-codes 5 (DescriptorMatcher.cpp:211) */
  uint32_t radius;      /* keep matches with distance <= radius; 0 = no radius cut.
                           unsigned int like DescriptorMatcher.cpp:257 */
  int32_t search_type;  /* TOD_SEARCH_* (DescriptorMatcher.cpp:175) */
  int32_t device;       /* CUDA device ordinal */
  int32_t shard_rank;   /* this handle holds rows [rank*ceil(N/count), ...) of the concatenated DB */
  int32_t shard_count;  /* 1 = whole DB on this GPU */
  int32_t kernel;       /* TOD_KERNEL_* : which K1 formulation to run */
  /* --- the two TODO blocks of DescriptorMatcher.cpp:223-229, implemented as opt-in extensions (default off: the
   *     reference's `ratio_` is an unsigned int, so the .ork value 0.8 truncates to 0 and its block is empty) --- */
  int32_t ratio_enabled;      /* 1: Lowe's ratio test on the two nearest neighbours, BEFORE the radius cut: a query
                                 keeps at most its best match, and only if distance0 < ratio * distance1 (float
                                 arithmetic, the OpenCV knnMatch(k=2) idiom); needs k >= 2.  A query with a single
                                 neighbour in the whole DB keeps it. */
  float ratio;                /* search_json_params.ratio as a real number (conf/detection.ork:39: 0.8) */
  int32_t remove_duplicates;  /* 1: "remove matches that match the same (common descriptors)" (:229): when several
                                 keypoints of one frame match the same DB descriptor, only the match with the
                                 smallest (distance, queryIdx) survives; lists are compacted, order kept. */
  int32_t frame_keypoints;    /* scope of remove_duplicates in a batched call: queries [f*n, (f+1)*n) are frame f;
                                 0 = the whole call is one frame */
} tod_matcher_params;

void tod_matcher_default_params(tod_matcher_params *p); /* k=5, radius=0, exact, device 0, 1 shard, auto, no ratio
                                                           test, no duplicate removal */

/* configure(): parses the reference's "search_json_params" string — fields type, radius, ratio, n_tables, key_size,
 * multi_probe_level (DescriptorMatcher.cpp:159-181) — into *p (k stays 5, as hard-coded in the reference).
 * type "LSH" is accepted and mapped to the exact search; unknown types fail with TOD_ERR_INVALID instead of the
 * reference's bare `throw;` -> std::terminate (:182-186).  `ratio` is stored as a real number but only applied when
 * the (non-reference) keys "ratio_enabled": true / "remove_duplicates": true are present. */
int tod_matcher_params_from_json(const char *search_json_params, tod_matcher_params *p);

int tod_matcher_create(const tod_matcher_params *p, tod_matcher **out);
void tod_matcher_destroy(tod_matcher *m);

/* parameter_callback(), one DB document per call (DescriptorMatcher.cpp:70-123): descriptors = n x 32 u8 row-major
 * (training.cpp:157), points = n x 3 f32 (the 1 x N CV_32FC3 "points" attachment, training.cpp:158). Data is copied.
 * The object's span (:106-121) is computed here.  imgIdx = order of calls.  At most 2^31 - 1 descriptors in all
 * (TOD_ERR_LIMIT beyond); above 2^23 the database is scanned in segments of 2^23 rows, transparently. */
int tod_matcher_add_object(tod_matcher *m, const char *object_id, const uint8_t *descriptors, const float *points,
                           int32_t n);
/* matcher_->clear() (:127) */
int tod_matcher_clear(tod_matcher *m);
/* matcher_->add(descriptors_db_) + (lazy) train (:128): concatenates objects in imgIdx order, selects this handle's
 * shard, uploads descriptors (sharded) and points/object table (replicated) to HBM. */
int tod_matcher_train(tod_matcher *m);

int32_t tod_matcher_num_objects(const tod_matcher *m);
int64_t tod_matcher_num_descriptors(const tod_matcher *m);       /* whole DB */
int64_t tod_matcher_shard_rows(const tod_matcher *m);            /* rows resident on this GPU */
const char *tod_matcher_object_id(const tod_matcher *m, int32_t object_index); /* outputs["object_ids"] (:248) */
float tod_matcher_span(const tod_matcher *m, int32_t object_index);            /* outputs["spans"] (:249) */
int32_t tod_matcher_k(const tod_matcher *m);

/* process() (DescriptorMatcher.cpp:195-252) with HOST buffers: knnMatch(k) -> radius cut -> matches_3d gather.
 *   descriptors: nq x 32 u8.   matches: nq x k (row q holds counts[q] valid entries, sorted by
 *   (distance, imgIdx, trainIdx) exactly like cv::BFMatcher(NORM_HAMMING).knnMatch).  counts: nq.
 *   points3d: nq x k x 3 f32 (matches_3d), may be NULL.
 * With shard_count > 1 the handle needs a communicator (tod_matcher_set_comm); every rank calls with the same
 * descriptors and every rank receives the same, complete result: K1 on the shard -> top-k reduction -> ncclAllGather
 * of the packed keys -> merge, all on the handle's stream. */
int tod_matcher_knn(tod_matcher *m, const uint8_t *descriptors, int32_t nq, tod_match *matches, int32_t *counts,
                    float *points3d);
/* The same call with DEVICE buffers on the handle's device (descriptors already in HBM, results stay in HBM),
 * enqueued on `stream` (a cudaStream_t; NULL = the handle's own stream) without synchronising it.  d_points3d may be
 * NULL.  Works for shard_count == 1 and, with a communicator, for sharded handles. */
int tod_matcher_knn_device(tod_matcher *m, const void *d_descriptors, int32_t nq, tod_match *d_matches,
                           int32_t *d_counts, float *d_points3d, void *stream);
/* Pre-size every per-call device buffer for up to max_nq queries, so that no cudaMalloc / cudaFree (an implicit device
 * synchronisation) can happen inside a streamed step.  Call after tod_matcher_train. */
int tod_matcher_reserve(tod_matcher *m, int32_t max_nq);

/* ---- communicator of a sharded matcher (one process per GPU) ----------------------------------------------------
 * tod_comm_unique_id: rank 0 creates a 128-byte NCCL unique id (ncclGetUniqueId) and hands it to the other ranks by
 * any host-side means (MPI, torch.distributed broadcast, a file).  tod_matcher_set_comm: collective over all
 * shard_count ranks — ncclCommInitRank on the handle's device with rank = shard_rank, world = shard_count.
 * libnccl.so.2 is loaded with dlopen at this point; the library has no link-time NCCL dependency.  Destroying the
 * handle tears the communicator down locally (ncclCommAbort after the stream has drained): no rank waits for another. */
#define TOD_COMM_ID_BYTES 128
int tod_comm_unique_id(void *id_out);
int tod_matcher_set_comm(tod_matcher *m, const void *unique_id);
/* 0 = no communicator, 1 = NCCL communicator attached (exchange = ncclAllGather of the packed keys), 2 = NCCL
 * communicator + peer-memory exchange: at the first sharded call the ranks trade CUDA IPC handles through the
 * communicator and map each other's exchange buffers; from then on the top-k reduction kernel stores every rank's keys
 * straight into all the other GPUs over NVLink and raises a flag there, and the merge kernel starts as soon as every
 * flag has arrived — no NCCL kernel, no copy engine on the data path.  If any rank cannot map a peer (no peer access,
 * one process holding several handles) all ranks stay in mode 1.  Results are identical in both modes. */
int32_t tod_matcher_comm_mode(const tod_matcher *m);
/* peer_memory = 0 keeps the ncclAllGather path (A/B measurements); 1 (default) lets the next sharded call try the
 * peer-memory exchange.  Collective in effect: call it with the same value on every rank, between steps. */
int tod_matcher_set_exchange(tod_matcher *m, int32_t peer_memory);
/* 1 if a merge kernel of the peer-memory exchange gave up waiting for a rank (about 4 s) since the communicator was
 * attached — the results of that call are invalid; tod_matcher_knn checks it itself and returns TOD_ERR_STATE. */
int32_t tod_matcher_exchange_error(const tod_matcher *m);
/* Device time (ms, CUDA events on the launch stream) of the ncclAllGather of the last sharded call; < 0 if none.
 * The two extra event records cost a few microseconds per step, so they are off unless stage timing is switched on
 * (profiling runs); the K1 events behind tod_matcher_last_k1_ms are always recorded. */
void tod_matcher_set_stage_timing(tod_matcher *m, int32_t on);
float tod_matcher_last_exchange_ms(const tod_matcher *m);

/* Sharding of the concatenated DB over `shard_count` GPUs (host-only, no device needed): rank r holds the contiguous
 * global rows [*begin, *begin + *rows), ceil(total/shard_count) rows each except the last ranks.  tod_matcher_train
 * uses exactly this split.  tod_pack_key builds the u32 candidate key (distance << 23 | global row) that the stage
 * calls below exchange; it addresses 2^23 rows, so those calls return TOD_ERR_LIMIT on a wider database (the
 * whole-call entry points handle it: their keys are local to a 2^23-row segment). */
int tod_shard_range(int64_t total_rows, int32_t shard_rank, int32_t shard_count, int64_t *begin, int64_t *rows);
uint32_t tod_pack_key(uint32_t distance, uint32_t global_row);

/* Device-resident stages of the same call, for the sharded (one process per GPU) path. All pointers are device
 * pointers on the handle's device; `stream` is a cudaStream_t (NULL = the handle's own stream).
 *   knn_keys: this shard's top-k per query as packed keys (distance << 23 | global_row), ascending, padded with
 *             0xFFFFFFFF.  d_keys: nq x k u32.  min() over keys == the (distance, imgIdx, trainIdx) order.
 *   merge:    d_keys_all = n_src x nq x k gathered keys (e.g. after an NCCL all-gather) -> final matches.
 */
int tod_matcher_knn_keys_device(tod_matcher *m, const void *d_descriptors, int32_t nq, uint32_t *d_keys, void *stream);
int tod_matcher_merge_device(tod_matcher *m, const uint32_t *d_keys_all, int32_t n_src, int32_t nq,
                             tod_match *d_matches, int32_t *d_counts, float *d_points3d, void *stream);
/* Device time (ms, CUDA events on the launch stream) of the K1 kernel alone in the last knn call; <0 if unknown. */
float tod_matcher_last_k1_ms(const tod_matcher *m);
/* The same for the call made `calls_ago` calls before the last one (0 = the last; the library keeps 64 event pairs):
 * a streaming caller reads the K1 time of every step after its loop instead of synchronising inside it. */
float tod_matcher_k1_ms_ago(const tod_matcher *m, int32_t calls_ago);
/* Name of the K1 formulation actually used by the last call ("popc" | "mma"). */
const char *tod_matcher_last_kernel(const tod_matcher *m);

/* ================================================================================================================
 * Trained-DB snapshot — what parameter_callback (DescriptorMatcher.cpp:60-129) reads from the object DB, as one flat
 * mmap-able file (layout: tod_b200/csrc/snapshot.cpp).  Host-only entry points: no GPU needed.
 * ============================================================================================================== */

typedef struct tod_snapshot tod_snapshot;

/* Write n_objects models: object_ids[o], descriptors[o] = rows[o] x 32 u8 (attachment "descriptors",
 * training.cpp:157), points[o] = rows[o] x 3 f32 (attachment "points", training.cpp:158).  Spans are stored too. */
int tod_snapshot_write(const char *path, int32_t n_objects, const char *const *object_ids,
                       const uint8_t *const *descriptors, const float *const *points, const int32_t *rows);
int tod_snapshot_open(const char *path, tod_snapshot **out);   /* mmap + validation */
void tod_snapshot_close(tod_snapshot *s);
int32_t tod_snapshot_num_objects(const tod_snapshot *s);
int64_t tod_snapshot_num_descriptors(const tod_snapshot *s);
/* Pointers into the mapping (valid until close); any output may be NULL. */
int tod_snapshot_object(const tod_snapshot *s, int32_t index, const char **object_id, const uint8_t **descriptors,
                        const float **points, int32_t *rows, float *span);
/* parameter_callback from a snapshot: clear(), then add_object() for every model in file order; call train() next. */
int tod_matcher_load_snapshot(tod_matcher *m, const char *path);

/* ================================================================================================================
 * Feature stage in front of the hot path: what TodDetector wires before the DescriptorMatcher —
 * ecto_opencv's FeatureDescriptor cell = cv::ORB (python/object_recognition_tod/detector.py:27,74;
 * conf/detection.ork:23-31: n_features 5000, n_levels 3, scale_factor 1.2) and DepthTo3d (detector.py:62-69).
 * Detection (FAST + Harris), orientation, descriptors and depth -> 3-D on the GPU, bit-exact against cv2.ORB (same
 * keypoint set, angles and responses as exact floats, descriptors bit for bit; tests/test_orb_gpu.py).  The descriptors
 * stay in HBM: *d_descriptors can be handed to tod_matcher_knn_device.
 * ============================================================================================================== */

typedef struct tod_orb tod_orb;
typedef struct tod_orb_params {
  int32_t n_levels;     /* pyramid levels, default 3 (conf/detection.ork:27) */
  float scale_factor;   /* default 1.2 (conf/detection.ork:28) */
  int32_t device;
  int32_t reserved;
} tod_orb_params;

void tod_orb_default_params(tod_orb_params *p);
int tod_orb_create(const tod_orb_params *p, tod_orb **out);
void tod_orb_destroy(tod_orb *o);
/* image: height x width u8 (host).  keypoints (host): x, y in level-0 pixel coordinates and octave are read; with
 * compute_angles != 0 the angle field (degrees) is overwritten with ORB's intensity-centroid orientation, else it is
 * used as given.  Keypoints closer than 22 pixels to the border of their pyramid level are refused (cv::ORB drops
 * them: edgeThreshold 31).  descriptors (host, n x 32 u8) may be NULL; *d_descriptors (may be NULL) receives the device
 * pointer of the same n x 32 bytes, valid until the next call on this handle. */
int tod_orb_describe(tod_orb *o, const uint8_t *image, int32_t height, int32_t width, tod_keypoint *keypoints,
                     int32_t n, int32_t compute_angles, uint8_t *descriptors, const void **d_descriptors);
/* The whole cell: cv::ORB::detectAndCompute with ORB's defaults (FAST threshold 20, edgeThreshold 31, Harris score,
 * patch 31) — FAST-9/16 + 3 x 3 non-maximum suppression on every level, the 2 N best FAST scores, Harris responses, the N
 * best per level, orientation, descriptors.  keypoints (capacity max_keypoints; ties at the cut can exceed n_features)
 * are returned ordered by (octave, row, column) — cv2's own order is an artefact of nth_element — with x, y in level-0
 * coordinates, size = 31 * scale, angle, response = Harris, octave, class_id = -1.  The SET of keypoints, every field
 * and every descriptor equal cv2.ORB_create(n_features, scale_factor, n_levels).detectAndCompute (tests/test_orb_gpu.py). */
int tod_orb_detect_and_compute(tod_orb *o, const uint8_t *image, int32_t height, int32_t width, int32_t n_features,
                               tod_keypoint *keypoints, int32_t max_keypoints, int32_t *n_keypoints,
                               uint8_t *descriptors, const void **d_descriptors);
/* The same with a detection mask (u8, height x width; keypoints on zero pixels are dropped level by level on ORB's mask
 * pyramid — cv::ORB::detectAndCompute(image, mask, ...)) and with grey (channels = 1) or BGR (channels = 3: converted
 * like cv::cvtColor(BGR2GRAY)) frames.  mask may be NULL. */
int tod_orb_detect_and_compute_masked(tod_orb *o, const uint8_t *image, int32_t channels, int32_t height,
                                      int32_t width, const uint8_t *mask, int32_t n_features, tod_keypoint *keypoints,
                                      int32_t max_keypoints, int32_t *n_keypoints, uint8_t *descriptors,
                                      const void **d_descriptors);
/* Reads back one pyramid level of the last processed frame (for stage-by-stage parity tests): kind 0 = the level as
 * resized, 1 = smoothed, 2 = its FAST corner scores (0 = no corner).  out = height x width u8, may be NULL to query the
 * level size only. */
int tod_orb_read_level(tod_orb *o, int32_t level, int32_t kind, uint8_t *out, int32_t *height, int32_t *width);
/* DepthTo3d: depth = height x width float32 metres (NaN = invalid) or, with depth_is_u16, uint16 millimetres (0 =
 * invalid); K = 3 x 3 row-major camera matrix (float); points3d = height x width x 3 f32 (host):
 * x = ((u - cx) (1 / fx)) z, y = ((v - cy) (1 / fy)) z, z — NaN where the depth is invalid: the `points3d` input of the
 * GuessGenerator. */
int tod_depth_to_3d(int32_t device, const void *depth, int32_t depth_is_u16, int32_t height, int32_t width,
                    const float *K, float *points3d);

/* ================================================================================================================
 * Offline training path: the Trainer cell (src/training/Trainer.cpp:121-187) with its helpers
 * (training.cpp:57-195) — per observation of an object: ORB features on the masked image, depth rescaled to the image,
 * keypoints validated against the eroded mask and the depth, back-projection, camera -> object frame, and the views
 * stacked into the model the DescriptorMatcher loads (descriptors N x 32 u8, points N x 3 f32: ModelFiller.cpp:23-24).
 * The DB view iteration (Trainer.cpp:124-133) is the caller's loop.  Keypoints come out ordered by (octave, row,
 * column), so the model holds the reference's (descriptor, point) pairs in a different row order.
 * ============================================================================================================== */

typedef struct tod_trainer tod_trainer;
typedef struct tod_trainer_params {
  int32_t n_features;   /* 500: Trainer.cpp:142-150 builds cv::ORB without parameters (its TODO), so OpenCV's defaults */
  int32_t n_levels;     /* 8 */
  float scale_factor;   /* 1.2 */
  int32_t device;
} tod_trainer_params;

void tod_trainer_default_params(tod_trainer_params *p);
int tod_trainer_create(const tod_trainer_params *p, tod_trainer **out);
void tod_trainer_destroy(tod_trainer *t);
/* One observation (the body of the loop, Trainer.cpp:134-171): image = height x width x channels u8 (1 grey, 3 BGR),
 * mask = height x width u8 (object pixels non-zero), depth = depth_height x depth_width float32 metres or uint16
 * millimetres (rescaled to the image like rescale_depth, :63-81), K = 3 x 3 camera matrix, R (3 x 3) and T (3) = the
 * observation's pose (cameraToWorld computes (p - T) * R, training.cpp:175-195).  *n_added = points appended. */
int tod_trainer_add_observation(tod_trainer *t, const uint8_t *image, int32_t channels, int32_t height, int32_t width,
                                const uint8_t *mask, const void *depth, int32_t depth_is_u16, int32_t depth_height,
                                int32_t depth_width, const float *K, const float *R, const float *T,
                                int32_t *n_added);
/* mergePoints (training.cpp:147-173): the stacked model so far; pointers stay valid until the next add / clear. */
int tod_trainer_model(const tod_trainer *t, const uint8_t **descriptors, const float **points, int64_t *n);
int64_t tod_trainer_num_points(const tod_trainer *t);
int tod_trainer_clear(tod_trainer *t);

/* ================================================================================================================
 * Geometry stages (adjacency_ransac.cpp, sac_model_registration_graph.h) — exposed for parity tests and reuse
 * ============================================================================================================== */

/* Words (u32) per bit-matrix row for a cluster of n correspondences: ceil(n/32) rounded up to a multiple of 4. */
int32_t tod_adjacency_row_words(int32_t n);

/* K2 — AdjacencyRansac::FillAdjacency (adjacency_ransac.cpp:127-172) for a batch of clusters, HOST buffers.
 *   cluster c owns correspondences [offsets[c], offsets[c+1]).
 *   query_pts/train_pts: N x 3 f32; pixels: N x 2 f32 (keypoints[query_indices_[i]].pt); spans: per cluster.
 *   physical/sample: per cluster a full symmetric n_c x row_words(n_c) u32 bit-matrix, clusters concatenated in
 *   order (bit j of row i set <=> the reference's neighbors(i) contains j).  matrix_offsets (n_clusters+1, in u32
 *   words, out) may be NULL. */
int tod_fill_adjacency(int32_t device, int32_t n_clusters, const int32_t *offsets, const float *query_pts,
                       const float *train_pts, const float *pixels, const float *spans, float sensor_error,
                       uint32_t *physical, uint32_t *sample, int64_t *matrix_offsets);

/* K3 — one RANSAC iteration body for a batch of hypotheses of ONE cluster, HOST buffers:
 * computeModelCoefficients (sac_model_registration_graph.h:271-288 -> :304-347) + the candidate/inlier part of
 * selectWithinDistance (:171-200), i.e. the PRE-gate inlier count.
 *   physical: n x row_words(n) bit-matrix; valid: row_words(n) bit-vector of still-valid correspondences;
 *   triples: H x 3 sample indices; threshold: the RANSAC distance threshold, +inf (or >= 1e150) reproduces the
 *   reference's never-set DBL_MAX (sac.h:70).
 *   counts: H (|candidates passing| incl. the 3 samples).  R: H x 9, T: H x 3 (query -> training frame), may be NULL. */
int tod_score_hypotheses(int32_t device, int32_t n, const float *query_pts, const float *train_pts,
                         const uint32_t *physical, const uint32_t *valid, int32_t n_hyp, const uint32_t *triples,
                         double threshold, int32_t *counts, float *R, float *T);

/* Device time (ms, CUDA events on the launch stream) of the K2 / K3 kernel launched by the last tod_fill_adjacency /
 * tod_score_hypotheses call on this thread; < 0 if unknown. */
float tod_last_stage_ms(void);

/* Host-only pieces of the guess generator (no GPU needed), exposed so that they can be pinned against the reference's
 * known-answer tests and compiled sources on a CPU-only machine:
 *   tod_clique_find  the bounded MaxCliqueDyn search of the clique gate (maximum_clique.cpp:286-369 as restated in
 *                    tod_b200/csrc/clique.h) on a graph given as an edge list (edges added in the given order, like
 *                    Graph::AddEdgeSorted).  Returns the clique size (-1 on bad input), the vertices in out_vertices
 *                    (capacity n_vertices, may be NULL); *finds_more (may be NULL) = the decision-only variant the
 *                    gate actually runs: would the search return MORE than minimal_size vertices?
 *   tod_rigid_fit    the m-point Kabsch fit of the refinement loop (estimateRigidTransformationSVD,
 *                    sac_model_registration_graph.h:304-347): R (9, row-major) and T (3) map query -> training. */
int32_t tod_clique_find(int32_t n_vertices, const int32_t *edges, int32_t n_edges, uint32_t minimal_size,
                        int32_t *out_vertices, int32_t *finds_more);
/*   tod_clique_gate_small  the fixed-capacity form of the same search that K5 runs on the GPU for induced sample
 *                    sub-graphs of at most 256 vertices (tod_b200/csrc/clique_small.h, one source for host and
 *                    device): the gate's question at minimal size 7 (sac_model_registration_graph.h:258-265).
 *                    Returns 1 (more than 7 vertices would be returned), 0 (not), -1 (step_cap reached: the caller
 *                    falls back to tod_clique_find), -2 on bad input; *steps (may be NULL) = search steps taken. */
int32_t tod_clique_gate_small(int32_t n_vertices, const int32_t *edges, int32_t n_edges, int32_t step_cap,
                              int32_t *steps);
/* Stage-level call of K5 (needs a B200): the same search on the device for a batch of graphs of 1..256 vertices given
 * as concatenated edge lists (graph g = edges[edge_offsets[g] .. edge_offsets[g + 1]), pairs of vertex ids) — same
 * kernel, job queue layout and 512-step cap as inside tod_guess_process, where K4 fills the queue.
 * results[g] = 1 (gate passes), 0 (fails), -1 (step cap reached: left to the host search). */
int tod_gate_search_device(int32_t device, int32_t n_graphs, const int32_t *n_vertices, const int32_t *edge_offsets,
                           const int32_t *edges, int32_t *results);
int tod_rigid_fit(const float *query_pts, const float *train_pts, const uint32_t *indices, int32_t m, float *R, float *T);
/*   tod_sample_triples  getSamples (sac_model_registration_graph.h:141-168): n_hyp sample triples from the sample
 *                    graph (n x tod_adjacency_row_words(n) bit-matrix) restricted to the valid mask, consuming the
 *                    stream *rng_state (tod_rng_seed / tod_rng_next).  Returns the number of triples produced. */
int32_t tod_sample_triples(int32_t n, const uint32_t *sample_bits, const uint32_t *valid_bits, uint64_t *rng_state,
                           int32_t n_hyp, uint32_t *triples);
/*   tod_select_inliers  selectWithinDistance (sac_model_registration_graph.h:171-269) with the reference's never-set
 *                    threshold: candidates from the physical bit-rows of the triple, then the clique gate on the
 *                    sample graph exactly as tod_guess_process runs it.  Returns the inlier count (sorted list in
 *                    `inliers`, capacity n + 3); 0 = the gate cleared the list; -1 on bad input. */
int32_t tod_select_inliers(int32_t n, const uint32_t *physical_bits, const uint32_t *sample_bits,
                           const uint32_t *valid_bits, const uint32_t *triple, uint32_t *inliers);

/* ================================================================================================================
 * GuessGenerator  (GuessGenerator.cpp)
 * ============================================================================================================== */

typedef struct tod_guess tod_guess;

typedef struct tod_guess_params {
  uint32_t min_inliers;          /* default 15 (GuessGenerator.cpp:74) */
  uint32_t n_ransac_iterations;  /* default 1000 (:75-76) */
  float sensor_error;            /* default 0.01 (:77) */
  int32_t device;
  double ransac_threshold;       /* +inf = reference-faithful (sac.h:70, quirk Q3) */
  uint64_t seed;                 /* sampler stream seed; stream is re-seeded per (object, round), see DESIGN.md */
  int32_t host_threads;          /* host threads for the per-object work (sampler, replay, clique gate, refinement);
                                    0 = min(16, hardware threads).  Results do not depend on it. */
  int32_t reserved;
} tod_guess_params;

void tod_guess_default_params(tod_guess_params *p);
int tod_guess_create(const tod_guess_params *p, tod_guess **out);
void tod_guess_destroy(tod_guess *g);

/* process() (GuessGenerator.cpp:127-250) with HOST buffers.
 *   keypoints: n_kp; cloud: H x W x 3 f32 ("points3d"), NaN where invalid;
 *   matches/counts/points3d: outputs of tod_matcher_knn for the same n_kp queries (row stride k);
 *   spans: per object index (n_objects).  poses: capacity max_poses; *n_poses = number produced (in the
 *   reference's emission order: ascending object index, then RANSAC round).
 *   inlier_keypoints (optional, may be NULL): concatenated sorted inlier keypoint indices of each pose, capacity
 *   max_inlier_total; pose p owns n_inliers entries in order. */
int tod_guess_process(tod_guess *g, const tod_keypoint *keypoints, int32_t n_kp, const float *cloud, int32_t height,
                      int32_t width, const tod_match *matches, const int32_t *counts, int32_t k,
                      const float *points3d, const float *spans, int32_t n_objects, tod_pose *poses,
                      int32_t max_poses, int32_t *n_poses, int32_t *inlier_keypoints, int32_t max_inlier_total);

/* The same for a BATCH of frames in one call (BASELINE config C4: a 64-frame stream): frame f owns keypoints
 * [kp_offsets[f], kp_offsets[f+1]) of the concatenated keypoints / matches / counts / points3d arrays and the f-th
 * height x width x 3 cloud of `clouds`.  All (frame, object) clusters go through K2 in one launch and through every
 * K3 round together, and their host work is spread over the handle's threads.  Results equal n_frames separate
 * tod_guess_process calls (the sampler stream depends on (seed, object, round) only): poses in (frame, object, round)
 * order, pose_frames[i] (may be NULL) = frame of pose i, inlier keypoint indices relative to the pose's frame. */
int tod_guess_process_batch(tod_guess *g, int32_t n_frames, const int32_t *kp_offsets, const tod_keypoint *keypoints,
                            const float *clouds, int32_t height, int32_t width, const tod_match *matches,
                            const int32_t *counts, int32_t k, const float *points3d, const float *spans,
                            int32_t n_objects, tod_pose *poses, int32_t *pose_frames, int32_t max_poses,
                            int32_t *n_poses, int32_t *inlier_keypoints, int32_t max_inlier_total);

/* Sampler stream used by tod_guess_process (public so a test harness can drive the reference's rand() with the
 * same numbers): state = tod_rng_seed(seed, object_index, round); tod_rng_next(&state) in [0, 2^31). */
uint64_t tod_rng_seed(uint64_t seed, uint32_t object_index, uint32_t round);
int32_t tod_rng_next(uint64_t *state);

/* Device time (ms) spent in K2 / K3 kernels and number of K3 hypotheses scored during the last process call. */
void tod_guess_last_stats(const tod_guess *g, float *k2_ms, float *k3_ms, int64_t *n_hypotheses, int32_t *n_rounds);
/* Shape of the clique gate's work in the last process call (24 counters): [0..7] gates that reached the induced
 * sub-graph stage by sub-graph size (buckets <= 16, 32, 64, 128, 256, 512, 1024, more), [8..15] the same by the size
 * of the sub-graph's 7-core, [16] settled because the core has fewer than 8 vertices, [17] settled by the colouring
 * bound, [18] bounded searches run, [19] of those, passes, [20] search steps summed, [21] hypotheses the replay
 * settled with a K4 "fails" verdict, [22] hypotheses K4 left undecided that the replay sent to the host search,
 * [23] device time of the K4 launches in microseconds. */
void tod_guess_last_gate_stats(const tod_guess *g, int64_t *out24);
/* K5 (the reference's bounded clique search stepped exactly on the GPU for induced sub-graphs of <= 256 vertices) in
 * the last process call: [0] hypotheses the replay settled with a K5 "passes" verdict, [1] with a K5 "fails" verdict,
 * [2] hypotheses left to the host search (larger graphs, full queue, step cap), [3] device time of the K5 launches in
 * microseconds (included in counter [23] of tod_guess_last_gate_stats). */
void tod_guess_last_k5_stats(const tod_guess *g, int64_t *out4);
/* Algorithmic bytes (SURVEY.md §8d units) moved by the K2 launch (32 n in + two n x W bit-matrices out, summed over
 * the clusters) and by all K3 launches ((3 rows + 2 masks) x W x 4 + 140 per hypothesis) of the last process call,
 * with the number of (frame, object) clusters and correspondences: bench.py divides them by k2_ms / k3_ms. */
void tod_guess_last_traffic(const tod_guess *g, double *k2_bytes, double *k3_bytes, int64_t *n_clusters,
                            int64_t *n_correspondences);
/* Host wall-clock profile (ms) of the last process call: [0] ClusterPerObject + upload + K2 + bit-matrix download,
 * [1] sampler, [2] K3 launches incl. copies and sync, [3] replay + inlier lists + clique gate, [4] refinement +
 * invalidation, [5] clique-gate evaluations (a count), [6] of those, settled by the exact no-8-clique proof without
 * running the search (a count), [7] total, [8] gate set-up, [9] gate proof, [10] gate search (summed over the host
 * threads), [11] gate evaluations settled by the round's 7-core test (a count).  ms12 must hold 12 doubles. */
void tod_guess_last_profile(const tod_guess *g, double *ms12);

#ifdef __cplusplus
}
#endif
#endif /* TOD_B200_H_ */
