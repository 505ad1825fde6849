"""Python view of the two cells of the hot path, named and parameterised like the reference's ecto cells
(src/detection/DescriptorMatcher.cpp:131-152, src/detection/GuessGenerator.cpp:71-99).  Pure plumbing over the C-ABI
(tod_b200/capi.py): numpy arrays in, numpy arrays out, all compute inside libtod_b200.so on the GPU."""
import ctypes
import json

import numpy as np

from . import capi


class DescriptorMatcher:
    """Mirror of tod::DescriptorMatcher.

    params: search_json_params (str, like conf/detection.ork's `search:` subtree serialised to JSON), or explicit
            k / radius.  inputs: descriptors (nq x 32 u8).  outputs: matches, matches_3d, object_ids, spans.
    """

    def __init__(self, search_json_params=None, k=None, radius=None, device=0, shard_rank=0, shard_count=1,
                 kernel=capi.TOD_KERNEL_AUTO, ratio=None, remove_duplicates=None, frame_keypoints=0):
        lib = capi.load()
        p = capi.MatcherParams()
        lib.tod_matcher_default_params(ctypes.byref(p))
        if search_json_params is not None:
            if not isinstance(search_json_params, str):
                search_json_params = json.dumps(search_json_params)
            capi.check(lib.tod_matcher_params_from_json(search_json_params.encode(), ctypes.byref(p)))
        if k is not None:
            p.k = int(k)
        if radius is not None:
            p.radius = int(radius)
        p.device, p.shard_rank, p.shard_count, p.kernel = int(device), int(shard_rank), int(shard_count), int(kernel)
        if ratio is not None:        # opt-in extension (the reference's ratio block is an empty TODO)
            p.ratio_enabled, p.ratio = (1, float(ratio)) if ratio else (0, 0.0)
        if remove_duplicates is not None:
            p.remove_duplicates = 1 if remove_duplicates else 0
        p.frame_keypoints = int(frame_keypoints)
        self.params = p
        self._h = ctypes.c_void_p()
        capi.check(lib.tod_matcher_create(ctypes.byref(p), ctypes.byref(self._h)))
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tod_matcher_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    @property
    def k(self):
        return int(self.params.k)

    # -- parameter_callback -------------------------------------------------------------------------------------
    def add_object(self, object_id, descriptors, points):
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        if d.shape[0] != p.shape[0]:
            raise ValueError("descriptors and points disagree: %d vs %d" % (d.shape[0], p.shape[0]))
        capi.check(self._lib.tod_matcher_add_object(self._h, str(object_id).encode(), capi._ptr(d), capi._ptr(p),
                                                    d.shape[0]))

    def clear(self):
        capi.check(self._lib.tod_matcher_clear(self._h))

    def load_snapshot(self, path):
        """parameter_callback from a flat DB snapshot (tod_b200.dbio.write_snapshot); call train() next."""
        capi.check(self._lib.tod_matcher_load_snapshot(self._h, str(path).encode()))

    def train(self):
        capi.check(self._lib.tod_matcher_train(self._h))

    @property
    def object_ids(self):
        return [self._lib.tod_matcher_object_id(self._h, i).decode()
                for i in range(self._lib.tod_matcher_num_objects(self._h))]

    @property
    def spans(self):
        """map object_id -> span, like outputs["spans"]."""
        return {oid: float(self._lib.tod_matcher_span(self._h, i)) for i, oid in enumerate(self.object_ids)}

    @property
    def spans_by_index(self):
        return np.array([self._lib.tod_matcher_span(self._h, i)
                         for i in range(self._lib.tod_matcher_num_objects(self._h))], np.float32)

    @property
    def num_descriptors(self):
        return int(self._lib.tod_matcher_num_descriptors(self._h))

    @property
    def shard_rows(self):
        return int(self._lib.tod_matcher_shard_rows(self._h))

    # -- process ------------------------------------------------------------------------------------------------
    def process(self, descriptors, out=None):
        """Host-buffer call.  Returns dict(matches[nq,k] MATCH_DTYPE, counts[nq], matches_3d[nq,k,3], object_ids,
        spans).  `out` may hold preallocated (pinned) arrays under the same keys."""
        q = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        nq, k = q.shape[0], self.k
        if out is None:
            out = {}
        m = out.get("matches")
        if m is None:
            m = np.empty((nq, k), capi.MATCH_DTYPE)
        c = out.get("counts")
        if c is None:
            c = np.empty(nq, np.int32)
        p3 = out.get("matches_3d")
        if p3 is None:
            p3 = np.empty((nq, k, 3), np.float32)
        capi.check(self._lib.tod_matcher_knn(self._h, capi._ptr(q), nq, capi._ptr(m), capi._ptr(c), capi._ptr(p3)))
        return {"matches": m, "counts": c, "matches_3d": p3, "object_ids": self.object_ids, "spans": self.spans}

    def reserve(self, max_nq):
        """Pre-size every per-call device buffer (no cudaMalloc / cudaFree inside a streamed step afterwards)."""
        capi.check(self._lib.tod_matcher_reserve(self._h, int(max_nq)))

    def set_comm(self, unique_id):
        """Collective over the shard_count ranks: attaches an NCCL communicator to the handle."""
        capi.preload_nccl()
        buf = ctypes.create_string_buffer(bytes(unique_id), capi.TOD_COMM_ID_BYTES)
        capi.check(self._lib.tod_matcher_set_comm(self._h, buf))

    @property
    def comm_mode(self):
        """0 none, 1 NCCL all-gather, 2 peer-memory exchange (after the first sharded call mapped the peers)."""
        return int(self._lib.tod_matcher_comm_mode(self._h))

    def set_exchange(self, peer_memory):
        capi.check(self._lib.tod_matcher_set_exchange(self._h, 1 if peer_memory else 0))

    @property
    def exchange_error(self):
        return int(self._lib.tod_matcher_exchange_error(self._h))

    def process_device(self, d_query_ptr, nq, d_matches_ptr, d_counts_ptr, d_points3d_ptr, stream=None):
        """DescriptorMatcher.process on device buffers (tod_matcher_knn_device), enqueued on `stream`."""
        capi.check(self._lib.tod_matcher_knn_device(self._h, ctypes.c_void_p(d_query_ptr), int(nq),
                                                    ctypes.c_void_p(d_matches_ptr), ctypes.c_void_p(d_counts_ptr),
                                                    ctypes.c_void_p(d_points3d_ptr) if d_points3d_ptr else None,
                                                    ctypes.c_void_p(stream) if stream else None))

    def knn_keys_device(self, d_query_ptr, nq, d_keys_ptr, stream=None):
        capi.check(self._lib.tod_matcher_knn_keys_device(self._h, ctypes.c_void_p(d_query_ptr), int(nq),
                                                         ctypes.c_void_p(d_keys_ptr),
                                                         ctypes.c_void_p(stream) if stream else None))

    def merge_device(self, d_keys_all_ptr, n_src, nq, d_matches_ptr, d_counts_ptr, d_points3d_ptr, stream=None):
        capi.check(self._lib.tod_matcher_merge_device(self._h, ctypes.c_void_p(d_keys_all_ptr), int(n_src), int(nq),
                                                      ctypes.c_void_p(d_matches_ptr), ctypes.c_void_p(d_counts_ptr),
                                                      ctypes.c_void_p(d_points3d_ptr) if d_points3d_ptr else None,
                                                      ctypes.c_void_p(stream) if stream else None))

    @property
    def last_k1_ms(self):
        return float(self._lib.tod_matcher_last_k1_ms(self._h))

    def k1_ms_history(self, n):
        """K1 kernel times (ms) of the last n calls, oldest first (the library keeps 64 event pairs)."""
        return [float(self._lib.tod_matcher_k1_ms_ago(self._h, i)) for i in range(int(n) - 1, -1, -1)]

    def set_stage_timing(self, on):
        self._lib.tod_matcher_set_stage_timing(self._h, 1 if on else 0)

    @property
    def last_exchange_ms(self):
        return float(self._lib.tod_matcher_last_exchange_ms(self._h))

    @property
    def last_kernel(self):
        return self._lib.tod_matcher_last_kernel(self._h).decode()


def fill_adjacency(offsets, query_pts, train_pts, pixels, spans, sensor_error, device=0):
    """K2 for a batch of clusters (AdjacencyRansac::FillAdjacency, adjacency_ransac.cpp:127-172).
    Returns (physical, sample, matrix_offsets): flat u32 arrays holding one n_c x row_words(n_c) bit-matrix per cluster."""
    lib = capi.load()
    offsets = np.ascontiguousarray(offsets, np.int32)
    nc = offsets.shape[0] - 1
    q = np.ascontiguousarray(query_pts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(train_pts, np.float32).reshape(-1, 3)
    px = np.ascontiguousarray(pixels, np.float32).reshape(-1, 2)
    sp = np.ascontiguousarray(spans, np.float32).reshape(-1)
    sizes = np.diff(offsets).astype(np.int64)
    words = np.array([capi.adjacency_row_words(int(n)) for n in sizes], np.int64)
    total = int((sizes * words).sum())
    physical = np.zeros(max(total, 1), np.uint32)
    sample = np.zeros(max(total, 1), np.uint32)
    mo = np.zeros(nc + 1, np.int64)
    capi.check(lib.tod_fill_adjacency(int(device), nc, capi._ptr(offsets), capi._ptr(q), capi._ptr(t), capi._ptr(px),
                                      capi._ptr(sp), ctypes.c_float(sensor_error), capi._ptr(physical),
                                      capi._ptr(sample), capi._ptr(mo)))
    return physical[:total], sample[:total], mo


def score_hypotheses(query_pts, train_pts, physical, valid, triples, threshold=float("inf"), device=0,
                     want_pose=True):
    """K3 for one cluster.  Returns (counts[H], R[H,9] or None, T[H,3] or None)."""
    lib = capi.load()
    q = np.ascontiguousarray(query_pts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(train_pts, np.float32).reshape(-1, 3)
    n = q.shape[0]
    P = np.ascontiguousarray(physical, np.uint32)
    V = np.ascontiguousarray(valid, np.uint32)
    tr = np.ascontiguousarray(triples, np.uint32).reshape(-1, 3)
    H = tr.shape[0]
    counts = np.zeros(H, np.int32)
    R = np.zeros((H, 9), np.float32) if want_pose else None
    T = np.zeros((H, 3), np.float32) if want_pose else None
    capi.check(lib.tod_score_hypotheses(int(device), n, capi._ptr(q), capi._ptr(t), capi._ptr(P), capi._ptr(V), H,
                                        capi._ptr(tr), ctypes.c_double(threshold), capi._ptr(counts), capi._ptr(R),
                                        capi._ptr(T)))
    return counts, R, T


class GuessGenerator:
    """Mirror of tod::GuessGenerator: params min_inliers, n_ransac_iterations, sensor_error (GuessGenerator.cpp:74-80);
    inputs points3d, keypoints, matches, matches_3d, spans, object_ids; outputs pose_results, Rs, Ts."""

    def __init__(self, min_inliers=15, n_ransac_iterations=1000, sensor_error=0.01, ransac_threshold=float("inf"),
                 seed=0, device=0, host_threads=0):
        lib = capi.load()
        p = capi.GuessParams()
        lib.tod_guess_default_params(ctypes.byref(p))
        p.min_inliers, p.n_ransac_iterations = int(min_inliers), int(n_ransac_iterations)
        p.sensor_error, p.ransac_threshold = float(sensor_error), float(ransac_threshold)
        p.seed, p.device, p.host_threads = int(seed), int(device), int(host_threads)
        self.params = p
        self._h = ctypes.c_void_p()
        capi.check(lib.tod_guess_create(ctypes.byref(p), ctypes.byref(self._h)))
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tod_guess_destroy(self._h)
            self._h = None

    __del__ = close

    def process(self, keypoints, points3d, matches, counts, matches_3d, spans_by_index, max_poses=256):
        """keypoints: KEYPOINT_DTYPE[n] or (n,2) float pixel coords; points3d: H x W x 3 f32 cloud.
        Returns dict(pose_results POSE_DTYPE[n_poses], Rs, Ts, inliers list of arrays)."""
        if keypoints.dtype != capi.KEYPOINT_DTYPE:
            xy = np.asarray(keypoints, np.float32).reshape(-1, 2)
            kp = np.zeros(xy.shape[0], capi.KEYPOINT_DTYPE)
            kp["x"], kp["y"] = xy[:, 0], xy[:, 1]
        else:
            kp = np.ascontiguousarray(keypoints)
        cloud = np.ascontiguousarray(points3d, np.float32)
        H, W = cloud.shape[0], cloud.shape[1]
        m = np.ascontiguousarray(matches)
        assert m.dtype == capi.MATCH_DTYPE
        k = m.shape[1]
        c = np.ascontiguousarray(counts, np.int32)
        p3 = np.ascontiguousarray(matches_3d, np.float32)
        sp = np.ascontiguousarray(spans_by_index, np.float32)
        poses = np.zeros(max_poses, capi.POSE_DTYPE)
        n_poses = ctypes.c_int32(0)
        cap = max(1, kp.shape[0] * max(1, min(k, sp.shape[0])))   # a keypoint can be an inlier of one pose per object
        inl = np.zeros(cap, np.int32)
        capi.check(self._lib.tod_guess_process(self._h, capi._ptr(kp), kp.shape[0], capi._ptr(cloud), H, W,
                                               capi._ptr(m), capi._ptr(c), k, capi._ptr(p3), capi._ptr(sp),
                                               sp.shape[0], capi._ptr(poses), max_poses, ctypes.byref(n_poses),
                                               capi._ptr(inl), cap))
        poses = poses[:n_poses.value]
        inliers, o = [], 0
        for p in poses:
            inliers.append(inl[o:o + int(p["n_inliers"])].copy())
            o += int(p["n_inliers"])
        return {"pose_results": poses, "Rs": poses["R"].reshape(-1, 3, 3).copy(), "Ts": poses["T"].copy(),
                "inliers": inliers}

    @staticmethod
    def pack_keypoints(keypoints_list):
        """Per-frame (n_f, 2) pixel coordinates (or KEYPOINT_DTYPE arrays) -> (one KEYPOINT_DTYPE array, int32 offsets):
        the form tod_guess_process_batch takes.  A caller that already holds its keypoints as cv::KeyPoint-shaped
        records (what the reference's feature cell emits) passes that tuple to process_batch directly."""
        sizes = [a.shape[0] for a in keypoints_list]
        off = np.zeros(len(sizes) + 1, np.int32)
        np.cumsum(sizes, out=off[1:])
        if keypoints_list and all(a.dtype == capi.KEYPOINT_DTYPE for a in keypoints_list):
            return np.ascontiguousarray(np.concatenate(keypoints_list)), off
        raw = np.zeros((int(off[-1]), 7), np.float32)             # x, y, size, angle, response, octave, class_id
        if sizes:
            raw[:, :2] = np.concatenate([np.asarray(a, np.float32).reshape(-1, 2) for a in keypoints_list])
        return raw.view(capi.KEYPOINT_DTYPE).reshape(-1), off

    def process_batch(self, keypoints_list, clouds, matches, counts, matches_3d, spans_by_index, max_poses=None):
        """A batch of frames in one call (tod_guess_process_batch).  keypoints_list: per frame (n_f, 2) pixel coords or
        KEYPOINT_DTYPE arrays, or the tuple (KEYPOINT_DTYPE array of all frames, int32 offsets[F + 1]) from
        pack_keypoints(); clouds: F x H x W x 3 f32; matches / counts / matches_3d: concatenated over the frames'
        keypoints (the matcher's output for the concatenated descriptors).  Returns a list of per-frame dicts like
        process()."""
        if isinstance(keypoints_list, tuple):
            kp, off = keypoints_list
            assert kp.dtype == capi.KEYPOINT_DTYPE and kp.flags.c_contiguous
            off = np.ascontiguousarray(off, np.int32)
        else:
            kp, off = self.pack_keypoints(keypoints_list)
        clouds = np.ascontiguousarray(clouds, np.float32)
        F, H, W = clouds.shape[0], clouds.shape[1], clouds.shape[2]
        assert F == off.shape[0] - 1
        m = np.ascontiguousarray(matches)
        assert m.dtype == capi.MATCH_DTYPE and m.shape[0] == kp.shape[0]
        k = m.shape[1]
        c = np.ascontiguousarray(counts, np.int32)
        p3 = np.ascontiguousarray(matches_3d, np.float32)
        sp = np.ascontiguousarray(spans_by_index, np.float32)
        if max_poses is None:
            max_poses = 64 * F
        cap = max(1, kp.shape[0] * max(1, min(k, sp.shape[0])))
        key = (max_poses, cap)
        if getattr(self, "_out_key", None) != key:               # output buffers are kept between calls
            self._out = (np.zeros(max_poses, capi.POSE_DTYPE), np.zeros(max_poses, np.int32), np.zeros(cap, np.int32))
            self._out_key = key
        poses, frames, inl = self._out
        n_poses = ctypes.c_int32(0)
        capi.check(self._lib.tod_guess_process_batch(self._h, F, capi._ptr(off), capi._ptr(kp), capi._ptr(clouds), H, W,
                                                     capi._ptr(m), capi._ptr(c), k, capi._ptr(p3), capi._ptr(sp),
                                                     sp.shape[0], capi._ptr(poses), capi._ptr(frames), max_poses,
                                                     ctypes.byref(n_poses), capi._ptr(inl), cap))
        n = n_poses.value
        poses, frames = poses[:n].copy(), frames[:n].copy()
        ends = np.cumsum(poses["n_inliers"]) if n else np.zeros(0, np.int64)
        starts = ends - poses["n_inliers"] if n else ends
        first = np.searchsorted(frames, np.arange(F + 1))        # poses come out in frame order
        out = []
        for f in range(F):
            lo, hi = int(first[f]), int(first[f + 1])
            pr = poses[lo:hi]
            out.append({"pose_results": pr, "Rs": pr["R"].reshape(-1, 3, 3), "Ts": pr["T"],
                        "inliers": [inl[int(starts[i]):int(ends[i])].copy() for i in range(lo, hi)]})
        return out

    def last_stats(self):
        k2, k3 = ctypes.c_float(), ctypes.c_float()
        nh, nr = ctypes.c_int64(), ctypes.c_int32()
        self._lib.tod_guess_last_stats(self._h, ctypes.byref(k2), ctypes.byref(k3), ctypes.byref(nh), ctypes.byref(nr))
        prof = (ctypes.c_double * 12)()
        self._lib.tod_guess_last_profile(self._h, prof)
        b2, b3 = ctypes.c_double(), ctypes.c_double()
        ncl, ncor = ctypes.c_int64(), ctypes.c_int64()
        self._lib.tod_guess_last_traffic(self._h, ctypes.byref(b2), ctypes.byref(b3), ctypes.byref(ncl),
                                         ctypes.byref(ncor))
        gh = (ctypes.c_int64 * 24)()
        self._lib.tod_guess_last_gate_stats(self._h, gh)
        gh = [int(x) for x in gh]
        k5 = (ctypes.c_int64 * 4)()
        self._lib.tod_guess_last_k5_stats(self._h, k5)
        return {"gate_shape": {"by_graph_size": gh[0:8], "by_core_size": gh[8:16], "core_too_small": gh[16],
                               "colour_bound": gh[17], "searches": gh[18], "search_passes": gh[19],
                               "search_steps": gh[20], "k4_fails_used": gh[21], "k4_undecided_used": gh[22],
                               "k4_kernel_ms": gh[23] / 1000.0, "k5_passes_used": int(k5[0]),
                               "k5_fails_used": int(k5[1]), "k5_kernel_ms": int(k5[3]) / 1000.0},
                "k2_ms": k2.value, "k3_ms": k3.value, "k2_bytes": b2.value, "k3_bytes": b3.value,
                "n_clusters": ncl.value, "n_correspondences": ncor.value, "n_hypotheses": nh.value, "n_rounds": nr.value,
                "host_ms": {"cluster_k2": prof[0], "sampler": prof[1], "k3_launch_sync": prof[2],
                            "replay_gate": prof[3], "refine_invalidate": prof[4], "total": prof[7]},
                "gate_calls": int(prof[5]), "gate_proved_empty": int(prof[6]), "gate_core_rejects": int(prof[11]),
                "gate_thread_ms": {"setup": prof[8], "proof": prof[9], "search": prof[10]}}


class FeatureDescriptor:
    """ecto_opencv's FeatureDescriptor cell (cv::ORB, detector.py:27,74; conf/detection.ork:23-31) on the GPU: pyramid,
    FAST + Harris detection, smoothing, orientation and rBRIEF descriptors — bit-exact against cv2.ORB.  The
    descriptors also stay on the device (`last_device_descriptors`) for DescriptorMatcher.process_device."""

    def __init__(self, n_features=5000, n_levels=3, scale_factor=1.2, device=0):
        self.n_features = int(n_features)
        lib = capi.load()
        p = capi.OrbParams()
        lib.tod_orb_default_params(ctypes.byref(p))
        p.n_levels, p.scale_factor, p.device = int(n_levels), float(scale_factor), int(device)
        self._h = ctypes.c_void_p()
        capi.check(lib.tod_orb_create(ctypes.byref(p), ctypes.byref(self._h)))
        self._lib = lib
        self.last_device_descriptors = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tod_orb_destroy(self._h)
            self._h = None

    __del__ = close

    def process(self, image):
        """inputs["image"] (H x W u8) -> keypoints (KEYPOINT_DTYPE, ordered by octave, row, column), descriptors."""
        img = np.ascontiguousarray(image, np.uint8)
        assert img.ndim == 2
        cap = 2 * self.n_features + 1024
        kp = np.zeros(cap, capi.KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = ctypes.c_int32(0)
        dptr = ctypes.c_void_p()
        capi.check(self._lib.tod_orb_detect_and_compute(self._h, capi._ptr(img), img.shape[0], img.shape[1],
                                                        self.n_features, capi._ptr(kp), cap, ctypes.byref(n),
                                                        capi._ptr(desc), ctypes.byref(dptr)))
        self.last_device_descriptors = dptr.value
        return kp[:n.value].copy(), desc[:n.value].copy()

    def process_masked(self, image, mask):
        """cv::ORB::detectAndCompute(image, mask): image H x W (grey) or H x W x 3 (BGR) u8, mask H x W u8 or None."""
        img = np.ascontiguousarray(image, np.uint8)
        ch = 1 if img.ndim == 2 else int(img.shape[2])
        mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        cap = 2 * self.n_features + 1024
        kp = np.zeros(cap, capi.KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = ctypes.c_int32(0)
        dptr = ctypes.c_void_p()
        capi.check(self._lib.tod_orb_detect_and_compute_masked(self._h, capi._ptr(img), ch, img.shape[0], img.shape[1],
                                                               capi._ptr(mk), self.n_features, capi._ptr(kp), cap,
                                                               ctypes.byref(n), capi._ptr(desc), ctypes.byref(dptr)))
        self.last_device_descriptors = dptr.value
        return kp[:n.value].copy(), desc[:n.value].copy()

    def read_level(self, level, kind=0):
        """One pyramid level of the last frame: kind 0 = resized, 1 = smoothed, 2 = FAST corner scores."""
        h, w = ctypes.c_int32(0), ctypes.c_int32(0)
        capi.check(self._lib.tod_orb_read_level(self._h, int(level), int(kind), None, ctypes.byref(h), ctypes.byref(w)))
        out = np.zeros((h.value, w.value), np.uint8)
        capi.check(self._lib.tod_orb_read_level(self._h, int(level), int(kind), capi._ptr(out), None, None))
        return out

    def describe(self, image, keypoints, compute_angles=True):
        """image: H x W u8; keypoints: KEYPOINT_DTYPE array (x, y, octave read; angle written when compute_angles).
        Returns (keypoints, descriptors[n, 32] u8)."""
        img = np.ascontiguousarray(image, np.uint8)
        kp = np.ascontiguousarray(keypoints).copy()
        assert kp.dtype == capi.KEYPOINT_DTYPE and img.ndim == 2
        desc = np.zeros((kp.shape[0], 32), np.uint8)
        dptr = ctypes.c_void_p()
        capi.check(self._lib.tod_orb_describe(self._h, capi._ptr(img), img.shape[0], img.shape[1], capi._ptr(kp),
                                              kp.shape[0], 1 if compute_angles else 0, capi._ptr(desc),
                                              ctypes.byref(dptr)))
        self.last_device_descriptors = dptr.value
        return kp, desc


class Trainer:
    """Mirror of the reference's Trainer cell (src/training/Trainer.cpp:83-187): observations of one object in, the
    stacked `descriptors` (N x 32 u8) and `points` (N x 3 f32, object frame) out — what ModelFiller stores
    (ModelFiller.cpp:23-24) and DescriptorMatcher.add_object / dbio.write_snapshot take."""

    def __init__(self, n_features=500, n_levels=8, scale_factor=1.2, device=0):
        lib = capi.load()
        p = capi.TrainerParams()
        lib.tod_trainer_default_params(ctypes.byref(p))
        p.n_features, p.n_levels, p.scale_factor, p.device = int(n_features), int(n_levels), float(scale_factor), int(device)
        self._h = ctypes.c_void_p()
        capi.check(lib.tod_trainer_create(ctypes.byref(p), ctypes.byref(self._h)))
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tod_trainer_destroy(self._h)
            self._h = None

    __del__ = close

    def add_observation(self, image, mask, depth, K, R, T):
        """One observation (obs.image, obs.mask, obs.depth, obs.K, obs.R, obs.T).  Returns the number of points added."""
        img = np.ascontiguousarray(image, np.uint8)
        ch = 1 if img.ndim == 2 else int(img.shape[2])
        mk = np.ascontiguousarray(mask, np.uint8)
        d = np.ascontiguousarray(depth)
        assert d.dtype in (np.float32, np.uint16) and mk.shape == img.shape[:2]
        k = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
        r = np.ascontiguousarray(np.asarray(R, np.float32).reshape(9))
        t = np.ascontiguousarray(np.asarray(T, np.float32).reshape(3))
        n = ctypes.c_int32(0)
        capi.check(self._lib.tod_trainer_add_observation(self._h, capi._ptr(img), ch, img.shape[0], img.shape[1],
                                                         capi._ptr(mk), capi._ptr(d), 1 if d.dtype == np.uint16 else 0,
                                                         d.shape[0], d.shape[1], capi._ptr(k), capi._ptr(r),
                                                         capi._ptr(t), ctypes.byref(n)))
        return n.value

    def model(self):
        """(descriptors N x 32 u8, points N x 3 f32): mergePoints over the observations added so far."""
        dp, pp, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64(0)
        capi.check(self._lib.tod_trainer_model(self._h, ctypes.byref(dp), ctypes.byref(pp), ctypes.byref(n)))
        if n.value == 0:
            return np.zeros((0, 32), np.uint8), np.zeros((0, 3), np.float32)
        d = np.ctypeslib.as_array(ctypes.cast(dp, ctypes.POINTER(ctypes.c_uint8)), (n.value, 32)).copy()
        p = np.ctypeslib.as_array(ctypes.cast(pp, ctypes.POINTER(ctypes.c_float)), (n.value, 3)).copy()
        return d, p

    def clear(self):
        capi.check(self._lib.tod_trainer_clear(self._h))


def depth_to_3d(depth, K, device=0, out=None):
    """DepthTo3d (detector.py:62-69) on the GPU: depth H x W float32 metres or uint16 millimetres -> H x W x 3 f32.
    `out` may be a preallocated (ideally pinned) H x W x 3 float32 array: the 12 bytes per pixel then come back at
    PCIe speed instead of through the driver's pageable staging path."""
    d = np.ascontiguousarray(depth)
    assert d.dtype in (np.float32, np.uint16) and d.ndim == 2
    k = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
    if out is None:
        out = np.empty((d.shape[0], d.shape[1], 3), np.float32)
    assert out.dtype == np.float32 and out.shape == (d.shape[0], d.shape[1], 3) and out.flags.c_contiguous
    capi.check(capi.load().tod_depth_to_3d(int(device), capi._ptr(d), 1 if d.dtype == np.uint16 else 0, d.shape[0],
                                           d.shape[1], capi._ptr(k), capi._ptr(out)))
    return out


# ---- the `.ork` pipeline parameters, as TodDetector forwards them (python/object_recognition_tod/detector.py:34-62) ----
def ork_parameters(parameters):
    """`parameters` = the `pipelineN.parameters` subtree of a detection `.ork` file (conf/detection.ork:21-46) as a
    dict.  Returns (MatcherParams, guess_kwargs): the `search` subtree goes to the matcher as a JSON string, exactly
    like TodDetector.configure does (detector.py:57-60 json-dumps it into `search_json_params`), and
    n_ransac_iterations / min_inliers / sensor_error go to the guess generator (detector.py:37-39).  Host-only."""
    lib = capi.load()
    p = capi.MatcherParams()
    lib.tod_matcher_default_params(ctypes.byref(p))
    capi.check(lib.tod_matcher_params_from_json(json.dumps(parameters["search"]).encode(), ctypes.byref(p)))
    guess = {}
    for key, cast in (("n_ransac_iterations", int), ("min_inliers", int), ("sensor_error", float)):
        if key in parameters:
            guess[key] = cast(parameters[key])
    return p, guess


def detector_from_ork(parameters, device=0, seed=0):
    """(DescriptorMatcher, GuessGenerator) configured from a detection `.ork` `parameters` subtree; add the models
    (add_object / load_snapshot) and train() before processing frames."""
    _, guess = ork_parameters(parameters)
    m = DescriptorMatcher(search_json_params=json.dumps(parameters["search"]), device=device)
    g = GuessGenerator(device=device, seed=seed, **guess)
    return m, g


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it and hands it to the other ranks by any host-side means)."""
    capi.preload_nccl()
    buf = ctypes.create_string_buffer(capi.TOD_COMM_ID_BYTES)
    capi.check(capi.load().tod_comm_unique_id(buf))
    return buf.raw
