"""TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/_ref/libtod_ref.so — the reference's own src/common sources compiled
unmodified (oracle/build_ref.py) behind the small C harness oracle/ref_harness.cpp."""
import ctypes
import os

import numpy as np

from . import build_ref

_LIB = None
MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
POSE_DTYPE = np.dtype([("R", "<f4", (9,)), ("T", "<f4", (3,)), ("object_index", "<i4"), ("n_inliers", "<i4")])


def available():
    return build_ref.build() is not None


def lib():
    global _LIB
    if _LIB is None:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref/libtod_ref.so is missing and /root/reference is not present to build it")
        L = ctypes.CDLL(path)
        P, U64, I, U, F, D = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_double
        sig = {
            "ref_set_rng": (None, [U64]),
            "ref_rng_seed": (U64, [U64, ctypes.c_uint32, ctypes.c_uint32]),
            "ref_find_clique": (I, [I, P, I, I, P, I, U, P]),
            "ref_ar_new": (P, []),
            "ref_ar_free": (None, [P]),
            "ref_ar_add": (None, [P, P, P, U]),
            "ref_ar_fill": (None, [P, P, I, F, F]),
            "ref_ar_size": (I, [P]),
            "ref_ar_neighbors": (I, [P, I, U, P, I]),
            "ref_ar_valid": (I, [P, P, I]),
            "ref_ar_invalidate_query": (None, [P, P, I]),
            "ref_ar_get_samples": (I, [P, U64, I, P]),
            "ref_ar_select": (I, [P, P, D, P, I, P, P]),
            "ref_ar_kabsch": (I, [P, P, I, P, P]),
            "ref_ar_compute_model": (I, [P, U, U64, D, P, I, P]),
            "ref_ar_ransac": (I, [P, F, U, U64, P, I, P, P]),
            "ref_process": (I, [P, I, P, I, I, P, P, I, P, P, I, U, U, F, U64, P, I, P, I]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def find_clique(n, edges, minimal_size=0xFFFFFFFF, sorted_insert=False, deleted=()):
    e = np.ascontiguousarray(np.array(edges, np.int32).reshape(-1, 2))
    d = np.ascontiguousarray(np.array(deleted, np.int32).reshape(-1, 2))
    out = np.zeros(max(n, 1), np.uint32)
    k = lib().ref_find_clique(n, _p(e), e.shape[0], int(sorted_insert), _p(d), d.shape[0], minimal_size, _p(out))
    return [int(x) for x in out[:k]]


class RefAdjacencyRansac:
    """The reference's tod::AdjacencyRansac object."""

    def __init__(self):
        self.h = ctypes.c_void_p(lib().ref_ar_new())
        self.n = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_ar_free(self.h)
            self.h = None

    def add_points(self, train, query, query_index):
        t = np.ascontiguousarray(train, np.float32)
        q = np.ascontiguousarray(query, np.float32)
        lib().ref_ar_add(self.h, _p(t), _p(q), int(query_index))
        self.n += 1

    def fill_adjacency(self, keypoints_xy, span, sensor_error):
        kp = np.ascontiguousarray(keypoints_xy, np.float32).reshape(-1, 2)
        lib().ref_ar_fill(self.h, _p(kp), kp.shape[0], float(span), float(sensor_error))

    def neighbors(self, which, i):
        out = np.zeros(max(self.n, 1), np.uint32)
        k = lib().ref_ar_neighbors(self.h, 1 if which == "sample" else 0, int(i), _p(out), out.shape[0])
        return [int(x) for x in out[:k]]

    def dense(self, which):
        M = np.zeros((self.n, self.n), bool)
        for i in range(self.n):
            M[i, self.neighbors(which, i)] = True
        return M

    def valid(self):
        out = np.zeros(max(self.n, 1), np.uint32)
        k = lib().ref_ar_valid(self.h, _p(out), out.shape[0])
        return [int(x) for x in out[:k]]

    def invalidate_query_indices(self, q):
        q = np.ascontiguousarray(q, np.uint32)
        lib().ref_ar_invalidate_query(self.h, _p(q), q.shape[0])

    def get_samples(self, rng_state, n_hyp):
        out = np.zeros((n_hyp, 3), np.uint32)
        k = lib().ref_ar_get_samples(self.h, rng_state, n_hyp, _p(out))
        return out[:k]

    def select(self, triple, threshold=float("inf")):
        t = np.ascontiguousarray(triple, np.uint32)
        out = np.zeros(self.n + 3, np.uint32)
        R = np.zeros(9, np.float32)
        T = np.zeros(3, np.float32)
        thr = threshold if np.isfinite(threshold) else np.finfo(np.float64).max
        k = lib().ref_ar_select(self.h, _p(t), thr, _p(out), out.shape[0], _p(R), _p(T))
        return [int(x) for x in out[:k]], R.reshape(3, 3), T

    def kabsch(self, idx):
        idx = np.ascontiguousarray(idx, np.uint32)
        R = np.zeros(9, np.float32)
        T = np.zeros(3, np.float32)
        rc = lib().ref_ar_kabsch(self.h, _p(idx), idx.shape[0], _p(R), _p(T))
        assert rc == 0
        return R.reshape(3, 3), T

    def compute_model(self, max_iterations, rng_state, threshold=float("inf")):
        out = np.zeros(self.n + 3, np.uint32)
        it = ctypes.c_int(0)
        thr = threshold if np.isfinite(threshold) else np.finfo(np.float64).max
        k = lib().ref_ar_compute_model(self.h, int(max_iterations), rng_state, thr, _p(out), out.shape[0],
                                       ctypes.addressof(it))
        return [int(x) for x in out[:k]], it.value

    def ransac(self, sensor_error, n_iterations, rng_state):
        out = np.zeros(self.n + 3, np.uint32)
        R = np.zeros(9, np.float32)
        T = np.zeros(3, np.float32)
        k = lib().ref_ar_ransac(self.h, float(sensor_error), int(n_iterations), rng_state, _p(out), out.shape[0],
                                _p(R), _p(T))
        return [int(x) for x in out[:k]], R.reshape(3, 3), T


def process(keypoints_xy, cloud, matches, counts, points3d, spans, min_inliers, n_ransac_iterations, sensor_error,
            seed, max_poses=256):
    """ClusterPerObject + the GuessGenerator::process loop on the reference's code.
    Returns list of (object_index, R, T, inlier_keypoints)."""
    kp = np.ascontiguousarray(keypoints_xy, np.float32).reshape(-1, 2)
    cl = np.ascontiguousarray(cloud, np.float32)
    m = np.ascontiguousarray(matches)
    assert m.dtype.itemsize == 16
    c = np.ascontiguousarray(counts, np.int32)
    p3 = np.ascontiguousarray(points3d, np.float32)
    sp = np.ascontiguousarray(spans, np.float32)
    poses = np.zeros(max_poses, POSE_DTYPE)
    cap = max(1, kp.shape[0] * 2)
    inl = np.zeros(cap, np.int32)
    n = lib().ref_process(_p(kp), kp.shape[0], _p(cl), cl.shape[0], cl.shape[1], _p(m), _p(c), m.shape[1], _p(p3),
                          _p(sp), sp.shape[0], int(min_inliers), int(n_ransac_iterations), float(sensor_error),
                          int(seed), _p(poses), max_poses, _p(inl), cap)
    assert 0 <= n <= max_poses, n
    out, o = [], 0
    for p in poses[:n]:
        k = int(p["n_inliers"])
        out.append((int(p["object_index"]), p["R"].reshape(3, 3).copy(), p["T"].copy(), [int(x) for x in inl[o:o + k]]))
        o += k
    return out
