"""Oracle for K1: exact k-NN Hamming matching over 256-bit descriptors.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates what `matcher_->knnMatch(descriptors, matches, 5)` + the radius cut compute at
src/detection/DescriptorMatcher.cpp:211-220 when the matcher is the exact one north_star names,
cv::BFMatcher(NORM_HAMMING).  OpenCV is a third-party dependency that is NOT vendored in the reference tree
(package.xml:12-22 pins no version); its published behaviour, verified here against cv2 4.13.0
(tests/golden/make_golden.py):

  * train set = list of per-object descriptor matrices in add() order; DMatch.imgIdx = object index,
    DMatch.trainIdx = row inside that object (DescriptorMatcher.cpp:127-128);
  * per query the k smallest entries under the lexicographic key (distance, imgIdx, trainIdx);
  * distance = popcount(a XOR b) over the 32 bytes, returned as a float holding an exact integer;
  * radius cut: the list is truncated at the first match with distance > radius (DescriptorMatcher.cpp:212-220).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def concat_objects(objects):
    """Concatenate per-object descriptor matrices in imgIdx order; return (db, offsets[n_obj+1])."""
    offsets = np.zeros(len(objects) + 1, dtype=np.int64)
    for i, o in enumerate(objects):
        offsets[i + 1] = offsets[i] + o.shape[0]
    if len(objects):
        db = np.ascontiguousarray(np.concatenate([np.asarray(o, dtype=np.uint8).reshape(-1, 32) for o in objects]))
    else:
        db = np.zeros((0, 32), np.uint8)
    return db, offsets


def hamming_matrix(query, db):
    """nq x ndb uint16 matrix of popcount(q XOR d).  (cv::normHamming over 32 bytes.)"""
    q = np.ascontiguousarray(query, dtype=np.uint8).view(np.uint64).reshape(-1, 1, 4)
    d = np.ascontiguousarray(db, dtype=np.uint8).view(np.uint64).reshape(1, -1, 4)
    return np.bitwise_count(q ^ d).sum(axis=2, dtype=np.uint16)


def knn_numpy(query, objects, k, radius=0, chunk=256):
    """Brute-force oracle.  Returns (matches[nq,k] structured (queryIdx,trainIdx,imgIdx,distance), counts[nq])."""
    db, offsets = concat_objects(objects)
    nq, ndb = query.shape[0], db.shape[0]
    out = np.zeros((nq, k), dtype=[("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
    out["queryIdx"] = -1
    out["trainIdx"] = -1
    out["imgIdx"] = -1
    counts = np.zeros(nq, dtype=np.int32)
    if ndb == 0:
        return out, counts
    kk = min(k, ndb)
    rows = np.arange(ndb, dtype=np.int64)
    for q0 in range(0, nq, chunk):
        dist = hamming_matrix(query[q0:q0 + chunk], db).astype(np.int64)
        key = (dist << 32) | rows[None, :]          # (distance, global row) == (distance, imgIdx, trainIdx)
        part = np.partition(key, kk - 1, axis=1)[:, :kk]
        part.sort(axis=1)
        d = part >> 32
        g = part & 0xFFFFFFFF
        img = np.searchsorted(offsets, g, side="right") - 1
        trn = g - offsets[img]
        n = np.full(d.shape[0], kk, dtype=np.int32)
        if radius:
            over = d > radius
            first = np.where(over.any(axis=1), over.argmax(axis=1), kk)
            n = first.astype(np.int32)
        sl = slice(q0, q0 + d.shape[0])
        out["queryIdx"][sl, :kk] = np.arange(q0, q0 + d.shape[0])[:, None]
        out["trainIdx"][sl, :kk] = trn
        out["imgIdx"][sl, :kk] = img
        out["distance"][sl, :kk] = d
        counts[sl] = n
    # blank the entries cut by the radius / short DB so that comparisons are well defined
    mask = np.arange(k)[None, :] >= counts[:, None]
    for f, v in (("queryIdx", -1), ("trainIdx", -1), ("imgIdx", -1), ("distance", 0.0)):
        out[f][mask] = v
    return out, counts


def gather_points3d(matches, counts, points_per_object):
    """matches_3d gather, DescriptorMatcher.cpp:232-244: features3d_db_[imgIdx](0, trainIdx)."""
    nq, k = matches.shape
    out = np.zeros((nq, k, 3), np.float32)
    for q in range(nq):
        for j in range(counts[q]):
            out[q, j] = points_per_object[matches["imgIdx"][q, j]][matches["trainIdx"][q, j]]
    return out


def object_span(points):
    """Span of one object, DescriptorMatcher.cpp:106-121: sqrt of the squared bbox diagonal, all in float."""
    p = np.asarray(points, np.float32).reshape(-1, 3)
    if p.shape[0] == 0:
        # min = FLT_MAX, max = -FLT_MAX -> (max-min) overflows to -inf, squared -> +inf
        return np.float32(np.inf)
    ext = (p.max(axis=0) - p.min(axis=0)).astype(np.float32)
    s = np.float32(ext[0] * ext[0])
    s = np.float32(s + np.float32(ext[1] * ext[1]))
    s = np.float32(s + np.float32(ext[2] * ext[2]))
    return np.float32(np.sqrt(s))


# ---------------------------------------------------------------------------------------------------------------
# C port (oracle/hamming_knn.c) — same algorithm, OpenMP over queries; used for big cases and as bench.py's
# cpu_baseline of kind "port" when cv2 is unavailable.
# ---------------------------------------------------------------------------------------------------------------
_LIB = None


def build_c(force=False):
    src = os.path.join(_HERE, "hamming_knn.c")
    out = os.path.join(_HERE, "libtod_oracle_knn.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-mpopcnt", "-fopenmp", "-shared", "-fPIC", "-o", out, src])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.oracle_knn_hamming.restype = ctypes.c_int
        _LIB.oracle_knn_hamming.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def knn_c(query, objects, k, radius=0, threads=0):
    """Same contract as knn_numpy, computed by the C port."""
    db, offsets = concat_objects(objects)
    query = np.ascontiguousarray(query, np.uint8)
    nq = query.shape[0]
    keys = np.zeros((nq, k), np.uint64)
    cnt = np.zeros(nq, np.int32)
    rc = _lib().oracle_knn_hamming(query.ctypes.data, nq, db.ctypes.data, db.shape[0], k, threads,
                                   keys.ctypes.data, cnt.ctypes.data)
    assert rc == 0
    out = np.zeros((nq, k), dtype=[("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
    d = (keys >> np.uint64(32)).astype(np.int64)
    g = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    if radius:
        over = (d > radius) | (np.arange(k)[None, :] >= cnt[:, None])
        first = np.where(over.any(axis=1), over.argmax(axis=1), k)
        cnt = np.minimum(cnt, first.astype(np.int32))
    img = np.searchsorted(offsets, g, side="right") - 1
    trn = g - offsets[np.clip(img, 0, max(len(offsets) - 2, 0))]
    mask = np.arange(k)[None, :] >= cnt[:, None]
    out["queryIdx"] = np.where(mask, -1, np.arange(nq)[:, None])
    out["trainIdx"] = np.where(mask, -1, trn)
    out["imgIdx"] = np.where(mask, -1, img)
    out["distance"] = np.where(mask, 0, d)
    return out, cnt.astype(np.int32)
