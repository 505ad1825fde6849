#!/usr/bin/env python
"""Generate tests/golden/knn_*.npz from the REAL matcher the reference calls: OpenCV's cv::BFMatcher(NORM_HAMMING)
(through cv2 4.13.0, importable in the build container; the reference's call site is
src/detection/DescriptorMatcher.cpp:127-128,211-220).  Run once in the build container:

    python tests/golden/make_golden.py

The .npz files are committed; tests never need cv2 or /root/reference at run time.
Each file: query[nq,32] u8, db[ndb,32] u8, sizes[n_obj] (rows per object, add() order), k, radius,
           trainIdx/imgIdx/distance[nq,k] (-1 / -1 / 0 padded), counts[nq].
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cv2_knn(query, objects, k, radius):
    m = cv2.BFMatcher(cv2.NORM_HAMMING)
    m.add([np.ascontiguousarray(o) for o in objects])
    res = m.knnMatch(np.ascontiguousarray(query), k)
    nq = query.shape[0]
    trn = -np.ones((nq, k), np.int32)
    img = -np.ones((nq, k), np.int32)
    dist = np.zeros((nq, k), np.float32)
    cnt = np.zeros(nq, np.int32)
    for q, lst in enumerate(res):
        lst = list(lst)
        if radius:  # DescriptorMatcher.cpp:212-220: truncate at the first distance > radius
            for j, dm in enumerate(lst[:5] if k >= 5 else lst):
                if dm.distance > radius:
                    lst = lst[:j]
                    break
        cnt[q] = len(lst)
        for j, dm in enumerate(lst):
            assert dm.queryIdx == q
            trn[q, j], img[q, j], dist[q, j] = dm.trainIdx, dm.imgIdx, dm.distance
    return trn, img, dist, cnt


def save(name, query, objects, k, radius):
    trn, img, dist, cnt = cv2_knn(query, objects, k, radius)
    np.savez_compressed(os.path.join(HERE, name), query=query, db=np.concatenate(objects),
                        sizes=np.array([o.shape[0] for o in objects], np.int32), k=np.int32(k),
                        radius=np.int32(radius), trainIdx=trn, imgIdx=img, distance=dist, counts=cnt)
    print(name, "nq", query.shape[0], "ndb", sum(o.shape[0] for o in objects), "k", k, "radius", radius,
          "mean count %.2f" % cnt.mean())


def main():
    rng = np.random.default_rng(0x70D)
    # 1. tie stress: only two low-entropy bytes differ -> many equal distances, exercises (dist, imgIdx, trainIdx)
    def low_entropy(n):
        d = np.zeros((n, 32), np.uint8)
        d[:, 3] = rng.integers(0, 4, n)
        d[:, 17] = rng.integers(0, 4, n) << 4
        return d
    save("knn_ties_k5.npz", low_entropy(64), [low_entropy(97), low_entropy(33), low_entropy(150)], 5, 0)
    # 2. uniform random, k = 2 and k = 5, ragged object sizes.  NOTE: every object has >= k rows — cv2's BFMatcher returns
    #    garbage (e.g. trainIdx 1111 for a 1-row image) whenever a train image of a multi-image set has fewer than k rows
    #    (probed on cv2 4.13.0: sizes (7,3,9), (300,1,517), (3,300) all wrong at k=5), so that case cannot be pinned.
    objs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (300, 5, 517, 64, 129)]
    q = rng.integers(0, 256, (130, 32), dtype=np.uint8)
    save("knn_random_k2.npz", q, objs, 2, 0)
    save("knn_random_k5.npz", q, objs, 5, 0)
    # 3. true matches + clutter with the .ork radius (35): DB rows with 4% bit flips, duplicates across objects
    objs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (400, 400, 400)]
    objs[2][:50] = objs[0][:50]          # duplicated descriptors on two objects -> cross-object ties at equal distance
    db = np.concatenate(objs)
    pick = rng.integers(0, db.shape[0], 96)
    flips = (rng.random((96, 256)) < 0.04)
    tq = db[pick] ^ np.packbits(flips, axis=1)
    cq = rng.integers(0, 256, (32, 32), dtype=np.uint8)
    save("knn_radius35_k5.npz", np.concatenate([tq, cq]), objs, 5, 35)
    # 4. DB smaller than k (single object: the one short-DB case cv2 handles correctly)
    save("knn_tiny_db_k5.npz", rng.integers(0, 256, (9, 32), dtype=np.uint8),
         [rng.integers(0, 256, (3, 32), dtype=np.uint8)], 5, 0)


if __name__ == "__main__":
    main()
