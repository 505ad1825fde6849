"""Seeded synthetic inputs of the shapes BASELINE.json names (SURVEY.md §8d recipe).  numpy only; used by tests/ and
bench.py on both sides of every comparison (CUDA path, oracle, reference arm)."""
import numpy as np

BASE_SEED = 0x70D


def make_db(n_objects, rows_per_object, seed=BASE_SEED, span=0.25):
    """Per-object descriptors (rows x 32 u8, i.i.d. uniform bits) and 3-D points (rows x 3 f32, object frame, metres,
    uniform in a box whose diagonal is ~span)."""
    rng = np.random.default_rng(seed)
    if np.isscalar(rows_per_object):
        rows_per_object = [int(rows_per_object)] * n_objects
    descs, points = [], []
    side = span / np.sqrt(3.0)
    for n in rows_per_object:
        descs.append(rng.integers(0, 256, (n, 32), dtype=np.uint8))
        points.append(((rng.random((n, 3)) - 0.5) * side).astype(np.float32))
    return descs, points


def make_queries(descs, nq, seed=BASE_SEED + 1, true_fraction=0.5, flip_p=0.04):
    """nq query descriptors: a fraction are DB rows with each bit flipped w.p. flip_p (mean Hamming ~10 < radius 35),
    the rest fresh uniform clutter.  Returns (queries, src_object[nq] (-1 clutter), src_row[nq])."""
    rng = np.random.default_rng(seed)
    sizes = np.array([d.shape[0] for d in descs])
    off = np.concatenate([[0], np.cumsum(sizes)])
    n_true = int(round(nq * true_fraction))
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    src_obj = -np.ones(nq, np.int64)
    src_row = -np.ones(nq, np.int64)
    if n_true and off[-1] > 0:
        g = rng.integers(0, off[-1], n_true)
        o = np.searchsorted(off, g, side="right") - 1
        r = g - off[o]
        flips = np.packbits(rng.random((n_true, 256)) < flip_p, axis=1)
        for i in range(n_true):
            q[i] = descs[o[i]][r[i]] ^ flips[i]
        src_obj[:n_true], src_row[:n_true] = o, r
    perm = rng.permutation(nq)
    return np.ascontiguousarray(q[perm]), src_obj[perm], src_row[perm]


def random_rotation(rng):
    a = rng.normal(size=(3, 3))
    qm, r = np.linalg.qr(a)
    qm = qm * np.sign(np.diag(r))
    if np.linalg.det(qm) < 0:
        qm[:, 0] = -qm[:, 0]
    return qm


def make_frame(descs, points, visible_objects, n_keypoints, height=480, width=640, seed=BASE_SEED + 2,
               noise_sigma=0.002, flip_p=0.04, clutter_fraction=0.3, duplicate_outliers=0.0):
    """One synthetic RGB-D frame: for each visible object a random rigid pose at 0.6-1.2 m, its model
    points projected with a Kinect-like K; keypoints = projected pixels (+U(0,1)), descriptors = model descriptors with
    bit flips, cloud = H x W x 3 f32 (NaN where no point).  Clutter keypoints get random descriptors and random depth.
    Returns dict(keypoints_xy[n,2] f32, descriptors[n,32] u8, cloud[H,W,3] f32, poses {obj: (R, T)} object->camera)."""
    rng = np.random.default_rng(seed)
    f = 525.0 * width / 640.0
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    cloud = np.full((height, width, 3), np.nan, np.float32)
    kps, ds = [], []
    poses = {}
    n_clutter = int(n_keypoints * clutter_fraction)
    per_obj = max(1, (n_keypoints - n_clutter) // max(1, len(visible_objects)))
    taken = set()
    for o in visible_objects:
        R = random_rotation(rng)
        T = np.array([rng.uniform(-0.25, 0.25), rng.uniform(-0.2, 0.2), rng.uniform(0.6, 1.2)])
        poses[o] = (R.astype(np.float32), T.astype(np.float32))
        idx = rng.permutation(points[o].shape[0])
        got = 0
        for i in idx:
            if got >= per_obj:
                break
            pc = R @ points[o][i].astype(np.float64) + T
            if pc[2] <= 0.1:
                continue
            u, v = f * pc[0] / pc[2] + cx, f * pc[1] / pc[2] + cy
            xi, yi = int(u), int(v)
            if not (0 <= xi < width - 1 and 0 <= yi < height - 1) or (yi, xi) in taken:
                continue
            taken.add((yi, xi))
            cloud[yi, xi] = (pc + rng.normal(0, noise_sigma, 3)).astype(np.float32)
            kps.append((xi + rng.random() * 0.99, yi + rng.random() * 0.99))
            flips = np.packbits(rng.random(256) < flip_p)
            ds.append(descs[o][i] ^ flips)
            got += 1
    while len(kps) < n_keypoints:
        xi, yi = int(rng.integers(0, width - 1)), int(rng.integers(0, height - 1))
        if (yi, xi) in taken:
            continue
        taken.add((yi, xi))
        z = rng.uniform(0.5, 2.0)
        cloud[yi, xi] = np.array([(xi - cx) * z / f, (yi - cy) * z / f, z], np.float32)
        kps.append((xi + rng.random() * 0.99, yi + rng.random() * 0.99))
        if duplicate_outliers > 0 and rng.random() < duplicate_outliers and visible_objects:
            o = visible_objects[int(rng.integers(0, len(visible_objects)))]
            i = int(rng.integers(0, descs[o].shape[0]))
            ds.append(descs[o][i] ^ np.packbits(rng.random(256) < flip_p))  # right descriptor, wrong 3-D place
        else:
            ds.append(rng.integers(0, 256, 32, dtype=np.uint8))
    perm = rng.permutation(len(kps))
    return {"keypoints_xy": np.array(kps, np.float32)[perm], "descriptors": np.array(ds, np.uint8)[perm],
            "cloud": cloud, "poses": poses}


def make_cluster(n, inlier_fraction=0.5, seed=BASE_SEED + 3, span=0.25, noise_sigma=0.002, width=640, height=480):
    """One (frame, object) cluster of n correspondences at the GuessGenerator boundary: query (camera-frame) points,
    training (object-frame) points, pixels.  A fraction are consistent with one rigid pose, the rest are random."""
    rng = np.random.default_rng(seed)
    side = span / np.sqrt(3.0)
    train = ((rng.random((n, 3)) - 0.5) * side).astype(np.float32)
    R = random_rotation(rng)
    T = np.array([0.05, -0.02, 0.9])
    query = (train.astype(np.float64) @ R.T + T + rng.normal(0, noise_sigma, (n, 3)))
    out = rng.random(n) >= inlier_fraction
    query[out] = (rng.random((int(out.sum()), 3)) - 0.5) * side * 1.5 + T
    query = query.astype(np.float32)
    f = 525.0 * width / 640.0
    px = np.stack([f * query[:, 0] / query[:, 2] + (width - 1) / 2.0,
                   f * query[:, 1] / query[:, 2] + (height - 1) / 2.0], axis=1).astype(np.float32)
    return query, train, px, (R.astype(np.float32), T.astype(np.float32)), ~out


def make_guess_inputs(n_objects, n_per_object, inlier_fraction, seed=BASE_SEED + 4, span=0.25, noise_sigma=0.002,
                      height=480, width=640, k=2):
    """Inputs injected directly at the GuessGenerator boundary (BASELINE config C5: outlier-heavy matches): every
    keypoint carries one match to one object; a fraction of each object's correspondences follow the object's planted
    rigid pose, the rest are geometrically inconsistent.  Returns dict(keypoints_xy, cloud, matches (MATCH dtype
    [n_kp,k]), counts, points3d [n_kp,k,3], spans [n_objects], poses {obj: (R, T)} object->camera)."""
    rng = np.random.default_rng(seed)
    f = 525.0 * width / 640.0
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    side = span / np.sqrt(3.0)
    n_kp = n_objects * n_per_object
    cloud = np.full((height, width, 3), np.nan, np.float32)
    mdt = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
    matches = np.zeros((n_kp, k), mdt)
    matches["queryIdx"] = -1
    matches["trainIdx"] = -1
    matches["imgIdx"] = -1
    counts = np.zeros(n_kp, np.int32)
    p3 = np.zeros((n_kp, k, 3), np.float32)
    kps = np.zeros((n_kp, 2), np.float32)
    poses = {}
    taken = set()
    order = rng.permutation(n_kp)
    qi = 0
    for o in range(n_objects):
        R = random_rotation(rng)
        T = np.array([rng.uniform(-0.2, 0.2), rng.uniform(-0.15, 0.15), rng.uniform(0.7, 1.1)])
        poses[o] = (R.astype(np.float32), T.astype(np.float32))
        for i in range(n_per_object):
            tp = (rng.random(3) - 0.5) * side
            if rng.random() < inlier_fraction:
                qp = R @ tp + T + rng.normal(0, noise_sigma, 3)
            else:
                qp = (rng.random(3) - 0.5) * side * 1.6 + T
            while True:
                u, v = f * qp[0] / qp[2] + cx, f * qp[1] / qp[2] + cy
                xi, yi = int(u), int(v)
                if 0 <= xi < width and 0 <= yi < height and (yi, xi) not in taken:
                    break
                qp = qp + rng.normal(0, 0.01, 3)   # nudge until it lands on a free pixel
            taken.add((yi, xi))
            q = int(order[qi])
            qi += 1
            cloud[yi, xi] = qp.astype(np.float32)
            kps[q] = (xi + rng.random() * 0.99, yi + rng.random() * 0.99)
            matches[q, 0] = (q, i, o, float(rng.integers(0, 30)))
            counts[q] = 1
            p3[q, 0] = tp.astype(np.float32)
    spans = np.full(n_objects, np.float32(span), np.float32)
    return {"keypoints_xy": kps, "cloud": cloud, "matches": matches, "counts": counts, "points3d": p3,
            "spans": spans, "poses": poses}


def make_textured_image(height=480, width=640, seed=BASE_SEED + 6, n_shapes=400):
    """Deterministic grey image with corners for the feature stage (numpy integer arithmetic only, so every machine
    regenerates the same bytes): overlapping random rectangles on a soft gradient, speckle, one 3 x 3 box pass."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width]
    img = ((xx * 40) // max(1, width) + (yy * 30) // max(1, height) + 60).astype(np.int64)
    for _ in range(n_shapes):
        w, h = int(rng.integers(6, 60)), int(rng.integers(6, 60))
        x, y = int(rng.integers(0, width - 5)), int(rng.integers(0, height - 5))
        img[y:y + h, x:x + w] = int(rng.integers(0, 256))
    img += rng.integers(-6, 7, img.shape)
    img = np.clip(img, 0, 255)
    p = np.pad(img, 1, mode="edge")
    box = sum(p[dy:dy + height, dx:dx + width] for dy in range(3) for dx in range(3))
    return ((box + 4) // 9).astype(np.uint8)


def make_depth_image(height=480, width=640, seed=BASE_SEED + 7, invalid_fraction=0.1):
    """Float32 depth in metres (a tilted plane with bumps, 0.5 - 2 m) with NaN holes, and its uint16 millimetre twin
    (0 = invalid)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    z = (0.8 + 0.6 * xx / width + 0.3 * yy / height + 0.05 * np.sin(xx / 17.0) * np.cos(yy / 23.0)).astype(np.float32)
    hole = rng.random((height, width)) < invalid_fraction
    zf = z.copy()
    zf[hole] = np.nan
    mm = np.where(hole, 0, np.rint(z * 1000.0)).astype(np.uint16)
    return zf, mm


def make_training_views(n_views=3, height=480, width=640, seed=BASE_SEED + 8, depth_scale=1, bgr=True):
    """Synthetic observations of one object for the training path (Trainer.cpp:134-171): per view a textured BGR (or
    grey) image, an object mask (a disc with a hole), a depth image (float32 metres with NaN holes; optionally at
    1/depth_scale of the image resolution), the camera matrix and a pose (R, T).  numpy only."""
    rng = np.random.default_rng(seed)
    views = []
    for v in range(n_views):
        g = make_textured_image(height, width, seed=seed + 10 * v + 1)
        if bgr:
            img = np.stack([g, np.roll(g, 3, axis=1), np.roll(g, -2, axis=0)], axis=2).astype(np.uint8)
        else:
            img = g
        yy, xx = np.mgrid[0:height, 0:width]
        cx, cy = width // 2 + int(rng.integers(-40, 40)), height // 2 + int(rng.integers(-30, 30))
        mask = (((xx - cx) ** 2 + (yy - cy) ** 2) < (min(height, width) * 0.38) ** 2).astype(np.uint8) * 255
        mask[cy - 20:cy + 25, cx - 60:cx - 20] = 0
        dh, dw = height // depth_scale, width // depth_scale
        zf, _ = make_depth_image(dh, dw, seed=seed + 10 * v + 2, invalid_fraction=0.05)
        f = 525.0 * width / 640.0
        K = np.array([[f, 0, (width - 1) / 2.0], [0, f, (height - 1) / 2.0], [0, 0, 1]], np.float32)
        R = random_rotation(rng).astype(np.float32)
        T = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0.6, 0.9)], np.float32)
        views.append({"image": img, "mask": mask, "depth": zf, "K": K, "R": R, "T": T})
    return views
