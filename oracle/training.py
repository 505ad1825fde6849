"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's offline training path:
src/training/Trainer.cpp:63-81 (rescale_depth), :121-187 (Trainer::process) and src/training/training.cpp:53-195
(validateKeyPoints, mergePoints, cameraToWorld).  Only tests/ and bench legs may import this module.

The cells this path leans on are OpenCV's (un-vendored): cv::ORB with its DEFAULT parameters (Trainer.cpp:142-150
ignores the json parameters: 500 features, 8 levels, scale 1.2) restated in oracle/orb.py; cv::erode (3 x 3, 4
iterations, border = +inf), cv::cvtColor(BGR2GRAY) (15-bit fixed point), cv::resize(INTER_NEAREST), rescaleDepth and
depthTo3dSparse of the rgbd module (published formulas; this image's cv2 has no rgbd), and the float matrix product of
cameraToWorld (OpenCV's gemm accumulates these three products in float, left to right) — each pinned against the cv2
binary where cv2 exposes it (tests/test_training_oracle.py).
"""
import numpy as np

from oracle import orb as oo

F32 = np.float32


def bgr_to_gray(bgr):
    """cv::cvtColor(COLOR_BGR2GRAY) on 8-bit images: (B 3735 + G 19235 + R 9798 + 2^14) >> 15."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def erode3x3(mask, iterations=4):
    """cv::erode(mask, mask, cv::Mat(), Point(-1,-1), 4) (training.cpp:70): 3 x 3 minimum, pixels outside the image do
    not erode (morphologyDefaultBorderValue)."""
    m = mask.copy()
    h, w = m.shape
    for _ in range(iterations):
        p = np.pad(m, 1, constant_values=255)
        out = m.copy()
        for dy in range(3):
            for dx in range(3):
                out = np.minimum(out, p[dy:dy + h, dx:dx + w])
        m = out
    return m


def rescale_depth(depth, image_hw):
    """Trainer.cpp:63-81: depth to float32 metres (uint16 millimetres, 0 -> NaN), then — when the depth image is
    smaller than the colour image — nearest-neighbour resize into the top `dsize.height * factor` rows of an
    image-sized NaN canvas."""
    if depth.dtype == np.uint16:
        d = np.where(depth == 0, np.nan, depth.astype(F32) * F32(0.001)).astype(F32)
    else:
        d = depth.astype(F32)
    ih, iw = image_hw
    dh, dw = d.shape
    if (dh, dw) == (ih, iw):
        return d
    factor = F32(iw) / F32(dw)
    out = np.full((ih, iw), np.nan, F32)
    sub_h = int(F32(dh) * factor)                     # rowRange(0, dsize.height * factor): truncation
    sy = np.minimum(np.floor(np.arange(sub_h) * (dh / sub_h)).astype(np.int64), dh - 1)
    sx = np.minimum(np.floor(np.arange(iw) * (dw / iw)).astype(np.int64), dw - 1)
    out[:sub_h] = d[sy][:, sx]
    return out


def _round_within(v, lo, hi):
    return int(min(max(int(np.rint(F32(v))), lo), hi))


def validate_keypoints(xs, ys, mask, depth):
    """training.cpp:57-145.  Returns (kept keypoint indices, integer pixel (x, y) per kept keypoint): a keypoint is kept
    when it (or, failing that, the nearest masked pixel of its 5 x 5 neighbourhood) lies in the eroded mask and the
    depth there is valid.  Out-of-range accesses the reference would make at x == width / y == height (its clamp is
    inclusive) count as "not in the mask"."""
    m = erode3x3(np.where(mask != 0, 255, 0).astype(np.uint8) if mask.dtype != np.uint8 else mask, 4)
    h, w = m.shape

    def in_mask(y, x):
        return 0 <= y < h and 0 <= x < w and m[y, x] != 0
    kept, pix = [], []
    for i, (fx, fy) in enumerate(zip(xs, ys)):
        x, y = _round_within(fx, 0, w), _round_within(fy, 0, h)
        good = in_mask(y, x)
        if not good:
            best = np.inf
            x0, y0 = x, y
            for ii in range(max(x0 - 2, 0), min(x0 + 2, w) + 1):
                for jj in range(max(y0 - 2, 0), min(y0 + 2, h) + 1):
                    if in_mask(jj, ii):
                        d2 = F32(F32(F32(ii) - F32(fx)) * F32(F32(ii) - F32(fx))) + \
                            F32(F32(F32(jj) - F32(fy)) * F32(F32(jj) - F32(fy)))
                        if d2 < best:
                            best, x, y, good = d2, ii, jj, True
        if not good:
            continue
        z = depth[y, x]
        if np.isnan(z):                               # cv::isValidDepth(float)
            continue
        kept.append(i)
        pix.append((x, y))
    return np.array(kept, np.int64), np.array(pix, np.int64).reshape(-1, 2)


def depth_to_3d_sparse(depth, K, pix):
    """cv::rgbd::depthTo3dSparse at integer pixels: z = depth(v, u), x = ((u - cx) / fx) z, y = ((v - cy) / fy) z
    (published formula; cv2 here has no rgbd module: parity unpinned for this function)."""
    K = np.asarray(K, F32).reshape(3, 3)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u, v = pix[:, 0], pix[:, 1]
    z = depth[v, u].astype(F32)
    x = ((u.astype(F32) - cx).astype(F32) / fx).astype(F32) * z       # depthTo3d_from_uvz: (u - cx) / fx, then .mul(z)
    y = ((v.astype(F32) - cy).astype(F32) / fy).astype(F32) * z
    return np.stack([x.astype(F32), y.astype(F32), z], axis=1)


def camera_to_world(R, T, pts):
    """training.cpp:175-195: (p - T) * R with float accumulation, the three products added left to right."""
    R = np.asarray(R, F32).reshape(3, 3)
    a = (pts.astype(F32) - np.asarray(T, F32).reshape(1, 3)).astype(F32)
    out = np.zeros_like(a)
    for j in range(3):
        s = (a[:, 0] * R[0, j]).astype(F32)
        s = (s + (a[:, 1] * R[1, j]).astype(F32)).astype(F32)
        s = (s + (a[:, 2] * R[2, j]).astype(F32)).astype(F32)
        out[:, j] = s
    return out


def train_observation(image, mask, depth, K, R, T, n_features=500, n_levels=8, scale_factor=1.2):
    """One pass of the loop of Trainer::process (Trainer.cpp:134-171).  Returns (descriptors n x 32, points n x 3,
    keypoints as (octave, x, y) level-0 coordinates) in the order (octave, row, column) of the detected keypoints."""
    gray = bgr_to_gray(image) if image.ndim == 3 else image
    kps = oo.detect(gray, n_features, n_levels, scale_factor, mask=mask)
    sc = oo.level_scales(n_levels, scale_factor)
    xs = np.array([F32(x) * sc[l] for l, x, y, r in kps], F32)
    ys = np.array([F32(y) * sc[l] for l, x, y, r in kps], F32)
    oc = np.array([l for l, x, y, r in kps], np.int64)
    _, desc = oo.describe(gray, xs, ys, oc, n_levels=n_levels, scale_factor=scale_factor)
    d = rescale_depth(depth, gray.shape)
    kept, pix = validate_keypoints(xs, ys, mask, d)
    if kept.size == 0:
        return np.zeros((0, 32), np.uint8), np.zeros((0, 3), F32), []
    p3 = depth_to_3d_sparse(d, K, pix)
    world = camera_to_world(R, T, p3)
    return desc[kept], world, [(int(oc[i]), float(xs[i]), float(ys[i])) for i in kept]


def merge_points(desc_list, points_list):
    """training.cpp:147-173: views stacked in order."""
    if not desc_list:
        return np.zeros((0, 32), np.uint8), np.zeros((0, 3), F32)
    return np.concatenate(desc_list), np.concatenate(points_list).astype(F32)
