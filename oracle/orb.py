"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the feature stage that precedes the hot path in the reference's
detection pipeline (python/object_recognition_tod/detector.py:26-27, 62-75): ecto_opencv's FeatureDescriptor cell =
cv::ORB (n_features 5000, n_levels 3, scale_factor 1.2 — conf/detection.ork:23-31) and DepthTo3d.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module.

OpenCV is un-vendored (package.xml:12-22 names no version) and its sources are not in this image; the arithmetic
below was established against the cv2 4.13.0 binary of this image, stage by stage, and is pinned by
tests/test_orb_oracle.py (bit-exact against cv2.ORB on synthetic frames) and the committed golden vectors
tests/golden/orb_*.npz (made by tests/golden/make_orb_golden.py):

  pyramid     level l = resize(level l-1, INTER_LINEAR_EXACT) to cvRound(size / scale_l), scale_l = (float) 1.2f^l:
              two passes with 8-bit fixed-point weights, (sum + 2^15) >> 16.
  smoothing   GaussianBlur(7 x 7, sigma 2, BORDER_REFLECT_101) on a sub-matrix takes OpenCV's float separable filter:
              row pass s = k0 x0, s = fma(x_j, k_j, s) left to right; column pass centre first, then
              fma(x_{+j} + x_{-j}, k_j, s); round half to even.  (FMA = what cv2 does on an AVX2/FMA host.)
  orientation intensity centroid over the radius-15 disc (umax table), fastAtan2's degree-7 polynomial WITHOUT
              contraction, result in degrees.
  detection   FAST-9/16 score (threshold 20) with 3 x 3 non-maximum suppression per level, edgeThreshold 31, the 2 N
              best FAST scores, Harris response (block 7, k = 0.04, float arithmetic operation by operation), the N
              best responses per level (N from ORB's geometric series); keypoint sets and responses equal cv2's.
  descriptor  steered BRIEF: 256 comparisons of the smoothed level image at pattern points rotated by the angle
              (cosf / sinf of angle * pi/180 in float, products and differences in float, cvRound = half to even);
              pattern = oracle/orb_pattern.npy (recovered from cv2, tools/recover_orb_pattern.py).
  depth -> 3D x = ((u - cx) (1 / fx)) z, y = ((v - cy) (1 / fy)) z, z = depth (metres; uint16 = millimetres, 0 = invalid
              -> NaN): the published formula of cv::rgbd::depthTo3d, which this image's cv2 does not ship — parity
              unpinned for this one function (checked against a float64 evaluation only).
"""
import ctypes
import math
import os

import numpy as np

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN = np.load(os.path.join(_HERE, "orb_pattern.npy")).astype(np.int64)     # 256 x (x_a, y_a, x_b, y_b)
HALF_PATCH = 15

# cv::getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) normalised in double, stored as float
_g = np.exp(-(np.arange(7) - 3.0) ** 2 / 8.0)
GAUSS7 = (_g / _g.sum()).astype(F32)

_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = _libm.sinf.restype = ctypes.c_float
_libm.cosf.argtypes = _libm.sinf.argtypes = [ctypes.c_float]


def _fma(a, b, c):
    """float32 fma(a, b, c) for arrays a, c and a scalar b (the product of two floats is exact in float64)."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(F32)


def level_scales(n_levels, scale_factor=1.2):
    sf = float(F32(scale_factor))                      # the parameter travels as a float, pow runs in double
    return [F32(sf ** l) for l in range(n_levels)]


def level_sizes(height, width, n_levels, scale_factor=1.2):
    sc = level_scales(n_levels, scale_factor)
    return [(int(np.rint(F32(height) / s)), int(np.rint(F32(width) / s))) for s in sc]


def linear_exact_table(src, dst):
    """Source indices and 8-bit weights of cv::resize(INTER_LINEAR_EXACT) along one axis."""
    scale = 1.0 / (dst / src)
    x = (np.arange(dst) + 0.5) * scale - 0.5
    ix = np.floor(x).astype(np.int64)
    f = np.floor((x - ix) * 256 + 0.5).astype(np.int64)
    return np.clip(ix, 0, src - 1), np.clip(ix + 1, 0, src - 1), f


def resize_linear_exact(img, dh, dw):
    sh, sw = img.shape
    x0, x1, fx = linear_exact_table(sw, dw)
    y0, y1, fy = linear_exact_table(sh, dh)
    im = img.astype(np.int64)
    h = im[:, x0] * (256 - fx)[None, :] + im[:, x1] * fx[None, :]
    v = h[y0, :] * (256 - fy)[:, None] + h[y1, :] * fy[:, None]
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def pyramid(img, n_levels=3, scale_factor=1.2):
    levels = [np.ascontiguousarray(img, np.uint8)]
    for l, (h, w) in enumerate(level_sizes(img.shape[0], img.shape[1], n_levels, scale_factor)):
        if l:
            levels.append(resize_linear_exact(levels[-1], h, w))
    return levels


def smooth(img):
    h_, w_ = img.shape
    p = np.pad(img, ((0, 0), (3, 3)), mode="reflect").astype(F32)
    s = (p[:, 0:w_] * GAUSS7[0]).astype(F32)
    for j in range(1, 7):
        s = _fma(p[:, j:j + w_], GAUSS7[j], s)
    q = np.pad(s, ((3, 3), (0, 0)), mode="reflect")
    v = (q[3:3 + h_] * GAUSS7[3]).astype(F32)
    for k in (1, 2, 3):
        v = _fma((q[3 + k:3 + k + h_] + q[3 - k:3 - k + h_]).astype(F32), GAUSS7[3 + k], v)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def umax_table(hp=HALF_PATCH):
    umax = [0] * (hp + 2)
    vmax = int(math.floor(hp * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(hp * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(math.sqrt(float(hp * hp - v * v))))
    v0 = 0
    for v in range(hp, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax[:hp + 1]


UMAX = umax_table()
_P1 = F32(0.9997878412794807) * F32(180 / np.pi)
_P3 = F32(-0.3258083974640975) * F32(180 / np.pi)
_P5 = F32(0.1555786518463281) * F32(180 / np.pi)
_P7 = F32(-0.04432655554792128) * F32(180 / np.pi)
_EPS = F32(2.220446049250313e-16)


def fast_atan2(y, x):
    """cv::fastAtan2 (degrees, [0, 360)): scalar float arithmetic, no contraction."""
    y, x = F32(y), F32(x)
    ax, ay = abs(x), abs(y)

    def poly(c):
        c2 = F32(c * c)
        r = F32(F32(_P7 * c2) + _P5)
        r = F32(F32(r * c2) + _P3)
        r = F32(F32(r * c2) + _P1)
        return F32(r * c)
    if ax >= ay:
        a = poly(F32(ay / F32(ax + _EPS)))
    else:
        a = F32(F32(90.0) - poly(F32(ax / F32(ay + _EPS))))
    if x < 0:
        a = F32(F32(180.0) - a)
    if y < 0:
        a = F32(F32(360.0) - a)
    return a


def level_center(x, y, scale):
    inv = F32(1.0) / scale
    return int(np.rint(F32(x) * inv)), int(np.rint(F32(y) * inv))


def ic_angle(level_img, cx, cy):
    """Orientation of the patch centred at (cx, cy) of an UNSMOOTHED pyramid level."""
    m01 = m10 = 0
    row = level_img[cy].astype(np.int64)
    u = np.arange(-HALF_PATCH, HALF_PATCH + 1)
    m10 += int((u * row[cx - HALF_PATCH:cx + HALF_PATCH + 1]).sum())
    for v in range(1, HALF_PATCH + 1):
        d = UMAX[v]
        uu = np.arange(-d, d + 1)
        vp = level_img[cy + v, cx - d:cx + d + 1].astype(np.int64)
        vm = level_img[cy - v, cx - d:cx + d + 1].astype(np.int64)
        m10 += int((uu * (vp + vm)).sum())
        m01 += v * int((vp - vm).sum())
    return fast_atan2(m01, m10)


def describe(img, xs, ys, octaves, angles=None, n_levels=3, scale_factor=1.2):
    """Angles (computed when `angles` is None) and 32-byte descriptors of keypoints given in level-0 coordinates."""
    levels = pyramid(img, n_levels, scale_factor)
    sc = level_scales(n_levels, scale_factor)
    smoothed = [smooth(l) for l in levels]
    n = len(xs)
    out_angles = np.zeros(n, F32)
    desc = np.zeros((n, 32), np.uint8)
    px = PATTERN[:, [0, 2]].reshape(-1).astype(F32)        # a0, b0, a1, b1, ...
    py = PATTERN[:, [1, 3]].reshape(-1).astype(F32)
    for i in range(n):
        l = int(octaves[i])
        cx, cy = level_center(xs[i], ys[i], sc[l])
        ang = ic_angle(levels[l], cx, cy) if angles is None else F32(angles[i])
        out_angles[i] = ang
        rad = F32(ang * F32(np.pi / 180.0))
        a, b = F32(_libm.cosf(rad)), F32(_libm.sinf(rad))
        x = ((px * a).astype(F32) - (py * b).astype(F32)).astype(F32)
        y = ((px * b).astype(F32) + (py * a).astype(F32)).astype(F32)
        ix, iy = np.rint(x).astype(np.int64), np.rint(y).astype(np.int64)
        v = smoothed[l][cy + iy, cx + ix]
        desc[i] = np.packbits((v[0::2] < v[1::2]).astype(np.uint8), bitorder="little")
    return out_angles, desc


# ---- detection: cv::ORB::detect = FAST-9/16 (threshold 20, 3 x 3 non-maximum suppression) on every pyramid level, border
# filter (edgeThreshold 31), the 2 N best FAST scores, Harris response (block 7, k 0.04), the N best Harris responses
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
        (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def fast_scores(img, threshold=20):
    """cv::FAST corner score of every pixel (0 = not a corner): the largest threshold that keeps it a corner."""
    h, w = img.shape
    im = img.astype(np.int64)
    c = im[3:h - 3, 3:w - 3]
    d = np.stack([im[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in RING]) - c[None]
    best = np.full(c.shape, -10 ** 9)
    for s in range(16):
        idx = [(s + j) % 16 for j in range(9)]
        best = np.maximum(best, np.maximum(d[idx].min(axis=0), (-d[idx]).min(axis=0)))
    out = np.zeros((h, w), np.int64)
    out[3:h - 3, 3:w - 3] = np.where(best > threshold, best - 1, 0)
    return out


def non_max_suppression(score):
    h, w = score.shape
    keep = score > 0
    p = np.pad(score, 1)
    for dy in range(3):
        for dx in range(3):
            if dx != 1 or dy != 1:
                keep &= score > p[dy:dy + h, dx:dx + w]
    return keep


def harris_responses(level, xs, ys):
    im = level.astype(np.int64)
    ix = np.zeros_like(im)
    iy = np.zeros_like(im)
    ix[1:-1, 1:-1] = (im[1:-1, 2:] - im[1:-1, :-2]) * 2 + (im[:-2, 2:] - im[:-2, :-2]) + (im[2:, 2:] - im[2:, :-2])
    iy[1:-1, 1:-1] = (im[2:, 1:-1] - im[:-2, 1:-1]) * 2 + (im[2:, :-2] - im[:-2, :-2]) + (im[2:, 2:] - im[:-2, 2:])
    scale = F32(1.0) / F32(F32(4 * 7) * F32(255.0))
    s4 = F32(F32(F32(scale * scale) * scale) * scale)
    out = np.zeros(len(xs), F32)
    for i, (x, y) in enumerate(zip(xs, ys)):
        bx, by = ix[y - 3:y + 4, x - 3:x + 4], iy[y - 3:y + 4, x - 3:x + 4]
        fa, fb, fc = F32(int((bx * bx).sum())), F32(int((by * by).sum())), F32(int((bx * by).sum()))
        t = F32(F32(F32(0.04) * F32(fa + fb)) * F32(fa + fb))
        out[i] = F32(F32(F32(F32(fa * fb) - F32(fc * fc)) - t) * s4)
    return out


def retain_best(resp, n):
    """KeyPointsFilter::retainBest as a mask: everything that reaches the n-th best response (ties are all kept)."""
    if len(resp) <= n:
        return np.ones(len(resp), bool)
    if n == 0:
        return np.zeros(len(resp), bool)
    return resp >= np.sort(resp)[::-1][n - 1]


def features_per_level(n_features, n_levels, scale_factor=1.2):
    f = F32(1.0 / float(F32(scale_factor)))
    nd = F32(F32(n_features) * F32(F32(1.0) - f) / F32(F32(1.0) - F32(float(f) ** n_levels)))
    out, s = [], 0
    for _ in range(n_levels - 1):
        out.append(int(np.rint(nd)))
        s += out[-1]
        nd = F32(nd * f)
    out.append(max(n_features - s, 0))
    return out


def mask_pyramid(mask, n_levels=3, scale_factor=1.2):
    """ORB's mask pyramid: every level is the previous one resized (INTER_LINEAR_EXACT), then values below 255 -> 0
    (cv::threshold(254, THRESH_TOZERO))."""
    levels = [np.ascontiguousarray(mask, np.uint8)]
    for l, (h, w) in enumerate(level_sizes(mask.shape[0], mask.shape[1], n_levels, scale_factor)):
        if l:
            m = resize_linear_exact(levels[-1], h, w)
            levels.append(np.where(m > 254, m, 0).astype(np.uint8))
    return levels


def detect(img, n_features=5000, n_levels=3, scale_factor=1.2, mask=None):
    """Keypoints of cv::ORB::detect as a list of (octave, x_level, y_level, harris_response), ordered by level, row,
    column (cv2's own order is an artefact of nth_element).  mask: u8 image, keypoints on zero pixels are dropped
    (KeyPointsFilter::runByPixelsMask on every level's resized mask)."""
    levels = pyramid(img, n_levels, scale_factor)
    masks = mask_pyramid(mask, n_levels, scale_factor) if mask is not None else None
    per_level = features_per_level(n_features, n_levels, scale_factor)
    out = []
    for l, lev in enumerate(levels):
        s = fast_scores(lev)
        keep = non_max_suppression(s)
        if masks is not None:
            keep &= masks[l] != 0
        ys, xs = np.nonzero(keep)
        h, w = lev.shape
        inside = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
        xs, ys = xs[inside], ys[inside]
        k1 = retain_best(s[ys, xs].astype(F32), 2 * per_level[l])
        xs, ys = xs[k1], ys[k1]
        hr = harris_responses(lev, xs, ys)
        k2 = retain_best(hr, per_level[l])
        out += [(l, int(x), int(y), F32(r)) for x, y, r in zip(xs[k2], ys[k2], hr[k2])]
    return out


def depth_to_3d(depth, K):
    """H x W x 3 float32 point image from a depth image (float32 metres, or uint16 millimetres) and the 3 x 3 camera
    matrix K: the cell detector.py:62-69 wires in front of the GuessGenerator.  Invalid depth (NaN, or 0 for uint16)
    -> NaN point."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    fx, fy, cx, cy = F32(K[0, 0]), F32(K[1, 1]), F32(K[0, 2]), F32(K[1, 2])
    if depth.dtype == np.uint16:
        z = np.where(depth == 0, np.nan, depth.astype(F32) * F32(0.001)).astype(F32)
    else:
        z = depth.astype(F32)
    h, w = z.shape
    u = np.arange(w, dtype=F32)[None, :]
    v = np.arange(h, dtype=F32)[:, None]
    # cv::rgbd::depthTo3d (dense, no mask) caches (u - cx) * (1 / fx) per column and (v - cy) * (1 / fy) per row, then
    # multiplies by z — restated from the published source, which this image does not hold
    inv_fx, inv_fy = F32(1.0) / fx, F32(1.0) / fy
    x = ((u - cx).astype(F32) * inv_fx).astype(F32) * z
    y = ((v - cy).astype(F32) * inv_fy).astype(F32) * z
    out = np.stack([x.astype(F32), y.astype(F32), z], axis=2)
    out[np.isnan(z)] = np.nan
    return out
