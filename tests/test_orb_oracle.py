"""CPU: the oracle of the feature stage (oracle/orb.py: ORB pyramid, smoothing, orientation, steered BRIEF; depth -> 3D)
against the cv2 of this image stage by stage, and against the committed golden vectors (tests/golden/orb_*.npz, made by
tests/golden/make_orb_golden.py with cv2.ORB_create(5000, 1.2, 3) — conf/detection.ork:23-31)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import orb as oo
from tod_b200 import synth

ORB_GOLDENS = ["orb_640x480.npz", "orb_333x517_ragged.npz"]


def test_pattern_table_is_the_one_compiled_into_the_library():
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "tod_b200", "csrc", "orb_pattern.h")).read()
    import re
    rows = re.findall(r"\{(-?\d+), (-?\d+), (-?\d+), (-?\d+)\}", src)
    assert len(rows) == 256
    assert (np.array(rows, np.int64) == oo.PATTERN).all()
    assert oo.PATTERN.min() == -13 and oo.PATTERN.max() == 13 or oo.PATTERN.max() == 12


@pytest.mark.parametrize("name", ORB_GOLDENS)
def test_oracle_reproduces_cv2_golden_descriptors_and_angles(name):
    g = np.load(os.path.join(GOLDEN, name))
    img = synth.make_textured_image(int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    ang, des = oo.describe(img, g["x"], g["y"], g["octave"])
    assert (ang == g["angle"]).all()                      # exact float equality (fastAtan2 restated operation by operation)
    assert (des == g["descriptors"]).all()
    ang2, des2 = oo.describe(img, g["x"], g["y"], g["octave"], angles=g["angle"])
    assert (des2 == des).all()


def test_stages_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)[:, 0]
    assert (k == oo.GAUSS7).all()
    for h, w in ((97, 131), (480, 643)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        # ORB smooths a sub-matrix of its pyramid buffer: OpenCV's float separable filter, not the fixed-point blur
        ref = cv2.sepFilter2D(img, cv2.CV_8U, k.reshape(-1, 1), k.reshape(-1, 1), borderType=cv2.BORDER_REFLECT_101)
        assert (oo.smooth(img) == ref).all()
        for (dh, dw) in oo.level_sizes(h, w, 3)[1:]:
            assert (oo.resize_linear_exact(img, dh, dw) ==
                    cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)).all()
    for _ in range(5000):
        y, x = float(rng.integers(-300000, 300000)), float(rng.integers(-300000, 300000))
        assert oo.fast_atan2(y, x) == np.float32(cv2.fastAtan2(y, x))
    assert oo.UMAX == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_full_frame_against_live_cv2_orb():
    cv2 = pytest.importorskip("cv2")
    img = synth.make_textured_image(480, 640, seed=21)
    kp, des = cv2.ORB_create(5000, 1.2, 3).detectAndCompute(img, None)
    assert len(kp) > 1500
    sel = np.arange(0, len(kp), 2)
    xs = np.array([kp[i].pt[0] for i in sel], np.float32)
    ys = np.array([kp[i].pt[1] for i in sel], np.float32)
    oc = np.array([kp[i].octave for i in sel])
    ang, mine = oo.describe(img, xs, ys, oc)
    assert (ang == np.array([kp[i].angle for i in sel], np.float32)).all()
    assert (mine == des[sel]).all()


@pytest.mark.parametrize("name", ORB_GOLDENS)
def test_oracle_detection_reproduces_cv2_keypoint_set(name):
    """cv::ORB::detect restated (FAST score + NMS, border, 2 N best, Harris, N best): the same SET of keypoints with the
    same Harris responses as the golden (cv2's order is an artefact of nth_element, so sets are compared)."""
    g = np.load(os.path.join(GOLDEN, name))
    img = synth.make_textured_image(int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    mine = oo.detect(img, 5000, 3, 1.2)
    sc = oo.level_scales(3)
    ref = {}
    for x, y, o, r in zip(g["x"], g["y"], g["octave"], g["response"]):
        cx, cy = oo.level_center(x, y, sc[int(o)])
        ref[(int(o), cx, cy)] = np.float32(r)
    got = {(l, x, y): r for l, x, y, r in mine}
    assert set(got) == set(ref) and len(mine) == len(got) == int(g["n_detected"])
    assert all(got[k] == ref[k] for k in ref)
    assert oo.features_per_level(5000, 3) == [1978, 1648, 1374]


def test_fast_and_nms_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    img = synth.make_textured_image(300, 400, seed=77)
    kps = cv2.FastFeatureDetector_create(20, True).detect(img, None)
    s = oo.fast_scores(img)
    ys, xs = np.nonzero(oo.non_max_suppression(s))
    assert {(int(k.pt[0]), int(k.pt[1])): int(k.response) for k in kps} == \
        {(int(x), int(y)): int(s[y, x]) for x, y in zip(xs, ys)}


def test_depth_to_3d():
    zf, mm = synth.make_depth_image(120, 160, seed=3)
    K = np.array([[525.0, 0, 79.5], [0, 525.0, 59.5], [0, 0, 1]])
    p = oo.depth_to_3d(zf, K)
    assert p.shape == (120, 160, 3) and p.dtype == np.float32
    hole = np.isnan(zf)
    assert np.isnan(p[hole]).all() and np.isfinite(p[~hole]).all()
    u, v = np.meshgrid(np.arange(160.0), np.arange(120.0))
    ref = np.stack([(u - 79.5) * zf / 525.0, (v - 59.5) * zf / 525.0, zf.astype(np.float64)], axis=2)
    assert np.nanmax(np.abs(p - ref)) < 1e-6
    q = oo.depth_to_3d(mm, K)                                   # uint16 millimetres, 0 = invalid
    assert np.isnan(q[mm == 0]).all() and np.nanmax(np.abs(q - p)) < 1e-3
