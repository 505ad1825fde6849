"""CPU: the oracle of the offline training path (oracle/training.py — Trainer.cpp:63-81,121-187, training.cpp:57-195)
against the cv2 of this image where cv2 exposes the operation (cvtColor, erode, resize INTER_NEAREST, gemm, ORB with a
mask and its default parameters)."""
import numpy as np
import pytest

from oracle import orb as oo
from oracle import training as ot
from tod_b200 import synth


def test_pieces_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    bgr = rng.integers(0, 256, (120, 170, 3), dtype=np.uint8)
    assert (ot.bgr_to_gray(bgr) == cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)).all()
    m = (rng.random((90, 130)) < 0.75).astype(np.uint8) * 255
    assert (ot.erode3x3(m, 4) == cv2.erode(m, None, iterations=4)).all()
    d = rng.random((48, 64)).astype(np.float32)
    out = ot.rescale_depth(d, (96, 128))                          # depth at half resolution: nearest neighbour
    assert (out == cv2.resize(d, (128, 96), interpolation=cv2.INTER_NEAREST)).all()
    mm = (d * 1000).astype(np.uint16)
    mm[5, 7] = 0
    z = ot.rescale_depth(mm, (48, 64))
    assert np.isnan(z[5, 7]) and z[6, 7] == np.float32(mm[6, 7]) * np.float32(0.001)
    P, T = rng.random((300, 3)).astype(np.float32), rng.random(3).astype(np.float32)
    R = np.linalg.qr(rng.random((3, 3)))[0].astype(np.float32)
    assert (ot.camera_to_world(R, T, P) == cv2.gemm(P - T.reshape(1, 3), R, 1.0, None, 0.0)).all()


def test_masked_orb_with_default_parameters_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    v = synth.make_training_views(1, seed=5)[0]
    gray = ot.bgr_to_gray(v["image"])
    kps, des = cv2.ORB_create().detectAndCompute(v["image"], v["mask"])          # what Trainer.cpp:142-150 runs
    mine = oo.detect(gray, 500, 8, 1.2, mask=v["mask"])
    sc = oo.level_scales(8)
    ref = {}
    for i, k in enumerate(kps):
        cx, cy = oo.level_center(k.pt[0], k.pt[1], sc[k.octave])
        ref[(k.octave, cx, cy)] = (np.float32(k.response), i)
    got = {(l, x, y): r for l, x, y, r in mine}
    assert set(got) == set(ref) and len(kps) == 500
    assert all(got[k] == ref[k][0] for k in ref)
    xs = np.array([np.float32(x) * sc[l] for l, x, y, r in mine], np.float32)
    ys = np.array([np.float32(y) * sc[l] for l, x, y, r in mine], np.float32)
    oc = np.array([l for l, x, y, r in mine])
    _, d = oo.describe(gray, xs, ys, oc, n_levels=8)
    order = [ref[(l, x, y)][1] for l, x, y, r in mine]
    assert (d == des[order]).all()


def test_validate_keypoints_rules():
    mask = np.zeros((60, 80), np.uint8)
    mask[10:50, 10:70] = 255                                       # eroded 4 times: rows 14..45, columns 14..65
    depth = np.full((60, 80), 1.0, np.float32)
    depth[30, 30] = np.nan
    xs = np.array([40.2, 12.6, 13.4, 30.0, 5.0, 66.4], np.float32)
    ys = np.array([30.0, 30.0, 30.2, 30.0, 5.0, 30.0], np.float32)
    kept, pix = ot.validate_keypoints(xs, ys, mask, depth)
    # 0: inside; 1: rounds to x = 13, 5 x 5 window reaches 14..15 -> nearest masked pixel (14, 30); 2: x = 13 likewise;
    # 3: inside the mask but invalid depth; 4: far outside; 5: rounds to 66, window reaches 64..65 -> (65, 30)
    assert list(kept) == [0, 1, 2, 5]
    assert pix.tolist() == [[40, 30], [14, 30], [14, 30], [65, 30]]


def test_train_observation_and_merge():
    views = synth.make_training_views(2, seed=9)
    ds, ps = [], []
    for v in views:
        d, p, kp = ot.train_observation(v["image"], v["mask"], v["depth"], v["K"], v["R"], v["T"])
        assert 100 < d.shape[0] <= 500 and p.shape == (d.shape[0], 3) and np.isfinite(p).all()
        # camera -> object frame and back: R (p_obj) + T lands on the back-projected camera point
        cam = p.astype(np.float64) @ v["R"].astype(np.float64).T + v["T"].astype(np.float64)
        assert (cam[:, 2] > 0.5).all() and (cam[:, 2] < 2.5).all()
        ds.append(d)
        ps.append(p)
    D, P = ot.merge_points(ds, ps)
    assert D.shape[0] == sum(d.shape[0] for d in ds) and (D[:ds[0].shape[0]] == ds[0]).all()
