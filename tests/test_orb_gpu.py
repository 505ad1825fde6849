"""GPU: the feature stage (tod_orb_describe, tod_depth_to_3d) through the C-ABI — orientation as exact floats and
descriptors bit for bit against the committed cv2.ORB golden vectors and the oracle restatement; live against cv2.ORB at
1280x960 when cv2 is importable; and the descriptors handed to K1 on the device (no host round trip)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import hamming_knn as hk
from oracle import orb as oo
from tod_b200 import DescriptorMatcher, FeatureDescriptor, capi, depth_to_3d, synth

pytestmark = pytest.mark.gpu


def keypoint_array(x, y, octave, angle=None):
    kp = np.zeros(len(x), capi.KEYPOINT_DTYPE)
    kp["x"], kp["y"], kp["octave"] = x, y, octave
    kp["angle"] = -1.0 if angle is None else angle
    return kp


@pytest.mark.parametrize("name", ["orb_640x480.npz", "orb_333x517_ragged.npz"])
def test_cv2_golden_vectors(name):
    g = np.load(os.path.join(GOLDEN, name))
    img = synth.make_textured_image(int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    fd = FeatureDescriptor(n_levels=3, scale_factor=1.2)
    kp, desc = fd.describe(img, keypoint_array(g["x"], g["y"], g["octave"]))
    assert (kp["angle"] == g["angle"]).all()              # fastAtan2 restated operation by operation: exact floats
    assert (desc == g["descriptors"]).all()
    kp2, desc2 = fd.describe(img, keypoint_array(g["x"], g["y"], g["octave"], g["angle"]), compute_angles=False)
    assert (desc2 == desc).all() and (kp2["angle"] == g["angle"]).all()
    fd.close()


def test_against_oracle_on_random_keypoints_five_levels():
    """Keypoints cv2 would not pick (arbitrary positions, 5 levels, scale 1.3): the oracle restatement is the reference."""
    rng = np.random.default_rng(9)
    img = synth.make_textured_image(400, 520, seed=31)
    n_levels, sf = 5, 1.3
    sizes = oo.level_sizes(400, 520, n_levels, sf)
    sc = oo.level_scales(n_levels, sf)
    xs, ys, oc = [], [], []
    for l, (h, w) in enumerate(sizes):
        for _ in range(120):
            cx, cy = int(rng.integers(23, w - 23)), int(rng.integers(23, h - 23))
            xs.append(np.float32(cx) * sc[l])
            ys.append(np.float32(cy) * sc[l])
            oc.append(l)
    xs, ys, oc = np.array(xs, np.float32), np.array(ys, np.float32), np.array(oc, np.int32)
    ok = np.array([oo.level_center(x, y, sc[o]) for x, y, o in zip(xs, ys, oc)])
    keep = np.array([23 <= c[0] < sizes[o][1] - 23 and 23 <= c[1] < sizes[o][0] - 23 for c, o in zip(ok, oc)])
    xs, ys, oc = xs[keep], ys[keep], oc[keep]
    fd = FeatureDescriptor(n_levels=n_levels, scale_factor=sf)
    kp, desc = fd.describe(img, keypoint_array(xs, ys, oc))
    fd.close()
    ang, exp = oo.describe(img, xs, ys, oc, n_levels=n_levels, scale_factor=sf)
    assert (kp["angle"] == ang).all()
    assert (desc == exp).all()


def test_live_cv2_orb_1280x960():
    cv2 = pytest.importorskip("cv2")
    img = synth.make_textured_image(960, 1280, seed=41, n_shapes=1500)
    kps, des = cv2.ORB_create(5000, 1.2, 3).detectAndCompute(img, None)
    assert len(kps) >= 3000
    x = np.array([k.pt[0] for k in kps], np.float32)
    y = np.array([k.pt[1] for k in kps], np.float32)
    oc = np.array([k.octave for k in kps], np.int32)
    fd = FeatureDescriptor()
    kp, desc = fd.describe(img, keypoint_array(x, y, oc))
    fd.close()
    assert (kp["angle"] == np.array([k.angle for k in kps], np.float32)).all()
    assert (desc == des).all()


def test_pyramid_smoothing_and_fast_scores_stage_by_stage():
    """Every intermediate image of the feature stage, read back from the device, equals the oracle restatement byte for
    byte: resized levels (INTER_LINEAR_EXACT), smoothed levels, FAST corner scores."""
    img = synth.make_textured_image(333, 517, seed=12)
    fd = FeatureDescriptor(n_features=1000)
    fd.process(img)
    levels = oo.pyramid(img, 3)
    for l, lev in enumerate(levels):
        assert (fd.read_level(l, 0) == lev).all()
        assert (fd.read_level(l, 1) == oo.smooth(lev)).all()
        assert (fd.read_level(l, 2) == oo.fast_scores(lev)).all()
    fd.close()


def as_records(kp):
    return {(int(k["octave"]), float(k["x"]), float(k["y"])): (float(k["angle"]), float(k["response"]), float(k["size"]))
            for k in kp}


@pytest.mark.parametrize("name", ["orb_640x480.npz", "orb_333x517_ragged.npz"])
def test_detect_and_compute_equals_cv2_golden(name):
    """The whole cell on the GPU: same keypoint SET as cv2.ORB (positions, octave, angle, Harris response, size — exact
    floats) and, keypoint by keypoint, the same 32 descriptor bytes."""
    g = np.load(os.path.join(GOLDEN, name))
    img = synth.make_textured_image(int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    fd = FeatureDescriptor(n_features=5000)
    kp, desc = fd.process(img)
    fd.close()
    assert kp.shape[0] == int(g["n_detected"]) == desc.shape[0]
    ref = {(int(o), float(x), float(y)): (float(a), float(r), float(s), i) for i, (x, y, o, a, r, s) in
           enumerate(zip(g["x"], g["y"], g["octave"], g["angle"], g["response"], g["size"]))}
    got = as_records(kp)
    assert set(got) == set(ref)
    for key, (a, r, s) in got.items():
        assert (a, r, s) == ref[key][:3], key
    order = [ref[(int(k["octave"]), float(k["x"]), float(k["y"]))][3] for k in kp]
    assert (desc == g["descriptors"][order]).all()
    assert (kp["class_id"] == -1).all()
    key = kp["octave"].astype(np.int64)                          # ordered by octave
    assert (np.diff(key) >= 0).all()


def test_detect_and_compute_live_cv2_1280x960_n_features_cut():
    """At 1280 x 960 the frame holds more corners than n_features: the two retainBest cuts (FAST score, then Harris) must
    pick cv2's keypoints, ties included."""
    cv2 = pytest.importorskip("cv2")
    img = synth.make_textured_image(960, 1280, seed=41, n_shapes=1500)
    for nf in (5000, 1200):
        kps, des = cv2.ORB_create(nf, 1.2, 3).detectAndCompute(img, None)
        fd = FeatureDescriptor(n_features=nf)
        kp, desc = fd.process(img)
        fd.close()
        ref = {(k.octave, float(k.pt[0]), float(k.pt[1])): (float(k.angle), float(k.response), float(k.size), i)
               for i, k in enumerate(kps)}
        got = as_records(kp)
        assert set(got) == set(ref) and len(kps) == kp.shape[0]
        for key_, v in got.items():
            assert v == ref[key_][:3]
        order = [ref[(int(k["octave"]), float(k["x"]), float(k["y"]))][3] for k in kp]
        assert (desc == des[order]).all()


def test_border_keypoints_are_refused():
    img = synth.make_textured_image(200, 200, seed=1)
    fd = FeatureDescriptor()
    with pytest.raises(capi.TodError) as e:
        fd.describe(img, keypoint_array([10.0], [100.0], [0]))
    assert e.value.code == capi.TOD_ERR_INVALID
    with pytest.raises(capi.TodError):
        fd.describe(img, keypoint_array([100.0], [100.0], [3]))
    kp, desc = fd.describe(img, keypoint_array([], [], []))
    assert desc.shape == (0, 32)
    fd.close()


def test_descriptors_feed_k1_without_leaving_the_device():
    """Feature stage -> DescriptorMatcher on device buffers: the DB holds descriptors of one view of a textured scene,
    the query frame is the same scene with speckle; K1 reads the query descriptors straight from the ORB stage's HBM
    buffer and must answer exactly like the host-buffer call fed with the downloaded descriptors."""
    import torch
    g = np.load(os.path.join(GOLDEN, "orb_640x480.npz"))
    img = synth.make_textured_image(480, 640, seed=int(g["seed"]))
    fd = FeatureDescriptor()
    kp_db, db_desc = fd.describe(img, keypoint_array(g["x"], g["y"], g["octave"]))
    rng = np.random.default_rng(3)
    noisy = np.clip(img.astype(np.int64) + rng.integers(-4, 5, img.shape), 0, 255).astype(np.uint8)
    kp_q, q_desc = fd.describe(noisy, keypoint_array(g["x"], g["y"], g["octave"]))
    dptr = fd.last_device_descriptors
    n = q_desc.shape[0]
    pts = rng.random((db_desc.shape[0], 3)).astype(np.float32)
    m = DescriptorMatcher(k=2, radius=0)
    m.add_object("scene", db_desc, pts)
    m.train()
    host = m.process(q_desc)
    dev = torch.device("cuda", 0)
    md = torch.empty((n, 2, 4), dtype=torch.int32, device=dev)
    cd = torch.empty((n,), dtype=torch.int32, device=dev)
    pd = torch.empty((n, 2, 3), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    m.process_device(dptr, n, md.data_ptr(), cd.data_ptr(), pd.data_ptr())
    torch.cuda.synchronize()
    got = md.cpu().numpy().view(capi.MATCH_DTYPE).reshape(n, 2)
    assert (got == host["matches"]).all() and (cd.cpu().numpy() == host["counts"]).all()
    em, ec = hk.knn_c(q_desc, [db_desc], 2, 0)
    assert (host["matches"]["trainIdx"] == em["trainIdx"]).all() and (host["matches"]["distance"] == em["distance"]).all()
    assert (host["matches"]["trainIdx"][:, 0] == np.arange(n)).mean() > 0.9      # each keypoint finds itself
    m.close()
    fd.close()


def test_depth_to_3d_matches_oracle_bit_for_bit():
    zf, mm = synth.make_depth_image(480, 640, seed=3)
    K = np.array([[525.0, 0, 319.5], [0, 525.0, 239.5], [0, 0, 1]])
    for depth in (zf, mm):
        got = depth_to_3d(depth, K)
        exp = oo.depth_to_3d(depth, K)
        assert got.shape == exp.shape
        assert (np.isnan(got) == np.isnan(exp)).all()
        ok = ~np.isnan(exp)
        assert (got[ok] == exp[ok]).all()
