"""GPU: the training path through the C-ABI (tod_trainer_*: masked ORB, mask erosion, depth rescale, keypoint
validation, back-projection, camera -> object frame, stacking) against the oracle restatement of Trainer.cpp /
training.cpp — descriptors bit for bit, points as exact floats, same order — and the trained model driven through the
detection path."""
import numpy as np
import pytest

from oracle import orb as oo
from oracle import training as ot
from tod_b200 import DescriptorMatcher, FeatureDescriptor, Trainer, capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bgr,depth_scale,u16", [(True, 1, False), (False, 2, False), (True, 1, True)])
def test_trainer_equals_oracle(bgr, depth_scale, u16):
    views = synth.make_training_views(3, seed=21 + depth_scale, depth_scale=depth_scale, bgr=bgr)
    t = Trainer()
    ds, ps = [], []
    for v in views:
        depth = v["depth"]
        if u16:
            depth = np.where(np.isnan(depth), 0, np.rint(depth * 1000.0)).astype(np.uint16)
        n = t.add_observation(v["image"], v["mask"], depth, v["K"], v["R"], v["T"])
        d, p, _ = ot.train_observation(v["image"], v["mask"], depth, v["K"], v["R"], v["T"])
        assert n == d.shape[0] > 100
        ds.append(d)
        ps.append(p)
    D, P = t.model()
    eD, eP = ot.merge_points(ds, ps)
    assert D.shape == eD.shape and (D == eD).all()
    assert (P == eP).all()
    t.clear()
    assert t.model()[0].shape[0] == 0
    t.close()


def test_masked_detection_equals_live_cv2():
    cv2 = pytest.importorskip("cv2")
    v = synth.make_training_views(1, seed=5)[0]
    kps, des = cv2.ORB_create().detectAndCompute(v["image"], v["mask"])
    fd = FeatureDescriptor(n_features=500, n_levels=8, scale_factor=1.2)
    kp, desc = fd.process_masked(v["image"], v["mask"])
    fd.close()
    ref = {(k.octave, float(k.pt[0]), float(k.pt[1])): (float(k.angle), float(k.response), i) for i, k in enumerate(kps)}
    got = {(int(k["octave"]), float(k["x"]), float(k["y"])): (float(k["angle"]), float(k["response"])) for k in kp}
    assert set(got) == set(ref)
    assert all(got[k] == ref[k][:2] for k in got)
    order = [ref[(int(k["octave"]), float(k["x"]), float(k["y"]))][2] for k in kp]
    assert (desc == des[order]).all()


def test_trained_model_is_found_by_the_detection_path():
    """Train on three views, then match a fourth frame's descriptors (the same textures, speckled) against the model:
    the feature stage, the trainer and K1 chained."""
    views = synth.make_training_views(3, seed=33)
    t = Trainer()
    for v in views:
        t.add_observation(v["image"], v["mask"], v["depth"], v["K"], v["R"], v["T"])
    D, P = t.model()
    t.close()
    m = DescriptorMatcher(k=2, radius=35)
    m.add_object("trained", D, P)
    m.train()
    rng = np.random.default_rng(1)
    noisy = np.clip(views[0]["image"].astype(np.int64) + rng.integers(-3, 4, views[0]["image"].shape), 0, 255)
    fd = FeatureDescriptor(n_features=500, n_levels=8)
    kp, desc = fd.process_masked(noisy.astype(np.uint8), views[0]["mask"])
    fd.close()
    out = m.process(desc)
    m.close()
    assert (out["counts"] > 0).mean() > 0.5                        # most keypoints of the re-observed view are matched
