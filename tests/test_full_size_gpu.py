"""GPU: the geometry half at BASELINE.json's FULL sizes against the reference's own compiled sources (oracle/_ref).

  C4  64 frames x 4096 keypoints, 1280x960 clouds, 1M-descriptor DB, k = 5, radius 35 (conf/detection.ork), 2500
      iterations: the whole batch goes through DescriptorMatcher.process + the batched GuessGenerator; 4 of the 64
      frames are replayed by the compiled reference (GuessGenerator.cpp:170-235 order) — identical inlier keypoint
      sets, poses within 1e-4 — and every planted object of the batch must be recovered.
  C5  100 objects x 2000 correspondences, 90 % outlier matches, 4096 iterations per object: the first 5 objects are
      replayed by the compiled reference (objects are independent: the sampler stream depends on (seed, object, round)).
"""
import numpy as np
import pytest

from oracle import ref
from tod_b200 import DescriptorMatcher, GuessGenerator, synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtod_ref.so not built")]

POSE_TOL = 1e-4


def compare(got_poses, got_inliers, exp):
    assert len(got_poses) == len(exp), (len(got_poses), len(exp))
    for p, inl, (eo, eR, eT, einl) in zip(got_poses, got_inliers, exp):
        assert int(p["object_index"]) == eo
        assert list(inl) == list(einl)
        assert np.abs(p["R"].reshape(3, 3) - eR).max() < POSE_TOL
        assert np.abs(p["T"] - eT).max() < POSE_TOL


def test_c4_full_size_batch_against_compiled_reference():
    n_frames, n_kp, H, W, K, RADIUS, ITERS = 64, 4096, 960, 1280, 5, 35, 2500
    descs, points = synth.make_db(100, 10000, seed=synth.BASE_SEED + 2)
    rng = np.random.default_rng(4)
    frames = []
    for f in range(n_frames):
        vis = sorted(int(x) for x in rng.choice(100, 4, replace=False))
        frames.append(synth.make_frame(descs, points, vis, n_kp, height=H, width=W, seed=synth.BASE_SEED + 400 + f))
    m = DescriptorMatcher(k=K, radius=RADIUS)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("object_%03d" % i, d, p)
    m.train()
    out = m.process(np.ascontiguousarray(np.concatenate([f["descriptors"] for f in frames])))
    spans = m.spans_by_index
    m.close()
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=ITERS, sensor_error=0.01, seed=9)
    res = gg.process_batch([f["keypoints_xy"] for f in frames], np.stack([f["cloud"] for f in frames]),
                           out["matches"], out["counts"], out["matches_3d"], spans, max_poses=64 * n_frames)
    gg.close()
    want = got = 0
    for f, r in zip(frames, res):
        for o, (R, T) in f["poses"].items():
            want += 1
            got += any(int(p["object_index"]) == o and np.abs(p["R"].reshape(3, 3) - R).max() < 0.02 and
                       np.abs(p["T"] - T).max() < 0.01 for p in r["pose_results"])
    assert want == 4 * n_frames and got == want
    for f in (0, 21, 42, 63):                                   # 4 of the 64 frames through the reference's own code
        lo = f * n_kp
        mt = out["matches"][lo:lo + n_kp].copy()
        mt["queryIdx"] = np.where(mt["queryIdx"] >= 0, mt["queryIdx"] - lo, mt["queryIdx"])
        exp = ref.process(frames[f]["keypoints_xy"], frames[f]["cloud"], mt, out["counts"][lo:lo + n_kp],
                          out["matches_3d"][lo:lo + n_kp], spans, 15, ITERS, 0.01, seed=9)
        assert len(exp) >= 4
        compare(res[f]["pose_results"], res[f]["inliers"], exp)


def test_c5_full_size_objects_against_compiled_reference():
    n_obj, n_per, iters, n_ref = 100, 2000, 4096, 5
    g = synth.make_guess_inputs(n_obj, n_per, 0.1, seed=synth.BASE_SEED + 5, k=1, height=960, width=1280)
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=iters, sensor_error=0.01, seed=11)
    res = gg.process(g["keypoints_xy"], g["cloud"], g["matches"], g["counts"], g["points3d"], g["spans"],
                     max_poses=32 * n_obj)
    st = gg.last_stats()
    gg.close()
    assert len(set(int(p["object_index"]) for p in res["pose_results"])) == n_obj     # every planted object found
    assert st["n_hypotheses"] > n_obj * iters
    # the reference on the first n_ref objects only: drop every match to another object (objects are independent)
    counts = g["counts"].copy()
    counts[g["matches"]["imgIdx"][:, 0] >= n_ref] = 0
    exp = ref.process(g["keypoints_xy"], g["cloud"], g["matches"], counts, g["points3d"], g["spans"], 15, iters, 0.01,
                      seed=11)
    assert len(exp) >= n_ref
    sel = [i for i, p in enumerate(res["pose_results"]) if int(p["object_index"]) < n_ref]
    compare([res["pose_results"][i] for i in sel], [res["inliers"][i] for i in sel], exp)


def test_c1_frame_against_compiled_reference():
    """BASELINE config C1 (the reference's own CPU-runnable case): one 640x480 frame, one object of 5000 descriptors,
    5000 keypoints, the .ork parameters.  One frame-filling cluster of ~3500 correspondences: its filtered graph is
    far above K4's proof limit (kGateProofMax) and K5's 256 vertices, so the gate is decided by the host search on a
    view of the cluster's own bit-rows — same pose and inlier set as the reference's own code."""
    descs, points = synth.make_db(1, 5000, seed=synth.BASE_SEED)
    fr = synth.make_frame(descs, points, [0], 5000, seed=synth.BASE_SEED + 9)
    m = DescriptorMatcher(search_json_params='{"type": "LSH", "key_size": 16, "multi_probe_level": 1, "n_tables": 10, '
                                             '"radius": 35, "ratio": 0.8}')
    m.add_object("object_0", descs[0], points[0])
    m.train()
    out = m.process(fr["descriptors"])
    spans = m.spans_by_index
    hist = m.k1_ms_history(1)
    assert len(hist) == 1 and hist[0] > 0 and abs(hist[0] - m.last_k1_ms) < 1e-6     # the K1 event ring
    m.close()
    assert int(out["counts"].sum()) > 3000
    gg = GuessGenerator(min_inliers=8, n_ransac_iterations=2500, sensor_error=0.01, seed=5)
    got = gg.process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"], spans)
    st = gg.last_stats()
    gg.close()
    exp = ref.process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"], spans, 8, 2500,
                      0.01, seed=5)
    assert len(exp) >= 1
    compare(got["pose_results"], got["inliers"], exp)
    assert st["gate_shape"]["k4_undecided_used"] >= 1 and st["gate_shape"]["k5_passes_used"] == 0
