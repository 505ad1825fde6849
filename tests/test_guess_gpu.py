"""GPU: GuessGenerator.process through the C-ABI (K2 + K3 + host replay/gate/refinement) against the reference's own
code (oracle/_ref, compiled from /root/reference/src/common) and the oracle restatement, on a shared seeded sampler
stream: identical inlier keypoint sets, poses within 1e-4 (north_star tolerance)."""
import numpy as np
import pytest

from oracle import geometry as og
from oracle import hamming_knn as hk
from oracle import ref
from tod_b200 import DescriptorMatcher, GuessGenerator, synth

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-4


def compare(got, exp):
    assert len(got["pose_results"]) == len(exp), (len(got["pose_results"]), len(exp))
    for p, inl, (eo, eR, eT, einl) in zip(got["pose_results"], got["inliers"], exp):
        assert int(p["object_index"]) == eo
        assert list(inl) == list(einl)
        assert np.abs(p["R"].reshape(3, 3) - eR).max() < POSE_TOL
        assert np.abs(p["T"] - eT).max() < POSE_TOL


def scene(seed, n_objects=4, rows=400, visible=(0, 2), n_kp=300, k=5, radius=35, **kw):
    descs, points = synth.make_db(n_objects, rows, seed=seed)
    fr = synth.make_frame(descs, points, list(visible), n_kp, seed=seed + 1, **kw)
    m = DescriptorMatcher(k=k, radius=radius)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("o%d" % i, d, p)
    m.train()
    out = m.process(fr["descriptors"])
    spans = m.spans_by_index
    m.close()
    return fr, out, spans


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtod_ref.so not built")
@pytest.mark.parametrize("seed,visible,kw", [(5, (0, 2), {}), (6, (1,), {"duplicate_outliers": 0.5}),
                                             (7, (0, 1, 2, 3), {"clutter_fraction": 0.1})])
def test_detection_pipeline_matches_reference(seed, visible, kw):
    """DescriptorMatcher.process -> GuessGenerator.process on a synthetic frame, vs the reference's geometry code."""
    fr, out, spans = scene(seed, visible=visible, **kw)
    gg = GuessGenerator(min_inliers=8, n_ransac_iterations=500, sensor_error=0.01, seed=77)
    got = gg.process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"], spans)
    exp = ref.process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"], spans, 8, 500,
                      0.01, seed=77)
    compare(got, exp)
    assert sorted(int(p["object_index"]) for p in got["pose_results"]) == sorted(visible)
    for p in got["pose_results"]:                       # planted ground truth is recovered
        gR, gT = fr["poses"][int(p["object_index"])]
        assert np.abs(p["R"].reshape(3, 3) - gR).max() < 0.02 and np.abs(p["T"] - gT).max() < 0.01
    st = gg.last_stats()
    assert st["n_hypotheses"] > 0 and st["n_rounds"] >= 1


def test_oracle_restatement_agrees_too():
    fr, out, spans = scene(5)
    gg = GuessGenerator(min_inliers=8, n_ransac_iterations=300, sensor_error=0.01, seed=3)
    got = gg.process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"], spans)
    exp = og.guess_process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"], spans, 8,
                           300, 0.01, seed=3)
    compare(got, exp)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtod_ref.so not built")
@pytest.mark.parametrize("n_obj,n_per,frac,iters", [(6, 150, 0.5, 300), (10, 300, 0.1, 1000), (3, 40, 0.9, 100),
                                                    (3, 1000, 0.1, 600)])
def test_outlier_heavy_clusters_match_reference(n_obj, n_per, frac, iters):
    """BASELINE config C5 shape (scaled): matches injected at the GuessGenerator boundary, 50-90% outliers."""
    gi = synth.make_guess_inputs(n_obj, n_per, frac, seed=1000 + n_per)
    gg = GuessGenerator(min_inliers=8, n_ransac_iterations=iters, sensor_error=0.01, seed=11)
    got = gg.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], gi["counts"], gi["points3d"], gi["spans"])
    exp = ref.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], gi["counts"], gi["points3d"], gi["spans"], 8,
                      iters, 0.01, seed=11)
    compare(got, exp)
    assert len(exp) >= 1


def test_result_does_not_depend_on_host_threads():
    """The per-object host work runs on a thread pool; every object owns its sampler stream, so 1 and 8 threads must
    give identical poses and inlier sets."""
    gi = synth.make_guess_inputs(12, 200, 0.3, seed=77)
    res = []
    for threads in (1, 8):
        gg = GuessGenerator(min_inliers=8, n_ransac_iterations=400, sensor_error=0.01, seed=5, host_threads=threads)
        res.append(gg.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], gi["counts"], gi["points3d"],
                              gi["spans"]))
        st = gg.last_stats()
        assert st["gate_calls"] >= st["gate_proved_empty"] >= 0
    a, b = res
    assert len(a["pose_results"]) == len(b["pose_results"]) >= 1
    assert (a["pose_results"] == b["pose_results"]).all()
    for x, y in zip(a["inliers"], b["inliers"]):
        assert list(x) == list(y)


def test_edge_cases():
    gi = synth.make_guess_inputs(2, 30, 0.9, seed=4)
    gg = GuessGenerator(min_inliers=8, n_ransac_iterations=100, seed=1)
    # no matches at all -> no poses
    got = gg.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], np.zeros_like(gi["counts"]), gi["points3d"],
                     gi["spans"])
    assert len(got["pose_results"]) == 0
    # NaN depth under every keypoint -> every correspondence skipped (adjacency_ransac.cpp:189)
    cl = np.full_like(gi["cloud"], np.nan)
    got = gg.process(gi["keypoints_xy"], cl, gi["matches"], gi["counts"], gi["points3d"], gi["spans"])
    assert len(got["pose_results"]) == 0
    # fewer than 3 correspondences per object
    c = np.zeros_like(gi["counts"])
    c[np.nonzero(gi["matches"]["imgIdx"][:, 0] == 0)[0][:2]] = 1
    got = gg.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], c, gi["points3d"], gi["spans"])
    assert len(got["pose_results"]) == 0
    # min_inliers above what exists
    gg2 = GuessGenerator(min_inliers=1000, n_ransac_iterations=100, seed=1)
    got = gg2.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], gi["counts"], gi["points3d"], gi["spans"])
    assert len(got["pose_results"]) == 0


def test_batched_frames_equal_single_frame_calls():
    """tod_guess_process_batch (BASELINE config C4 shape, scaled down): all (frame, object) clusters share the K2 launch
    and every K3 round, yet the result must equal frame-by-frame calls exactly."""
    descs, points = synth.make_db(6, 500, seed=90)
    frames = [synth.make_frame(descs, points, vis, 350, seed=91 + i, duplicate_outliers=0.2)
              for i, vis in enumerate([(0, 3), (1,), (2, 4, 5), (), (0, 1, 2)])]
    m = DescriptorMatcher(k=5, radius=35)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("o%d" % i, d, p)
    m.train()
    out = m.process(np.concatenate([f["descriptors"] for f in frames]))
    spans = m.spans_by_index
    m.close()
    gg = GuessGenerator(min_inliers=8, n_ransac_iterations=400, sensor_error=0.01, seed=21)
    batch = gg.process_batch([f["keypoints_xy"] for f in frames], np.stack([f["cloud"] for f in frames]),
                             out["matches"], out["counts"], out["matches_3d"], spans)
    assert len(batch) == len(frames)
    o = 0
    total = 0
    for f, got in zip(frames, batch):
        n = f["keypoints_xy"].shape[0]
        mt = out["matches"][o:o + n].copy()
        mt["queryIdx"] = np.where(mt["queryIdx"] >= 0, mt["queryIdx"] - o, mt["queryIdx"])
        one = gg.process(f["keypoints_xy"], f["cloud"], mt, out["counts"][o:o + n], out["matches_3d"][o:o + n], spans)
        assert len(one["pose_results"]) == len(got["pose_results"])
        assert (one["pose_results"] == got["pose_results"]).all()
        for a, b in zip(one["inliers"], got["inliers"]):
            assert list(a) == list(b)
        total += len(got["pose_results"])
        o += n
    assert total >= 6 and len(batch[3]["pose_results"]) == 0     # the empty frame yields nothing
