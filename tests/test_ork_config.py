"""The reference's only config tests are `object_recognition_core_config_test(conf/*.ork)` (test/CMakeLists.txt:2-4):
each `.ork` must parse and the pipeline must be instantiable.  Same check for this library: a detection `.ork` with the
reference's parameter names and values (conf/detection.ork:21-46, conf/detection.ros.ork) is accepted, on CPU as far as
parsing goes and on the GPU by instantiating both cells and pushing a frame through them."""
import pytest
import yaml

from tod_b200 import capi, ork_parameters

DETECTION_ORK = """
source1:
  type: 'OpenNI'
  module: 'object_recognition_core.io.source'
pipeline1:
  type: 'TodDetector'
  module: 'object_recognition_tod'
  inputs: [source1]
  parameters:
    object_ids: "all"
    feature: {type: ORB, module: ecto_opencv.features2d, n_features: 5000, n_levels: 3, scale_factor: 1.2}
    descriptor: {type: ORB, module: ecto_opencv.features2d}
    search:
      type: LSH
      module: ecto_opencv.features2d
      key_size: 16
      multi_probe_level: 1
      n_tables: 10
      radius: %d
      ratio: 0.8
    n_ransac_iterations: 2500
    min_inliers: 8
    sensor_error: 0.01
    db: {type: CouchDB, root: 'http://localhost:5984', collection: object_recognition}
"""


@pytest.mark.parametrize("radius", [35, 55])      # detection.ork / detection.ros.ork
def test_ork_parameters_parse(radius):
    params = yaml.safe_load(DETECTION_ORK % radius)["pipeline1"]["parameters"]
    p, guess = ork_parameters(params)
    assert (p.k, p.radius, p.search_type) == (5, radius, capi.TOD_SEARCH_LSH)
    assert guess == {"n_ransac_iterations": 2500, "min_inliers": 8, "sensor_error": 0.01}


def test_unknown_search_type_is_an_error_not_a_terminate():
    params = yaml.safe_load(DETECTION_ORK % 35)["pipeline1"]["parameters"]
    params["search"]["type"] = "KDTREE"            # the reference prints and calls std::terminate (:182-186)
    with pytest.raises(capi.TodError) as e:
        ork_parameters(params)
    assert e.value.code == capi.TOD_ERR_INVALID


@pytest.mark.gpu
def test_pipeline_from_ork_detects_planted_objects():
    import numpy as np
    from tod_b200 import detector_from_ork, synth
    params = yaml.safe_load(DETECTION_ORK % 35)["pipeline1"]["parameters"]
    m, g = detector_from_ork(params, seed=3)
    descs, points = synth.make_db(5, 600, seed=70)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("model_%d" % i, d, p)
    m.train()
    fr = synth.make_frame(descs, points, [1, 3], 500, seed=71)
    out = m.process(fr["descriptors"])
    res = g.process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"],
                    m.spans_by_index)
    found = sorted(int(p["object_index"]) for p in res["pose_results"])
    assert found == [1, 3]
    for p in res["pose_results"]:
        R, T = fr["poses"][int(p["object_index"])]
        assert np.abs(p["R"].reshape(3, 3) - R).max() < 0.02 and np.abs(p["T"] - T).max() < 0.01
    m.close()
    g.close()
