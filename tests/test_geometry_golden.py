"""Committed golden vectors produced by the REFERENCE'S OWN geometry code (tests/golden/make_geometry_golden.py ran
src/common/*.cpp of wg-perception/tod compiled unmodified): on CPU they pin the oracle restatement, on the GPU the
product path (K2 bit-matrices; GuessGenerator.process poses and inlier sets) — with neither /root/reference nor
oracle/_ref present at run time."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import geometry as og

POSE_TOL = 1e-4   # north_star: poses agree within 1e-4 rotation and 1e-4 m translation


def adjacency_goldens():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "geom_adjacency_*.npz")))


def guess_goldens():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "geom_guess_*.npz")))


def load_adjacency(name):
    g = np.load(os.path.join(GOLDEN, name))
    n = int(g["n"])
    P = np.unpackbits(g["physical"], axis=1)[:, :n].astype(bool)
    S = np.unpackbits(g["sample"], axis=1)[:, :n].astype(bool)
    return g, n, P, S


def load_guess(name):
    g = np.load(os.path.join(GOLDEN, name))
    H, W = [int(x) for x in g["cloud_shape"]]
    cloud = np.full((H, W, 3), np.nan, np.float32)
    cloud[g["cloud_y"], g["cloud_x"]] = g["cloud_v"]
    exp, o = [], 0
    for i in range(len(g["pose_object"])):
        k = int(g["pose_n_inliers"][i])
        exp.append((int(g["pose_object"][i]), g["pose_R"][i], g["pose_T"][i], [int(x) for x in g["inliers"][o:o + k]]))
        o += k
    return g, cloud, exp


def check_poses(got_poses, got_inliers, exp):
    """got_poses: [(object_index, R, T)], got_inliers: [sorted keypoint indices]; exp: the reference's."""
    assert len(got_poses) == len(exp)
    for (obj, R, T), inl, (eo, eR, eT, einl) in zip(got_poses, got_inliers, exp):
        assert obj == eo
        assert list(inl) == list(einl)
        assert np.abs(np.asarray(R).reshape(3, 3) - eR).max() < POSE_TOL
        assert np.abs(np.asarray(T) - eT).max() < POSE_TOL


def test_goldens_exist():
    assert len(adjacency_goldens()) >= 3 and len(guess_goldens()) >= 2


@pytest.mark.parametrize("name", adjacency_goldens())
def test_oracle_adjacency_equals_reference_golden(name):
    g, n, P, S = load_adjacency(name)
    oP, oS = og.fill_adjacency_dense(g["query"], g["train"], g["pixels"], float(g["span"]), float(g["sensor_error"]))
    assert (oP == P).all() and (oS == S).all()
    assert (P == P.T).all() and not P.diagonal().any()


@pytest.mark.parametrize("name", guess_goldens())
def test_oracle_guess_equals_reference_golden(name):
    g, cloud, exp = load_guess(name)
    got = og.guess_process(g["keypoints_xy"], cloud, g["matches"], g["counts"], g["points3d"], g["spans"],
                           int(g["min_inliers"]), int(g["n_ransac_iterations"]), float(g["sensor_error"]),
                           seed=int(g["seed"]))
    check_poses([(e[0], e[1], e[2]) for e in got], [e[3] for e in got], exp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", adjacency_goldens())
def test_k2_equals_reference_golden(name):
    from tod_b200 import fill_adjacency
    g, n, P, S = load_adjacency(name)
    gP, gS, _ = fill_adjacency([0, n], g["query"], g["train"], g["pixels"], [float(g["span"])],
                               float(g["sensor_error"]))
    assert (gP.reshape(n, -1) == og.pack_bits(P)).all()
    assert (gS.reshape(n, -1) == og.pack_bits(S)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", guess_goldens())
def test_guess_generator_equals_reference_golden(name):
    from tod_b200 import GuessGenerator
    g, cloud, exp = load_guess(name)
    gg = GuessGenerator(min_inliers=int(g["min_inliers"]), n_ransac_iterations=int(g["n_ransac_iterations"]),
                        sensor_error=float(g["sensor_error"]), seed=int(g["seed"]))
    got = gg.process(g["keypoints_xy"], cloud, g["matches"], g["counts"], g["points3d"], g["spans"])
    check_poses([(int(p["object_index"]), p["R"], p["T"]) for p in got["pose_results"]], got["inliers"], exp)
