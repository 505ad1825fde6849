import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `-m gpu`)")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    sizes = g["sizes"]
    off = np.concatenate([[0], np.cumsum(sizes)])
    objs = [np.ascontiguousarray(g["db"][off[i]:off[i + 1]]) for i in range(len(sizes))]
    return g, objs


def golden_names():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "knn_*.npz")))


def assert_matches_equal(m, c, exp_trn, exp_img, exp_dist, exp_cnt):
    assert (np.asarray(c) == exp_cnt).all()
    assert (m["trainIdx"] == exp_trn).all()
    assert (m["imgIdx"] == exp_img).all()
    assert (m["distance"] == exp_dist).all()
    k = m.shape[1]
    mask = np.arange(k)[None, :] < np.asarray(c)[:, None]
    assert (m["queryIdx"][mask] == np.broadcast_to(np.arange(m.shape[0])[:, None], m.shape)[mask]).all()


@pytest.fixture(scope="session")
def lib():
    from tod_b200 import capi
    return capi.load()
