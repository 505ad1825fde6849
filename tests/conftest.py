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


def _gpu_unavailable_reason():
    """None when an sm_100 device is usable, else why not.  Asked of the CUDA runtime through ctypes (no torch import,
    no CUDA context when there is no driver)."""
    import ctypes
    for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            rt = None
    if rt is None:
        return "libcudart not found"
    n = ctypes.c_int(0)
    if rt.cudaGetDeviceCount(ctypes.byref(n)) != 0 or n.value <= 0:
        return "no CUDA device"
    major = ctypes.c_int(0)
    if rt.cudaDeviceGetAttribute(ctypes.byref(major), 75, 0) != 0 or major.value != 10:   # 75 = ComputeCapabilityMajor
        return "device 0 is not sm_100 (compute capability major %d)" % major.value
    return None


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a machine without a B200, so a plain `pytest` run is green on
    CPU-only boxes; on the GPU box they run."""
    if not any("gpu" in item.keywords for item in items):
        return
    why = _gpu_unavailable_reason()
    if why is None:
        return
    skip = pytest.mark.skip(reason="needs a B200: " + why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
