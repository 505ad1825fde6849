"""CPU: the K1 oracle (numpy restatement and its C port) against the golden vectors produced by the real
cv2.BFMatcher(NORM_HAMMING) (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import assert_matches_equal, golden_names, load_golden
from oracle import hamming_knn as hk


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_matches_cv2_golden(name, impl):
    g, objs = load_golden(name)
    fn = hk.knn_numpy if impl == "numpy" else hk.knn_c
    m, c = fn(g["query"], objs, int(g["k"]), int(g["radius"]))
    assert_matches_equal(m, c, g["trainIdx"], g["imgIdx"], g["distance"], g["counts"])


def test_numpy_and_c_agree_on_ties():
    rng = np.random.default_rng(5)
    objs = [rng.integers(0, 2, (n, 32), dtype=np.uint8) for n in (50, 70, 9)]   # 32 significant bits: tie-heavy
    q = rng.integers(0, 2, (40, 32), dtype=np.uint8)
    for k in (1, 2, 5, 8):
        a, ca = hk.knn_numpy(q, objs, k)
        b, cb = hk.knn_c(q, objs, k)
        assert (ca == cb).all()
        for f in ("trainIdx", "imgIdx", "distance"):
            assert (a[f] == b[f]).all()


def test_span_matches_reference_formula():
    rng = np.random.default_rng(7)
    p = rng.normal(size=(100, 3)).astype(np.float32)
    ext = p.max(0) - p.min(0)
    assert abs(float(hk.object_span(p)) - float(np.sqrt((ext.astype(np.float64) ** 2).sum()))) < 1e-5
