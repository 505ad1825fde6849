"""GPU: K5 (k5_search_kernel, the clique gate's bounded search stepped exactly on the device) as a stage of its own —
same kernel, job queue layout and step cap as inside tod_guess_process — against the host compilation of the same
source (tod_clique_gate_small), the product's CliqueFinder and the reference's own compiled maximum_clique.cpp."""
import ctypes

import numpy as np
import pytest

from oracle import ref
from tod_b200 import capi

pytestmark = pytest.mark.gpu


def _graphs(seed, count, lo, hi):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(count):
        n = int(rng.integers(lo, hi + 1))
        if t % 2:
            a = np.triu(rng.random((n, n)) < rng.choice([0.4, 0.6, 0.8, 0.85, 0.9, 0.97]), 1)
            if n > 12 and rng.random() < 0.4:                   # low-degree fringe, like the gate's filtered graphs
                low = rng.choice(n, n // 4, replace=False)
                a[low, :] &= rng.random((len(low), n)) < 0.3
        else:
            a = np.triu(rng.random((n, n)) < rng.uniform(0.2, 0.62), 1)     # clique number around 7-8: the search steps
        out.append((n, np.argwhere(a).astype(np.int32)))
    return out


def _device(graphs):
    lib = capi.load()
    nv = np.array([g[0] for g in graphs], np.int32)
    off = np.zeros(len(graphs) + 1, np.int32)
    off[1:] = np.cumsum([g[1].shape[0] for g in graphs])
    edges = np.ascontiguousarray(np.concatenate([g[1].reshape(-1, 2) for g in graphs]) if graphs else np.zeros((0, 2)),
                                 np.int32)
    res = np.full(len(graphs), -9, np.int32)
    capi.check(lib.tod_gate_search_device(0, len(graphs), capi._ptr(nv), capi._ptr(off), capi._ptr(edges),
                                          capi._ptr(res)))
    return res


def _host(n, edges, cap=512):
    lib = capi.load()
    e = np.ascontiguousarray(edges, np.int32).reshape(-1, 2)
    return int(lib.tod_clique_gate_small(n, capi._ptr(e), e.shape[0], cap, None))


@pytest.mark.parametrize("lo,hi,count", [(1, 64, 1500), (65, 128, 700), (129, 256, 300), (1, 256, 900)])
def test_k5_equals_host_compilation(lo, hi, count):
    """Every row width (1, 2 and 4 words) and mixed batches: the device verdict of each graph is the host's."""
    graphs = _graphs(100 + lo, count, lo, hi)
    got = _device(graphs)
    exp = np.array([_host(n, e) for n, e in graphs], np.int32)
    assert (got == exp).all(), np.flatnonzero(got != exp)[:10]
    assert set(np.unique(got)) <= {0, 1}            # nothing near the 512-step cap on these graphs
    assert (got == 1).any() and (got == 0).any()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtod_ref.so not built")
def test_k5_equals_compiled_reference():
    graphs = _graphs(7, 240, 8, 256)
    got = _device(graphs)
    for (n, e), r in zip(graphs, got):
        exp = ref.find_clique(n, [tuple(x) for x in e], 7, sorted_insert=True)
        assert int(r) == int(len(exp) > 7)


def test_k5_edge_cases():
    k8 = np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], np.int32)
    k7 = np.array([(i, j) for i in range(7) for j in range(i + 1, 7)], np.int32)
    got = _device([(8, k8), (7, k7), (1, np.zeros((0, 2), np.int32)), (256, np.zeros((0, 2), np.int32)), (200, k8)])
    assert list(got) == [1, 0, 0, 0, 1]
    lib = capi.load()
    one = np.array([300], np.int32)
    off = np.zeros(2, np.int32)
    res = np.zeros(1, np.int32)
    assert lib.tod_gate_search_device(0, 1, capi._ptr(one), capi._ptr(off), capi._ptr(np.zeros((0, 2), np.int32)),
                                      capi._ptr(res)) != 0       # more than 256 vertices: refused
