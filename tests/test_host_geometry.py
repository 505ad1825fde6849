"""CPU: the host-side pieces of the product's guess generator (tod_b200/csrc/clique.h, host_geometry.h) through their
host-only C-ABI entry points, against the reference's known-answer tests (test/test_maximum_clique.cpp:7-53) and the
reference's own compiled sources (oracle/_ref)."""
import ctypes

import numpy as np
import pytest

from oracle import geometry as og
from oracle import ref
from tod_b200 import capi, synth
from test_oracle_geometry import KAT1_EDGES

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtod_ref.so not built")


def clique_find(n, edges, minimal=0xFFFFFFFF):
    lib = capi.load()
    e = np.ascontiguousarray(edges, np.int32).reshape(-1, 2)
    out = np.zeros(max(n, 1), np.int32)
    more = ctypes.c_int32(-1)
    k = lib.tod_clique_find(n, capi._ptr(e), e.shape[0], minimal, capi._ptr(out), ctypes.byref(more))
    assert k >= 0
    return [int(x) for x in out[:k]], bool(more.value)


def test_reference_kats_on_the_product_clique_finder():
    got, _ = clique_find(10, KAT1_EDGES)                        # test_maximum_clique.cpp:7-38
    assert len(got) == 4
    full = [(i, j) for i in range(10) for j in range(i + 1, 10) if (i, j) != (0, 1)]
    got, _ = clique_find(10, full)                              # :40-53 (K10 minus one edge)
    assert len(got) == 9
    for clique, edges in ((clique_find(10, KAT1_EDGES)[0], KAT1_EDGES), (got, full)):
        es = set(map(tuple, edges)) | set((b, a) for a, b in edges)
        assert all((a, b) in es for a in clique for b in clique if a != b)      # it IS a clique


@needs_ref
def test_product_clique_finder_equals_compiled_reference():
    """Gate mode (minimal size 7, sac_model_registration_graph.h:259): identical vertex lists, and the decision-only
    variant the gate runs answers exactly `len(reference result) > 7` — sparse, dense and near-complete graphs."""
    rng = np.random.default_rng(1)
    for trial in range(250):
        n = int(rng.integers(5, 60))
        p = rng.choice([0.15, 0.4, 0.7, 0.9, 0.97])
        edges = [(i, j) for i in range(n) for j in range(i + 1, n) if rng.random() < p]
        exp = ref.find_clique(n, edges, 7, sorted_insert=True)
        got, more = clique_find(n, edges, 7)
        assert got == exp, trial
        assert more == (len(exp) > 7), trial


@needs_ref
@pytest.mark.parametrize("n,p,seed", [(150, 0.85, 1), (300, 0.83, 2), (420, 0.9, 3), (260, 0.6, 4), (500, 0.97, 5)])
def test_product_clique_finder_on_large_dense_graphs(n, p, seed):
    """The gate's dense regime (inlier graphs are ~85 % dense, hundreds of vertices): the class-by-class colouring
    must reproduce the reference's sequential ColorSort exactly — identical vertex lists in gate mode."""
    rng = np.random.default_rng(seed)
    a = np.triu(rng.random((n, n)) < p, 1)
    low = rng.choice(n, n // 10, replace=False)                # a few low-degree vertices: the search visits them first
    for v in low:
        keep = rng.random(n) < 0.06
        a[v, :] &= keep
        a[:, v] &= keep
    edges = np.argwhere(a)
    exp = ref.find_clique(n, edges, 7, sorted_insert=True)
    got, more = clique_find(n, edges, 7)
    assert got == exp
    assert more == (len(exp) > 7)


@needs_ref
def test_product_rigid_fit_equals_reference():
    lib = capi.load()
    for n, seed in ((3, 1), (10, 2), (200, 3)):
        q, t, px, _, _ = synth.make_cluster(n, 1.0, seed=seed)
        ar = ref.RefAdjacencyRansac()
        for i in range(n):
            ar.add_points(t[i], q[i], i)
        idx = np.arange(n, dtype=np.uint32)
        eR, eT = ar.kabsch(idx)
        R, T = np.zeros(9, np.float32), np.zeros(3, np.float32)
        capi.check(lib.tod_rigid_fit(capi._ptr(q), capi._ptr(t), capi._ptr(idx), n, capi._ptr(R), capi._ptr(T)))
        assert np.abs(R.reshape(3, 3) - eR).max() < 1e-5 and np.abs(T - eT).max() < 1e-5
        oR, oT = og.kabsch(q, t, list(range(n)))
        assert np.abs(R.reshape(3, 3) - oR).max() < 1e-5 and np.abs(T - oT).max() < 1e-5


@needs_ref
@pytest.mark.parametrize("n,frac,seed", [(60, 0.8, 1), (300, 0.4, 2), (700, 0.15, 3)])
def test_product_sampler_equals_reference_sampler(n, frac, seed):
    """The library's getSamples on bit-rows draws the same triples, in the same (s3, s2, s1) order, as the reference's
    drawIndexSampleHelper on sorted neighbour lists when both consume the same stream (libc rand() in the reference is
    redirected to tod_rng_* by the harness), also after part of the cluster has been invalidated."""
    lib = capi.load()
    q, t, px, _, _ = synth.make_cluster(n, frac, seed=seed)
    ar = ref.RefAdjacencyRansac()
    for i in range(n):
        ar.add_points(t[i], q[i], i)
    ar.fill_adjacency(px, 0.25, 0.01)
    S = og.pack_bits(ar.dense("sample"))
    W = capi.adjacency_row_words(n)
    assert S.shape == (n, W)
    for invalidate in (False, True):
        if invalidate:
            ar.invalidate_query_indices(np.arange(0, n, 7, dtype=np.uint32))
        valid = np.zeros(W * 32, bool)
        valid[ar.valid()] = True
        V = np.packbits(valid, bitorder="little").view("<u4")
        state0 = int(lib.tod_rng_seed(1234, 5, 1 if invalidate else 0))
        exp = ar.get_samples(state0, 400)
        st = ctypes.c_uint64(state0)
        out = np.zeros((400, 3), np.uint32)
        k = lib.tod_sample_triples(n, capi._ptr(np.ascontiguousarray(S)), capi._ptr(V), ctypes.byref(st), 400,
                                   capi._ptr(out))
        assert k == len(exp)
        assert (out[:k] == exp).all()


@needs_ref
@pytest.mark.parametrize("n,frac,seed", [(80, 0.9, 4), (300, 0.5, 5), (600, 0.2, 6), (900, 0.1, 7)])
def test_product_select_within_distance_equals_reference(n, frac, seed):
    """selectWithinDistance with the never-set threshold (quirk Q3): candidate list + clique gate.  The product's host
    code (7-core test, no-8-clique proofs, decision-only bounded search) must return exactly the reference's inlier
    list for every sampled triple — outlier-heavy clusters make most gates fail, inlier-rich ones make them pass."""
    lib = capi.load()
    q, t, px, _, _ = synth.make_cluster(n, frac, seed=seed)
    ar = ref.RefAdjacencyRansac()
    for i in range(n):
        ar.add_points(t[i], q[i], i)
    ar.fill_adjacency(px, 0.25, 0.01)
    P = np.ascontiguousarray(og.pack_bits(ar.dense("physical")))
    S = np.ascontiguousarray(og.pack_bits(ar.dense("sample")))
    W = capi.adjacency_row_words(n)
    valid = np.zeros(W * 32, bool)
    valid[ar.valid()] = True
    V = np.packbits(valid, bitorder="little").view("<u4")
    triples = ar.get_samples(int(lib.tod_rng_seed(99, seed, 0)), 150)
    assert len(triples) > 20
    out = np.zeros(n + 3, np.uint32)
    passed = failed = 0
    for tr in triples:
        exp, _, _ = ar.select(tr)
        tri = np.ascontiguousarray(tr, np.uint32)
        k = lib.tod_select_inliers(n, capi._ptr(P), capi._ptr(S), capi._ptr(V), capi._ptr(tri), capi._ptr(out))
        assert k >= 0
        assert [int(x) for x in out[:k]] == exp
        passed += k > 7
        failed += k == 0
    assert passed + failed > 0


def clique_gate_small(n, edges, cap=100000):
    lib = capi.load()
    e = np.ascontiguousarray(edges, np.int32).reshape(-1, 2)
    steps = ctypes.c_int32(0)
    r = lib.tod_clique_gate_small(n, capi._ptr(e), e.shape[0], cap, ctypes.byref(steps))
    return int(r), int(steps.value)


def test_small_gate_search_equals_the_host_clique_finder():
    """clique_small.h (the search K5 runs on the GPU, compiled here for the host) answers the gate's question exactly
    like CliqueFinder::finds_more_than(7) on graphs of 1..256 vertices: dense graphs with a low-degree fringe (the
    gate's regime) and the densities where the clique number sits around 7-8 and the search has to step.  The entry
    point also runs the 64- and 128-bit instantiations where the graph fits and fails if they disagree."""
    rng = np.random.default_rng(5)
    for trial in range(1500):
        n = int(rng.integers(1, 257)) if trial % 3 else int(rng.integers(1, 129))
        if trial % 2:
            p = rng.choice([0.15, 0.4, 0.6, 0.7, 0.8, 0.85, 0.9, 0.97])
            a = np.triu(rng.random((n, n)) < p, 1)
            if rng.random() < 0.3 and n > 12:
                low = rng.choice(n, n // 4, replace=False)
                a[low, :] &= rng.random((len(low), n)) < 0.3
        else:
            a = np.triu(rng.random((n, n)) < rng.uniform(0.2, 0.62), 1)
        edges = np.argwhere(a)
        r, steps = clique_gate_small(n, edges)
        _, more = clique_find(n, edges, 7)
        assert r == int(more), (trial, n)
        assert 0 <= steps <= 100000


@needs_ref
def test_small_gate_search_equals_compiled_reference():
    rng = np.random.default_rng(6)
    for trial in range(150):
        n = int(rng.integers(8, 257))
        a = np.triu(rng.random((n, n)) < rng.uniform(0.2, 0.9), 1)
        edges = np.argwhere(a)
        exp = ref.find_clique(n, [tuple(x) for x in edges], 7, sorted_insert=True)
        r, _ = clique_gate_small(n, edges)
        assert r == int(len(exp) > 7), (trial, n)


def test_small_gate_search_limits():
    rng = np.random.default_rng(7)
    edges = np.argwhere(np.triu(rng.random((100, 100)) < 0.45, 1))
    assert clique_gate_small(100, edges, cap=5)[0] == -1            # step cap reached: the caller falls back
    assert clique_gate_small(0, np.zeros((0, 2), np.int32))[0] == 0
    assert clique_gate_small(257, np.zeros((0, 2), np.int32))[0] == -2
    full = [(i, j) for i in range(8) for j in range(i + 1, 8)]
    assert clique_gate_small(8, full)[0] == 1                       # K8
    assert clique_gate_small(7, [(i, j) for i in range(7) for j in range(i + 1, 7)])[0] == 0   # K7: exactly 7
