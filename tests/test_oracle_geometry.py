"""CPU: pin the geometry oracle (oracle/geometry.py) against
  (1) the reference's own known-answer tests (test/test_maximum_clique.cpp:7-53), and
  (2) the reference's own src/common sources compiled unmodified into oracle/_ref/libtod_ref.so."""
import numpy as np
import pytest

from oracle import geometry as og
from oracle import hamming_knn as hk
from oracle import ref
from tod_b200 import synth

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtod_ref.so not built")

KAT1_EDGES = [(4, 1), (4, 3), (5, 3), (6, 1), (6, 4), (7, 0), (7, 2), (7, 3), (7, 4), (7, 5), (8, 0), (8, 2), (8, 3),
              (8, 5), (8, 6), (9, 0), (9, 1), (9, 2), (9, 3), (9, 4), (9, 6), (9, 7), (9, 8)]


def test_reference_kat_graph1_restatement():           # test_maximum_clique.cpp:7-38 -> size 4
    g = og.Graph(10)
    for a, b in KAT1_EDGES:
        g.add_edge(a, b)
    assert len(g.find_maximum_clique()) == 4


def test_reference_kat_graph2_restatement():           # test_maximum_clique.cpp:40-53 -> size 9
    g = og.Graph(10)
    for i in range(10):
        for j in range(i + 1, 10):
            g.add_edge(i, j)
    g.delete_edge(0, 1)
    assert len(g.find_maximum_clique()) == 9


@needs_ref
def test_reference_kats_on_compiled_reference():
    assert len(ref.find_clique(10, KAT1_EDGES)) == 4
    full = [(i, j) for i in range(10) for j in range(i + 1, 10)]
    assert len(ref.find_clique(10, full, deleted=[(0, 1)])) == 9


@needs_ref
def test_clique_gate_mode_matches_reference_on_random_graphs():
    """FindClique(vertices, 7) — the only mode the hot path uses (sac_model_registration_graph.h:259)."""
    rng = np.random.default_rng(0)
    for trial in range(150):
        n = int(rng.integers(5, 40))
        p = rng.uniform(0.2, 0.9)
        edges = [(i, j) for i in range(n) for j in range(i + 1, n) if rng.random() < p]
        g = og.Graph(n)
        for a, b in edges:
            g.add_edge_sorted(a, b)
        assert g.find_clique(7) == ref.find_clique(n, edges, 7, sorted_insert=True), trial


def build_pair(n, frac, seed):
    q, t, px, _, _ = synth.make_cluster(n, frac, seed=seed)
    ar, oa = ref.RefAdjacencyRansac(), og.AdjacencyRansac()
    for i in range(n):
        ar.add_points(t[i], q[i], i)
        oa.add_points(t[i], q[i], i)
    ar.fill_adjacency(px, 0.25, 0.01)
    oa.fill_adjacency(px, 0.25, 0.01)
    return q, t, px, ar, oa


@needs_ref
@pytest.mark.parametrize("n,span,err", [(5, 0.25, 0.01), (50, 0.25, 0.01), (257, 0.1, 0.003), (300, 0.4, 0.05)])
def test_fill_adjacency_bit_identical_to_reference(n, span, err):
    q, t, px, _, _ = synth.make_cluster(n, 0.6, seed=n)
    ar = ref.RefAdjacencyRansac()
    for i in range(n):
        ar.add_points(t[i], q[i], i)
    ar.fill_adjacency(px, span, err)
    P, S = og.fill_adjacency_dense(q, t, px, span, err)
    assert (ar.dense("physical") == P).all()
    assert (ar.dense("sample") == S).all()
    assert (og.unpack_bits(og.pack_bits(P), n) == P).all()


@needs_ref
@pytest.mark.parametrize("n,frac,seed", [(60, 0.7, 1), (200, 0.4, 2), (400, 0.25, 3)])
def test_sampler_select_ransac_match_reference(n, frac, seed):
    q, t, px, ar, oa = build_pair(n, frac, seed)
    st = og.rng_seed(42, 0, 0)
    # the reference's own sampler on the shared stream == the restated sampler
    tr = ar.get_samples(st, 200)
    rng = og.Rng(st)
    assert [list(x) for x in tr] == [oa.get_samples(rng) for _ in range(len(tr))]
    for h in range(0, len(tr), 4):
        ri, R, T = ar.select(tr[h])
        oR, oT = og.kabsch(oa.query_points, oa.training_points, list(tr[h]))
        assert ri == oa.select_within_distance(list(tr[h]), oR, oT, float("inf"), {"best": 8})
        assert np.abs(R - oR).max() < 1e-4 and np.abs(T - oT).max() < 1e-4
        ri, R, T = ar.select(tr[h], 0.02)               # finite-threshold extension, reference poses
        assert ri == oa.select_within_distance(list(tr[h]), R, T, 0.02, {"best": 8})
    ri, it = ar.compute_model(300, st)
    trace = []
    oi, _, _ = oa.compute_model(og.Rng(st), 300, trace=trace)
    assert sorted(ri) == sorted(oi) and it == len(trace)
    ri, R, T = ar.ransac(0.01, 300, st)
    oi, oR, oT = oa.ransac(0.01, 300, og.Rng(st))
    assert ri == oi
    if oR is not None:
        assert np.abs(R - oR).max() < 1e-4 and np.abs(T - oT).max() < 1e-4
    if len(ri) >= 8:
        ar.invalidate_query_indices(ri)
        oa.invalidate_query_indices(oi)
        assert ar.valid() == oa.valid_indices
        st1 = og.rng_seed(42, 0, 1)
        ri, _, _ = ar.ransac(0.01, 300, st1)
        oi, _, _ = oa.ransac(0.01, 300, og.Rng(st1))
        assert ri == oi


def make_scene(seed, n_objects=4, rows=400, visible=(0, 2), n_kp=300, k=5, radius=35):
    descs, points = synth.make_db(n_objects, rows, seed=seed)
    fr = synth.make_frame(descs, points, list(visible), n_kp, seed=seed + 1)
    m, c = hk.knn_numpy(fr["descriptors"], descs, k, radius)
    p3 = hk.gather_points3d(m, c, points)
    spans = np.array([hk.object_span(p) for p in points], np.float32)
    return fr, m, c, p3, spans


@needs_ref
def test_full_guess_generation_matches_reference():
    """ClusterPerObject + FillAdjacency + RANSAC rounds + invalidation: restatement vs the reference's code, on a
    synthetic frame with two visible objects; poses also recover the planted ground truth."""
    fr, m, c, p3, spans = make_scene(5)
    r = ref.process(fr["keypoints_xy"], fr["cloud"], m, c, p3, spans, 8, 500, 0.01, seed=7)
    o = og.guess_process(fr["keypoints_xy"], fr["cloud"], m, c, p3, spans, 8, 500, 0.01, seed=7)
    assert len(r) == len(o) >= 2
    for (ro, rR, rT, ri), (oo, oR, oT, oi) in zip(r, o):
        assert ro == oo and ri == oi
        assert np.abs(rR - oR).max() < 1e-4 and np.abs(rT - oT).max() < 1e-4
    for obj, R, T, inl in r:
        gR, gT = fr["poses"][obj]
        assert np.abs(R - gR).max() < 0.02 and np.abs(T - gT).max() < 0.01


def test_shim_arithmetic_against_cv2():
    """The shim restates three OpenCV behaviours; check them against the real library where cv2 is importable."""
    cv2 = pytest.importorskip("cv2")
    v = np.array([1e4, 1e-1, 0], np.float32)
    assert cv2.norm(v) == float(np.sqrt(np.float64(v[0]) ** 2 + np.float64(v[1]) ** 2))   # double accumulation (Q8)
    rng = np.random.default_rng(3)
    for _ in range(20):
        n = int(rng.integers(3, 30))
        q = rng.normal(size=(n, 3)).astype(np.float32)
        Rg = synth.random_rotation(rng)
        t = (q @ Rg.T + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
        R, T = og.kabsch(q, t, list(range(n)))
        # the reference's recipe on real OpenCV: H via gemm(GEMM_1_T), SVDecomp, det fix, U * Vt
        ct = t.mean(axis=0, dtype=np.float32)
        cq = q.mean(axis=0, dtype=np.float32)
        H = cv2.gemm((t - ct).astype(np.float32), (q - cq).astype(np.float32), 1.0, None, 0.0, flags=cv2.GEMM_1_T)
        w, U, Vt = cv2.SVDecomp(H)
        if cv2.determinant(U) * cv2.determinant(Vt) < 0:
            Vt[2, :] *= -1
        Rc = U @ Vt
        assert np.abs(R - Rc).max() < 1e-4
