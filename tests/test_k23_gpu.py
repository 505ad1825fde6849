"""GPU: K2 (adjacency bit-matrices) and K3 (hypothesis scoring) through the C-ABI against the oracle restatement."""
import numpy as np
import pytest

from oracle import geometry as og
from tod_b200 import fill_adjacency, score_hypotheses, synth

pytestmark = pytest.mark.gpu


def oracle_bits(q, t, px, span, err):
    P, S = og.fill_adjacency_dense(q, t, px, span, err)
    return og.pack_bits(P), og.pack_bits(S), P, S


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 129, 1024, 1025, 2500])
def test_k2_single_cluster_bit_exact(n):
    q, t, px, _, _ = synth.make_cluster(n, 0.6, seed=100 + n)
    P, S, mo = fill_adjacency([0, n], q, t, px, [0.25], 0.01)
    eP, eS, _, _ = oracle_bits(q, t, px, 0.25, 0.01)
    assert (P.reshape(n, -1) == eP).all()
    assert (S.reshape(n, -1) == eS).all()
    assert mo[1] == n * og.row_words(n)


def test_k2_batched_clusters_and_thresholds():
    sizes = [5, 0, 300, 64, 1, 777, 33]
    qs, ts, ps, spans = [], [], [], []
    for i, n in enumerate(sizes):
        q, t, px, _, _ = synth.make_cluster(max(n, 1), 0.5, seed=500 + i, span=0.1 + 0.05 * i)
        qs.append(q[:n]); ts.append(t[:n]); ps.append(px[:n]); spans.append(0.1 + 0.05 * i)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    for err in (0.01, 0.003, 0.05):
        P, S, mo = fill_adjacency(off, np.concatenate(qs), np.concatenate(ts), np.concatenate(ps), spans, err)
        for c, n in enumerate(sizes):
            if n == 0:
                continue
            eP, eS, _, _ = oracle_bits(qs[c], ts[c], ps[c], spans[c], err)
            assert (P[mo[c]:mo[c + 1]].reshape(n, -1) == eP).all(), (c, err)
            assert (S[mo[c]:mo[c + 1]].reshape(n, -1) == eS).all(), (c, err)


def test_k2_nan_points_follow_reference_polarity():
    q, t, px, _, _ = synth.make_cluster(200, 0.7, seed=9)
    q[5, 1] = np.nan          # ClusterPerObject only rejects NaN in x (adjacency_ransac.cpp:189)
    q[77, 2] = np.nan
    P, S, _ = fill_adjacency([0, 200], q, t, px, [0.25], 0.01)
    eP, eS, dP, _ = oracle_bits(q, t, px, 0.25, 0.01)
    assert (P.reshape(200, -1) == eP).all() and (S.reshape(200, -1) == eS).all()
    assert dP[5].sum() == 199  # `NaN > x` is false twice: the pair is kept in the physical graph, as in the reference


def make_triples(S, rng, H):
    """Random triangles of the sample graph (like the reference's sampler would draw)."""
    n = S.shape[0]
    out = []
    tries = 0
    while len(out) < H and tries < 100000:
        tries += 1
        a = int(rng.integers(0, n))
        na = np.nonzero(S[a])[0]
        if len(na) == 0:
            continue
        b = int(na[rng.integers(0, len(na))])
        nab = np.nonzero(S[a] & S[b])[0]
        if len(nab) == 0:
            continue
        out.append((int(nab[rng.integers(0, len(nab))]), b, a))
    return np.array(out, np.uint32).reshape(-1, 3)


@pytest.mark.parametrize("n,frac", [(64, 0.8), (400, 0.5), (2100, 0.3)])
def test_k3_counts_reference_faithful_mode(n, frac):
    """threshold = +inf (the reference never sets it, sac.h:70): count = |common physical neighbours| + 3."""
    rng = np.random.default_rng(n)
    q, t, px, _, _ = synth.make_cluster(n, frac, seed=700 + n)
    eP, eS, dP, dS = oracle_bits(q, t, px, 0.25, 0.01)
    valid = np.ones(n, bool)
    valid[rng.integers(0, n, n // 10)] = False
    V = og.pack_bits(valid[None, :].repeat(1, 0))[0] if False else np.packbits(
        np.concatenate([valid, np.zeros(og.row_words(n) * 32 - n, bool)]), bitorder="little").view("<u4")
    tri = make_triples(dS, rng, 500)
    assert len(tri) > 50
    counts, R, T = score_hypotheses(q, t, eP, V, tri)
    exp = np.array([(dP[a] & dP[b] & dP[c] & valid).sum() + 3 for a, b, c in tri])
    assert (counts == exp).all()
    # rigid fit of the 3 samples agrees with the oracle's Kabsch (pose tolerance 1e-4, north_star)
    for h in range(0, len(tri), 25):
        eR, eT = og.kabsch(q, t, list(tri[h]))
        assert np.abs(R[h].reshape(3, 3) - eR).max() < 1e-4
        assert np.abs(T[h] - eT).max() < 1e-4


def test_k3_finite_threshold_extension():
    n = 600
    rng = np.random.default_rng(1)
    q, t, px, _, inl = synth.make_cluster(n, 0.5, seed=801)
    eP, eS, dP, dS = oracle_bits(q, t, px, 0.25, 0.01)
    V = np.packbits(np.concatenate([np.ones(n, bool), np.zeros(og.row_words(n) * 32 - n, bool)]),
                    bitorder="little").view("<u4")
    tri = make_triples(dS, rng, 120)
    thr = 0.02
    counts, R, T = score_hypotheses(q, t, eP, V, tri, threshold=thr)
    for h in range(len(tri)):
        a, b, c = [int(x) for x in tri[h]]
        cand = list(np.nonzero(dP[a] & dP[b] & dP[c])[0]) + [a, b, c]
        d2 = np.array([float(og.transform_dist_sq(R[h].reshape(3, 3), T[h], q[i], t[i])) for i in cand])
        margin = np.abs(np.sqrt(d2) - thr)
        exp_lo = int((d2[margin > 1e-5] < thr * thr).sum())
        assert exp_lo <= counts[h] <= exp_lo + int((margin <= 1e-5).sum())


def test_k3_nonfinite_points_never_count():
    n = 128
    rng = np.random.default_rng(2)
    q, t, px, _, _ = synth.make_cluster(n, 0.9, seed=901)
    eP, eS, dP, dS = oracle_bits(q, t, px, 0.25, 0.01)
    tri = make_triples(dS, rng, 64)
    q2 = q.copy()
    bad = [int(x) for x in set(range(n)) - set(tri.ravel().tolist())][:5]
    q2[bad, 2] = np.nan
    V = np.packbits(np.concatenate([np.ones(n, bool), np.zeros(og.row_words(n) * 32 - n, bool)]),
                    bitorder="little").view("<u4")
    counts, _, _ = score_hypotheses(q2, t, eP, V, tri)
    fin = np.ones(n, bool)
    fin[bad] = False
    exp = np.array([(dP[a] & dP[b] & dP[c] & fin).sum() + 3 for a, b, c in tri])
    assert (counts == exp).all()
