"""GPU: K1 (exact k-NN Hamming) through the C-ABI, bit-exact against the cv2 golden vectors and the oracle."""
import numpy as np
import pytest

from conftest import assert_matches_equal, golden_names, load_golden
from oracle import hamming_knn as hk
from tod_b200 import DescriptorMatcher, capi, synth

pytestmark = pytest.mark.gpu

KERNELS = {"popc": capi.TOD_KERNEL_POPC, "mma": capi.TOD_KERNEL_MMA}
both_kernels = pytest.mark.parametrize("kernel", ["popc", "mma"])


def run_matcher(query, descs, points, k, radius, kernel="popc", **kw):
    m = DescriptorMatcher(k=k, radius=radius, kernel=KERNELS[kernel], **kw)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("obj%d" % i, d, p)
    m.train()
    out = m.process(query)
    out["kernel"] = m.last_kernel
    assert query.shape[0] == 0 or m.num_descriptors == 0 or out["kernel"] == kernel
    out["span_idx"] = m.spans_by_index
    m.close()
    return out


def check_against_oracle(query, descs, points, k, radius, kernel="popc"):
    out = run_matcher(query, descs, points, k, radius, kernel)
    em, ec = hk.knn_c(query, descs, k, radius)
    assert_matches_equal(out["matches"], out["counts"], em["trainIdx"], em["imgIdx"], em["distance"], ec)
    e3 = hk.gather_points3d(em, ec, points)
    mask = np.arange(k)[None, :] < ec[:, None]
    assert (out["matches_3d"][mask] == e3[mask]).all()
    for i, p in enumerate(points):
        assert out["span_idx"][i] == hk.object_span(p)
    return out


@both_kernels
@pytest.mark.parametrize("name", golden_names())
def test_cv2_golden_vectors(name, kernel):
    g, objs = load_golden(name)
    pts = [np.zeros((o.shape[0], 3), np.float32) for o in objs]
    out = run_matcher(g["query"], objs, pts, int(g["k"]), int(g["radius"]), kernel)
    assert_matches_equal(out["matches"], out["counts"], g["trainIdx"], g["imgIdx"], g["distance"], g["counts"])
    assert out["kernel"] in ("popc", "mma")


@both_kernels
@pytest.mark.parametrize("nq", [1, 31, 255, 256, 257, 513, 1025, 2000])
def test_ragged_query_counts(nq, kernel):
    descs, points = synth.make_db(3, [700, 1300, 555], seed=11)
    q, _, _ = synth.make_queries(descs, nq, seed=nq)
    check_against_oracle(q, descs, points, 5, 0, kernel)


@both_kernels
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
def test_every_k(k, kernel):
    descs, points = synth.make_db(4, 2500, seed=21)
    q, _, _ = synth.make_queries(descs, 300, seed=22)
    check_against_oracle(q, descs, points, k, 0, kernel)


@both_kernels
def test_tie_heavy_descriptors(kernel):
    rng = np.random.default_rng(3)
    descs = [np.zeros((n, 32), np.uint8) for n in (3000, 2000, 4100)]
    for d in descs:
        d[:, 5] = rng.integers(0, 8, d.shape[0])       # only 3 significant bits: thousands of exact ties
    points = [rng.random((d.shape[0], 3)).astype(np.float32) for d in descs]
    q = np.zeros((200, 32), np.uint8)
    q[:, 5] = rng.integers(0, 8, 200)
    check_against_oracle(q, descs, points, 5, 0, kernel)
    check_against_oracle(q, descs, points, 5, 1, kernel)


@both_kernels
def test_config_c2_shape_k2(kernel):
    """BASELINE configs[1]: 10-object DB (50k descriptors), 1k query keypoints, k=2."""
    descs, points = synth.make_db(10, 5000, seed=synth.BASE_SEED + 1)
    q, _, _ = synth.make_queries(descs, 1000, seed=synth.BASE_SEED + 101)
    check_against_oracle(q, descs, points, 2, 0, kernel)


@both_kernels
def test_radius_cut_like_detection_ork(kernel):
    descs, points = synth.make_db(10, 5000, seed=31)
    q, src_obj, src_row = synth.make_queries(descs, 1000, seed=32)
    out = check_against_oracle(q, descs, points, 5, 35, kernel)
    true = src_obj >= 0
    assert (out["counts"][true] >= 1).mean() > 0.99      # 4% flips ~ 10 bits < radius 35
    assert (out["counts"][~true] == 0).all()             # random clutter never gets within 35 bits
    hit = out["matches"][true][:, 0]
    assert (hit["imgIdx"] == src_obj[true]).mean() > 0.99


@both_kernels
def test_config_c3_full_size_2k_by_1m(kernel):
    """north_star size: 2k keypoints x 1M descriptors (100 objects x 10k), k=2 — bit-exact against the C oracle on a
    query subset, plus size-independent properties on all queries."""
    descs, points = synth.make_db(100, 10000, seed=synth.BASE_SEED + 2)
    q, src_obj, src_row = synth.make_queries(descs, 2000, seed=synth.BASE_SEED + 102)
    out = run_matcher(q, descs, points, 2, 0, kernel)
    m, c = out["matches"], out["counts"]
    assert (c == 2).all()
    assert (m["distance"][:, 0] <= m["distance"][:, 1]).all()            # sortedness
    true = src_obj >= 0                                                  # planted rows are found as the best match
    assert (m["imgIdx"][true, 0] == src_obj[true]).mean() > 0.999
    assert (m["trainIdx"][true, 0] == src_row[true]).mean() > 0.999
    sub = np.arange(0, 2000, 8)
    em, ec = hk.knn_c(q[sub], descs, 2, 0)
    for f in ("trainIdx", "imgIdx", "distance"):
        assert (m[f][sub] == em[f]).all()
    # idempotence: distances recomputed from the returned indices
    db, off = hk.concat_objects(descs)
    g = off[m["imgIdx"]] + m["trainIdx"]
    d = np.bitwise_count(q.view(np.uint64)[:, None, :] ^ db.view(np.uint64)[g]).sum(axis=2)
    assert (d == m["distance"]).all()


def test_empty_and_error_paths():
    from tod_b200 import capi
    m = DescriptorMatcher(k=5)
    with pytest.raises(capi.TodError) as e:
        m.process(np.zeros((4, 32), np.uint8))
    assert e.value.code == capi.TOD_ERR_STATE                            # knn before train
    m.train()                                                            # empty DB
    out = m.process(np.zeros((4, 32), np.uint8))
    assert (out["counts"] == 0).all()
    descs, points = synth.make_db(1, 3, seed=1)                          # DB smaller than k
    m.add_object("tiny", descs[0], points[0])
    m.train()
    out = m.process(np.zeros((2, 32), np.uint8))
    assert (out["counts"] == 3).all()
    out = m.process(np.zeros((0, 32), np.uint8))                         # no queries
    assert out["matches"].shape == (0, 5)
    with pytest.raises(capi.TodError):
        DescriptorMatcher(k=9)
    m.close()


@both_kernels
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_db_merge_equals_unsharded(world, kernel):
    """The multi-GPU path on one device: `world` matcher handles, each holding one row range of the DB, produce packed
    keys (knn_keys_device); the concatenation (what the NCCL all-gather delivers) goes through merge_device.  Must
    equal the single-handle result bit for bit — shard boundaries cut through objects and through runs of ties."""
    import torch
    rng = np.random.default_rng(world)
    descs, points = synth.make_db(5, [1500, 700, 2300, 41, 999], seed=60 + world)
    descs[1][:, :] = 0
    descs[1][:, 7] = rng.integers(0, 4, descs[1].shape[0])           # a tie-heavy object in the middle
    q, _, _ = synth.make_queries(descs, 333, seed=61)
    q[:50] = 0
    k, radius = 5, 0
    dev = torch.device("cuda", 0)
    nq = q.shape[0]
    q_dev = torch.from_numpy(q).to(dev)
    keys_all = torch.empty((world, nq, k), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    # one explicit stream for every handle: a NULL stream argument means "the handle's own stream", and the handles'
    # streams are not ordered with respect to each other
    stream = torch.cuda.Stream()
    sptr = stream.cuda_stream
    assert sptr != 0
    handles = []
    for r in range(world):
        m = DescriptorMatcher(k=k, radius=radius, kernel=KERNELS[kernel], shard_rank=r, shard_count=world)
        for i, (d, p) in enumerate(zip(descs, points)):
            m.add_object("o%d" % i, d, p)
        m.train()
        m.knn_keys_device(q_dev.data_ptr(), nq, keys_all[r].data_ptr(), sptr)
        handles.append(m)
    assert sum(h.shard_rows for h in handles) == sum(d.shape[0] for d in descs)
    matches = torch.empty((nq, k, 4), dtype=torch.int32, device=dev)
    counts = torch.empty((nq,), dtype=torch.int32, device=dev)
    pts = torch.empty((nq, k, 3), dtype=torch.float32, device=dev)
    handles[0].merge_device(keys_all.data_ptr(), world, nq, matches.data_ptr(), counts.data_ptr(), pts.data_ptr(),
                            sptr)
    torch.cuda.synchronize()
    got = matches.cpu().numpy().view(capi.MATCH_DTYPE).reshape(nq, k)
    em, ec = hk.knn_c(q, descs, k, radius)
    assert_matches_equal(got, counts.cpu().numpy(), em["trainIdx"], em["imgIdx"], em["distance"], ec)
    e3 = hk.gather_points3d(em, ec, points)
    assert (pts.cpu().numpy() == e3).all()
    for h in handles:
        h.close()


def test_maximum_db_size_and_limit():
    """The packed key addresses 2^23 rows: a DB of exactly that size works (bit-exact on planted and random queries,
    first / last rows included), one more row is refused with TOD_ERR_LIMIT."""
    rng = np.random.default_rng(99)
    n_obj, rows = 32, 262144                                   # 32 x 2^18 = 2^23 descriptors (268 MB packed)
    base = rng.integers(0, 256, (rows, 32), dtype=np.uint8)
    pts = rng.random((rows, 3)).astype(np.float32)
    m = DescriptorMatcher(k=2, radius=0)
    descs = []
    for o in range(n_obj):
        d = base.copy()
        d[:, 0] ^= np.uint8(o)                                 # objects differ in a few bits of byte 0
        d[:, 31] = rng.integers(0, 256, rows, dtype=np.uint8)
        descs.append(d)
        m.add_object("o%d" % o, d, pts)
    assert m.num_descriptors == 1 << 23
    with pytest.raises(capi.TodError) as e:
        m.add_object("one_too_many", base[:1], pts[:1])
    assert e.value.code == capi.TOD_ERR_LIMIT
    m.train()
    # queries: exact copies of chosen rows (first row of the DB, last row of the DB, random ones) + random clutter
    picks = [(0, 0), (n_obj - 1, rows - 1)] + [(int(rng.integers(0, n_obj)), int(rng.integers(0, rows)))
                                              for _ in range(30)]
    q = np.stack([descs[o][r] for o, r in picks] + [rng.integers(0, 256, 32, dtype=np.uint8) for _ in range(32)])
    out = m.process(q)
    assert m.last_kernel == "mma"
    em, ec = hk.knn_c(q, descs, 2, 0)
    assert_matches_equal(out["matches"], out["counts"], em["trainIdx"], em["imgIdx"], em["distance"], ec)
    for i, (o, r) in enumerate(picks):
        assert out["matches"]["distance"][i, 0] == 0
        # the same row of a lower-numbered object can tie at distance 0 only if byte 0 and 31 agree; the oracle decides
        assert (int(out["matches"]["imgIdx"][i, 0]), int(out["matches"]["trainIdx"][i, 0])) <= (o, r)
    m.close()
