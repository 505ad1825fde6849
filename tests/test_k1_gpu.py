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
    """north_star size: 2k keypoints x 1M descriptors (100 objects x 10k), k=2 — bit-exact against the C oracle on ALL
    2000 queries (matches, counts, matches_3d), plus size-independent properties."""
    descs, points = synth.make_db(100, 10000, seed=synth.BASE_SEED + 2)
    q, src_obj, src_row = synth.make_queries(descs, 2000, seed=synth.BASE_SEED + 102)
    out = check_against_oracle(q, descs, points, 2, 0, kernel)
    m, c = out["matches"], out["counts"]
    assert (c == 2).all()
    assert (m["distance"][:, 0] <= m["distance"][:, 1]).all()            # sortedness
    true = src_obj >= 0                                                  # planted rows are found as the best match
    assert (m["imgIdx"][true, 0] == src_obj[true]).mean() > 0.999
    assert (m["trainIdx"][true, 0] == src_row[true]).mean() > 0.999
    # idempotence: distances recomputed from the returned indices
    db, off = hk.concat_objects(descs)
    g = off[m["imgIdx"]] + m["trainIdx"]
    d = np.bitwise_count(q.view(np.uint64)[:, None, :] ^ db.view(np.uint64)[g]).sum(axis=2)
    assert (d == m["distance"]).all()


def test_timed_bench_shape_128k_queries_by_1m():
    """The exact plan bench.py times: one call with 64 frames x 2000 = 128 000 queries against the 1M-descriptor DB
    (500 query tiles x db chunks), host-buffer call and device-buffer call.  Bit-exact against the C oracle on a
    sample that covers the first and last query tiles and every frame; sortedness / distance recomputation on all."""
    import torch
    descs, points = synth.make_db(100, 10000, seed=synth.BASE_SEED + 2)
    q = np.ascontiguousarray(np.concatenate(
        [synth.make_queries(descs, 2000, seed=synth.BASE_SEED + 102 + f)[0] for f in range(64)]))
    nq, k = q.shape[0], 2
    m = DescriptorMatcher(k=k, radius=0)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("obj%d" % i, d, p)
    m.train()
    m.reserve(nq)
    out = m.process(q)
    assert m.last_kernel == "mma"
    mm, c = out["matches"], out["counts"]
    assert (c == k).all() and (mm["distance"][:, 0] <= mm["distance"][:, 1]).all()
    rng = np.random.default_rng(7)
    idx = np.unique(np.concatenate([np.arange(256), np.arange(nq - 256, nq), rng.choice(nq, 768, replace=False),
                                    np.arange(64) * 2000, np.arange(64) * 2000 + 1999]))
    em, ec = hk.knn_c(q[idx], descs, k, 0)
    assert (c[idx] == ec).all()
    for f in ("trainIdx", "imgIdx", "distance"):
        assert (mm[f][idx] == em[f]).all(), f
    assert (mm["queryIdx"][idx] == idx[:, None]).all()
    assert (out["matches_3d"][idx] == hk.gather_points3d(em, ec, points)).all()
    db, off = hk.concat_objects(descs)                                   # every distance recomputed from its indices
    for lo in range(0, nq, 16000):
        g = off[mm["imgIdx"][lo:lo + 16000]] + mm["trainIdx"][lo:lo + 16000]
        d = np.bitwise_count(q[lo:lo + 16000].view(np.uint64)[:, None, :] ^ db.view(np.uint64)[g]).sum(axis=2)
        assert (d == mm["distance"][lo:lo + 16000]).all()
    # the device-buffer entry point (what bench.py's `value` times) returns the same bytes
    dev = torch.device("cuda", 0)
    qd = torch.from_numpy(q).to(dev)
    md = torch.empty((nq, k, 4), dtype=torch.int32, device=dev)
    cd = torch.empty((nq,), dtype=torch.int32, device=dev)
    pd = torch.empty((nq, k, 3), dtype=torch.float32, device=dev)
    st = torch.cuda.Stream()
    qd.record_stream(st)
    torch.cuda.synchronize()
    for _ in range(2):
        m.process_device(qd.data_ptr(), nq, md.data_ptr(), cd.data_ptr(), pd.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    assert (md.cpu().numpy().view(capi.MATCH_DTYPE).reshape(nq, k) == mm).all()
    assert (cd.cpu().numpy() == c).all() and (pd.cpu().numpy() == out["matches_3d"]).all()
    m.close()


# ---- the two TODO blocks of DescriptorMatcher.cpp:223-229 as opt-in extensions ----------------------------------
def lowe_ratio_numpy(em, ec, ratio, radius):
    """numpy restatement: on the k-NN lists (distances pinned to cv2 by the oracle), BEFORE the radius cut, keep the
    best match only, and only if distance0 < ratio * distance1 in float32 (a lone neighbour is kept)."""
    nq, k = em.shape
    out = np.zeros_like(em)
    out["queryIdx"] = out["trainIdx"] = out["imgIdx"] = -1
    cnt = np.zeros(nq, np.int32)
    for i in range(nq):
        if ec[i] == 0:
            continue
        d0 = np.float32(em["distance"][i, 0])
        ok = True if ec[i] < 2 else bool(d0 < np.float32(ratio) * np.float32(em["distance"][i, 1]))
        if ok and (radius == 0 or d0 <= radius):
            out[i, 0] = em[i, 0]
            cnt[i] = 1
    return out, cnt


def dedupe_numpy(em, ec, offsets, frame_kp):
    """Inside each frame a DB descriptor keeps only the match with the smallest (distance, queryIdx); lists compacted."""
    nq, k = em.shape
    best = {}
    for i in range(nq):
        f = i // frame_kp if frame_kp else 0
        for j in range(int(ec[i])):
            key = (f, int(offsets[em["imgIdx"][i, j]] + em["trainIdx"][i, j]))
            val = (int(em["distance"][i, j]), i)
            if key not in best or val < best[key]:
                best[key] = val
    out = np.zeros_like(em)
    out["queryIdx"] = out["trainIdx"] = out["imgIdx"] = -1
    cnt = np.zeros(nq, np.int32)
    for i in range(nq):
        f = i // frame_kp if frame_kp else 0
        for j in range(int(ec[i])):
            key = (f, int(offsets[em["imgIdx"][i, j]] + em["trainIdx"][i, j]))
            if best[key] == (int(em["distance"][i, j]), i):
                out[i, cnt[i]] = em[i, j]
                cnt[i] += 1
    return out, cnt


def oracle_lists(q, descs, k):
    em, ec = hk.knn_c(q, descs, k, 0)
    full = np.zeros(em.shape, capi.MATCH_DTYPE)
    for f in ("trainIdx", "imgIdx", "distance"):
        full[f] = em[f]
    full["queryIdx"] = np.arange(q.shape[0])[:, None]
    return full, ec


@both_kernels
@pytest.mark.parametrize("ratio,radius", [(0.8, 0), (0.6, 35), (0.95, 55)])
def test_ratio_test_extension(kernel, ratio, radius):
    descs, points = synth.make_db(5, 1500, seed=71)
    q, _, _ = synth.make_queries(descs, 600, seed=72, flip_p=0.12)
    q[:40] = descs[2][:40]                                      # exact duplicates: distance0 == 0
    descs[3][:40] = descs[2][:40]                               # ... present twice in the DB: 0 < ratio * 0 fails
    m = DescriptorMatcher(k=5, radius=radius, kernel=KERNELS[kernel], ratio=ratio)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("o%d" % i, d, p)
    m.train()
    out = m.process(q)
    m.close()
    full, ec = oracle_lists(q, descs, 5)
    exp, cnt = lowe_ratio_numpy(full, ec, ratio, radius)
    assert (out["counts"] == cnt).all()
    assert (out["counts"][:40] == 0).all()
    keep = cnt == 1
    for f in ("queryIdx", "trainIdx", "imgIdx", "distance"):
        assert (out["matches"][f][keep, 0] == exp[f][keep, 0]).all(), f
    assert 0 < keep.sum() < q.shape[0]
    # default (reference-faithful): the same .ork string leaves the lists untouched — the reference's block is empty
    m2 = DescriptorMatcher(search_json_params='{"type": "LSH", "radius": %d, "ratio": %g}' % (radius, ratio),
                           kernel=KERNELS[kernel])
    for i, (d, p) in enumerate(zip(descs, points)):
        m2.add_object("o%d" % i, d, p)
    m2.train()
    plain = m2.process(q)
    m2.close()
    em, ec2 = hk.knn_c(q, descs, 5, radius)
    assert_matches_equal(plain["matches"], plain["counts"], em["trainIdx"], em["imgIdx"], em["distance"], ec2)


@pytest.mark.parametrize("frame_kp", [0, 100])
def test_duplicate_match_removal_extension(frame_kp):
    descs, points = synth.make_db(4, 800, seed=81)
    rng = np.random.default_rng(82)
    base, _, _ = synth.make_queries(descs, 100, seed=83)
    q = np.concatenate([base, base, base])                       # every true match is claimed three times
    flips = np.packbits(rng.random((300, 256)) < 0.01, axis=1)
    q = np.ascontiguousarray(q ^ flips)
    m = DescriptorMatcher(k=5, radius=35, remove_duplicates=True, frame_keypoints=frame_kp)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("o%d" % i, d, p)
    m.train()
    out = m.process(q)
    m.close()
    em, ec = hk.knn_c(q, descs, 5, 35)
    full = np.zeros(em.shape, capi.MATCH_DTYPE)
    for f in ("trainIdx", "imgIdx", "distance"):
        full[f] = em[f]
    full["queryIdx"] = np.arange(q.shape[0])[:, None]
    _, off = hk.concat_objects(descs)
    exp, cnt = dedupe_numpy(full, ec, off, frame_kp)
    assert (out["counts"] == cnt).all()
    mask = np.arange(5)[None, :] < cnt[:, None]
    for f in ("queryIdx", "trainIdx", "imgIdx", "distance"):
        assert (out["matches"][f][mask] == exp[f][mask]).all(), f
    e3 = hk.gather_points3d(exp, cnt, points)
    assert (out["matches_3d"][mask] == e3[mask]).all()
    assert cnt.sum() < ec.sum()                                   # something was removed
    if frame_kp:                                                  # frames are independent: each keeps its own winner
        assert cnt[:100].sum() > 0 and cnt[100:200].sum() > 0 and cnt[200:].sum() > 0


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


def _big_db(n_obj, rows, seed):
    """n_obj objects of `rows` descriptors that differ from one another in byte 0 only (plus a random last byte)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (rows, 32), dtype=np.uint8)
    pts = rng.random((rows, 3)).astype(np.float32)
    descs = []
    for o in range(n_obj):
        d = base.copy()
        d[:, 0] ^= np.uint8(o)                                 # objects differ in a few bits of byte 0
        d[:, 31] = rng.integers(0, 256, rows, dtype=np.uint8)
        descs.append(d)
    return descs, pts, rng


def test_one_segment_database_2_pow_23_rows():
    """2^23 rows — the most one packed key addresses — in one segment: bit-exact on planted and random queries, first
    and last rows included."""
    n_obj, rows = 32, 262144                                   # 32 x 2^18 = 2^23 descriptors (268 MB packed)
    descs, pts, rng = _big_db(n_obj, rows, 99)
    m = DescriptorMatcher(k=2, radius=0)
    for o in range(n_obj):
        m.add_object("o%d" % o, descs[o], pts)
    assert m.num_descriptors == 1 << 23
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


@pytest.mark.parametrize("kernel", ["mma", "popc"])
def test_wide_database_is_scanned_in_segments(kernel):
    """More than 2^23 rows: the shard is scanned in segments of 2^23 rows whose keys are local to the segment, and the
    merge orders (distance, global row) on 64 bits.  33 x 2^18 rows = two segments; queries planted on both sides of
    the boundary (its last and first rows included), and an object copied into BOTH segments so that equal distances
    meet across the boundary: cv::BFMatcher's order — the lower imgIdx first — must hold, bit for bit vs the oracle."""
    n_obj, rows = 33, 262144
    descs, pts, rng = _big_db(n_obj, rows, 100)
    descs[32] = descs[3].copy()                                # object 32 (second segment) == object 3 (first segment)
    m = DescriptorMatcher(k=3, radius=0, kernel=capi.TOD_KERNEL_MMA if kernel == "mma" else capi.TOD_KERNEL_POPC)
    for o in range(n_obj):
        m.add_object("o%d" % o, descs[o], pts)
    assert m.num_descriptors == (1 << 23) + rows
    m.train()
    picks = [(0, 0), (31, rows - 1), (32, 0), (32, rows - 1), (3, 77), (32, 77)] + \
            [(int(rng.integers(0, n_obj)), int(rng.integers(0, rows))) for _ in range(26)]
    nq = 64 if kernel == "mma" else 40
    q = np.stack([descs[o][r] for o, r in picks] + [rng.integers(0, 256, 32, dtype=np.uint8)
                                                    for _ in range(nq - len(picks))])
    q[8, 5] ^= np.uint8(3)                                     # a near copy: distance 2 to its source row
    out = m.process(q)
    assert m.last_kernel == kernel
    em, ec = hk.knn_c(q, descs, 3, 0)
    assert_matches_equal(out["matches"], out["counts"], em["trainIdx"], em["imgIdx"], em["distance"], ec)
    e3 = hk.gather_points3d(em, ec, [pts] * n_obj)
    assert (out["matches_3d"] == e3).all()
    # the twin objects: the copy of row 77 is found in object 3 first, then in object 32, both at distance 0
    for i in (4, 5):
        assert list(out["matches"]["imgIdx"][i, :2]) == [3, 32] and list(out["matches"]["trainIdx"][i, :2]) == [77, 77]
        assert list(out["matches"]["distance"][i, :2]) == [0, 0]
    # a second call with a radius and fewer queries (ragged), same handle
    out2 = m.process(q[:17])
    assert_matches_equal(out2["matches"], out2["counts"], em["trainIdx"][:17], em["imgIdx"][:17], em["distance"][:17],
                         ec[:17])
    # the stage calls carry 32-bit global keys: refused on a wide database
    import torch
    keys = torch.empty((17, 3), dtype=torch.int32, device="cuda")
    with pytest.raises(capi.TodError) as e:
        m.knn_keys_device(torch.from_numpy(q[:17]).cuda().data_ptr(), 17, keys.data_ptr())
    assert e.value.code == capi.TOD_ERR_LIMIT
    m.close()
