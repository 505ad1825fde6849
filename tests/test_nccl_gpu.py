"""GPU, >= 2 devices: the sharded matcher with the exchange INSIDE the library — one process per GPU, each handle holds
one row range of the DB and an NCCL communicator (tod_matcher_set_comm); tod_matcher_knn / tod_matcher_knn_device then
run K1 -> top-k reduction -> exchange -> merge on the handle's stream and every rank must return the complete,
bit-exact result (oracle = exact Hamming k-NN over the whole DB), step after step with changing query sets — with the
peer-memory exchange (reduce_push_kernel storing into the other GPUs over NVLink; comm_mode 2, the default where the
GPUs can map each other) and with the ncclAllGather exchange (comm_mode 1), including a query set larger than the
reserved size, which makes the ranks re-map their exchange buffers.

Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_nccl_gpu.py -m gpu`."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _workload():
    from tod_b200 import synth
    rng = np.random.default_rng(17)
    descs, points = synth.make_db(7, [9000, 1200, 30000, 41, 15000, 7777, 22000], seed=301)
    descs[1][:, :] = 0
    descs[1][:, 9] = rng.integers(0, 4, descs[1].shape[0])            # tie-heavy object straddling a shard boundary
    steps = []
    for s, nq in enumerate([700, 2000, 333, 2000, 1, 4096, 5000]):    # ragged, changing sizes step after step
        q, _, _ = synth.make_queries(descs, nq, seed=400 + s)
        if s == 2:
            q[:60] = 0
        steps.append(q)
    return descs, points, steps


def _rank_main(rank, world, uid, k, radius, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    from oracle import hamming_knn as hk
    from tod_b200 import DescriptorMatcher, capi
    torch.cuda.set_device(rank)
    descs, points, steps = _workload()
    m = DescriptorMatcher(k=k, radius=radius, device=rank, shard_rank=rank, shard_count=world)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("o%d" % i, d, p)
    m.train()
    m.reserve(4096)
    # a sharded handle without a communicator must refuse the whole-call entry point
    try:
        m.process(steps[0])
        raise AssertionError("expected TOD_ERR_STATE")
    except capi.TodError as e:
        assert e.code == capi.TOD_ERR_STATE
    m.set_comm(uid)
    mode = m.comm_mode
    assert mode == 1
    dev = torch.device("cuda", rank)
    ok = True
    modes = []
    for rep in range(3):
        if rep == 2:
            m.set_exchange(False)                                      # the ncclAllGather path, same results
        for q in steps:
            out = m.process(q)                                         # host buffers: H2D + K1 + all-gather + merge + D2H
            em, ec = hk.knn_c(q, descs, k, radius)
            same = (out["counts"] == ec).all()
            mask = np.arange(k)[None, :] < ec[:, None]
            for f in ("trainIdx", "imgIdx", "distance"):
                same = same and (out["matches"][f][mask] == em[f][mask]).all()
            e3 = hk.gather_points3d(em, ec, points)
            same = same and (out["matches_3d"][mask] == e3[mask]).all()
            # device buffers, user stream
            nq = q.shape[0]
            qd = torch.from_numpy(q).to(dev)
            md = torch.empty((nq, k, 4), dtype=torch.int32, device=dev)
            cd = torch.empty((nq,), dtype=torch.int32, device=dev)
            pd = torch.empty((nq, k, 3), dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            m.process_device(qd.data_ptr(), nq, md.data_ptr(), cd.data_ptr(), pd.data_ptr())
            torch.cuda.synchronize()
            same = same and (md.cpu().numpy().view(capi.MATCH_DTYPE).reshape(nq, k) == out["matches"]).all()
            same = same and (cd.cpu().numpy() == out["counts"]).all()
            ok = ok and bool(same)
        modes.append(m.comm_mode)
        ok = ok and m.exchange_error == 0
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([int(ok), mode, m.shard_rows] + modes))
    if rank == 0:
        import time
        time.sleep(2.0)            # ranks tear their handles down at different times: nobody may wait for anybody
    m.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("k,radius", [(2, 0), (5, 35)])
def test_sharded_matcher_with_in_library_nccl(tmp_path, k, radius):
    import torch.multiprocessing as mp
    from tod_b200 import comm_unique_id
    world = min(_n_gpus(), 4)
    uid = comm_unique_id()
    mp.spawn(_rank_main, args=(world, uid, k, radius, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r)) for r in range(world)]
    assert all(int(r[0]) == 1 for r in res), res
    assert all(int(r[5]) == 1 for r in res), res                       # third pass ran on ncclAllGather
    assert len(set(int(r[3]) for r in res)) == 1                       # every rank made the same choice of exchange
    print("exchange modes per pass:", [int(x) for x in res[0][3:6]])
    descs, _, _ = _workload()
    assert sum(int(r[2]) for r in res) == sum(d.shape[0] for d in descs)


def _rank_main_wide(rank, world, uid, out_dir):
    """A database of more than 2^23 rows over `world` GPUs: keys are local to their shard's segment, the exchange
    carries them as they are and the merge orders (distance, global row) on 64 bits."""
    sys.path.insert(0, ROOT)
    import torch
    from oracle import hamming_knn as hk
    from tod_b200 import DescriptorMatcher
    torch.cuda.set_device(rank)
    rng = np.random.default_rng(5)
    n_obj, rows = 33, 262144                                    # 2^23 + 2^18 rows
    base = rng.integers(0, 256, (rows, 32), dtype=np.uint8)
    pts = rng.random((rows, 3)).astype(np.float32)
    descs = []
    for o in range(n_obj):
        d = base.copy()
        d[:, 0] ^= np.uint8(o)
        d[:, 31] = rng.integers(0, 256, rows, dtype=np.uint8)
        descs.append(d)
    descs[32] = descs[3].copy()                                 # twins in the first and the last shard: ties across ranks
    m = DescriptorMatcher(k=3, radius=0, device=rank, shard_rank=rank, shard_count=world)
    for o in range(n_obj):
        m.add_object("o%d" % o, descs[o], pts)
    m.train()
    m.set_comm(uid)
    picks = [(0, 0), (32, rows - 1), (3, 77), (32, 77), (16, rows // 2)] + \
            [(int(rng.integers(0, n_obj)), int(rng.integers(0, rows))) for _ in range(27)]
    q = np.stack([descs[o][r] for o, r in picks] + [rng.integers(0, 256, 32, dtype=np.uint8) for _ in range(32)])
    ok = True
    for rep in range(2):
        out = m.process(q)
        em, ec = hk.knn_c(q, descs, 3, 0)
        same = (out["counts"] == ec).all()
        for f in ("trainIdx", "imgIdx", "distance"):
            same = same and (out["matches"][f] == em[f]).all()
        same = same and (out["matches_3d"] == hk.gather_points3d(em, ec, [pts] * n_obj)).all()
        same = same and list(out["matches"]["imgIdx"][2, :2]) == [3, 32]
        ok = ok and bool(same)
    np.save(os.path.join(out_dir, "wide_%d.npy" % rank), np.array([int(ok), m.comm_mode, m.shard_rows]))
    m.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_sharded_wide_database(tmp_path):
    import torch.multiprocessing as mp
    from tod_b200 import comm_unique_id
    world = 2
    uid = comm_unique_id()
    mp.spawn(_rank_main_wide, args=(world, uid, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(str(tmp_path), "wide_%d.npy" % r)) for r in range(world)]
    assert all(int(r[0]) == 1 for r in res), res
    assert all(int(r[1]) == 1 for r in res)                      # wide databases use the ncclAllGather exchange
    assert sum(int(r[2]) for r in res) == (1 << 23) + 262144
