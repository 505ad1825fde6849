"""CPU, world size 2 over gloo: the host-side protocol of the sharded matcher — every rank scores its
row range of the DB (here with the oracle standing in for K1), packs its per-query top-k as u32 keys, one all_gather
moves the keys, and a plain min-k merge on every rank must reproduce the single-node result bit for bit (including
ties across the shard boundary).  The shard split and the key packing come from the product library's host-only entry
points (tod_shard_range, tod_pack_key), so the test pins the same arithmetic tod_matcher_train uses."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY_ROW_BITS = 23
EMPTY = 0xFFFFFFFF


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _workload(tie_heavy):
    from tod_b200 import synth
    if tie_heavy:
        rng = np.random.default_rng(5)
        descs = [np.zeros((n, 32), np.uint8) for n in (700, 523, 901)]
        for d in descs:
            d[:, 3] = rng.integers(0, 4, d.shape[0])          # 2 significant bits: ties everywhere, across shards too
        q = np.zeros((96, 32), np.uint8)
        q[:, 3] = rng.integers(0, 4, 96)
        return descs, q
    descs, _ = synth.make_db(5, [400, 333, 512, 77, 1001], seed=41)
    q, _, _ = synth.make_queries(descs, 128, seed=42)
    return descs, q


def _rank_main(rank, world, port, k, radius, tie_heavy, out_dir, local_keys=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import hamming_knn as hk
    from tod_b200 import capi
    lib = capi.load()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    descs, q = _workload(tie_heavy)
    db, off = hk.concat_objects(descs)
    begin, rows = ctypes.c_int64(), ctypes.c_int64()
    b2, r2 = ctypes.c_int64(), ctypes.c_int64()
    assert lib.tod_shard_range(db.shape[0], rank, world, ctypes.byref(begin), ctypes.byref(rows)) == capi.TOD_OK
    begin, rows = begin.value, rows.value
    # this rank's candidates: exact top-k of its row range as ONE pseudo-object, re-based to global rows
    keys = np.full((q.shape[0], k), EMPTY, np.uint32)
    if rows > 0:
        m, c = hk.knn_c(q, [db[begin:begin + rows]], k, radius)
        for i in range(q.shape[0]):
            for j in range(int(c[i])):
                # wide databases: the key counts rows from the start of the rank's segment, not from row 0 of the DB
                keys[i, j] = lib.tod_pack_key(int(m["distance"][i, j]),
                                              int(m["trainIdx"][i, j]) + (0 if local_keys else begin))
    mine = torch.from_numpy(keys.view(np.int32).copy())
    gathered = torch.empty((world,) + tuple(mine.shape), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1))
    allk = gathered.numpy().view(np.uint32)                       # world x nq x k
    if local_keys:
        # the merge of a wide database: (distance, first row of the source + local row) on 64 bits, empty = all ones
        bases = []
        for r in range(world):
            assert lib.tod_shard_range(db.shape[0], r, world, ctypes.byref(b2), ctypes.byref(r2)) == capi.TOD_OK
            bases.append(b2.value)
        wide = np.full(allk.shape, np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
        for r in range(world):
            kk = allk[r].astype(np.uint64)
            ok = allk[r] != EMPTY
            wide[r][ok] = ((kk[ok] >> np.uint64(KEY_ROW_BITS)) << np.uint64(32)) | \
                          (np.uint64(bases[r]) + (kk[ok] & np.uint64((1 << KEY_ROW_BITS) - 1)))
        m64 = np.sort(wide.transpose(1, 0, 2).reshape(q.shape[0], -1), axis=1)[:, :k]
        merged = np.where(m64 == np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(EMPTY),
                          ((m64 >> np.uint64(32)) << np.uint64(KEY_ROW_BITS)) | (m64 & np.uint64(0xFFFFFFFF))).astype(np.uint32)
    else:
        merged = np.sort(allk.transpose(1, 0, 2).reshape(q.shape[0], -1), axis=1)[:, :k]
    np.save(os.path.join(out_dir, "merged_%d.npy" % rank), merged)
    np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array([begin, rows]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k,radius,tie_heavy,local_keys", [(2, 0, False, False), (5, 35, False, False),
                                                           (5, 0, True, False), (5, 0, True, True), (3, 35, False, True)])
def test_two_rank_key_exchange_equals_single_node(tmp_path, k, radius, tie_heavy, local_keys):
    """local_keys: the protocol of a wide database (more than 2^23 rows) — keys local to the rank's segment travel
    unchanged and the merge orders (distance, global row) on 64 bits from the table of segment bases."""
    from oracle import hamming_knn as hk
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), k, radius, tie_heavy, str(tmp_path), local_keys), nprocs=world,
             join=True)
    descs, q = _workload(tie_heavy)
    db, off = hk.concat_objects(descs)
    em, ec = hk.knn_c(q, descs, k, radius)
    exp = np.full((q.shape[0], k), EMPTY, np.uint32)
    for i in range(q.shape[0]):
        for j in range(int(ec[i])):
            g = int(off[em["imgIdx"][i, j]] + em["trainIdx"][i, j])
            exp[i, j] = (int(em["distance"][i, j]) << KEY_ROW_BITS) | g
    merged = [np.load(os.path.join(str(tmp_path), "merged_%d.npy" % r)) for r in range(world)]
    assert (merged[0] == merged[1]).all()                         # every rank ends with the same lists
    assert (merged[0] == exp).all()                               # ... equal to the unsharded exact result
    ranges = [np.load(os.path.join(str(tmp_path), "range_%d.npy" % r)) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[0][0] + ranges[0][1] == ranges[1][0]
    assert ranges[1][0] + ranges[1][1] == db.shape[0]             # contiguous, complete cover


def test_shard_range_properties():
    from tod_b200 import capi
    lib = capi.load()
    b, r = ctypes.c_int64(), ctypes.c_int64()
    for total in (0, 1, 7, 8, 1000000, 1000003):
        for count in (1, 2, 3, 4, 8):
            pos = 0
            for rank in range(count):
                assert lib.tod_shard_range(total, rank, count, ctypes.byref(b), ctypes.byref(r)) == capi.TOD_OK
                assert b.value == pos and r.value >= 0
                pos += r.value
            assert pos == total
    assert lib.tod_shard_range(10, 2, 2, ctypes.byref(b), ctypes.byref(r)) == capi.TOD_ERR_INVALID
    assert lib.tod_pack_key(3, 5) == (3 << KEY_ROW_BITS) | 5
