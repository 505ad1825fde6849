"""Trained-DB snapshot: host-only round trips on CPU, and on the GPU a matcher fed from a
snapshot must answer exactly like one fed object by object."""
import ctypes
import os

import numpy as np
import pytest

from oracle import hamming_knn as hk
from tod_b200 import capi, dbio, synth


def test_snapshot_round_trip(tmp_path):
    descs, points = synth.make_db(4, [300, 0, 1234, 77], seed=5)
    ids = ["obj_%d" % i for i in range(4)]
    ids[2] = "a much longer object id with spaces / 00ff"
    path = os.path.join(str(tmp_path), "db.todb")
    dbio.write_snapshot(path, ids, descs, points)
    rid, rd, rp, spans = dbio.read_snapshot(path)
    assert rid == ids
    for a, b in zip(rd, descs):
        assert a.shape == b.shape and (a == b).all()
    for a, b in zip(rp, points):
        assert a.shape == b.shape and (a == b).all()
    for i, p in enumerate(points):
        if p.shape[0]:
            assert spans[i] == hk.object_span(p)          # same float arithmetic as DescriptorMatcher.cpp:106-121
    # sections are where the header says, 64-byte aligned, descriptors contiguous in imgIdx order
    raw = np.fromfile(path, np.uint8)
    assert bytes(raw[:8]) == b"TODB200\x00"
    off_desc = int(raw[32:40].view("<u8")[0])
    assert off_desc % 64 == 0
    assert (raw[off_desc:off_desc + 300 * 32].reshape(300, 32) == descs[0]).all()


def test_snapshot_rejects_garbage(tmp_path):
    lib = capi.load()
    path = os.path.join(str(tmp_path), "bad.todb")
    h = ctypes.c_void_p()
    assert lib.tod_snapshot_open(path.encode(), ctypes.byref(h)) == capi.TOD_ERR_INVALID      # missing file
    open(path, "wb").write(b"TODB200\x00" + b"\x01" * 100)
    assert lib.tod_snapshot_open(path.encode(), ctypes.byref(h)) == capi.TOD_ERR_PARSE
    descs, points = synth.make_db(2, 50, seed=1)
    good = os.path.join(str(tmp_path), "good.todb")
    dbio.write_snapshot(good, ["a", "b"], descs, points)
    raw = bytearray(open(good, "rb").read())
    open(path, "wb").write(raw[:-7])                                                         # truncated
    assert lib.tod_snapshot_open(path.encode(), ctypes.byref(h)) == capi.TOD_ERR_PARSE
    with pytest.raises(ValueError):
        dbio.write_snapshot(good, ["a"], descs, points)


def test_snapshot_rejects_corrupted_header_offsets(tmp_path):
    """Header offsets that would wrap a u64 sum, point outside the file, or are misaligned must be refused — never
    dereferenced (the validation works by subtraction against the file size)."""
    lib = capi.load()
    descs, points = synth.make_db(1, 40, seed=2)
    good = os.path.join(str(tmp_path), "good.todb")
    dbio.write_snapshot(good, ["a"], descs, points)
    raw = bytearray(open(good, "rb").read())
    size = len(raw)
    # header: magic[8] version u32 n_objects u32 total_rows u64 off_table off_desc off_pts off_ids file_bytes (u64 each)
    OFF = {"total_rows": 16, "off_table": 24, "off_desc": 32, "off_pts": 40, "off_ids": 48}
    cases = [("off_table", 2 ** 64 - 24), ("off_table", size + 8), ("off_table", 65), ("off_table", 2 ** 63),
             ("off_desc", 2 ** 64 - 32), ("off_desc", size + 64), ("off_pts", 2 ** 64 - 16), ("off_pts", 0),
             ("off_ids", 2 ** 64 - 1), ("off_ids", size + 1), ("total_rows", 2 ** 59), ("total_rows", 2 ** 64 - 1),
             ("total_rows", 41)]
    path = os.path.join(str(tmp_path), "bad.todb")
    for field, value in cases:
        bad = bytearray(raw)
        bad[OFF[field]:OFF[field] + 8] = int(value).to_bytes(8, "little")
        open(path, "wb").write(bad)
        h = ctypes.c_void_p()
        assert lib.tod_snapshot_open(path.encode(), ctypes.byref(h)) == capi.TOD_ERR_PARSE, (field, value)
    bad = bytearray(raw)
    bad[12:16] = (2 ** 32 - 1).to_bytes(4, "little")                                         # n_objects
    open(path, "wb").write(bad)
    h = ctypes.c_void_p()
    assert lib.tod_snapshot_open(path.encode(), ctypes.byref(h)) == capi.TOD_ERR_PARSE
    h = ctypes.c_void_p()
    assert lib.tod_snapshot_open(good.encode(), ctypes.byref(h)) == capi.TOD_OK               # the original still opens
    lib.tod_snapshot_close(h)


def test_import_cv_filestorage(tmp_path):
    cv2 = pytest.importorskip("cv2")
    descs, points = synth.make_db(1, 222, seed=9)
    for ext, shape in ((".yml", (1, 222, 3)), (".xml.gz", (222, 1, 3))):
        path = os.path.join(str(tmp_path), "model" + ext)
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
        fs.write("descriptors", descs[0])
        fs.write("points", points[0].reshape(shape))          # 1 x N (training.cpp:158) or N x 1 (transposed on load)
        fs.release()
        d, p = dbio.import_cv_filestorage(path)
        assert (d == descs[0]).all() and (p == points[0]).all()


@pytest.mark.gpu
def test_matcher_from_snapshot_equals_matcher_from_arrays(tmp_path):
    from tod_b200 import DescriptorMatcher
    descs, points = synth.make_db(5, [900, 400, 1, 1500, 333], seed=12)
    ids = ["o%d" % i for i in range(5)]
    path = os.path.join(str(tmp_path), "db.todb")
    dbio.write_snapshot(path, ids, descs, points)
    q, _, _ = synth.make_queries(descs, 400, seed=13)
    a = DescriptorMatcher(k=5, radius=35)
    for i in range(5):
        a.add_object(ids[i], descs[i], points[i])
    a.train()
    b = DescriptorMatcher(k=5, radius=35)
    b.load_snapshot(path)
    b.train()
    ra, rb = a.process(q), b.process(q)
    assert (ra["matches"] == rb["matches"]).all() and (ra["counts"] == rb["counts"]).all()
    assert (ra["matches_3d"] == rb["matches_3d"])[np.arange(5)[None, :] < ra["counts"][:, None]].all()
    assert ra["object_ids"] == rb["object_ids"] == ids and ra["spans"] == rb["spans"]
    em, ec = hk.knn_c(q, descs, 5, 35)
    assert (rb["counts"] == ec).all() and (rb["matches"]["trainIdx"] == em["trainIdx"]).all()
    a.close()
    b.close()
