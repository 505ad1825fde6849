"""Trained-DB I/O for the detection hot path.

* write_snapshot / read_snapshot: the flat mmap-able file of include/tod_b200.h (tod_snapshot_*), which replaces the
  per-document CouchDB round trips of DescriptorMatcher::parameter_callback (src/detection/DescriptorMatcher.cpp:60-129
  of the reference).
* import_cv_filestorage: reads one model's `descriptors` (N x 32 CV_8U, training.cpp:157) and `points` (1 x N or
  N x 1 CV_32FC3, training.cpp:158; transposed like DescriptorMatcher.cpp:82-86) from an OpenCV FileStorage
  YAML / XML(.gz) file — the cv::Mat serialisation.  object_recognition_core (un-vendored) stores the two attachments
  of a TOD model document this way; check the attachment names against a real DB before relying on it.

Pure plumbing: no compute happens here."""
import ctypes

import numpy as np

from . import capi


def write_snapshot(path, object_ids, descriptors, points):
    """object_ids: list of str; descriptors[o]: rows x 32 u8; points[o]: rows x 3 f32."""
    lib = capi.load()
    n = len(object_ids)
    ds = [np.ascontiguousarray(d, np.uint8).reshape(-1, 32) for d in descriptors]
    ps = [np.ascontiguousarray(p, np.float32).reshape(-1, 3) for p in points]
    if not (len(ds) == len(ps) == n) or any(d.shape[0] != p.shape[0] for d, p in zip(ds, ps)):
        raise ValueError("object_ids, descriptors and points disagree")
    ids = (ctypes.c_char_p * n)(*[str(s).encode() for s in object_ids])
    dp = (ctypes.c_void_p * n)(*[d.ctypes.data for d in ds])
    pp = (ctypes.c_void_p * n)(*[p.ctypes.data for p in ps])
    rows = np.array([d.shape[0] for d in ds], np.int32)
    capi.check(lib.tod_snapshot_write(str(path).encode(), n, ids, dp, pp, capi._ptr(rows)))


def read_snapshot(path):
    """Returns (object_ids, descriptors, points, spans); arrays are copies (the mapping is closed on return)."""
    lib = capi.load()
    h = ctypes.c_void_p()
    capi.check(lib.tod_snapshot_open(str(path).encode(), ctypes.byref(h)))
    try:
        ids, ds, ps, spans = [], [], [], []
        for o in range(lib.tod_snapshot_num_objects(h)):
            oid, d, p = ctypes.c_char_p(), ctypes.c_void_p(), ctypes.c_void_p()
            rows, span = ctypes.c_int32(), ctypes.c_float()
            capi.check(lib.tod_snapshot_object(h, o, ctypes.byref(oid), ctypes.byref(d), ctypes.byref(p),
                                               ctypes.byref(rows), ctypes.byref(span)))
            n = rows.value
            ids.append(oid.value.decode())
            ds.append(np.ctypeslib.as_array(ctypes.cast(d, ctypes.POINTER(ctypes.c_uint8)), (n, 32)).copy()
                      if n else np.zeros((0, 32), np.uint8))
            ps.append(np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), (n, 3)).copy()
                      if n else np.zeros((0, 3), np.float32))
            spans.append(span.value)
        return ids, ds, ps, np.array(spans, np.float32)
    finally:
        lib.tod_snapshot_close(h)


def import_cv_filestorage(path, descriptors_key="descriptors", points_key="points"):
    """One model from an OpenCV FileStorage file -> (descriptors rows x 32 u8, points rows x 3 f32)."""
    import cv2
    fs = cv2.FileStorage(str(path), cv2.FILE_STORAGE_READ)
    if not fs.isOpened():
        raise IOError("cannot open %s" % path)
    try:
        d = fs.getNode(descriptors_key).mat()
        p = fs.getNode(points_key).mat()
    finally:
        fs.release()
    if d is None or p is None:
        raise ValueError("%s holds no '%s' / '%s' matrices" % (path, descriptors_key, points_key))
    d = np.ascontiguousarray(d, np.uint8)
    if d.ndim != 2 or d.shape[1] != 32:
        raise ValueError("descriptors must be N x 32 CV_8U (256-bit ORB), got %r" % (d.shape,))
    p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)     # 1 x N x 3 or N x 1 x 3 -> N x 3
    if p.shape[0] != d.shape[0]:
        raise ValueError("descriptors and points disagree: %d vs %d" % (d.shape[0], p.shape[0]))
    return d, p
