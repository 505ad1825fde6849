"""ctypes binding of include/tod_b200.h — the binding a maintainer of a Python-hosted pipeline would use, and the one
tests/ and bench.py call the library through.  No compute happens here; there is no CPU fallback: if the shared
library is missing, import fails loudly."""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOD_B200_LIB") or os.path.join(_PKG, "libtod_b200.so")  # override: instrumented builds only

TOD_OK, TOD_ERR_INVALID, TOD_ERR_STATE, TOD_ERR_CUDA, TOD_ERR_LIMIT, TOD_ERR_PARSE = range(6)
TOD_SEARCH_EXACT, TOD_SEARCH_LSH = 0, 1
TOD_KERNEL_AUTO, TOD_KERNEL_POPC, TOD_KERNEL_MMA = 0, 1, 2
TOD_MAX_K = 8
TOD_COMM_ID_BYTES = 128

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
POSE_DTYPE = np.dtype([("R", "<f4", (9,)), ("T", "<f4", (3,)), ("object_index", "<i4"), ("n_inliers", "<i4")])
assert MATCH_DTYPE.itemsize == 16 and KEYPOINT_DTYPE.itemsize == 28 and POSE_DTYPE.itemsize == 56


class MatcherParams(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int32), ("radius", ctypes.c_uint32), ("search_type", ctypes.c_int32),
                ("device", ctypes.c_int32), ("shard_rank", ctypes.c_int32), ("shard_count", ctypes.c_int32),
                ("kernel", ctypes.c_int32), ("ratio_enabled", ctypes.c_int32), ("ratio", ctypes.c_float),
                ("remove_duplicates", ctypes.c_int32), ("frame_keypoints", ctypes.c_int32)]


class GuessParams(ctypes.Structure):
    _fields_ = [("min_inliers", ctypes.c_uint32), ("n_ransac_iterations", ctypes.c_uint32),
                ("sensor_error", ctypes.c_float), ("device", ctypes.c_int32), ("ransac_threshold", ctypes.c_double),
                ("seed", ctypes.c_uint64), ("host_threads", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class OrbParams(ctypes.Structure):
    _fields_ = [("n_levels", ctypes.c_int32), ("scale_factor", ctypes.c_float), ("device", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class TrainerParams(ctypes.Structure):
    _fields_ = [("n_features", ctypes.c_int32), ("n_levels", ctypes.c_int32), ("scale_factor", ctypes.c_float),
                ("device", ctypes.c_int32)]


class TodError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tod_b200 error %d: %s" % (code, msg))
        self.code = code


# Every symbol include/tod_b200.h declares: (name, restype, argtypes).  tests/test_abi.py checks the header and this
# table agree and that the library exports them all.
_P = ctypes.c_void_p
_I32, _I64, _U32, _U64, _F, _D = (ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float,
                                  ctypes.c_double)
SIGNATURES = [
    ("tod_last_error", ctypes.c_char_p, []),
    ("tod_abi_version", ctypes.c_int, []),
    ("tod_kernel_launch_count", _U64, []),
    ("tod_matcher_default_params", None, [ctypes.POINTER(MatcherParams)]),
    ("tod_matcher_params_from_json", ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(MatcherParams)]),
    ("tod_matcher_create", ctypes.c_int, [ctypes.POINTER(MatcherParams), ctypes.POINTER(_P)]),
    ("tod_matcher_destroy", None, [_P]),
    ("tod_matcher_add_object", ctypes.c_int, [_P, ctypes.c_char_p, _P, _P, _I32]),
    ("tod_matcher_clear", ctypes.c_int, [_P]),
    ("tod_matcher_train", ctypes.c_int, [_P]),
    ("tod_matcher_num_objects", _I32, [_P]),
    ("tod_matcher_num_descriptors", _I64, [_P]),
    ("tod_matcher_shard_rows", _I64, [_P]),
    ("tod_matcher_object_id", ctypes.c_char_p, [_P, _I32]),
    ("tod_matcher_span", _F, [_P, _I32]),
    ("tod_matcher_k", _I32, [_P]),
    ("tod_matcher_knn", ctypes.c_int, [_P, _P, _I32, _P, _P, _P]),
    ("tod_matcher_knn_device", ctypes.c_int, [_P, _P, _I32, _P, _P, _P, _P]),
    ("tod_matcher_reserve", ctypes.c_int, [_P, _I32]),
    ("tod_comm_unique_id", ctypes.c_int, [_P]),
    ("tod_matcher_set_comm", ctypes.c_int, [_P, _P]),
    ("tod_matcher_set_exchange", ctypes.c_int, [_P, _I32]),
    ("tod_matcher_exchange_error", _I32, [_P]),
    ("tod_matcher_comm_mode", _I32, [_P]),
    ("tod_matcher_set_stage_timing", None, [_P, _I32]),
    ("tod_matcher_last_exchange_ms", _F, [_P]),
    ("tod_shard_range", ctypes.c_int, [_I64, _I32, _I32, ctypes.POINTER(_I64), ctypes.POINTER(_I64)]),
    ("tod_pack_key", _U32, [_U32, _U32]),
    ("tod_matcher_knn_keys_device", ctypes.c_int, [_P, _P, _I32, _P, _P]),
    ("tod_matcher_merge_device", ctypes.c_int, [_P, _P, _I32, _I32, _P, _P, _P, _P]),
    ("tod_matcher_last_k1_ms", _F, [_P]),
    ("tod_matcher_k1_ms_ago", ctypes.c_float, [_P, _I32]),
    ("tod_matcher_last_kernel", ctypes.c_char_p, [_P]),
    ("tod_snapshot_write", ctypes.c_int, [ctypes.c_char_p, _I32, _P, _P, _P, _P]),
    ("tod_snapshot_open", ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_P)]),
    ("tod_snapshot_close", None, [_P]),
    ("tod_snapshot_num_objects", _I32, [_P]),
    ("tod_snapshot_num_descriptors", _I64, [_P]),
    ("tod_snapshot_object", ctypes.c_int, [_P, _I32, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_P),
                                           ctypes.POINTER(_P), ctypes.POINTER(_I32), ctypes.POINTER(_F)]),
    ("tod_matcher_load_snapshot", ctypes.c_int, [_P, ctypes.c_char_p]),
    ("tod_orb_default_params", None, [ctypes.POINTER(OrbParams)]),
    ("tod_orb_create", ctypes.c_int, [ctypes.POINTER(OrbParams), ctypes.POINTER(_P)]),
    ("tod_orb_destroy", None, [_P]),
    ("tod_orb_describe", ctypes.c_int, [_P, _P, _I32, _I32, _P, _I32, _I32, _P, ctypes.POINTER(_P)]),
    ("tod_orb_detect_and_compute", ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _I32, ctypes.POINTER(_I32), _P,
                                                  ctypes.POINTER(_P)]),
    ("tod_orb_detect_and_compute_masked", ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _I32, _P, _I32,
                                                         ctypes.POINTER(_I32), _P, ctypes.POINTER(_P)]),
    ("tod_trainer_default_params", None, [ctypes.POINTER(TrainerParams)]),
    ("tod_trainer_create", ctypes.c_int, [ctypes.POINTER(TrainerParams), ctypes.POINTER(_P)]),
    ("tod_trainer_destroy", None, [_P]),
    ("tod_trainer_add_observation", ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _I32, _I32, _I32, _P, _P, _P,
                                                   ctypes.POINTER(_I32)]),
    ("tod_trainer_model", ctypes.c_int, [_P, ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(_I64)]),
    ("tod_trainer_num_points", _I64, [_P]),
    ("tod_trainer_clear", ctypes.c_int, [_P]),
    ("tod_orb_read_level", ctypes.c_int, [_P, _I32, _I32, _P, ctypes.POINTER(_I32), ctypes.POINTER(_I32)]),
    ("tod_depth_to_3d", ctypes.c_int, [_I32, _P, _I32, _I32, _I32, _P, _P]),
    ("tod_adjacency_row_words", _I32, [_I32]),
    ("tod_fill_adjacency", ctypes.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _F, _P, _P, _P]),
    ("tod_score_hypotheses", ctypes.c_int, [_I32, _I32, _P, _P, _P, _P, _I32, _P, _D, _P, _P, _P]),
    ("tod_last_stage_ms", _F, []),
    ("tod_clique_find", _I32, [_I32, _P, _I32, _U32, _P, ctypes.POINTER(_I32)]),
    ("tod_clique_gate_small", _I32, [_I32, _P, _I32, _I32, ctypes.POINTER(_I32)]),
    ("tod_gate_search_device", ctypes.c_int, [_I32, _I32, _P, _P, _P, _P]),
    ("tod_rigid_fit", ctypes.c_int, [_P, _P, _P, _I32, _P, _P]),
    ("tod_sample_triples", _I32, [_I32, _P, _P, ctypes.POINTER(_U64), _I32, _P]),
    ("tod_select_inliers", _I32, [_I32, _P, _P, _P, _P, _P]),
    ("tod_guess_default_params", None, [ctypes.POINTER(GuessParams)]),
    ("tod_guess_create", ctypes.c_int, [ctypes.POINTER(GuessParams), ctypes.POINTER(_P)]),
    ("tod_guess_destroy", None, [_P]),
    ("tod_guess_process", ctypes.c_int, [_P, _P, _I32, _P, _I32, _I32, _P, _P, _I32, _P, _P, _I32, _P, _I32,
                                         ctypes.POINTER(_I32), _P, _I32]),
    ("tod_guess_process_batch", ctypes.c_int, [_P, _I32, _P, _P, _P, _I32, _I32, _P, _P, _I32, _P, _P, _I32, _P, _P, _I32,
                                               ctypes.POINTER(_I32), _P, _I32]),
    ("tod_rng_seed", _U64, [_U64, _U32, _U32]),
    ("tod_rng_next", _I32, [ctypes.POINTER(_U64)]),
    ("tod_guess_last_gate_stats", None, [_P, ctypes.POINTER(_I64)]),
    ("tod_guess_last_k5_stats", None, [_P, ctypes.POINTER(_I64)]),
    ("tod_guess_last_traffic", None, [_P, ctypes.POINTER(_D), ctypes.POINTER(_D), ctypes.POINTER(_I64),
                                      ctypes.POINTER(_I64)]),
    ("tod_guess_last_profile", None, [_P, ctypes.POINTER(_D)]),
    ("tod_guess_last_stats", None, [_P, ctypes.POINTER(_F), ctypes.POINTER(_F), ctypes.POINTER(_I64),
                                    ctypes.POINTER(_I32)]),
]

_lib = None


def load():
    """Load libtod_b200.so (built by tod_b200._build / __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, res, args in SIGNATURES:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


_nccl_preloaded = False


def preload_nccl():
    """libtod_b200.so binds NCCL with dlopen("libnccl.so.2") at first use, which returns the copy already mapped in
    the process if there is one.  A Python process that will ALSO import torch must end up with torch's bundled NCCL
    (torch needs symbols of its own, newer version), so map that copy first when it exists — found through the
    `nvidia.nccl` wheel, without importing torch.  No-op when the wheel is absent (the system library is used)."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            for d in spec.submodule_search_locations:
                path = os.path.join(d, "lib", "libnccl.so.2")
                if os.path.exists(path):
                    ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
                    return
    except Exception:
        pass


def check(rc):
    if rc != TOD_OK:
        raise TodError(rc, load().tod_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def adjacency_row_words(n):
    return int(load().tod_adjacency_row_words(int(n)))
