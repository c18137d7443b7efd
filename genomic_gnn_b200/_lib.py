"""ctypes binding of libgg_b200.so (include/gg_b200.h).

The CUDA library IS the product: if it is missing and cannot be built this
module raises -- there is no CPU or PyTorch fallback behind it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgg_b200.so")

GG_OK = 0
GG_ABI_VERSION = 3
GG_SEP = 10
GG_MAX_K = 13
GG_KF_BLOCK = 1024
GG_NO_POS = 0xFFFFFFFFFFFFFFFF
ST_BAD_BYTES, ST_BRIDGES, ST_CURSOR, ST_REMAINING, ST_NONZERO, ST_WORDS = 0, 1, 2, 3, 4, 8
KF_IMPL = {"auto": 0, "dp4a": 1, "tcgen05": 2, "naive": 3, "tcgen05x2": 4}
KF_DEDUP_MAX_NODES = 1 << 20  # shared-memory bitmap over node space (gg_kf_dedup_*)


class KfParams(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("d", C.c_int32),
        ("iblock_begin", C.c_int32),
        ("iblock_end", C.c_int32),
        ("p_max", C.c_int32),
        ("fast_shift", C.c_int32),
        ("impl", C.c_int32),
        ("d_valid", C.c_int32),
        ("hist_m", C.c_int32),
    ]


class KfDedupExtra(C.Structure):
    _fields_ = [
        ("class_hist", C.c_void_p),
        ("hist_class_begin", C.c_int64),
        ("hist_class_end", C.c_int64),
        ("tie_out", C.c_void_p),
        ("tie_num", C.c_int64),
        ("tie_den", C.c_int64),
        ("row_lo", C.c_int64),
        ("row_hi", C.c_int64),
    ]


GG_MAX_PEERS = 16


class Peers(C.Structure):
    _fields_ = [("n", C.c_int), ("ptr", C.c_void_p * GG_MAX_PEERS)]


_P, _I64, _I32, _SZ, _DBL = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_double

# name -> (restype, argtypes); mirrors include/gg_b200.h one to one
SIGNATURES = {
    "gg_abi_version": (_I32, []),
    "gg_last_error": (C.c_char_p, []),
    "gg_build_id": (C.c_char_p, []),
    "gg_stream_padded_bytes": (_SZ, [_I64]),
    "gg_hist_bins": (_SZ, [_I32]),
    "gg_count_reset": (_I32, [_I32, _P, _P, _P, _P]),
    "gg_count_kp1mers": (_I32, [_P, _I64, _I64, _I32, _I64, _P, _P, _I64, _P, _P]),
    "gg_count_smem_workspace_bytes": (_SZ, [_I64, _I32]),
    "gg_count_kp1mers_smem": (_I32, [_P, _I64, _I64, _I32, _I64, _P, _P, _I64, _P, _P, _SZ, _P]),
    "gg_count_part_workspace_bytes": (_SZ, [_I64, _I32]),
    "gg_count_kp1mers_part": (_I32, [_P, _I64, _I64, _I32, _I64, _P, _P, _I64, _P, _P, _SZ, _P]),
    "gg_count_merge": (_I32, [_P, _I32, _I64, _I32, _P, _P, _P]),
    "gg_first_occurrence": (_I32, [_P, _I64, _I64, _I32, _I64, _P, _P, _P, _P]),
    "gg_db_workspace_bytes": (_SZ, [_I32, _I64]),
    "gg_db_build": (_I32, [_P, _P, _P, _I64, _I32, _I32, _I32, _DBL, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _SZ, _I32, _P]),
    "gg_kf_counts": (_I32, [_P, _I64, _I32, _I32, _P, _P, _P, _P]),
    "gg_kf_counts_upto": (_I32, [_P, _I64, _P, _I32, _I32, _P, _P, _P, _P]),
    "gg_kf_features_f32": (_I32, [_P, _I64, _I32, _P, _I32, _P, _P, _P]),
    "gg_kf_rowcnt_elems": (_SZ, [_I64, _I32, _I32]),
    "gg_kf_emit": (_I32, [C.POINTER(KfParams), _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "gg_kf_class_hist_packed": (_I32, [_P, _I64, _P, _I32, _P, _P]),
    "gg_kf_select_packed_workspace_bytes": (_SZ, [_I64]),
    "gg_kf_select_packed": (_I32, [_P, _I64, _P, _I32, _P, C.c_uint64, _P, _P, _P, _P, _P, _I64, _P, _SZ, _P]),
    "gg_kf_place_workspace_bytes": (_SZ, [_I64, _I32, _I32]),
    "gg_kf_place": (_I32, [C.POINTER(KfParams), _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "gg_kf_expand": (_I32, [_P, _I32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _P, _P, _P, _P, _P, _P, _P]),
    "gg_kf_finalize_workspace_bytes": (_SZ, [_I64, _I32, _I32, _I64]),
    "gg_kf_finalize": (_I32, [C.POINTER(KfParams), _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gg_kf_dedup_classes_workspace_bytes": (_SZ, [_I64]),
    "gg_kf_dedup_classes": (_I32, [_P, _I32, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "gg_kf_dedup_adj_words": (_SZ, [_I64]),
    "gg_kf_dedup_score": (_I32, [_P, _I32, _I32, _P, _P, _P, _I64, _P, _I32, _P, C.POINTER(KfDedupExtra), _P]),
    "gg_kf_dedup_count_workspace_bytes": (_SZ, [_I64, _I32, _I32]),
    "gg_kf_dedup_count": (_I32, [_I64, _I32, _I32, _P, _P, _P, _I64, _P, _P, _P, _SZ, _P]),
    "gg_kf_dedup_place": (_I32, [_I64, _I32, _I32, _P, _P, _P, _I64, _P, _P, _I32, _I32, _P, _I64, _P, _SZ, _P]),
    "gg_kf_class_hist": (_I32, [_P, _P, _P, _I64, _P, _I32, _P, _P]),
    "gg_count_pairs_stride": (_I32, [_P, _P, _P, _I64, _I32, _I32, _I64, _P, _I64, _P, _P]),
    "gg_db_prune_workspace_bytes": (_SZ, [_I64]),
    "gg_db_prune_quantile": (_I32, [_P, _P, _P, _I64, _DBL, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gg_kmer_index": (_I32, [_P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _P, _P, _P]),
    "gg_host_features_f32": (_I32, [_P, _I64, _P, _P, _I32]),
    "gg_host_features_f32_start": (_P, [_P, _I64, _P, _P, _I32, _I32, _P]),
    "gg_host_job_wait": (_I32, [_P]),
    "gg_spmm_csr_f32": (_I32, [_P, _P, _P, _I64, _P, _I32, _P, _P]),
    "gg_rw_pairs": (_I32, [_P, _P, _P, _I64, _P, _I64, _I32, _I32, _I32, _DBL, _DBL, C.c_uint64, _I32, _P, _P, _P, _P, _I64,
                           _P, _P]),
    "gg_knn_set_impl": (_I32, [_I32]),
    "gg_knn_count": (_I32, [_P, _I32, _I64, _I64, _I64, _P, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "gg_knn_plan_workspace_bytes": (_SZ, [_I64]),
    "gg_knn_plan": (_I32, [_I64, _I64, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _SZ, _P]),
    "gg_kf_emit_sparse": (_I32, [_P, _SZ, _I32, _I64, _I32, _I32, _P, _P, _P, _I64, _P, _P, _P]),
    "gg_knn_sparse_eligible": (_I32, [_I32, _I64, _I32, _I32]),
    "gg_knn_sparse_index_bytes": (_SZ, [_I64, _I32]),
    "gg_knn_sparse_index": (_I32, [_P, _I32, _I64, _P, _P, _I32, _I32, _P, _SZ, _P, _P]),
    "gg_knn_sparse_count": (_I32, [_P, _SZ, _I32, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _P, _P]),
    "gg_knn_sparse_write": (_I32, [_P, _SZ, _I32, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P,
                                   _P, _P, _P, _I64, _P]),
    "gg_knn_write_scratch_words": (_SZ, [_I32, _I64, _I32, _I32, _I32, _I64, _I64]),
    "gg_knn_write": (_I32, [_P, _I32, _I64, _I64, _I64, _P, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P,
                            _I64, _P, _P, _P, _I64, _P]),
    "gg_peer_push": (_I32, [_P, _I64, C.POINTER(Peers), _P]),
    "gg_peer_signal": (_I32, [C.POINTER(Peers), _I32, C.c_uint32, _P]),
    "gg_peer_wait": (_I32, [_P, _I32, _I32, _I32, C.c_uint32, _P, _P]),
    "gg_probe_smem_atomics": (_I32, [_I32, _I32, _P, C.POINTER(C.c_double), _P]),
    "gg_probe_tensor_i8": (_I32, [_I32, _I32, _P, C.POINTER(C.c_double), _P]),
    "gg_kf_select_workspace_bytes": (_SZ, [_I64]),
    "gg_kf_select": (_I32, [_P, _P, _P, _P, _I64, _P, _I32, _P, _I64, _P, _P, _P, _P, _P, _P, _SZ, _P]),
}

_lib = None


class GGError(RuntimeError):
    pass


def load():
    """Load (building first if the .so is absent) and type the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    # build() is a no-op when the .so carries the digest of the sources next to it; a library built from older
    # sources (stale signatures behind an unchanged symbol table) is rebuilt, never loaded
    try:
        _build.build()
    except Exception as exc:  # loud: no fallback path exists
        raise GGError(f"libgg_b200.so is missing or stale and could not be built: {exc}") from exc
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.gg_abi_version() != GG_ABI_VERSION:
        raise GGError("libgg_b200.so ABI version mismatch")
    _lib = lib
    return lib


_pyhost = None


def load_pyhost():
    """libgg_pyhost.so (csrc/gg_pyhost.c): list[str] -> byte stream.  PyDLL: its calls take Python objects and keep
    the GIL until they release it themselves."""
    global _pyhost
    if _pyhost is None:
        from . import build as _build

        lib = C.PyDLL(_build.build_pyhost())
        lib.gg_py_reads_scan.restype = C.c_void_p
        lib.gg_py_reads_scan.argtypes = [C.py_object, C.c_int, C.POINTER(C.c_int64)]
        lib.gg_py_reads_copy.restype = C.c_int
        lib.gg_py_reads_copy.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        lib.gg_py_reads_offset.restype = C.c_int64
        lib.gg_py_reads_offset.argtypes = [C.c_void_p, C.c_int64]
        lib.gg_py_reads_free.restype = None
        lib.gg_py_reads_free.argtypes = [C.c_void_p]
        _pyhost = lib
    return _pyhost


def check(rc, what):
    if rc != GG_OK:
        msg = load().gg_last_error().decode("utf-8", "replace")
        raise GGError(f"{what} failed with status {rc}: {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None (a ready-made c_void_p passes through)."""
    if t is None or isinstance(t, C.c_void_p):
        return t
    return C.c_void_p(t.data_ptr())
