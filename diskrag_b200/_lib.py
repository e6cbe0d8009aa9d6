"""ctypes binding of libdiskrag_b200.so (the C ABI declared in include/diskrag_b200.h).

The library is the product: if it is missing, or no CUDA device is present when a compute call is
made, this module raises — there is no CPU path to fall back to.
"""
import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
import os as _os
# DISKRAG_B200_LIB: load another build of the same library (A/B runs of kernel variants, scripts/build_variants.py)
LIB_PATH = Path(_os.environ["DISKRAG_B200_LIB"]) if _os.environ.get("DISKRAG_B200_LIB") else _HERE / "libdiskrag_b200.so"

DR_DIST_PQ, DR_DIST_EXACT, DR_DIST_COSINE = 0, 1, 2
DR_ADC_SEQ, DR_ADC_TREE = 0, 1
DR_LUT_F32, DR_LUT_U8, DR_LUT_U8_TC = 0, 1, 2
DR_ST_VISITED_OVERFLOW, DR_ST_TIE_OVERFLOW = 1, 2


class SearchParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("L", C.c_int32), ("W", C.c_int32), ("dist", C.c_int32),
                ("adc_order", C.c_int32), ("rerank", C.c_int32), ("sqrt_out", C.c_int32),
                ("hash_cap", C.c_int32), ("chunk", C.c_int32), ("threads", C.c_int32), ("lut_fmt", C.c_int32),
                ("prefetch", C.c_int32), ("start_plus1", C.c_int32), ("w_after_empty", C.c_int32), ("ignore_deleted", C.c_int32)]


_vp, _i32, _i64, _u64, _f32, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
_PP = C.POINTER

# name -> (restype, argtypes); mirrors include/diskrag_b200.h one to one
SIGNATURES = {
    "dr_abi_version": (C.c_int, []),
    "dr_last_error": (C.c_char_p, []),
    "dr_device_count": (C.c_int, [_PP(C.c_int)]),
    "dr_device_info": (C.c_int, [C.c_int, _PP(C.c_int), _PP(C.c_int), _PP(_i64)]),
    "dr_index_create_from_records": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _i64, C.c_int, _PP(_vp)]),
    "dr_index_create": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i64, C.c_int, _PP(_vp)]),
    "dr_index_create_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i64, C.c_int, _PP(_vp)]),
    "dr_index_destroy": (C.c_int, [_vp]),
    "dr_index_info": (C.c_int, [_vp, _PP(_i64), _PP(_i32), _PP(_i32), _PP(_i32), _PP(_i64), _PP(C.c_int)]),
    "dr_index_export_records": (C.c_int, [_vp, _vp]),
    "dr_search_batch": (C.c_int, [_vp, _vp, _i64, _PP(SearchParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "dr_search_batch_dev": (C.c_int, [_vp, _vp, _i64, _PP(SearchParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "dr_beam_search_c": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp]),
    "dr_launch_count": (_i64, []),
    "dr_search_kernel_timing": (C.c_int, [_vp, C.c_int, _PP(_dbl), _PP(_i64)]),
    "dr_lut_build": (C.c_int, [_vp, _vp, _i64, _vp]),
    "dr_lut_build_dev": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "dr_pq_lut": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, C.c_int]),
    "dr_pq_lut_u8": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, C.c_int]),
    "dr_pq_train_tensor_cores": (C.c_int, [C.c_int]),
    "dr_pq_train_kmeanspp": (C.c_int, [C.c_int]),
    "dr_pq_train": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _u64, _vp, _PP(_dbl), C.c_int]),
    "dr_pq_train_dev": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _u64, _vp, _PP(_dbl), C.c_int, _vp]),
    "dr_pq_encode": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, C.c_int]),
    "dr_pq_encode_dev": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, C.c_int, _vp]),
    "dr_pq_decode": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, C.c_int]),
    "dr_adc": (C.c_int, [_vp, _vp, _i64, _i32, _vp, C.c_int]),
    "dr_l2sq_batch": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, C.c_int]),
    "dr_dot_batch": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, C.c_int]),
    "dr_cosine_batch": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, C.c_int]),
    "dr_pq_sdc_batch": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, C.c_int]),
    "dr_medoid": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _PP(_i64), C.c_int]),
    "dr_medoid_dev": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _PP(_i64), C.c_int, _vp]),
    "dr_vamana_build": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _f32, _i64, _u64, _vp, _vp, C.c_int]),
    "dr_vamana_build_dev": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _f32, _i64, _u64, _vp, _vp, C.c_int, _vp]),
    "dr_vamana_build_last_truncated": (_i64, []),
    "dr_robust_prune": (C.c_int, [_vp, _vp, _i32, _i32, _f32, _i32, _vp, _PP(_i32), C.c_int]),
    "dr_index_set_deleted": (C.c_int, [_vp, _vp]),
    "dr_index_set_start": (C.c_int, [_vp, _i64]),
    "dr_index_append": (C.c_int, [_vp, _vp, _vp, _i64]),
    "dr_index_patch_rows": (C.c_int, [_vp, _vp, _i64, _vp]),
    "dr_index_patch_vectors": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "dr_index_set_deleted_rows": (C.c_int, [_vp, _vp, _i64, _vp]),
    "dr_topk_merge_dev": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, C.c_int, _vp]),
    "dr_topk_pack_dev": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _i32, _vp, C.c_int, _vp]),
    "dr_topk_merge_keys_dev": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp, C.c_int, _vp]),
    "dr_index_set_peer_route": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _i64]),
    "dr_dev_alloc": (C.c_int, [C.c_int, _i64, _PP(_vp), _vp]),
    "dr_dev_free": (C.c_int, [C.c_int, _vp]),
    "dr_ipc_open": (C.c_int, [C.c_int, _vp, _PP(_vp)]),
    "dr_ipc_close": (C.c_int, [C.c_int, _vp]),
    "dr_dev_memset": (C.c_int, [C.c_int, _vp, C.c_int, _i64, _vp]),
    "dr_dev_upload": (C.c_int, [C.c_int, _vp, _vp, _i64]),
}

_lib = None


class DiskragError(RuntimeError):
    pass


def lib():
    """Load the shared library (built by diskrag_b200/build_ext.py).  Raises if it is absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DiskragError(
                f"{LIB_PATH} is missing: run `python -m diskrag_b200.build_ext` (needs nvcc). "
                "diskrag_b200 has no CPU fallback.")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().dr_last_error().decode("utf-8", "replace")
        if rc == 2:
            raise ValueError(f"{what}: {msg}" if what else msg)
        raise DiskragError(f"{what}: {msg}" if what else msg)


def device_count() -> int:
    n = C.c_int(0)
    lib().dr_device_count(C.byref(n))
    return n.value


def require_gpu():
    if device_count() < 1:
        raise DiskragError("no CUDA device visible; diskrag_b200 runs on the GPU only (no CPU fallback)")


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)
