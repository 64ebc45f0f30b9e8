"""ctypes binding of libkmx_sm100.so (include/kmx.h).  No fallback: a missing library or a
missing sm_100 device is a hard error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkmx_sm100.so")

KMX_OK, KMX_ERR_ARG, KMX_ERR_CUDA, KMX_ERR_FORMAT, KMX_ERR_NOMEM, KMX_ERR_STATE = range(6)
KEY_KMER, KEY_HASH = 0, 1
FMT_COUNT, FMT_PA, FMT_BF, FMT_BFT = 0, 1, 2, 3
PROF_KINDS = ["fq_index", "s1_superk", "hash_hist", "hash_emit", "expand", "radix_sort", "rle", "merge", "transpose", "fill", "exchange"]


class KmxParams(C.Structure):
    _fields_ = [("kmer_size", C.c_uint32), ("minim_size", C.c_uint32), ("nb_partitions", C.c_uint32),
                ("key_kind", C.c_uint32), ("window_bits", C.c_uint64),
                ("repart_table", C.POINTER(C.c_uint16)), ("nb_samples", C.c_uint32), ("reserved", C.c_uint32)]


class KmxMergeParams(C.Structure):
    _fields_ = [("soft_min", C.POINTER(C.c_uint32)), ("recurrence_min", C.c_uint32), ("share_min", C.c_uint32),
                ("format", C.c_uint32), ("emit_all", C.c_uint32)]


class KmxMergeResult(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("row_bytes", C.c_uint64), ("n_union", C.c_uint64)]


# every symbol include/kmx.h declares -> (restype, argtypes)
_vp, _u64, _u32, _sz, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_int
SYMBOLS = {
    "kmx_create": (_i, [_i, C.POINTER(KmxParams), C.POINTER(_vp)]),
    "kmx_destroy": (None, [_vp]),
    "kmx_last_error": (C.c_char_p, [_vp]),
    "kmx_launch_count": (_u64, [_vp]),
    "kmx_sync": (_i, [_vp]),
    "kmx_stream": (_vp, [_vp]),
    "kmx_superk_begin": (_i, [_vp]),
    "kmx_superk_push_fastq": (_i, [_vp, _vp, _sz, _i]),
    "kmx_superk_push_reads": (_i, [_vp, _vp, C.POINTER(_u64), _sz]),
    "kmx_superk_end": (_i, [_vp, C.POINTER(_u64)]),
    "kmx_count_sample": (_i, [_vp, _u32, _u32]),
    "kmx_run_samples": (_i, [_vp, _u32, C.POINTER(C.c_void_p), C.POINTER(_sz), _i, C.POINTER(_u32), C.POINTER(_u32), _u32, C.POINTER(_u64)]),
    "kmx_dist_unique_id": (_i, [_vp]),
    "kmx_dist_init": (_i, [_vp, _i, _i, _u32, _vp]),
    "kmx_dist_owner": (_i, [_vp, _u32, _i]),
    "kmx_dist_set_lanes": (_i, [_vp, _u32]),
    "kmx_dist_run_batch": (_i, [_vp, _u32, C.POINTER(C.c_void_p), C.POINTER(_sz), _i, C.POINTER(_u32), _u32, _u32, C.POINTER(_u64)]),
    "kmx_minimizer_load_enable": (_i, [_vp, _i]),
    "kmx_minimizer_load_get": (_i, [_vp, _vp]),
    "kmx_lanes": (_i, [_vp, _u32]),
    "kmx_lane_superk_begin": (_i, [_vp, _u32]),
    "kmx_lane_superk_push_fastq": (_i, [_vp, _u32, _vp, _sz, _i]),
    "kmx_lane_superk_push_reads": (_i, [_vp, _u32, _vp, _vp, _sz]),
    "kmx_lane_superk_end": (_i, [_vp, _u32, C.POINTER(_u64)]),
    "kmx_lane_count_sample": (_i, [_vp, _u32, _u32, _u32]),
    "kmx_lane_count_sample_hist": (_i, [_vp, _u32, _u32, _u32, _u32, _u32, _vp]),
    "kmx_dist_run_samples": (_i, [_vp, _u32, C.POINTER(C.c_void_p), C.POINTER(_sz), _i, C.POINTER(_u32), C.POINTER(_u64)]),
    "kmx_counts_size": (_i, [_vp, _u32, _u32, C.POINTER(_u64)]),
    "kmx_counts_get": (_i, [_vp, _u32, _u32, _vp, _vp]),
    "kmx_counts_put": (_i, [_vp, _u32, _u32, _vp, _vp, _u64]),
    "kmx_counts_vector": (_i, [_vp, _u32, _u32, _vp]),
    "kmx_merge_partition": (_i, [_vp, _u32, C.POINTER(KmxMergeParams), C.POINTER(KmxMergeResult)]),
    "kmx_merge_get": (_i, [_vp, _vp, _vp, _vp]),
    "kmx_merge_body_device": (_vp, [_vp]),
    "kmx_transpose_bits": (_i, [_vp, _vp, _u64, _u64, _vp]),
    "kmx_synth_fastq": (_i, [_vp, _u64, _u32, _u64, _u64, _u32, _u64, C.c_double, C.c_double, _i, _vp]),
    "kmx_dev_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "kmx_dev_free": (_i, [_vp, _vp]),
    "kmx_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "kmx_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "kmx_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "kmx_host_free": (_i, [_vp]),
    "kmx_reset": (_i, [_vp]),
    "kmx_set_merge_output": (_i, [_vp, _vp, _sz]),
    "kmx_profile_enable": (_i, [_vp, _i]),
    "kmx_profile_reset": (_i, [_vp]),
    "kmx_profile_get": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "kmx_device_bytes": (_u64, [_vp]),
    "kmx_stat": (_u64, [_vp, _i]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run ./build.sh (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
