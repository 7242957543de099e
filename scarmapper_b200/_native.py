"""
_native.py — ctypes binding of libscar_b200.so (include/scar_b200.h).

There is no fallback: if the shared library is missing or cannot be loaded, importing the
compute path raises.  Build it with `python -m scarmapper_b200.build` (or __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libscar_b200.so")

SCAR_N_COUNTERS = 16
SCAR_PHASE_BINS = 34
SAMPLE_NONE = 0xFFFF

RECORD_DTYPE = np.dtype([
    ("status", "u1"), ("flags", "u1"), ("sample", "<u2"),
    ("clj", "<u2"), ("crj", "<u2"), ("tlj", "<u2"), ("trj", "<u2"),
    ("hr_n", "<u2"), ("phase_r1", "i1"), ("phase_r2", "i1"),
])
ENTRY_DTYPE = np.dtype([
    ("h1", "<u8"), ("h2", "<u8"), ("first_idx", "<u8"), ("count", "<u4"), ("sample", "<u2"), ("pad", "<u2"),
])
assert RECORD_DTYPE.itemsize == 16 and ENTRY_DTYPE.itemsize == 32

ST_UNASSIGNED, ST_FILTERED, ST_NO_JUNCTION, ST_NO_CUT, ST_N_INSERTION, ST_SCAR = range(6)
FL_LEFT_DEL, FL_RIGHT_DEL, FL_INSERTION, FL_MICROHOM, FL_HR, FL_CUTWINDOW = 1, 2, 4, 8, 16, 32
(CT_ASSIGNED, CT_PASSING, CT_LEFT_DEL, CT_RIGHT_DEL, CT_INSERTION, CT_MICROHOM, CT_NO_JUNCTION, CT_NO_CUT,
 CT_FILTERED, CT_HR_FIRST, CT_HR_MORE, CT_N_INSERTION, CT_SCAR) = range(13)
PLATFORMS = {"Illumina": 0, "TruSeq": 1, "Ramsden": 2}

ERRORS = {-1: "SCAR_ERR_CUDA", -2: "SCAR_ERR_INVALID", -3: "SCAR_ERR_UNSUPPORTED", -4: "SCAR_ERR_TABLE_FULL",
          -5: "SCAR_ERR_NOMEM", -6: "SCAR_ERR_HASH_COLLISION"}


ERR_NOMEM = -5


class ScarError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("{} ({}): {}".format(ERRORS.get(code, "SCAR_ERR"), code, text))
        self.code = code


class Ragged(C.Structure):
    """struct scar_ragged"""
    _fields_ = [("bytes", C.c_void_p), ("offsets", C.c_void_p), ("lengths", C.c_void_p), ("stride", C.c_int64)]


class Config(C.Structure):
    """struct scar_config"""
    _fields_ = [
        ("device", C.c_int32), ("platform", C.c_int32), ("mismatch", C.c_int32), ("n_targets", C.c_int32),
        ("target_bytes", C.c_void_p), ("target_off", C.c_void_p), ("target_cutsite", C.c_void_p),
        ("n_samples", C.c_int32), ("sample_target", C.c_void_p),
        ("left_index_bytes", C.c_void_p), ("left_index_off", C.c_void_p),
        ("right_index_bytes", C.c_void_p), ("right_index_off", C.c_void_p),
        ("phase_off", C.c_void_p),
        ("phase_r1_bytes", C.c_void_p), ("phase_r1_off", C.c_void_p),
        ("phase_r2_bytes", C.c_void_p), ("phase_r2_off", C.c_void_p),
        ("hr_donor", C.c_void_p), ("hr_len", C.c_int32),
        ("n_limit", C.c_double), ("min_length", C.c_int32), ("max_read_len", C.c_int32),
        ("table_capacity", C.c_int64), ("cutwindow_bytes", C.c_void_p),
    ]


class FrequencyIn(C.Structure):
    """struct scar_frequency_in"""
    _fields_ = [("n", C.c_int64), ("counts", C.c_void_p), ("first", C.c_void_p), ("reads", C.c_void_p), ("read_off", C.c_void_p),
                ("pattern", C.c_void_p), ("region", C.c_void_p), ("region_len", C.c_int32), ("cutsite", C.c_int32),
                ("reverse_comp", C.c_int32), ("passing", C.c_int64), ("no_cut", C.c_int64), ("pattern_threshold", C.c_double)]


class FastqRun(C.Structure):
    """struct scar_fastq_run"""
    _fields_ = [("n_records", C.c_int64), ("text_bytes", C.c_int64), ("text", C.c_void_p), ("seq_off", C.c_void_p),
                ("seq_len", C.c_void_p), ("records", C.c_void_p), ("max_seq_len", C.c_int32), ("source_kind", C.c_int32),
                ("threads", C.c_int32), ("reserved", C.c_int32), ("wait_seconds", C.c_double), ("total_seconds", C.c_double),
                ("owner", C.c_void_p)]


# every symbol include/scar_b200.h declares: (restype, argtypes)
_P = C.c_void_p
_RP = C.POINTER(Ragged)
PROTOTYPES = {
    "scar_last_error": (C.c_char_p, []),
    "scar_version": (C.c_int, []),
    "scar_cutsite_search": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "scar_ctx_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "scar_ctx_destroy": (None, [_P]),
    "scar_cells_per_read": (C.c_int, [_P]),
    "scar_ctx_info": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "scar_window_counts": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "scar_k3_plan": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "scar_pack_dev": (C.c_int, [_P, _RP, C.c_int64, _P, _P, _P]),
    "scar_demux_dev": (C.c_int, [_P, _RP, _RP, _RP, C.c_int64, _P, _P]),
    "scar_classify_dev": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, _P]),
    "scar_aggregate_dev": (C.c_int, [_P, _RP, _P, C.c_int64, C.c_uint64, _P]),
    "scar_classify_raw_dev": (C.c_int, [_P, _RP, _P, C.c_int64, C.c_uint64, _P, C.c_int32, _P]),
    "scar_fused_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "scar_run_dev": (C.c_int, [_P, _RP, _RP, _RP, C.c_int64, C.c_uint64, _P, _P, _P, _P, _P]),
    "scar_process_host": (C.c_int, [_P, _RP, _RP, _RP, C.c_int64, C.c_uint64, _P]),
    "scar_counters_read": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "scar_phase_hist_read": (C.c_int, [_P, _P]),
    "scar_table_size": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "scar_table_read": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "scar_verify_dev": (C.c_int, [_P, _RP, _P, C.c_int64, C.c_uint64, C.POINTER(C.c_int64), _P]),
    "scar_reset": (C.c_int, [_P]),
    "scar_table_export_dev": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64), _P]),
    "scar_table_merge_dev": (C.c_int, [_P, _P, C.c_int64, _P]),
    "scar_table_export_parts_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P]),
    "scar_table_export_p2p_dev": (C.c_int, [_P, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64, _P]),
    "scar_table_clear_dev": (C.c_int, [_P, _P]),
    "scar_table_merge_parts_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P]),
    "scar_reset_dev": (C.c_int, [_P, _P]),
    "scar_ctx_set_verify": (C.c_int, [_P, C.c_int32]),
    "scar_device_memory": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "scar_ctx_set_sm_limit": (C.c_int, [_P, C.c_int32]),
    "scar_launch_count": (C.c_int64, [_P]),
    "scar_accumulator_words": (C.c_int64, [_P]),
    "scar_accumulators_export_dev": (C.c_int, [_P, _P, _P]),
    "scar_accumulators_import_dev": (C.c_int, [_P, _P, _P]),
    "scar_fastq_index": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, _P, _P, _P,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "scar_process_fastq_host": (C.c_int, [_P, _P, C.c_int64, C.c_uint64, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "scar_fastq_records_fetch": (C.c_int, [_P, _P, C.c_int64]),
    "scar_raw_rows_host": (C.c_int, [_P, _RP, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64,
                                     C.POINTER(C.c_int64)]),
    "scar_raw_rows_indexed_host": (C.c_int, [_P, _RP, _P, C.c_int64, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64,
                                             C.POINTER(C.c_int64)]),
    "scar_demux_fastq_rows_host": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int64, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "scar_stored_gather_host": (C.c_int, [_RP, _P, C.c_int64, C.c_int32, _P, _P, C.c_int64, _P]),
    "scar_homeology_batch": (C.c_int, [C.c_int32, _RP, _RP, _P, _P, _P, _P, _P, _RP, _RP, C.c_int32, C.c_int64, _P]),
    "scar_frequency_rows_host": (C.c_int, [C.POINTER(FrequencyIn), C.POINTER(C.c_int64), _P, _P, _P, _P, _P, C.c_int64,
                                          C.POINTER(C.c_int64), _P]),
    "scar_frequency_rows_fetch": (C.c_int, [_P, C.c_int64]),
    "scar_process_fastq_file": (C.c_int, [_P, C.c_char_p, C.c_int64, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(FastqRun)]),
    "scar_fastq_run_release": (None, [C.POINTER(FastqRun)]),
    "scar_fastq_gather_host": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int64, _P]),
    "scar_gunzip_host": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, _P, C.c_int64, C.POINTER(C.c_int64), _P]),
    "scar_table_reserve": (C.c_int, [_P, C.c_int64]),
    "scar_alignment_batch": (C.c_int, [C.c_int32, _RP, _RP, _P, _P, _P, C.c_int64, C.c_int32, _P, _P]),
    "scar_match_maker_batch": (C.c_int, [C.c_int32, _RP, _RP, C.c_int64, _P]),
}

_lib = None


def lib():
    """Load libscar_b200.so and bind every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libscar_b200.so is not built ({}); run `python -m scarmapper_b200.build`. "
                              "There is no CPU fallback.".format(LIB_PATH))
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise ScarError(rc, lib().scar_last_error().decode("utf-8", "replace"))
