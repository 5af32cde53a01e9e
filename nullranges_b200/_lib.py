"""ctypes binding of libnrb200.so (include/nrb200.h).

There is NO fallback: if the shared library is missing it is built with nvcc, and if that
fails (or no GPU is present at run time) the caller gets an error, never a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

i32p, i64p, i8p, dp, vp = (C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int8),
                           C.POINTER(C.c_double), C.c_void_p)

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4, -5
UNSEG_ACROSS, UNSEG_WITHIN, SEG_PROP, SEG_NOPROP, PERMUTE_WITHIN, PERMUTE_ACROSS = range(6)
RNG_HOST, RNG_PHILOX = 0, 1
FLAG_FORCE_STREAM = 1
EXCLUDE_DROP, EXCLUDE_TRIM = 0, 1


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("stream", C.c_void_p),
                ("n_devices", C.c_int32), ("devices", C.POINTER(C.c_int32))]


class PlanSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("block_length", C.c_int64), ("n_seg", C.c_int64),
                ("seg_seqid", i32p), ("seg_start", i32p), ("seg_end", i32p), ("seg_state", i32p)]


class Draws(C.Structure):
    _fields_ = [("n_iter", C.c_int64), ("iter0", C.c_int64), ("rng_mode", C.c_int32), ("seed", C.c_uint64),
                ("bait_seqid", i32p), ("bait_start", i32p), ("dst_tile", i32p)]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("n_launches", C.c_int64),
                ("n_hits", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("rows_kernel_ms", C.c_double), ("count_kernel_ms", C.c_double), ("n_windows", C.c_int64)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/nrb200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("nrb200_create", C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
    ("nrb200_n_devices", C.c_int, [vp]),
    ("nrb200_destroy", None, [vp]),
    ("nrb200_last_error", C.c_char_p, [vp]),
    ("nrb200_version", C.c_char_p, []),
    ("nrb200_zero_width_rule", C.c_int, []),
    ("nrb200_set_ranges", C.c_int, [vp, C.c_int64, C.c_int32, i64p, i32p, i32p, i32p, i8p]),
    ("nrb200_set_ranges_begin", C.c_int, [vp, C.c_int64, C.c_int32, i64p, i32p, i32p, i32p, i8p]),
    ("nrb200_set_ranges_end", C.c_int, [vp]),
    ("nrb200_ranges_info", C.c_int, [vp, i32p, i32p, i32p]),
    ("nrb200_set_query", C.c_int, [vp, C.c_int64, i32p, i32p, i32p, i8p]),
    ("nrb200_set_query_ex", C.c_int, [vp, C.c_int64, i32p, i32p, i32p, i8p, C.c_int32, C.c_int32]),
    ("nrb200_set_exclude", C.c_int, [vp, C.c_int64, i32p, i32p, i32p, i8p]),
    ("nrb200_set_exclude_option", C.c_int, [vp, C.c_int32]),
    ("nrb200_set_plan", C.c_int, [vp, C.POINTER(PlanSpec), i64p]),
    ("nrb200_get_tiles", C.c_int, [vp, i32p, i32p, i32p]),
    ("nrb200_run_counts", C.c_int, [vp, C.POINTER(Draws), i64p, i64p, i32p, C.POINTER(Stats)]),
    ("nrb200_run_counts_u16", C.c_int, [vp, C.POINTER(Draws), i64p, i64p, C.POINTER(C.c_uint16), i32p, C.POINTER(Stats)]),
    ("nrb200_run_counts_resident", C.c_int, [vp, C.POINTER(Draws), C.c_int32, C.POINTER(Stats)]),
    ("nrb200_resident_pointers", C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    ("nrb200_run_counts_gathered", C.c_int, [vp, C.POINTER(Draws), C.POINTER(Stats)]),
    ("nrb200_gathered_pointers", C.c_int, [vp, C.c_int32, i32p, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    ("nrb200_moments_reset", C.c_int, [vp]),
    ("nrb200_run_moments", C.c_int, [vp, C.POINTER(Draws), C.c_int64, i32p, i64p, i64p, C.POINTER(Stats)]),
    ("nrb200_moments_fetch", C.c_int, [vp, i64p, i64p, i64p]),
    ("nrb200_set_weights", C.c_int, [vp, dp]),
    ("nrb200_run_weighted", C.c_int, [vp, C.POINTER(Draws), dp, dp, C.POINTER(Stats)]),
    ("nrb200_run_weighted_moments", C.c_int, [vp, C.POINTER(Draws), dp, dp, dp, dp, i64p, C.POINTER(Stats)]),
    ("nrb200_count_observed", C.c_int, [vp, i32p, i64p]),
    ("nrb200_count_observed_ex", C.c_int, [vp, C.c_int32, i32p, i64p]),
    ("nrb200_run_materialize", C.c_int, [vp, C.POINTER(Draws), i64p, C.POINTER(Stats)]),
    ("nrb200_fetch_rows", C.c_int, [vp, C.c_int64, C.c_int64, i32p, i32p, i32p, i32p, i32p]),
    ("nrb200_fetch_dst_tiles", C.c_int, [vp, C.c_int64, C.c_int64, i32p]),
    ("nrb200_measure_l2_sector_rate", C.c_int, [vp, C.c_int64, dp]),
    ("nrb200_rrng_create", vp, [C.c_uint32]),
    ("nrb200_rrng_destroy", None, [vp]),
    ("nrb200_rrng_unif_rand", C.c_double, [vp]),
    ("nrb200_rrng_runif", C.c_int, [vp, C.c_int64, dp, C.c_int64, dp, C.c_int64, dp]),
    ("nrb200_rrng_sample", C.c_int, [vp, C.c_int32, C.c_int64, C.c_int32, dp, i32p]),
    ("nrb200_r_round", C.c_double, [C.c_double]),
    ("nrb200_expand_runs", None, [i64p, C.c_int64, C.c_int32, i32p]),
]

_lib = None


def so_path() -> str:
    return _build.SO


def load():
    """Loads (building first if the .so is missing or stale) and types every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.SO) or (_build.stale() and _build.nvcc_available()):
        _build.build()
    lib = C.CDLL(_build.SO)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError = header and library disagree: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))
