"""Handle-level wrapper of the C ABI: one Engine = one GPU = one nrb200_handle."""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import ptr


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libnrb200 error {code}: {msg}")
        self.code = code


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _i8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int8)


@dataclass
class Draws:
    """Block draws for a run: Philox (device) or host arrays of shape [n_iter, n_blocks]."""
    n_iter: int
    rng_mode: int = _lib.RNG_PHILOX
    seed: int = 0
    iter0: int = 0
    bait_seqid: np.ndarray | None = None
    bait_start: np.ndarray | None = None
    dst_tile: np.ndarray | None = None

    def to_c(self):
        self._keep = (_i32(self.bait_seqid), _i32(self.bait_start), _i32(self.dst_tile))
        return _lib.Draws(int(self.n_iter), int(self.iter0), int(self.rng_mode), C.c_uint64(int(self.seed) & (2**64 - 1)),
                          ptr(self._keep[0], C.c_int32), ptr(self._keep[1], C.c_int32), ptr(self._keep[2], C.c_int32))


class Engine:
    """One nrb200_handle: a single GPU, or -- ``device`` a list of ordinals -- a device group that shards the
    iterations of every run over those GPUs and delivers one result (include/nrb200.h, nrb200_config.n_devices)."""

    def __init__(self, device=0, stream: int | None = None, force_stream: bool = False):
        """force_stream: always stream the sampled ranges instead of counting by ranks (tests, profiling)."""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.lock = threading.RLock()  # the handle is stateful: a front-end call holds it from set_ranges to the result
        devices = [int(d) for d in device] if isinstance(device, (list, tuple)) else None
        flags = _lib.FLAG_FORCE_STREAM if force_stream else 0
        if devices is not None:
            if not devices:
                raise ValueError("Engine: empty device list")
            arr = (C.c_int32 * len(devices))(*devices)
            cfg = _lib.Config(devices[0], flags, C.c_void_p(stream) if stream else None, len(devices), arr)
        else:
            cfg = _lib.Config(int(device), flags, C.c_void_p(stream) if stream else None, 0, None)
        rc = self._lib.nrb200_create(C.byref(cfg), C.byref(self._h))
        if rc:
            raise EngineError(rc, self._lib.nrb200_last_error(None).decode())
        self.device = devices[0] if devices is not None else int(device)
        self.devices = devices if devices is not None else [int(device)]
        self.force_stream = bool(force_stream)
        self.n_blocks = 0
        self.n_query = 0
        self.n_seq = 0

    @property
    def n_devices(self) -> int:
        return int(self._lib.nrb200_n_devices(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nrb200_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc: int):
        if rc:
            raise EngineError(rc, self._lib.nrb200_last_error(self._h).decode())

    # ---- inputs
    def set_ranges(self, seqlengths, seqid, start, end, strand=None, defer: bool = False):
        """defer=True: nrb200_set_ranges_begin() only -- the upload and the checks are queued; set_query / set_exclude /
        set_plan may follow, then finish_ranges() (which reports what set_ranges would have reported)."""
        sl = np.ascontiguousarray(seqlengths, dtype=np.int64)
        a, b, c, d = _i32(seqid), _i32(start), _i32(end), _i8(strand)
        self._pending = None
        fn = self._lib.nrb200_set_ranges_begin if defer else self._lib.nrb200_set_ranges
        self._check(fn(self._h, a.size, sl.size, ptr(sl, C.c_int64), ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(c, C.c_int32),
                       ptr(d, C.c_int8)))
        if defer:
            self._pending = (sl, a, b, c, d)  # the arrays stay alive (and must stay unchanged) until finish_ranges()
        self.n_seq = sl.size
        self.n_query = 0
        self.n_blocks = 0

    def finish_ranges(self):
        try:
            self._check(self._lib.nrb200_set_ranges_end(self._h))
        finally:
            self._pending = None

    def ranges_info(self) -> dict:
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._lib.nrb200_ranges_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"sorted": bool(a.value), "stranded": bool(b.value), "max_width": c.value}

    def set_query(self, seqid, start, end, strand=None, maxgap: int = -1, minoverlap: int = 0):
        """maxgap / minoverlap as in countOverlaps(x, boots, maxgap=, minoverlap=)."""
        a, b, c, d = _i32(seqid), _i32(start), _i32(end), _i8(strand)
        self._check(self._lib.nrb200_set_query_ex(self._h, a.size, ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(c, C.c_int32),
                                                  ptr(d, C.c_int8), int(maxgap), int(minoverlap)))
        self.n_query = a.size

    def set_exclude(self, seqid, start, end, strand=None, option: str = "drop"):
        """option: excludeOption of bootRanges() -- "drop" removes overlapping ranges, "trim" cuts them to the gaps."""
        a, b, c, d = _i32(seqid), _i32(start), _i32(end), _i8(strand)
        self._check(self._lib.nrb200_set_exclude(self._h, a.size, ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(c, C.c_int32), ptr(d, C.c_int8)))
        self._check(self._lib.nrb200_set_exclude_option(self._h, {"drop": _lib.EXCLUDE_DROP, "trim": _lib.EXCLUDE_TRIM}[option]))

    def clear_exclude(self):
        self._check(self._lib.nrb200_set_exclude(self._h, 0, None, None, None, None))
        self._check(self._lib.nrb200_set_exclude_option(self._h, _lib.EXCLUDE_DROP))

    def clear_query(self):
        self._check(self._lib.nrb200_set_query(self._h, 0, None, None, None, None))
        self.n_query = 0

    def set_plan(self, kind: int, block_length: int, seg=None) -> int:
        """seg: dict(seqid, start, end, state) for the segmented kinds."""
        if seg is not None:
            keep = tuple(_i32(seg[k]) for k in ("seqid", "start", "end", "state"))
            spec = _lib.PlanSpec(kind, int(block_length), keep[0].size, *(ptr(k, C.c_int32) for k in keep))
        else:
            spec = _lib.PlanSpec(kind, int(block_length), 0, None, None, None, None)
        nb = C.c_int64()
        self._check(self._lib.nrb200_set_plan(self._h, C.byref(spec), C.byref(nb)))
        self.n_blocks = nb.value
        return nb.value

    def tiles(self):
        B = self.n_blocks
        ts, tt, bw = (np.empty(B, np.int32) for _ in range(3))
        self._check(self._lib.nrb200_get_tiles(self._h, ptr(ts, C.c_int32), ptr(tt, C.c_int32), ptr(bw, C.c_int32)))
        return ts, tt, bw

    # ---- fused counting
    def run_counts(self, draws: Draws, want_counts: bool = True, out: dict | None = None, counts_dtype=np.int32) -> dict:
        """out: optional preallocated (e.g. pinned) C-contiguous result arrays keyed iter_nranges /
        iter_totals [R] int64 and feature_counts [R, G] int32; every element is overwritten.
        counts_dtype=np.uint16: the matrix comes back as uint16 (nrb200_run_counts_u16: half the bytes over PCIe and into
        host memory); result['overflow'] tells whether some count was saturated at 65535."""
        R, G = int(draws.n_iter), self.n_query
        out = out or {}
        if np.dtype(counts_dtype) == np.uint16 and want_counts and G:
            nr = out.get("iter_nranges", np.empty(R, np.int64))
            tot = out.get("iter_totals", np.empty(R, np.int64))
            cnt = out.get("feature_counts")
            if cnt is None:
                cnt = np.empty((R, G), np.uint16)
            if cnt.shape != (R, G) or cnt.dtype != np.uint16 or not cnt.flags.c_contiguous:
                raise ValueError("out['feature_counts'] must be C-contiguous uint16[R, G]")
            st, ov, d = _lib.Stats(), C.c_int32(0), draws.to_c()
            self._check(self._lib.nrb200_run_counts_u16(self._h, C.byref(d), ptr(nr, C.c_int64), ptr(tot, C.c_int64), ptr(cnt, C.c_uint16),
                                                        C.byref(ov), C.byref(st)))
            return {"iter_nranges": nr, "iter_totals": tot, "feature_counts": cnt, "overflow": bool(ov.value), "stats": st.as_dict()}

        def buf(key, shape, dtype):
            a = out.get(key)
            if a is None:
                return np.empty(shape, dtype)
            if a.shape != shape or a.dtype != dtype or not a.flags.c_contiguous:
                raise ValueError(f"out[{key!r}] must be C-contiguous {dtype.__name__}{list(shape)}")
            return a

        nr, tot = buf("iter_nranges", (R,), np.int64), buf("iter_totals", (R,), np.int64)
        cnt = buf("feature_counts", (R, G), np.int32) if (want_counts and G) else None
        if not G:
            tot[:] = 0
        st = _lib.Stats()
        d = draws.to_c()
        self._check(self._lib.nrb200_run_counts(self._h, C.byref(d), ptr(nr, C.c_int64), ptr(tot, C.c_int64), ptr(cnt, C.c_int32), C.byref(st)))
        return {"iter_nranges": nr, "iter_totals": tot, "feature_counts": cnt, "stats": st.as_dict()}

    def run_counts_resident(self, draws: Draws, want_counts: bool = True, want_stats: bool = True) -> dict | None:
        st = _lib.Stats()
        d = draws.to_c()
        self._check(self._lib.nrb200_run_counts_resident(self._h, C.byref(d), int(want_counts), C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def run_counts_gathered(self, draws: Draws, want_stats: bool = True) -> dict | None:
        """Device group: every GPU ends up with the whole [R, G] matrix and [R] vectors in its own memory (the count
        kernels store into each other's memory over NVLink); see gathered_pointers()."""
        st = _lib.Stats()
        d = draws.to_c()
        self._check(self._lib.nrb200_run_counts_gathered(self._h, C.byref(d), C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def gathered_pointers(self, member: int):
        """(device ordinal, iter_nranges ptr, iter_totals ptr, feature_counts ptr) of a member's copy of the gathered result."""
        dev, a, b, c = C.c_int32(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._lib.nrb200_gathered_pointers(self._h, int(member), C.byref(dev), C.byref(a), C.byref(b), C.byref(c)))
        return dev.value, a.value, b.value, c.value

    def resident_pointers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._lib.nrb200_resident_pointers(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def run_moments(self, draws: Draws, observed=None, batch_iter: int = 0, reset: bool = True) -> dict:
        if reset:
            self._check(self._lib.nrb200_moments_reset(self._h))
        R, G = int(draws.n_iter), self.n_query
        nr, tot = np.zeros(R, np.int64), np.zeros(R, np.int64)
        obs = _i32(observed)
        st = _lib.Stats()
        d = draws.to_c()
        self._check(self._lib.nrb200_run_moments(self._h, C.byref(d), int(batch_iter), ptr(obs, C.c_int32), ptr(nr, C.c_int64), ptr(tot, C.c_int64), C.byref(st)))
        s, q, ge = (np.zeros(G, np.int64) for _ in range(3))
        self._check(self._lib.nrb200_moments_fetch(self._h, ptr(s, C.c_int64), ptr(q, C.c_int64), ptr(ge, C.c_int64)))
        return {"iter_nranges": nr, "iter_totals": tot, "sum": s, "sumsq": q, "n_ge_observed": ge, "stats": st.as_dict()}

    def set_weights(self, weights):
        """One numeric value per range of y (the caller's order) for run_weighted(); None clears."""
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._check(self._lib.nrb200_set_weights(self._h, ptr(w, C.c_double)))

    def run_weighted(self, draws: Draws, want_matrix: bool = True) -> dict:
        """Sum of the weights of the bootstrapped ranges overlapping each feature, per iteration."""
        R, G = int(draws.n_iter), self.n_query
        it = np.zeros(R, np.float64)
        fs = np.zeros((R, G), np.float64) if want_matrix else None
        st = _lib.Stats()
        d = draws.to_c()
        self._check(self._lib.nrb200_run_weighted(self._h, C.byref(d), ptr(it, C.c_double), ptr(fs, C.c_double), C.byref(st)))
        return {"iter_sums": it, "feature_sums": fs, "stats": st.as_dict()}

    def run_weighted_moments(self, draws: Draws, observed=None) -> dict:
        """Per-feature moments of the weighted sums over the iterations (no [R, G] matrix): sum, sum of squares and,
        with observed, #{iterations with sum >= observed}."""
        R, G = int(draws.n_iter), self.n_query
        it = np.zeros(R, np.float64)
        s, q, ge = np.zeros(G, np.float64), np.zeros(G, np.float64), np.zeros(G, np.int64)
        obs = None if observed is None else np.ascontiguousarray(observed, dtype=np.float64)
        st = _lib.Stats()
        d = draws.to_c()
        self._check(self._lib.nrb200_run_weighted_moments(self._h, C.byref(d), ptr(obs, C.c_double), ptr(it, C.c_double), ptr(s, C.c_double),
                                                          ptr(q, C.c_double), ptr(ge, C.c_int64), C.byref(st)))
        return {"iter_sums": it, "sum": s, "sumsq": q, "n_ge_observed": ge, "stats": st.as_dict()}

    def count_observed(self, minoverlap: int = 0):
        """countOverlaps(query, y[, minoverlap]) on the ranges as given (no bootstrap)."""
        cnt = np.zeros(self.n_query, np.int32)
        tot = C.c_int64()
        self._check(self._lib.nrb200_count_observed_ex(self._h, int(minoverlap), ptr(cnt, C.c_int32), C.byref(tot)))
        return cnt, tot.value

    # ---- materialising
    def run_materialize(self, draws: Draws) -> dict:
        R = int(draws.n_iter)
        offs = np.zeros(R + 1, np.int64)
        st = _lib.Stats()
        d = draws.to_c()
        self._check(self._lib.nrb200_run_materialize(self._h, C.byref(d), ptr(offs, C.c_int64), C.byref(st)))
        n = int(offs[-1])
        cols = {k: np.empty(n, np.int32) for k in ("seqid", "start", "end", "src", "block")}
        self._check(self._lib.nrb200_fetch_rows(self._h, 0, n, *(ptr(cols[k], C.c_int32) for k in ("seqid", "start", "end", "src", "block"))))
        cols["iter_offsets"] = offs
        cols["stats"] = st.as_dict()
        return cols


    def expand_runs(self, offsets, first: int = 1) -> np.ndarray:
        """rep(first:(first + R - 1), diff(offsets)) as int32 (the dense `iter` column), filled by all host cores."""
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        out = np.empty(int(offs[-1] - offs[0]), np.int32)
        self._lib.nrb200_expand_runs(ptr(offs, C.c_int64), offs.size - 1, int(first), ptr(out, C.c_int32))
        return out


    def measure_l2_sector_rate(self, buffer_bytes: int = 0) -> float:
        """GB/s (sectors x 32 B) the L2 serves to scattered 8-byte loads on this GPU: the count kernel's ceiling."""
        v = C.c_double()
        self._check(self._lib.nrb200_measure_l2_sector_rate(self._h, int(buffer_bytes), C.byref(v)))
        return v.value

    def fetch_dst_tiles(self, r0: int, n: int) -> np.ndarray:
        """Permute variants: the tile every block of iterations [r0, r0 + n) of the last run moved to, [n, n_blocks]."""
        out = np.empty((int(n), self.n_blocks), np.int32)
        self._check(self._lib.nrb200_fetch_dst_tiles(self._h, int(r0), int(n), ptr(out, C.c_int32)))
        return out


_engines: dict[tuple, Engine] = {}
_engines_lock = threading.Lock()


def resolve_devices(device=0, gpus=None):
    """device: one ordinal or a list of them; gpus: N -> the first N GPUs, or a list of ordinals (overrides device)."""
    if gpus is not None:
        device = list(range(int(gpus))) if isinstance(gpus, (int, np.integer)) else [int(d) for d in gpus]
    if isinstance(device, (list, tuple)):
        device = [int(d) for d in device]
        if not device:
            raise ValueError("gpus: no device given")
        return device if len(device) > 1 else device[0]
    return int(device)


def get_engine(device=0, gpus=None) -> Engine:
    """Process-wide engine per GPU -- or per device group -- (device buffers are reused across calls)."""
    dev = resolve_devices(device, gpus)
    key = tuple(dev) if isinstance(dev, list) else (dev,)
    with _engines_lock:
        e = _engines.get(key)
        if e is None or e._h is None:
            e = _engines[key] = Engine(dev)
    return e
