"""ctypes front-end of the CPU oracle (oracle/nr_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py -- never from nullranges_b200/.
See the header of nr_oracle.c for what it restates (R/bootRanges.R:72-331,
R/bootUnsegmented.R:1-191) and for the parity status ("parity unpinned" at value level).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnr_oracle.so")

UNSEG_ACROSS, UNSEG_WITHIN, SEG_PROP, SEG_NOPROP, PERMUTE_WITHIN, PERMUTE_ACROSS = range(6)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        rule = os.environ.get("NRB200_ZERO_WIDTH_RULE")  # development: flip the zero-width rule together with the engine's
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ([f"EXTRA=-DORC_ZERO_WIDTH_RULE={int(rule)}"] if rule else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, i32p, i64p, i8p, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int8), C.POINTER(C.c_double)
    u32p = C.POINTER(C.c_uint32)
    L.orc_last_error.restype = C.c_char_p
    L.orc_zero_width_rule.restype = C.c_int
    L.orc_rng_new.restype = vp
    L.orc_rng_new.argtypes = [C.c_uint32]
    L.orc_rng_free.argtypes = [vp]
    L.orc_unif_rand.restype = C.c_double
    L.orc_unif_rand.argtypes = [vp]
    L.orc_runif.argtypes = [vp, C.c_int64, dp, C.c_int64, dp, C.c_int64, dp]
    L.orc_round.restype = C.c_double
    L.orc_round.argtypes = [C.c_double]
    L.orc_sample_replace.argtypes = [vp, C.c_double, C.c_int64, i32p]
    L.orc_sample_noreplace.argtypes = [vp, C.c_int32, C.c_int32, i32p]
    L.orc_sample_prob.argtypes = [vp, C.c_int32, dp, C.c_int64, i32p]
    L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.orc_plan_new.restype = vp
    L.orc_plan_new.argtypes = [C.c_int, C.c_int, i64p, C.c_int64, C.c_int, i32p, i32p, i32p, i32p]
    L.orc_plan_free.argtypes = [vp]
    L.orc_plan_nblocks.restype = C.c_int64
    L.orc_plan_nblocks.argtypes = [vp]
    L.orc_plan_nstates.argtypes = [vp]
    L.orc_plan_state_block_length.restype = C.c_int64
    L.orc_plan_state_block_length.argtypes = [vp, C.c_int]
    L.orc_plan_tiles.argtypes = [vp, i32p, i32p, i32p]
    L.orc_plan_draw_R.argtypes = [vp, vp, i32p, i32p, i32p]
    L.orc_plan_draw_philox.argtypes = [vp, C.c_uint64, C.c_uint64, i32p, i32p, i32p]
    L.orc_index_new.restype = vp
    L.orc_index_new.argtypes = [C.c_int64, C.c_int, i32p, i32p, i32p, i8p]
    L.orc_index_free.argtypes = [vp]
    L.orc_index_set_maxgap.argtypes = [vp, C.c_int64]
    L.orc_index_set_minoverlap.argtypes = [vp, C.c_int64]
    L.orc_iteration_rows.restype = C.c_int64
    L.orc_iteration_rows.argtypes = [vp, vp, vp, C.c_int, i32p, i32p, i32p, C.c_int64, i32p, i32p, i32p, i32p, i32p]
    L.orc_count_overlaps.restype = C.c_int64
    L.orc_count_overlaps.argtypes = [vp, C.c_int64, i32p, i32p, i32p, i8p, i32p]
    L.orc_boot_counts.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int64, C.c_int, C.c_uint64, C.c_uint64,
                                  i32p, i32p, i32p, i64p, i64p, i32p, C.c_int]
    L.orc_max_threads.restype = C.c_int
    _lib = L
    return L


def _err() -> str:
    return lib().orc_last_error().decode()


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class RRng:
    """Base-R RNG stream under ``set.seed(seed)`` (Mersenne-Twister/Inversion/Rejection)."""

    def __init__(self, seed: int):
        self._h = lib().orc_rng_new(C.c_uint32(seed & 0xFFFFFFFF))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_rng_free(self._h)
            self._h = None

    def unif_rand(self) -> float:
        return lib().orc_unif_rand(self._h)

    def runif(self, n: int, a=0.0, b=1.0) -> np.ndarray:
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        out = np.empty(n, dtype=np.float64)
        lib().orc_runif(self._h, n, _p(a, C.c_double), a.size, _p(b, C.c_double), b.size, _p(out, C.c_double))
        return out

    def sample(self, n: int, size: int | None = None, replace: bool = False, prob=None) -> np.ndarray:
        """``sample.int(n, size, replace, prob)`` -- 1-based like R."""
        size = n if size is None else size
        out = np.empty(size, dtype=np.int32)
        if prob is not None:
            if not replace:
                raise NotImplementedError("weighted sampling without replacement is not on the path")
            prob = np.ascontiguousarray(prob, dtype=np.float64)
            if lib().orc_sample_prob(self._h, n, _p(prob, C.c_double), size, _p(out, C.c_int32)):
                raise ValueError(_err())
        elif replace:
            lib().orc_sample_replace(self._h, float(n), size, _p(out, C.c_int32))
        else:
            lib().orc_sample_noreplace(self._h, n, size, _p(out, C.c_int32))
        return out


def r_round(x: float) -> float:
    return lib().orc_round(x)


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


class Plan:
    """Block tables of one sampler (tiles are iteration-invariant)."""

    def __init__(self, kind: int, seqlengths, block_length: int, seg=None):
        self.kind = kind
        self.seqlengths = np.ascontiguousarray(seqlengths, dtype=np.int64)
        self.block_length = int(block_length)
        if seg is None:
            nseg, sc, ss, se, st = 0, None, None, None, None
        else:
            sc, ss, se, st = (_i32(seg[k]) for k in ("chrom", "start", "end", "state"))
            nseg = sc.size
        self._keep = (sc, ss, se, st)
        self._h = lib().orc_plan_new(kind, self.seqlengths.size, _p(self.seqlengths, C.c_int64), self.block_length,
                                     nseg, _p(sc, C.c_int32), _p(ss, C.c_int32), _p(se, C.c_int32), _p(st, C.c_int32))
        if not self._h:
            raise ValueError(_err())
        self.B = lib().orc_plan_nblocks(self._h)
        self.S = lib().orc_plan_nstates(self._h)
        self.tile_chrom = np.empty(self.B, np.int32)
        self.tile_start = np.empty(self.B, np.int32)
        self.bait_width = np.empty(self.B, np.int32)
        lib().orc_plan_tiles(self._h, _p(self.tile_chrom, C.c_int32), _p(self.tile_start, C.c_int32), _p(self.bait_width, C.c_int32))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_plan_free(self._h)
            self._h = None

    def state_block_length(self, s: int) -> int:
        return lib().orc_plan_state_block_length(self._h, s)

    def draw_R(self, rng: RRng, R: int = 1):
        """R iterations of block draws on the R RNG stream -> (bait_chrom, bait_start, dst_tile), each [R, B]."""
        bc = np.empty((R, self.B), np.int32)
        bs = np.empty((R, self.B), np.int32)
        dt = np.empty((R, self.B), np.int32)
        for r in range(R):
            if lib().orc_plan_draw_R(self._h, rng._h, _p(bc[r], C.c_int32), _p(bs[r], C.c_int32), _p(dt[r], C.c_int32)):
                raise ValueError(_err())
        return bc, bs, dt

    def draw_philox(self, seed: int, iter0: int, R: int = 1):
        bc = np.empty((R, self.B), np.int32)
        bs = np.empty((R, self.B), np.int32)
        dt = np.empty((R, self.B), np.int32)
        for r in range(R):
            if lib().orc_plan_draw_philox(self._h, seed, iter0 + r, _p(bc[r], C.c_int32), _p(bs[r], C.c_int32), _p(dt[r], C.c_int32)):
                raise ValueError(_err())
        return bc, bs, dt


class Index:
    """A set of ranges (y, exclude or query) with its per-chromosome overlap index."""

    def __init__(self, n_chrom: int, chrom, start, end, strand=None, maxgap: int = -1, minoverlap: int = 0):
        """maxgap / minoverlap: only meaningful for a query index -- countOverlaps(x, boots, maxgap=, minoverlap=)."""
        self.chrom, self.start, self.end = _i32(chrom), _i32(start), _i32(end)
        self.strand = None if strand is None else np.ascontiguousarray(strand, dtype=np.int8)
        self.n = self.chrom.size
        self._h = lib().orc_index_new(self.n, n_chrom, _p(self.chrom, C.c_int32), _p(self.start, C.c_int32),
                                      _p(self.end, C.c_int32), _p(self.strand, C.c_int8))
        if not self._h:
            raise ValueError(_err())
        if maxgap != -1:
            lib().orc_index_set_maxgap(self._h, int(maxgap))
        if minoverlap:
            lib().orc_index_set_minoverlap(self._h, int(minoverlap))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_index_free(self._h)
            self._h = None


def gaps(n_chrom: int, seqlengths, chrom, start, end, strand=None) -> "Index":
    """gaps(exclude, end = seqlengths(y)) %>% filter(strand == "*")  (R/bootRanges.R:112-113): the
    complement, inside [1, seqlength], of the '*'-strand exclude ranges, per sequence, as an Index.
    The reference insists on all(strand(exclude) == "*") for excludeOption = "trim" (:84)."""
    if strand is not None and np.any(np.asarray(strand) != 0):
        raise ValueError('all(strand(exclude) == "*") is not TRUE')
    chrom, start, end = np.asarray(chrom), np.asarray(start, dtype=np.int64), np.asarray(end, dtype=np.int64)
    gc, gs, ge = [], [], []
    for c in range(n_chrom):
        L = int(seqlengths[c])
        m = chrom == c
        o = np.argsort(start[m], kind="stable")
        pos = 1  # first position not yet covered
        for s, e in zip(start[m][o], end[m][o]):
            if s > pos and pos <= L:
                gc.append(c); gs.append(pos); ge.append(min(int(s) - 1, L))
            pos = max(pos, int(e) + 1)
        if pos <= L:
            gc.append(c); gs.append(pos); ge.append(L)
    return Index(n_chrom, np.array(gc, np.int32), np.array(gs, np.int32), np.array(ge, np.int32))


EXCLUDE_DROP, EXCLUDE_TRIM = 0, 1


def iteration_rows(plan: Plan, y: Index, bait_chrom, bait_start, dst_tile, exclude: Index | None = None,
                   exclude_mode: int = EXCLUDE_DROP) -> dict:
    """Rows of one iteration in ascending (block, y-index[, gap]) order.  exclude_mode EXCLUDE_TRIM:
    ``exclude`` must be the gaps() index."""
    bc, bs, dt = _i32(bait_chrom), _i32(bait_start), _i32(dst_tile)
    ex = exclude._h if exclude is not None else None
    args = (plan._h, y._h, ex, exclude_mode, _p(bc, C.c_int32), _p(bs, C.c_int32), _p(dt, C.c_int32))
    n = lib().orc_iteration_rows(*args, 0, None, None, None, None, None)
    out = {k: np.empty(n, np.int32) for k in ("chrom", "start", "end", "src", "block")}
    lib().orc_iteration_rows(*args, n, *(_p(out[k], C.c_int32) for k in ("chrom", "start", "end", "src", "block")))
    return out


def count_overlaps(query: Index, chrom, start, end, strand=None):
    """countOverlaps(query, rows) -> (counts[G], total)."""
    counts = np.zeros(query.n, np.int32)
    ch, st, en = _i32(chrom), _i32(start), _i32(end)
    sd = None if strand is None else np.ascontiguousarray(strand, dtype=np.int8)
    tot = lib().orc_count_overlaps(query._h, ch.size, _p(ch, C.c_int32), _p(st, C.c_int32), _p(en, C.c_int32), _p(sd, C.c_int8), _p(counts, C.c_int32))
    return counts, tot


def boot_counts(plan: Plan, y: Index, query: Index | None, R: int, *, exclude: Index | None = None,
                exclude_mode: int = EXCLUDE_DROP, draws=None, seed: int = 0, iter0: int = 0, want_counts: bool = True, nthreads: int = 1) -> dict:
    """Fused bootstrap -> countOverlaps over R iterations.

    ``draws=(bait_chrom, bait_start, dst_tile)`` ([R, B] each) selects validation mode;
    ``draws=None`` uses Philox draws for iterations iter0..iter0+R-1 under ``seed``.
    """
    G = query.n if query is not None else 0
    nr = np.zeros(R, np.int64)
    tot = np.zeros(R, np.int64)
    cnt = np.zeros((R, G), np.int32) if (want_counts and query is not None) else None
    if draws is not None:
        bc, bs, dt = (_i32(a) for a in draws)
        assert bc.shape == (R, plan.B)
        mode = 0
    else:
        bc = bs = dt = None
        mode = 1
    rc = lib().orc_boot_counts(plan._h, y._h, exclude._h if exclude is not None else None, exclude_mode,
                               query._h if query is not None else None, R, mode, seed, iter0,
                               _p(bc, C.c_int32), _p(bs, C.c_int32), _p(dt, C.c_int32),
                               _p(nr, C.c_int64), _p(tot, C.c_int64), _p(cnt, C.c_int32), nthreads)
    if rc:
        raise ValueError(_err())
    return {"iter_nranges": nr, "iter_totals": tot, "feature_counts": cnt}


def max_threads() -> int:
    return lib().orc_max_threads()


def boot_weighted(plan: Plan, y: Index, query: Index, weights, R: int, *, exclude: Index | None = None,
                  exclude_mode: int = EXCLUDE_DROP, draws=None, seed: int = 0, iter0: int = 0, maxgap: int = -1,
                  minoverlap: int = 0) -> dict:
    """Sum of ``weights`` (one per range of y) over the bootstrapped ranges overlapping each feature, per
    iteration: the vignette's summarize(sum(signal)) per (x_id, iter) (vignettes/bootRanges.Rmd:532-540).
    Plain Python / numpy over the oracle's rows: small cases only."""
    w = np.asarray(weights, dtype=np.float64)
    if draws is None:
        draws = plan.draw_philox(seed, iter0, R)
    ystr = y.strand if y.strand is not None else np.zeros(y.n, np.int8)
    qstr = query.strand if query.strand is not None else np.zeros(query.n, np.int8)
    m1 = maxgap + 1
    sums = np.zeros((R, query.n), np.float64)
    for r in range(R):
        rows = iteration_rows(plan, y, draws[0][r], draws[1][r], draws[2][r], exclude=exclude, exclude_mode=exclude_mode)
        for c, s, e, src in zip(rows["chrom"], rows["start"], rows["end"], rows["src"]):
            if e < s:
                continue
            ov = np.minimum(e, query.end) - np.maximum(s, query.start) + 1
            hit = (query.chrom == c) & (query.start - m1 <= e) & (query.end + m1 >= s) & ((ystr[src] == 0) | (qstr == 0) | (ystr[src] == qstr))
            if minoverlap > 0:
                hit &= ov >= minoverlap
            sums[r, hit] += w[src]
    return {"feature_sums": sums, "iter_sums": sums.sum(axis=1)}


def zero_width_rule() -> int:
    """0: a zero-width bootstrapped range overlaps nothing; 1: IRanges' plain overlap test (nr_oracle.c, THE ZERO-WIDTH RULE)."""
    return int(lib().orc_zero_width_rule())
