"""bootRanges(): the reference's front end (R/bootRanges.R:72-128), served by the B200 engine.

Same arguments, same checks, same result columns.  New, default-off arguments:
``rng`` ("R" = draw blocks on R's stream under set_seed(), bit-comparable with the
reference; "philox" = draw on the GPU), ``seed``/``iter0`` for Philox, ``device``.
``bootCounts()`` is the fused path: bootstrap and countOverlaps(x, boots) per iteration
without ever building the R x N table (vignettes/bootRanges.Rmd:497-503).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib, samplers
from .engine import Draws, Engine, get_engine
from .granges import BootRanges, GRanges
from .rrng import RRng, global_rng


def _match_arg(value, choices, name):
    if isinstance(value, (tuple, list)):  # R's match.arg(): the default is the first choice
        value = value[0]
    if value not in choices:
        raise ValueError(f"'{name}' should be one of {', '.join(repr(c) for c in choices)}")
    return value


def _plan_kind(seg, proportionLength, type_, withinChrom):
    if seg is not None:  # type / withinChrom are ignored when seg is given (R/bootRanges.R:91-98)
        return _lib.SEG_PROP if proportionLength else _lib.SEG_NOPROP
    if type_ == "bootstrap":
        return _lib.UNSEG_WITHIN if withinChrom else _lib.UNSEG_ACROSS
    return _lib.PERMUTE_WITHIN if withinChrom else _lib.PERMUTE_ACROSS


def _seg_arrays(seg: GRanges, y: GRanges):
    if "state" not in seg.mcols:
        raise ValueError('"state" %in% names(mcols(seg)) is not TRUE')  # R/bootRanges.R:87
    state = np.asarray(seg.mcols["state"])
    if not np.all(state == np.round(state)):
        raise ValueError("seg$state must be integer valued")
    lut = {s: i for i, s in enumerate(y.seqlevels)}
    try:
        seqid = np.array([lut[s] for s in seg.seqnames.tolist()], dtype=np.int32)
    except KeyError as e:  # factor(levels = seqlevels(y)) would give NA (R/bootRanges.R:161-163)
        raise ValueError(f"seqnames(seg) {e} not in seqlevels(y)") from None
    return {"seqid": seqid, "start": seg.start, "end": seg.end, "state": state.astype(np.int32)}


def _side_arrays(x: GRanges, y: GRanges, what: str):
    """(seqid in y's seqlevels, start, end, strand or None when all '*', positions kept) of the ranges of x that lie
    on sequences y has."""
    strand = x.strand if np.any(x.strand) else None
    if x.seqlevels == y.seqlevels:
        return x.seqid, x.start, x.end, strand, None
    lut = {s: i for i, s in enumerate(y.seqlevels)}
    ids = np.array([lut.get(s, -1) for s in x.seqlevels], dtype=np.int32)[x.seqid]
    keep = ids >= 0  # ranges on sequences y does not have can never be hit
    if keep.all():
        return ids, x.start, x.end, strand, None
    return ids[keep], x.start[keep], x.end[keep], None if strand is None else strand[keep], np.flatnonzero(keep)


def _draw_host(kind, rng: RRng, R: int, y: GRanges, L_b: int, seg, eng: Engine):
    L_s = y.seqlengths.astype(np.float64)
    tile_seq, tile_start, _ = eng.tiles()
    out = []
    for _ in range(R):
        if kind == _lib.UNSEG_ACROSS:
            d = samplers.unseg_bootstrap_across_chrom(rng, L_s, L_b)
        elif kind == _lib.UNSEG_WITHIN:
            d = samplers.unseg_bootstrap_within_chrom(rng, L_s, L_b)
        elif kind == _lib.PERMUTE_WITHIN:
            d = samplers.unseg_permute_within_chrom(rng, L_s, L_b, tile_start)
        elif kind == _lib.PERMUTE_ACROSS:
            d = samplers.unseg_permute_across_chrom(rng, tile_seq, tile_start)
        elif kind == _lib.SEG_PROP:
            d = samplers.seg_bootstrap_prop(rng, seg["seqid"], seg["start"], seg["end"], seg["state"], L_b, L_s)
        else:
            d = samplers.seg_bootstrap_noprop(rng, seg["seqid"], seg["start"], seg["end"], seg["state"], L_b)
        out.append(d)
    bait_seq = np.stack([d[0] for d in out]) if R else np.empty((0, eng.n_blocks), np.int32)
    bait_start = np.stack([d[1] for d in out]) if R else np.empty((0, eng.n_blocks), np.int32)
    dst = np.stack([d[2] for d in out]) if (R and out[0][2] is not None) else None
    return bait_seq, bait_start, dst


def _setup(eng: Engine, y: GRanges, blockLength, seg, exclude, excludeOption, proportionLength, type_, withinChrom, query=None):
    """Argument checks of R/bootRanges.R:80-88 and R/bootUnsegmented.R:8,31, inputs to the engine.  y goes first and
    asynchronously (nrb200_set_ranges_begin): exclude, plan and query are prepared while it is on the wire.
    query: (x, maxgap, minoverlap) of bootCounts(); returns (kind, L_b, seg arrays, positions of x kept).  The caller holds
    eng.lock: the handle is stateful from here to the result."""
    chr_lens = y.seqlengths
    if np.any(chr_lens < 0):
        raise ValueError("all(!is.na(chr_lens)) is not TRUE")  # R/bootRanges.R:81
    excludeOption = _match_arg(excludeOption, ("drop", "trim"), "excludeOption")
    type_ = _match_arg(type_, ("bootstrap", "permute"), "type")
    if excludeOption == "trim":
        if exclude is None:
            raise ValueError("excludeOption='trim' needs exclude")  # strand(NULL) errors in the reference (:84)
        if np.any(exclude.strand != 0):
            raise ValueError('all(strand(exclude) == "*") is not TRUE')
    if float(blockLength) != int(blockLength):
        raise ValueError("blockLength must be integer valued")
    L_b = int(blockLength)
    kind = _plan_kind(seg, proportionLength, type_, withinChrom)
    segd = _seg_arrays(seg, y) if seg is not None else None
    stranded = bool(np.any(y.strand))
    eng.set_ranges(chr_lens, y.seqid, y.start, y.end, y.strand if stranded else None, defer=True)
    qkeep = None
    try:
        if exclude is not None:
            xi, xs, xe, xstr, _ = _side_arrays(exclude, y, "exclude")
            eng.set_exclude(xi, xs, xe, xstr, option=excludeOption)
        else:
            eng.clear_exclude()
        eng.set_plan(kind, L_b, segd)
        if query is not None:
            x, maxgap, minoverlap = query
            qi, qs, qe, qstr, qkeep = _side_arrays(x, y, "x")
            eng.set_query(qi, qs, qe, qstr, maxgap=maxgap, minoverlap=minoverlap)
    finally:
        eng.finish_ranges()  # y's own problems are reported first, as nrb200_set_ranges() would
    if kind in (_lib.UNSEG_ACROSS, _lib.PERMUTE_ACROSS):
        # stopifnot(all(y == GenomicRanges::sort(y)))  R/bootUnsegmented.R:31.  For an unstranded y the
        # engine's canonical-order test (made on the GPU while loading) is that check.
        if not (eng.ranges_info()["sorted"] if not stranded else y.is_sorted()):
            raise ValueError("all(y == GenomicRanges::sort(y)) is not TRUE")
    return kind, L_b, segd, qkeep


def _make_draws(eng, kind, R, y, L_b, segd, rng, seed, iter0):
    rng = _match_arg(rng, ("R", "philox"), "rng")
    if rng == "philox":
        return Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=int(seed or 0), iter0=int(iter0))
    stream = global_rng() if seed is None else RRng(seed)
    bs, bt, dt = _draw_host(kind, stream, R, y, L_b, segd, eng)
    return Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=bs, bait_start=bt, dst_tile=dt)


def bootRanges(y: GRanges, blockLength, R: int = 1, seg: GRanges | None = None, exclude: GRanges | None = None,
               excludeOption=("drop", "trim"), proportionLength: bool = True, type=("bootstrap", "permute"),
               withinChrom: bool = False, storeBlockLength: bool = False, *, rng="R", seed=None, iter0: int = 0,
               storeBlockStart: bool = False, device=0, gpus=None) -> BootRanges:
    """Block bootstrap of genomic ranges; returns all iterations concatenated with an ``iter`` column.
    gpus: as in bootCounts(); the materialised table itself is produced by the first GPU."""
    eng = get_engine(device, gpus)
    with eng.lock:
        return _boot_ranges(eng, y, blockLength, R, seg, exclude, excludeOption, proportionLength, type, withinChrom,
                            storeBlockLength, rng, seed, iter0, storeBlockStart)


def _boot_ranges(eng, y, blockLength, R, seg, exclude, excludeOption, proportionLength, type, withinChrom, storeBlockLength,
                 rng, seed, iter0, storeBlockStart) -> BootRanges:
    kind, L_b, segd, _ = _setup(eng, y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom)
    draws = _make_draws(eng, kind, int(R), y, L_b, segd, rng, seed, iter0)
    rows = eng.run_materialize(draws)
    lens = np.diff(rows["iter_offsets"])
    if not eng.ranges_info()["sorted"]:
        # The engine emits (iteration, block, canonical y order).  The reference's row order follows the order y was
        # GIVEN in: findOverlaps() lists a block's hits by subject index (R/bootUnsegmented.R:67, 133; R/bootRanges.R:166),
        # and the permute variants shift y in place, chromosome by chromosome (:28, :105, :164).  Only an unsorted y
        # (allowed within chromosomes and with seg) can tell the difference.
        it = np.repeat(np.arange(int(R)), lens)
        second = y.seqid[rows["src"]] if kind in (_lib.PERMUTE_WITHIN, _lib.PERMUTE_ACROSS) else rows["block"]
        o = np.lexsort((rows["start"], rows["src"], second, it))   # start: the pieces of a trimmed range stay ascending
        for k in ("seqid", "start", "end", "src", "block"):
            rows[k] = rows[k][o]
    src = rows["src"]
    br = object.__new__(BootRanges)
    br.seqlevels, br.seqlengths = y.seqlevels, y.seqlengths
    br.seqid, br.start, br.end = rows["seqid"], rows["start"], rows["end"]
    br.strand = y.strand[src] if np.any(y.strand) else np.zeros(src.size, np.int8)
    br.mcols = {k: v[src] for k, v in y.mcols.items() if k != "block"}
    # mcols(br)$iter <- Rle(factor(rep(seq_len(R), lens), levels = seq_len(R)))   R/bootRanges.R:122
    br.mcols["iter"] = eng.expand_runs(rows["iter_offsets"], first=1)
    if storeBlockLength:
        br.mcols["blockLength"] = np.full(src.size, L_b, dtype=np.int32)  # :123-125
    if storeBlockStart:
        _, tile_start, _ = eng.tiles()
        # where the block LANDED: bootstrap kinds, block b -> tile b; permute kinds, the tile it was drawn for (host
        # draws: the caller's dst_tile; Philox: the permutation the device made)
        dst_tile = draws.dst_tile
        if dst_tile is None and kind in (_lib.PERMUTE_WITHIN, _lib.PERMUTE_ACROSS):
            dst_tile = eng.fetch_dst_tiles(0, int(R))
        dst = rows["block"] if dst_tile is None else dst_tile[br.mcols["iter"] - 1, rows["block"]]
        br.mcols["blockStart"] = tile_start[dst]
    return br


@dataclass
class BootCounts:
    """Per-iteration results of the fused bootstrap -> countOverlaps path."""
    iter_totals: np.ndarray            # [R] sum over features of the overlap counts
    iter_nranges: np.ndarray           # [R] length of the iteration's bootstrap sample
    feature_counts: np.ndarray | None  # [R, length(x)] countOverlaps(x, boots[iter == r])
    stats: dict
    iter_sums: np.ndarray | None = None     # with weights=: [R] sum over features of feature_sums
    feature_sums: np.ndarray | None = None  # with weights=: [R, length(x)] sum of the column over the overlapping ranges


def bootCounts(y: GRanges, x: GRanges, blockLength, R: int = 1, seg: GRanges | None = None,
               exclude: GRanges | None = None, excludeOption=("drop", "trim"), proportionLength: bool = True,
               type=("bootstrap", "permute"), withinChrom: bool = False, *, rng="philox", seed=None, iter0: int = 0,
               perFeature: bool = True, device=0, out: dict | None = None, maxgap: int = -1,
               minoverlap: int = 0, weights: str | None = None, gpus=None, counts_dtype=np.int32) -> BootCounts:
    """countOverlaps(x, bootRanges(y, ...), maxgap = , minoverlap = ) per iteration, fused on the GPU.
    weights: name of a numeric metadata column of y; adds its sum over the overlapping bootstrapped ranges per
    (iteration, feature) -- summarize(sum(signal)) of vignettes/bootRanges.Rmd:532-540.
    gpus: N (the first N GPUs of the box) or a list of ordinals: the R iterations are sharded over them
    (replicate() at R/bootRanges.R:90 is R independent draws) and land in ONE result, bit-identical for every N.
    counts_dtype: np.int32 (R's integer, the default) or np.uint16 -- half the bytes over PCIe and into host memory;
    counts above 65535 are saturated and flagged in ``stats['overflow']``."""
    eng = get_engine(device, gpus)
    with eng.lock:
        return _boot_counts(eng, y, x, blockLength, R, seg, exclude, excludeOption, proportionLength, type, withinChrom, rng, seed,
                            iter0, perFeature, out, maxgap, minoverlap, weights, counts_dtype)


def _boot_counts(eng, y, x, blockLength, R, seg, exclude, excludeOption, proportionLength, type, withinChrom, rng, seed, iter0,
                 perFeature, out, maxgap, minoverlap, weights, counts_dtype=np.int32) -> BootCounts:
    kind, L_b, segd, qkeep = _setup(eng, y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom,
                                    query=(x, maxgap, minoverlap))
    draws = _make_draws(eng, kind, int(R), y, L_b, segd, rng, seed, iter0)
    if out is not None and qkeep is not None:
        raise ValueError("out= needs every seqlevel of x to be a seqlevel of y")
    res = eng.run_counts(draws, want_counts=perFeature, out=out, counts_dtype=counts_dtype)
    if "overflow" in res:
        res["stats"]["overflow"] = res["overflow"]
    fc = res["feature_counts"]
    if fc is None and perFeature:  # no feature lies on a sequence of y
        fc = np.zeros((int(R), len(x)), counts_dtype)
    elif fc is not None and qkeep is not None:  # features on sequences y lacks count 0
        full = np.zeros((int(R), len(x)), fc.dtype)
        full[:, qkeep] = fc
        fc = full
    bc = BootCounts(res["iter_totals"], res["iter_nranges"], fc, res["stats"])
    if weights is not None:
        if weights not in y.mcols:
            raise ValueError(f"{weights!r} is not a metadata column of y")
        eng.set_weights(y.mcols[weights])
        ws = eng.run_weighted(draws, want_matrix=perFeature)
        fs = ws["feature_sums"]
        if fs is not None and qkeep is not None:
            full = np.zeros((int(R), len(x)), np.float64)
            full[:, qkeep] = fs
            fs = full
        bc.iter_sums, bc.feature_sums = ws["iter_sums"], fs
    return bc


@dataclass
class BootMoments:
    """Per-feature summary of the bootstrap distribution of a statistic, without the [R, G] matrix."""
    n_iter: int
    mean: np.ndarray                  # [length(x)]
    sd: np.ndarray                    # sample standard deviation over the iterations
    observed: np.ndarray | None       # the statistic on y itself (counts: countOverlaps(x, y); weights: as given)
    z: np.ndarray | None              # (observed - mean) / sd
    p_ge: np.ndarray | None           # empirical P(bootstrap >= observed) = (#{>=} + 1) / (R + 1)
    sum: np.ndarray
    sumsq: np.ndarray
    stats: dict


def bootMoments(y: GRanges, x: GRanges, blockLength, R: int = 1, seg: GRanges | None = None,
                exclude: GRanges | None = None, excludeOption=("drop", "trim"), proportionLength: bool = True,
                type=("bootstrap", "permute"), withinChrom: bool = False, *, rng="philox", seed=None, iter0: int = 0,
                device=0, maxgap: int = -1, minoverlap: int = 0, weights: str | None = None,
                observed=None, gpus=None) -> BootMoments:
    """Mean / sd / z-score / empirical p-value per feature of countOverlaps(x, boots) -- or, with ``weights``, of the
    sum of that metadata column over the overlapping bootstrapped ranges -- over R iterations, accumulated on the GPU
    (the summary the vignette computes from the per-iteration table, vignettes/bootRanges.Rmd:497-560; no R x G
    matrix, so R = 10^4 with 10^5 features is fine).  observed: the statistic on the original ranges; for counts it
    defaults to countOverlaps(x, y), for weights it has to be given to get z and p.
    gpus: as in bootCounts(); every GPU reduces its shard of the iterations, the per-feature sums are added on the host."""
    eng = get_engine(device, gpus)
    with eng.lock:
        return _boot_moments(eng, y, x, blockLength, R, seg, exclude, excludeOption, proportionLength, type, withinChrom, rng, seed,
                             iter0, maxgap, minoverlap, weights, observed)


def _boot_moments(eng, y, x, blockLength, R, seg, exclude, excludeOption, proportionLength, type, withinChrom, rng, seed, iter0,
                  maxgap, minoverlap, weights, observed) -> BootMoments:
    kind, L_b, segd, qkeep = _setup(eng, y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom,
                                    query=(x, maxgap, minoverlap))
    if qkeep is not None:
        raise ValueError("bootMoments needs every seqlevel of x to be a seqlevel of y")
    draws = _make_draws(eng, kind, int(R), y, L_b, segd, rng, seed, iter0)
    if weights is None:
        obs = eng.count_observed(minoverlap=minoverlap)[0] if observed is None else np.asarray(observed, dtype=np.int32)
        m = eng.run_moments(draws, observed=obs)
    else:
        if weights not in y.mcols:
            raise ValueError(f"{weights!r} is not a metadata column of y")
        eng.set_weights(y.mcols[weights])
        obs = None if observed is None else np.asarray(observed, dtype=np.float64)
        m = eng.run_weighted_moments(draws, observed=obs)
    n = int(R)
    # (in place: with 10^5..10^6 features every temporary is megabytes of fresh pages)
    mean = np.divide(m["sum"], float(max(n, 1)))                   # exact int64 -> float64 below 2^53
    sd = np.multiply(mean, mean)
    sd *= n
    np.subtract(m["sumsq"], sd, out=sd)                           # q - n * mean^2
    np.maximum(sd, 0.0, out=sd)
    sd /= max(n - 1, 1)
    np.sqrt(sd, out=sd)
    z = p = None
    if obs is not None:
        with np.errstate(divide="ignore", invalid="ignore"):
            z = np.subtract(obs, mean)
            z /= sd
        p = np.add(m["n_ge_observed"], 1.0)
        p /= n + 1.0
    return BootMoments(n, mean, sd, obs, z, p, m["sum"], m["sumsq"], m["stats"])
