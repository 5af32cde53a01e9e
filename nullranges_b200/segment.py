"""GPU counting for the step that feeds bootRanges(): countOverlaps() on plain ranges, and the tile counts
of segmentDensity() (R/segmentDensity.R:51-78).  The segmentation itself (CBS / HMM + k-means,
R/segmentDensity.R:80-116) is third-party statistics run once and stays outside the engine."""
from __future__ import annotations

import numpy as np

from .engine import get_engine
from .granges import GRanges


def countOverlaps(query: GRanges, subject: GRanges, maxgap: int = -1, minoverlap: int = 0, device: int = 0) -> np.ndarray:
    """GenomicRanges::countOverlaps(query, subject, maxgap, minoverlap): per query range, the number of
    subject ranges on the same sequence with compatible strand that overlap it."""
    if maxgap != -1 and minoverlap > 0:
        raise ValueError("'maxgap' and 'minoverlap' cannot both be set")
    if np.any(subject.seqlengths < 0):
        raise ValueError("seqlengths(subject) must be known")
    eng = get_engine(device)
    eng.set_ranges(subject.seqlengths, subject.seqid, subject.start, subject.end, subject.strand if np.any(subject.strand) else None)
    lut = {s: i for i, s in enumerate(subject.seqlevels)}
    ids = np.array([lut.get(s, -1) for s in query.seqlevels], dtype=np.int32)[query.seqid]
    keep = np.flatnonzero(ids >= 0)
    out = np.zeros(len(query), np.int32)
    if keep.size:
        strand = query.strand[keep]
        eng.set_query(ids[keep], query.start[keep], query.end[keep], strand if np.any(strand) else None, maxgap=maxgap)
        out[keep] = eng.count_observed(minoverlap=minoverlap)[0]  # minoverlap is applied by the counting kernel itself
    return out


def tileGenome(seqlengths: dict, tilewidth: int) -> GRanges:
    """GenomicRanges::tileGenome(seqlengths, tilewidth = , cut.last.tile.in.chrom = TRUE)."""
    names, starts, ends = [], [], []
    for name, L in seqlengths.items():
        s = np.arange(1, int(L) + 1, int(tilewidth), dtype=np.int64)
        names.append(np.full(s.size, name, dtype=object))
        starts.append(s)
        ends.append(np.minimum(s + int(tilewidth) - 1, int(L)))
    return GRanges(np.concatenate(names), np.concatenate(starts), end=np.concatenate(ends), seqlengths=dict(seqlengths),
                   seqlevels=list(seqlengths))


def segmentTileCounts(x: GRanges, L_s: int = 1_000_000, exclude: GRanges | None = None, minoverlap: int = 8, device: int = 0):
    """The counting half of segmentDensity(x, n, L_s, exclude) (R/segmentDensity.R:53-78): tiles of width L_s,
    cut down to gaps(exclude) and kept when wider than L_s / 100, then countOverlaps(tiles, x, minoverlap = 8)
    and the counts standardised to the tile width.  Returns (tiles, counts_nostand, counts)."""
    present = [s for i, s in enumerate(x.seqlevels) if np.any(x.seqid == i)]  # seqlengths(x)[seqnames(x)@values]
    tiles = tileGenome({s: int(x.seqlengths[x.seqlevels.index(s)]) for s in present}, L_s)
    if exclude is not None:
        if np.any(exclude.strand != 0):
            raise ValueError('all(strand(exclude) == "*") is not TRUE')
        pieces = {"seq": [], "start": [], "end": []}
        for ci, name in enumerate(tiles.seqlevels):
            L = int(tiles.seqlengths[ci])
            m = exclude.seqnames == name
            xs, xe = exclude.start[m].astype(np.int64), exclude.end[m].astype(np.int64)
            o = np.argsort(xs, kind="stable")
            gaps, pos = [], 1
            for s, e in zip(xs[o], xe[o]):
                if s > pos and pos <= L:
                    gaps.append((pos, min(int(s) - 1, L)))
                pos = max(pos, int(e) + 1)
            if pos <= L:
                gaps.append((pos, L))
            t = tiles.seqid == ci
            for ts, te in zip(tiles.start[t], tiles.end[t]):   # join_overlap_intersect(query, gap)
                for gs, ge in gaps:
                    if gs <= te and ge >= ts:
                        s2, e2 = max(int(ts), gs), min(int(te), ge)
                        if e2 - s2 + 1 > L_s / 100:               # filter(width > L_s / 100)
                            pieces["seq"].append(name); pieces["start"].append(s2); pieces["end"].append(e2)
        tiles = GRanges(np.array(pieces["seq"], dtype=object), np.array(pieces["start"], np.int64), end=np.array(pieces["end"], np.int64),
                        seqlengths={s: int(L) for s, L in zip(tiles.seqlevels, tiles.seqlengths)}, seqlevels=tiles.seqlevels)
    counts_nostand = countOverlaps(tiles, x, minoverlap=minoverlap, device=device)
    return tiles, counts_nostand, counts_nostand / tiles.width * L_s


def oneRegionSegment(x: GRanges, seqlength: int) -> GRanges:
    """nullranges::oneRegionSegment(x, seqlength) (R/bootHelper.R:24-40): a three-range segmentation around one
    region -- state 1 before it, state 2 the region, state 1 after it -- for bootstrapping inside the region with
    bootRanges(..., seg = , proportionLength = FALSE) (vignettes/bootRanges.Rmd:733-749)."""
    if len(x) != 1:
        raise ValueError("length(x) == 1 is not TRUE")
    name = x.seqnames[0]
    s, e = int(x.start[0]), int(x.end[0])
    return GRanges([name] * 3, [1, s, e + 1], end=[s - 1, e, int(seqlength)], seqlengths={name: int(seqlength)},
                   state=np.array([1, 2, 1]))
