"""A minimal GRanges / BootRanges pair for the Python mirror of bootRanges().

Only what the hot path reads and returns: seqnames (as ids into seqlevels), ranges,
strand, seqlengths and metadata columns.  BootRanges is a GRanges with an ``iter`` column
(R/AllClasses.R:10-13; R/bootRanges.R:120-127).
"""
from __future__ import annotations

import numpy as np

STRAND_CODE = {"*": 0, "+": 1, "-": 2}
STRAND_CHAR = np.array(["*", "+", "-"])
# GenomicRanges orders strand levels "+" < "-" < "*"
_STRAND_RANK = np.array([2, 0, 1], dtype=np.int8)


class GRanges:
    def __init__(self, seqnames, start, end=None, width=None, strand=None, seqlengths=None, seqlevels=None, **mcols):
        start = np.asarray(start, dtype=np.int64)
        if (end is None) == (width is None):
            raise ValueError("give exactly one of end / width")
        end = np.asarray(end, dtype=np.int64) if end is not None else start + np.asarray(width, dtype=np.int64) - 1
        n = start.size
        seqnames = np.asarray(seqnames)
        if seqlevels is None:
            seqlevels = list(seqlengths.keys()) if isinstance(seqlengths, dict) else None
        if seqnames.dtype.kind in "iu":
            if seqlevels is None:
                raise ValueError("integer seqnames need seqlevels")
            seqid = seqnames.astype(np.int32)
        else:
            if seqlevels is None:
                seqlevels = list(dict.fromkeys(seqnames.tolist()))  # order of appearance
            lut = {s: i for i, s in enumerate(seqlevels)}
            try:
                seqid = np.fromiter((lut[s] for s in seqnames.tolist()), dtype=np.int32, count=n)
            except KeyError as e:
                raise ValueError(f"seqname {e} not in seqlevels") from None
        self.seqlevels = list(seqlevels)
        self.seqid = np.broadcast_to(seqid, (n,)).astype(np.int32) if seqid.size != n else seqid
        self.start = start.astype(np.int32)
        self.end = end.astype(np.int32)
        if strand is None:
            self.strand = np.zeros(n, dtype=np.int8)
        else:
            strand = np.asarray(strand)
            if strand.dtype.kind in "iu":
                self.strand = strand.astype(np.int8)
            else:
                self.strand = np.fromiter((STRAND_CODE[s] for s in np.broadcast_to(strand, (n,)).tolist()), dtype=np.int8, count=n)
        if seqlengths is None:
            self.seqlengths = np.full(len(self.seqlevels), -1, dtype=np.int64)  # NA
        elif isinstance(seqlengths, dict):
            self.seqlengths = np.array([seqlengths.get(s, -1) for s in self.seqlevels], dtype=np.int64)
        else:
            self.seqlengths = np.asarray(seqlengths, dtype=np.int64)
        self.mcols = {k: np.asarray(v) for k, v in mcols.items()}
        for k, v in self.mcols.items():
            if v.shape[0] != n:
                raise ValueError(f"metadata column {k!r} has the wrong length")

    # -- accessors in the reference's vocabulary
    def __len__(self) -> int:
        return int(self.start.size)

    @property
    def width(self) -> np.ndarray:
        return self.end - self.start + 1

    @property
    def seqnames(self) -> np.ndarray:
        return np.asarray(self.seqlevels, dtype=object)[self.seqid]

    def __getitem__(self, idx) -> "GRanges":
        out = object.__new__(type(self))
        out.seqlevels, out.seqlengths = self.seqlevels, self.seqlengths
        out.seqid, out.start, out.end, out.strand = self.seqid[idx], self.start[idx], self.end[idx], self.strand[idx]
        out.mcols = {k: v[idx] for k, v in self.mcols.items()}
        return out

    def sort_order(self) -> np.ndarray:
        """order(gr): seqnames, strand (+ < - < *), start, end -- GenomicRanges::sort."""
        return np.lexsort((self.end, self.start, _STRAND_RANK[self.strand], self.seqid))

    def sort(self) -> "GRanges":
        return self[self.sort_order()]

    def is_sorted(self) -> bool:
        """all(y == sort(y)): equal ranges may be in any order.  One O(N) pass over neighbours."""
        if len(self) < 2:
            return True
        keys = [self.seqid]
        if np.any(self.strand):
            keys.append(_STRAND_RANK[self.strand])
        keys += [self.start, self.end]
        undecided = np.ones(len(self) - 1, dtype=bool)  # pairs equal on every key so far
        for k in keys:
            a, b = k[:-1], k[1:]
            if np.any(undecided & (a > b)):
                return False
            undecided &= a == b
        return True

    def __repr__(self) -> str:
        return f"{type(self).__name__} with {len(self)} ranges on {len(self.seqlevels)} sequences, mcols={list(self.mcols)}"


class BootRanges(GRanges):
    """GRanges of all bootstrap iterations concatenated; ``mcols['iter']`` is 1..R."""

    @property
    def iter(self) -> np.ndarray:
        return self.mcols["iter"]
