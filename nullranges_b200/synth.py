"""Deterministic synthetic inputs for the BASELINE.json configs.

hg38 chr1-22 seqlengths; ranges drawn per chromosome in proportion to its length and
sorted by (chrom, start, end); all strands '*'.  numpy default_rng(seed) throughout.
"""
from __future__ import annotations

import numpy as np

from .granges import GRanges

HG38 = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468,
}
assert sum(HG38.values()) == 2_875_001_522


def _assemble(seqid, start, end, **mcols) -> GRanges:
    o = np.lexsort((end, start, seqid))
    return GRanges(seqid[o].astype(np.int32), start[o], end=end[o], seqlengths=HG38, seqlevels=list(HG38),
                   **{k: v[o] for k, v in mcols.items()})


def _chrom_of(rng, n):
    L = np.array(list(HG38.values()), dtype=np.float64)
    return rng.choice(L.size, size=n, p=L / L.sum()).astype(np.int32), L.astype(np.int64)


def uniform_ranges(n=10_000, seed=1) -> GRanges:
    """cfg 1: start ~ U(1, L_c - w), width ~ 200 + Geom(mean 300) capped at 2 kb."""
    rng = np.random.default_rng(seed)
    seqid, L = _chrom_of(rng, n)
    w = np.minimum(200 + rng.geometric(1 / 300, n), 2000)
    start = 1 + (rng.random(n) * (L[seqid] - w)).astype(np.int64)
    return _assemble(seqid, start, start + w - 1)


def atac_peaks(n=1_000_000, seed=3) -> GRanges:
    """cfg 2/3/5: clustered peaks -- centres ~ U, 5-50 peaks per cluster, N(0, 5 kb) jitter, widths 150-1500."""
    rng = np.random.default_rng(seed)
    sizes = []
    total = 0
    while total < n:
        k = rng.integers(5, 51, size=max(1024, n // 20))
        sizes.append(k)
        total += int(k.sum())
    sizes = np.concatenate(sizes)
    cut = int(np.searchsorted(np.cumsum(sizes), n)) + 1
    sizes = sizes[:cut]
    sizes[-1] -= int(sizes.sum()) - n
    seqid_c, L = _chrom_of(rng, sizes.size)
    centre = 1 + (rng.random(sizes.size) * (L[seqid_c] - 1)).astype(np.int64)
    seqid = np.repeat(seqid_c, sizes)
    w = rng.integers(150, 1501, size=n)
    mid = np.repeat(centre, sizes) + np.rint(rng.normal(0.0, 5000.0, n)).astype(np.int64)
    start = np.clip(mid - w // 2, 1, L[seqid] - w)
    return _assemble(seqid, start, start + w - 1)


def points(n=10_000_000, seed=6) -> GRanges:
    """cfg 4: width-1 SNP-like points, uniform per chromosome."""
    rng = np.random.default_rng(seed)
    seqid, L = _chrom_of(rng, n)
    start = 1 + (rng.random(n) * L[seqid]).astype(np.int64)
    start = np.minimum(start, L[seqid])
    return _assemble(seqid, start, start.copy())


def genes(n=20_000, seed=2) -> GRanges:
    """cfg 1/2/3/5 query: width ~ logN(median 2e4, sigma 1.2) clipped to [200, 2e6], end <= L_c - 2."""
    rng = np.random.default_rng(seed)
    seqid, L = _chrom_of(rng, n)
    w = np.clip(np.rint(np.exp(rng.normal(np.log(2e4), 1.2, n))), 200, 2_000_000).astype(np.int64)
    w = np.minimum(w, L[seqid] - 3)
    start = 1 + (rng.random(n) * (L[seqid] - 2 - w)).astype(np.int64)
    return _assemble(seqid, start, start + w - 1, x_id=np.arange(1, n + 1))


def enhancers(n=200_000, seed=7) -> GRanges:
    """cfg 4 query: width ~ U[200, 2000]."""
    rng = np.random.default_rng(seed)
    seqid, L = _chrom_of(rng, n)
    w = rng.integers(200, 2001, size=n)
    start = 1 + (rng.random(n) * (L[seqid] - 2 - w)).astype(np.int64)
    return _assemble(seqid, start, start + w - 1, x_id=np.arange(1, n + 1))


def segmentation(L_s=2_000_000, seed=4, stay=0.85) -> GRanges:
    """cfg 3: tile at L_s, states from a 3-state Markov chain, adjacent equal states merged."""
    rng = np.random.default_rng(seed)
    seqid, start, end, state = [], [], [], []
    for c, L in enumerate(HG38.values()):
        n = -(-L // L_s)
        st = np.empty(n, np.int32)
        st[0] = rng.choice(3, p=[0.5, 0.3, 0.2]) + 1
        for i in range(1, n):
            st[i] = st[i - 1] if rng.random() < stay else rng.choice(3, p=[0.5, 0.3, 0.2]) + 1
        edges = np.flatnonzero(np.diff(st)) + 1
        first = np.concatenate(([0], edges))
        last = np.concatenate((edges, [n]))
        seqid.append(np.full(first.size, c, np.int32))
        start.append(1 + first * L_s)
        end.append(np.minimum(last * L_s, L))
        state.append(st[first])
    return GRanges(np.concatenate(seqid), np.concatenate(start), end=np.concatenate(end), seqlengths=HG38,
                   seqlevels=list(HG38), state=np.concatenate(state))


def exclude_regions(n=1000, seed=5) -> GRanges:
    """cfg 3: reduced regions, widths log-uniform [500, 5e6], one centromere-like region per chromosome."""
    rng = np.random.default_rng(seed)
    seqid, L = _chrom_of(rng, n - len(HG38))
    w = np.exp(rng.uniform(np.log(500), np.log(5e6), seqid.size)).astype(np.int64)
    w[w > 2_000_000] = np.exp(rng.uniform(np.log(500), np.log(2e5), int((w > 2_000_000).sum()))).astype(np.int64)
    start = 1 + (rng.random(seqid.size) * (L[seqid] - w - 1)).astype(np.int64)
    cs = np.arange(len(HG38), dtype=np.int32)
    cw = np.full(cs.size, 3_000_000, np.int64)
    cstart = (L * 0.4).astype(np.int64)
    seqid = np.concatenate((seqid, cs))
    start = np.concatenate((start, cstart))
    end = np.concatenate((start[: -cs.size] + w - 1, cstart + cw - 1))
    # reduce(): merge overlapping/adjacent regions per chromosome
    o = np.lexsort((start, seqid))
    seqid, start, end = seqid[o], start[o], end[o]
    out = []
    for c in np.unique(seqid):
        m = seqid == c
        s, e = start[m], end[m]
        run_max = np.maximum.accumulate(e)
        new = np.concatenate(([True], s[1:] > run_max[:-1] + 1))
        grp = np.cumsum(new) - 1
        rs = s[new]
        re = np.zeros(rs.size, np.int64)
        np.maximum.at(re, grp, e)
        out.append((np.full(rs.size, c, np.int32), rs, re))
    return GRanges(np.concatenate([a for a, _, _ in out]), np.concatenate([b for _, b, _ in out]),
                   end=np.concatenate([c for _, _, c in out]), seqlengths=HG38, seqlevels=list(HG38))
