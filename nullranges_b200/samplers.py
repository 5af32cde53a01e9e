"""Validation-mode block draws: the sampling half of the reference's samplers, on R's RNG.

Each function mirrors one reference function up to the point where it would call
findOverlaps(): it returns where every block is sampled FROM (0-based seqname id, start) and,
for the permutation variants, which tile it moves TO.  The draws consume the R stream in
exactly the reference's order, so ``set_seed(s); bootRanges(...)`` is reproducible against
``set.seed(s); nullranges::bootRanges(...)``.  In the R package this file's role is played
by R code calling R's own sample()/runif() (r-package/R/bootRanges_b200.R).

Block order is the engine's (include/nrb200.h, nrb200_set_plan): unsegmented = seqlevel
major; segmented = state, then segment, then tile.
"""
from __future__ import annotations

import numpy as np

from .rrng import RRng, r_round


def unseg_bootstrap_across_chrom(rng: RRng, L_s: np.ndarray, L_b: int):
    """R/bootUnsegmented.R:116-127."""
    n_per_chrom = np.ceil(L_s / L_b)
    n = int(n_per_chrom.sum())
    random_chr = rng.sample(L_s.size, n, replace=True, prob=L_s)  # names(L_s) by position
    random_start = r_round(rng.runif(n, 1, (n_per_chrom[random_chr - 1] - 1) * L_b + 1))
    return random_chr - 1, random_start.astype(np.int32), None


def unseg_bootstrap_within_chrom(rng: RRng, L_s: np.ndarray, L_b: int):
    """R/bootUnsegmented.R:15-28 + :50-61: one runif() call per seqlevel, in level order."""
    chroms, starts = [], []
    for c, L in enumerate(L_s):
        n = int(np.ceil(L / L_b))
        starts.append(r_round(rng.runif(n, 1, (n - 1) * L_b + 1)))
        chroms.append(np.full(n, c, dtype=np.int32))
    return np.concatenate(chroms), np.concatenate(starts).astype(np.int32), None


def unseg_permute_within_chrom(rng: RRng, L_s: np.ndarray, L_b: int, tile_start: np.ndarray):
    """R/bootUnsegmented.R:83-101: block k of a chromosome moves to tile perm[k]."""
    chroms, dst = [], []
    off = 0
    for c, L in enumerate(L_s):
        n = int(np.ceil(L / L_b))
        perm = rng.sample(n)
        chroms.append(np.full(n, c, dtype=np.int32))
        dst.append(off + perm - 1)
        off += n
    return np.concatenate(chroms), tile_start.astype(np.int32), np.concatenate(dst).astype(np.int32)


def unseg_permute_across_chrom(rng: RRng, tile_seq: np.ndarray, tile_start: np.ndarray):
    """R/bootUnsegmented.R:150-157: rearranged_blocks <- blocks[sample(length(blocks))]."""
    perm = rng.sample(tile_seq.size)
    return tile_seq.astype(np.int32), tile_start.astype(np.int32), (perm - 1).astype(np.int32)


def seg_bootstrap_prop(rng: RRng, seg_seqid, seg_start, seg_end, state, L_b: int, chr_lens: np.ndarray):
    """R/bootRanges.R:179-233 (draws only)."""
    width = (seg_end - seg_start + 1).astype(np.float64)
    esses = range(1, int(state.max()) + 1)
    L_g = float(chr_lens.sum())
    L_s = np.array([width[state == s].sum() for s in esses])
    L_b_per_state = r_round(L_b * L_s / L_g)
    chroms, starts = [], []
    for s in esses:
        L_bs = L_b_per_state[s - 1]
        pick = state == s
        nblocks_per_rng = np.ceil(width[pick] / L_bs)
        n = int(nblocks_per_rng.sum())
        random_rng = rng.sample(int(pick.sum()), n, replace=True, prob=width[pick])
        rmin = seg_start[pick][random_rng - 1].astype(np.float64)
        rmax = (nblocks_per_rng[random_rng - 1] - 1) * L_bs + rmin
        starts.append(r_round(rng.runif(n, rmin, rmax)))
        chroms.append(seg_seqid[pick][random_rng - 1])
    return np.concatenate(chroms).astype(np.int32), np.concatenate(starts).astype(np.int32), None


def seg_bootstrap_noprop(rng: RRng, seg_seqid, seg_start, seg_end, state, L_b: int):
    """R/bootRanges.R:256-314 (draws only)."""
    width = (seg_end - seg_start + 1).astype(np.float64)
    esses = range(1, int(state.max()) + 1)
    nblocks_per_seg = np.ceil(width / L_b)
    n = int(nblocks_per_seg.sum())
    random_seg = rng.sample(seg_start.size, n, replace=True, prob=width)
    rmin = seg_start[random_seg - 1].astype(np.float64)
    rmax = (nblocks_per_seg[random_seg - 1] - 1) * L_b + rmin
    random_start = r_round(rng.runif(n, rmin, rmax))
    chroms, starts = [], []
    for s in esses:
        idx = np.flatnonzero(state[random_seg - 1] == s)
        rand_n = idx.size
        rearr_n = int(nblocks_per_seg[state == s].sum())
        if rand_n < rearr_n:
            if rand_n == 0:
                # the reference indexes with zeros and then fails its length check
                raise ValueError("length(rearranged_blocks_start) == length(random_blocks_start) is not TRUE")
            rand_idx = rng.sample(rand_n, rearr_n, replace=True) - 1
        else:
            rand_idx = np.arange(rearr_n)
        starts.append(random_start[idx][rand_idx])
        chroms.append(seg_seqid[random_seg[idx] - 1][rand_idx])
    return np.concatenate(chroms).astype(np.int32), np.concatenate(starts).astype(np.int32), None
