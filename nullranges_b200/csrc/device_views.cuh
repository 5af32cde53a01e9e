// Device-side views of the engine's buffers, the searches and direct-address rank tables, Philox block draws.
// (part of kernels.cuh)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/nrb200.h"
#include "philox.cuh"

namespace nrb {


constexpr int kWarp = 32;
constexpr int kMaxPeers = 16;  // GPUs of a device group whose count kernels store into each other's memory
constexpr unsigned kFull = 0xffffffffu;

// ---- THE ZERO-WIDTH RULE (one named constant; the oracle carries the same one, oracle/nr_oracle.c) -------------
// The within-chromosome variants keep ranges that trim() shrank to width 0 at [L + 1, L] (no width filter at
// R/bootUnsegmented.R:73, :107).  Whether such a range overlaps an exclude / query range is decided inside IRanges,
// which is neither in the reference tree nor runnable here ("R-check items"):
//   0  a zero-width range overlaps nothing                                              (what this build ships)
//   1  the plain test start_q <= end_s && end_q >= start_s also at width 0: a hit of every range that strictly
//      contains its position.  Only exclude / query ranges that reach past their sequence end can tell the two apart.
// Flip with -DNRB_ZERO_WIDTH_RULE=1 (NRB200_ZERO_WIDTH_RULE=1 python nullranges_b200/build.py --force).  Under rule 1
// the streaming kernels apply the test per range; the rank path, whose windows encode rule 0 (start' <= L), steps
// aside for exactly the inputs where the rules differ (rank_path_ok(), run.cuh).
#ifndef NRB_ZERO_WIDTH_RULE
#define NRB_ZERO_WIDTH_RULE 0
#endif
constexpr bool kZeroWidthStrictlyInside = NRB_ZERO_WIDTH_RULE != 0;

// ---- device views ---------------------------------------------------------------------

struct RangesView {            // canonical (seqname, start, end) order
    const int32_t *start;
    const int32_t *end;
    const int32_t *pmax;       // running max of end within the sequence
    const int8_t *strand;      // nullptr when every strand is '*'
    const int32_t *src;        // index in the caller's order; nullptr = identity (the input was already canonical)
    const int32_t *seq_off;    // [n_seq + 1]
    const int4 *seq_info;      // [n_seq] {seq_off[c], seq_off[c+1], max_pos[c], first table entry} in one 16-byte load
    const uint2 *rank_start;   // packed rank tables over start / pmax, see "direct-address rank tables" below
    const uint2 *rank_pmax;
    int rank_shift;
    // Rank path.  When strand can matter the ranges are ALSO kept grouped by strand class inside each
    // sequence ("virtual sequence" v = seq * n_cls + class), so that a count can be taken per class; class =
    // strand class ('*', '+', '-' when y is stranded) x width class (ranges narrower than the query's
    // minoverlap can never count).  With one class (n_cls = 1) these alias the canonical arrays.
    int n_cls;
    const int32_t *rstart, *rend, *rpmax;   // (seq, class, start, end) order
    const int32_t *end_sorted;               // ends ascending inside each virtual sequence
    int32_t end_shift;                       // all ranges equally wide (w): end order = start order, end_sorted aliases
                                             // rstart and #{end < b} = #{start < b - (w - 1)}; end_shift = w - 1, else 0
    const int4 *rseq_info;                   // [n_seq * n_cls]
    const uint2 *rrank_start, *rank_end;     // tables over rstart / end_sorted
};

struct SliceView {             // query or exclude (or, for "trim", the gaps), canonical order, plus a per-tile window
    const int32_t *start;
    const int32_t *end;
    const int32_t *pmax;       // running max of end inside the sequence
    const int8_t *strand;
    const int32_t *src;
    const int32_t *tile_lo;    // [n_blocks] first candidate of the tile
    const int32_t *tile_hi;    // [n_blocks] one past the last
};

struct SamplerView {
    int kind, n_seq, n_states;
    int64_t block_length, genome_length;
    const int64_t *seq_cum, *tiles_per_seq;
    const int64_t *state_block_off, *state_seg_off, *state_len, *state_block_len;
    const int64_t *sg_cum;
    const int32_t *sg_seq, *sg_start, *sg_tiles;
    const int32_t *pick_tab, *pick_shift;  // coarse pick tables (philox.cuh): [0] over seq_cum, [st] over state st's sg_cum
};

struct BootParams {
    RangesView y;
    SliceView query, excl;
    SamplerView sampler;
    // iteration-invariant block tables
    int64_t n_blocks;
    const int32_t *tile_seq, *tile_start, *bait_width;
    const int2 *tile_info;    // [n_blocks] {bait width, length of the tile's sequence}
    const int64_t *seqlen;
    int permute, drop_zero;
    int rows_windows_only;    // rank path, excludeOption = "trim": rows = sum of the tile's gap windows
    int32_t minoverlap;       // countOverlaps(minoverlap =): smallest intersection width that counts (0 = any)
    // draws
    int rng_mode;
    uint64_t seed;
    int64_t iter0;            // global index of local iteration 0 (Philox counter)
    const int32_t *bait_seq, *bait_start, *dst_tile;  // [n_iter x n_blocks] (host draws)
    // outputs
    int64_t n_iter, n_query;
    unsigned long long *iter_nranges, *iter_totals;  // [n_iter]
    int32_t *counts;          // [n_iter x n_query], caller's query order, or nullptr
    unsigned short *counts16; // the same as uint16 (rank path; takes precedence), with one overflow flag for the launch
    int *overflow16;
    // device group, gathered result: the rows of this GPU's iterations are stored into the [R x n_query] matrix of EVERY
    // member (peer-mapped pointers over NVLink, this GPU's own included), already offset to this launch's first row
    int n_peers;
    int32_t *peer_counts[kMaxPeers];
    int32_t *item_rows;       // [n_iter x n_blocks] surviving rows per item, or nullptr
    unsigned long long *n_hits;  // [kHitSlots] partial sums of the ranges sampled (byte-model H), slot = iteration % kHitSlots
    int2 *draw_tab;           // [n_iter x n_blocks] {bait seq, bait start}: rank path scratch
    int4 *item_rec;           // [n_iter x n_blocks] {lo, hi, bait seq, bait start}: written by the rows kernel for the fill pass
};

// ---- searches -------------------------------------------------------------------------

// first index in [a, b) with v[idx] > key
__device__ __forceinline__ int32_t first_gt(const int32_t *__restrict__ v, int32_t a, int32_t b, int32_t key) {
    while (a < b) {
        const int32_t m = (a + b) >> 1;
        if (__ldg(v + m) <= key) a = m + 1; else b = m;
    }
    return a;
}
// first index in [a, b) with v[idx] >= key
__device__ __forceinline__ int32_t first_ge(const int32_t *__restrict__ v, int32_t a, int32_t b, int32_t key) {
    while (a < b) {
        const int32_t m = (a + b) >> 1;
        if (__ldg(v + m) < key) a = m + 1; else b = m;
    }
    return a;
}

// ---- direct-address rank tables --------------------------------------------------------
// One table per ascending per-sequence array (start, running-max end, sorted end).  Entry u of a
// sequence describes bucket u = [u << shift, (u + 1) << shift) in 8 bytes:
//     .x = (first << 3) | min(count, 7)
//     .y = the bucket's first values themselves, as offsets from the bucket's lower end
// first = index, relative to the sequence's first range, of the first value >= u << shift;
// count = values inside the bucket (7 means "7 or more": the next entry holds the end).
// .y holds three 10-bit offsets when shift <= 10 (kInlineShift) and one 32-bit offset otherwise;
// unused slots are all ones, which no key offset exceeds.  A rank query is therefore ONE 8-byte
// load -- one 32-byte sector -- and touches the value array only when a bucket holds more values
// than are inlined and the key lies beyond the last inlined one.
typedef uint2 RankEntry;
constexpr int kHitSlots = 64;  // the H statistic is tallied in this many places so that no single address takes every atomic
constexpr int kRankCountBits = 3;
constexpr int32_t kRankCountCap = (1 << kRankCountBits) - 1;
constexpr int kInlineShift = 10;                 // up to this shift an entry inlines three values
constexpr uint32_t kInlineMask = (1u << kInlineShift) - 1u;

__device__ __forceinline__ int32_t rank_clamp_key(int32_t key, const int4 si) {
    return min(max(key, 0), si.z + 1);  // key <= 0 -> si.x and key > max_pos -> si.y fall out of the table
}

// Entry of a bucket from its values: vals[p .. q) are the values inside bucket `base`..; a = first index of the sequence.
__device__ __forceinline__ RankEntry rank_entry_pack(const int32_t *__restrict__ vals, int32_t a, int32_t p, int32_t q, int64_t base,
                                                     int shift) {
    RankEntry e;
    e.x = ((uint32_t)(p - a) << kRankCountBits) | (uint32_t)min(q - p, kRankCountCap);
    if (shift <= kInlineShift) {
        uint32_t o[3] = {kInlineMask, kInlineMask, kInlineMask};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (p + k < q) o[k] = (uint32_t)((int64_t)__ldg(vals + p + k) - base);
        e.y = o[0] | (o[1] << kInlineShift) | (o[2] << (2 * kInlineShift));
    } else {
        e.y = p < q ? (uint32_t)((int64_t)__ldg(vals + p) - base) : 0xffffffffu;
    }
    return e;
}

// Rank from a loaded entry: first index of the sequence whose value is >= key (key already clamped, u its entry).
__device__ __forceinline__ int32_t rank_from_entry(const int32_t *__restrict__ vals, const RankEntry *__restrict__ table, int rank_shift,
                                                   const int4 si, int32_t key, int32_t u, const RankEntry e) {
    const int32_t a = si.x + (int32_t)(e.x >> kRankCountBits), n = (int32_t)(e.x & kRankCountCap);
    const uint32_t koff = (uint32_t)key & ((1u << rank_shift) - 1u);
    int32_t below, inlined;
    bool beyond;  // every inlined value is below the key
    if (rank_shift <= kInlineShift) {
        const uint32_t o0 = e.y & kInlineMask, o1 = (e.y >> kInlineShift) & kInlineMask, o2 = (e.y >> (2 * kInlineShift)) & kInlineMask;
        below = (o0 < koff) + (o1 < koff) + (o2 < koff);
        beyond = o2 < koff;
        inlined = 3;
    } else {
        below = e.y < koff;
        beyond = below;
        inlined = 1;
    }
    if (n > inlined && beyond) {  // rare: the bucket holds more values than the entry inlines and the key lies beyond them
        const int32_t b = n < kRankCountCap ? a + n : si.x + (int32_t)(__ldg(table + u + 1).x >> kRankCountBits);
        return first_ge(vals, a + inlined, b, key);
    }
    return a + below;
}

// First index of a sequence (si = seq_info[c]) whose value in `vals` is >= key: one table probe.
__device__ __forceinline__ int32_t first_ge_ranked(const int32_t *__restrict__ vals, const RankEntry *__restrict__ table,
                                                   int rank_shift, const int4 si, int32_t key) {
    key = rank_clamp_key(key, si);
    const int32_t u = si.w + (key >> rank_shift);
    return rank_from_entry(vals, table, rank_shift, si, key, u, __ldg(table + u));
}

// Hits of the bait block [s, e] on sequence c are a window [lo, hi) of the canonical order:
// hi = #{start <= e}, lo = first index whose running-max end reaches s.
__device__ __forceinline__ void locate_block(const RangesView &y, int c, int32_t s, int64_t e,
                                             int32_t &lo, int32_t &hi) {
    const int4 si = __ldg(y.seq_info + c);
    hi = first_ge_ranked(y.start, y.rank_start, y.rank_shift, si, (int32_t)min(e, (int64_t)INT32_MAX - 1) + 1);
    lo = first_ge_ranked(y.pmax, y.rank_pmax, y.rank_shift, si, s);
    if (lo > hi) lo = hi;
}

// #{start <= a} and #{end < b} on a sequence, both as canonical-order indices (their difference
// is the number of ranges overlapping [b, a] when b <= a).  a, b are below 2^31 - 1.
__device__ __forceinline__ int32_t idx_start_le(const RangesView &y, const int4 si, int32_t a) {
    return first_ge_ranked(y.rstart, y.rrank_start, y.rank_shift, si, a + 1);
}
__device__ __forceinline__ int32_t idx_end_lt(const RangesView &y, const int4 si, int32_t b) {
    return first_ge_ranked(y.end_sorted, y.rank_end, y.rank_shift, si, b - y.end_shift);
}

__device__ __forceinline__ bool strands_match(int a, int b) { return a == 0 || b == 0 || a == b; }

// ---- block draws ----------------------------------------------------------------------

// Production-mode draw of block b of global iteration `iter` (same arithmetic on the host:
// philox.cuh).  Returns the sequence and start the block is sampled from.
__device__ __forceinline__ void draw_block(const BootParams &P, uint64_t iter, int64_t b, int &c, int &s) {
    const SamplerView &S = P.sampler;
    const Philox128 u = philox_draw(P.seed, iter, (uint32_t)b, 0u);
    if (S.kind == NRB200_UNSEG_ACROSS) {
        const int64_t pos = (int64_t)mul_hi_u64_any(u.lo, (uint64_t)S.genome_length);
        c = pick_by_length_tab(S.seq_cum, S.n_seq, pos, S.pick_tab, S.pick_shift[0]);
        s = (int32_t)rounded_uniform(u.hi, 1, (S.tiles_per_seq[c] - 1) * S.block_length);
    } else if (S.kind == NRB200_UNSEG_WITHIN) {
        c = __ldg(P.tile_seq + b);
        s = (int32_t)rounded_uniform(u.hi, 1, (S.tiles_per_seq[c] - 1) * S.block_length);
    } else if (S.kind == NRB200_PERMUTE_WITHIN || S.kind == NRB200_PERMUTE_ACROSS) {
        c = __ldg(P.tile_seq + b);    // the blocks ARE the tiles; the draw is where they go (dst_tile, permute_keys_kernel)
        s = __ldg(P.tile_start + b);
    } else {  // NRB200_SEG_PROP
        int st = 1;
        while (st < S.n_states && b >= S.state_block_off[st + 1]) ++st;
        const int64_t first = S.state_seg_off[st];
        const int n = (int)(S.state_seg_off[st + 1] - first);
        const int64_t pos = (int64_t)mul_hi_u64_any(u.lo, (uint64_t)S.state_len[st]);
        const int64_t j = first + pick_by_length_tab(S.sg_cum + first, n, pos, S.pick_tab + st * kPickBuckets, S.pick_shift[st]);
        c = S.sg_seq[j];
        s = (int32_t)rounded_uniform(u.hi, S.sg_start[j], (int64_t)(S.sg_tiles[j] - 1) * S.state_block_len[st]);
    }
}

}  // namespace nrb
