// Device side of the stream engine: one warp per (iteration, block) work item.
//
//   bait block (sampled or supplied)  ->  locate its hits in the start-sorted ranges
//   -> stream the hits, shift to the destination tile, trim, drop, exclude
//   -> count overlaps against the tile's slice of the query index.
//
// Reference code this replaces, per iteration: R/bootUnsegmented.R:116-143 / :50-75 /
// R/bootRanges.R:141-176 (sampler + findOverlaps + gather), R/bootUnsegmented.R:175-191
// (shift_and_swap_chrom), R/bootRanges.R:106-109 (exclude drop) and the user-side
// countOverlaps of vignettes/bootRanges.Rmd:497-503.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/nrb200.h"
#include "philox.cuh"

namespace nrb {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- device views ---------------------------------------------------------------------

struct RangesView {            // canonical (seqname, start, end) order
    const int32_t *start;
    const int32_t *end;
    const int32_t *pmax;       // running max of end within the sequence
    const int8_t *strand;      // nullptr when every strand is '*'
    const int32_t *src;        // index in the caller's order; nullptr = identity (the input was already canonical)
    const int32_t *seq_off;    // [n_seq + 1]
    const int4 *seq_info;      // [n_seq] {seq_off[c], seq_off[c+1], max_pos[c], first table entry} in one 16-byte load
    const uint32_t *rank_start;  // packed rank tables over start / pmax, see "direct-address rank tables" below
    const uint32_t *rank_pmax;
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
    const uint32_t *rrank_start, *rank_end;  // tables over rstart / end_sorted
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
};

struct BootParams {
    RangesView y;
    SliceView query, excl;
    SamplerView sampler;
    // iteration-invariant block tables
    int64_t n_blocks;
    const int32_t *tile_seq, *tile_start, *bait_width;
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
    int32_t *item_rows;       // [n_iter x n_blocks] surviving rows per item, or nullptr
    unsigned long long *n_hits;  // total ranges sampled (byte-model H)
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
// sequence describes bucket u = [u << shift, (u + 1) << shift):
//     entry = (first << 3) | min(count, 7)
// first = index, relative to the sequence's first range, of the first value >= u << shift;
// count = values inside the bucket.  One 4-byte probe therefore yields both ends of the bucket
// (count == 7 means "7 or more": the next entry holds the end).
constexpr int kRankCountBits = 3;
constexpr int32_t kRankCountCap = (1 << kRankCountBits) - 1;

__device__ __forceinline__ int32_t rank_clamp_key(int32_t key, const int4 si) {
    return min(max(key, 0), si.z + 1);  // key <= 0 -> si.x and key > max_pos -> si.y fall out of the table
}

// First index of a sequence (si = seq_info[c]) whose value in `vals` is >= key.  Branch-free in the
// common case: one table probe, then the first three values of the bucket read together.
__device__ __forceinline__ int32_t first_ge_ranked(const int32_t *__restrict__ vals, const uint32_t *__restrict__ table,
                                                   int rank_shift, const int4 si, int32_t key) {
    key = rank_clamp_key(key, si);
    const int32_t u = si.w + (key >> rank_shift);
    const uint32_t e = __ldg(table + u);
    const int32_t a = si.x + (int32_t)(e >> kRankCountBits), n = (int32_t)(e & kRankCountCap);
    const int32_t v0 = n > 0 ? __ldg(vals + a) : INT32_MAX;
    const int32_t v1 = n > 1 ? __ldg(vals + a + 1) : INT32_MAX;
    const int32_t v2 = n > 2 ? __ldg(vals + a + 2) : INT32_MAX;
    if (v2 < key) {  // rare: more than three values in the bucket and the first three are below the key
        const int32_t b = n < kRankCountCap ? a + n : si.x + (int32_t)(__ldg(table + u + 1) >> kRankCountBits);
        return first_ge(vals, a + 3, b, key);
    }
    return a + (v0 < key) + (v1 < key) + (v2 < key);
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
        const int64_t pos = (int64_t)mul_hi_u64(u.lo, (uint64_t)S.genome_length);
        c = pick_by_length(S.seq_cum, S.n_seq, pos);
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
        const int64_t pos = (int64_t)mul_hi_u64(u.lo, (uint64_t)S.state_len[st]);
        const int64_t j = first + pick_by_length(S.sg_cum + first, n, pos);
        c = S.sg_seq[j];
        s = (int32_t)rounded_uniform(u.hi, S.sg_start[j], (int64_t)(S.sg_tiles[j] - 1) * S.state_block_len[st]);
    }
}

// ---- the fused kernel -----------------------------------------------------------------

struct Item {
    int bait_seq, tile;        // sampled-from sequence; destination tile index
    int32_t bait_start;
    int64_t bait_end;
    int tile_seq;
    int32_t shift;             // tile_start - bait_start       (block_shift, :179)
    int32_t seq_len;           // seqlengths[tile_seq]
    int32_t lo, hi;            // hit window in canonical y
};

__device__ __forceinline__ Item load_item(const BootParams &P, int64_t r, int64_t b) {
    Item it;
    const int64_t flat = r * P.n_blocks + b;
    if (P.item_rec) {  // the rows kernel already drew the block and located its hit window
        const int4 rec = __ldg(P.item_rec + flat);
        it.lo = rec.x;
        it.hi = rec.y;
        it.bait_seq = rec.z;
        it.bait_start = rec.w;
        it.tile = P.dst_tile ? __ldg(P.dst_tile + flat) : (int)b;
        it.bait_end = (int64_t)it.bait_start + __ldg(P.bait_width + b) - 1;
        it.tile_seq = __ldg(P.tile_seq + it.tile);
        it.shift = __ldg(P.tile_start + it.tile) - it.bait_start;
        it.seq_len = (int32_t)__ldg(P.seqlen + it.tile_seq);
        return it;
    }
    if (P.draw_tab && !P.permute) {  // the rows kernel of the rank path already drew this block (bootstrap kinds: block b -> tile b)
        const int2 d = __ldg(P.draw_tab + flat);
        it.bait_seq = d.x;
        it.bait_start = d.y;
        it.tile = (int)b;
    } else if (P.rng_mode == NRB200_RNG_PHILOX) {
        draw_block(P, (uint64_t)(P.iter0 + r), b, it.bait_seq, it.bait_start);
        it.tile = P.dst_tile ? __ldg(P.dst_tile + flat) : (int)b;
    } else {
        it.bait_seq = __ldg(P.bait_seq + flat);
        it.bait_start = __ldg(P.bait_start + flat);
        it.tile = P.dst_tile ? __ldg(P.dst_tile + flat) : (int)b;
    }
    it.bait_end = (int64_t)it.bait_start + __ldg(P.bait_width + b) - 1;
    it.tile_seq = __ldg(P.tile_seq + it.tile);
    it.shift = __ldg(P.tile_start + it.tile) - it.bait_start;
    it.seq_len = (int32_t)__ldg(P.seqlen + it.tile_seq);
    locate_block(P.y, it.bait_seq, it.bait_start, it.bait_end, it.lo, it.hi);
    return it;
}

// shift() + trim() + the width filter for one candidate.  Returns false when the range is
// not a hit or is dropped.  (s2, e2) is the trimmed range; e2 < s2 marks a zero-width
// survivor of the within-chromosome variants.
__device__ __forceinline__ bool transform(const BootParams &P, const Item &it, int32_t st, int32_t en,
                                          int32_t &s2, int32_t &e2) {
    // bootstrap: any overlap with the bait block; permute: the tile that holds the start
    const bool hit = P.permute ? (st >= it.bait_start) : (en >= it.bait_start);
    if (!hit) return false;
    s2 = st + it.shift;
    e2 = en + it.shift;
    if (s2 > it.seq_len) {       // wholly beyond the sequence end -> [L + 1, L]
        if (P.drop_zero) return false;
        s2 = it.seq_len + 1;
        e2 = it.seq_len;
    } else {
        s2 = max(s2, 1);
        e2 = min(e2, it.seq_len);
    }
    return true;
}

template <bool STRANDED>
__device__ __forceinline__ bool excluded(const SliceView &X, int tile, int32_t s2, int32_t e2, int strand) {
    if (e2 < s2) return false;
    const int32_t a = __ldg(X.tile_lo + tile), b = __ldg(X.tile_hi + tile);
    for (int32_t k = a; k < b; ++k) {
        if (s2 <= __ldg(X.end + k) && e2 >= __ldg(X.start + k)) {
            if (!STRANDED || !X.strand || strands_match(strand, __ldg(X.strand + k))) return true;
        }
    }
    return false;
}

// Exclusion modes of the streaming kernels (template parameter EXCL)
constexpr int kExclNone = 0;  // no exclude set
constexpr int kExclDrop = 1;  // excludeOption = "drop": P.excl = exclude ranges, overlapping ranges vanish
constexpr int kExclTrim = 2;  // excludeOption = "trim": P.excl = gaps(exclude), ranges are cut to the gaps

// countOverlaps contribution of one (piece of a) bootstrapped range: features of the tile's slice
// with start <= pe, walked backwards under the running-max-of-end bound.
template <bool STRANDED>
__device__ __forceinline__ int count_piece(const BootParams &P, int64_t r, int32_t q_lo, int32_t q_hi, int32_t ps,
                                           int32_t pe, int strand) {
    int n = 0;
    for (int32_t k = first_gt(P.query.start, q_lo, q_hi, pe) - 1; k >= q_lo && __ldg(P.query.pmax + k) >= ps; --k) {
        const int32_t qe = __ldg(P.query.end + k);
        if (qe < ps) continue;
        if (STRANDED && P.query.strand && !strands_match(strand, __ldg(P.query.strand + k))) continue;
        if (P.minoverlap > 0 && min(pe, qe) - max(ps, __ldg(P.query.start + k)) + 1 < P.minoverlap) continue;
        ++n;
        if (P.counts) atomicAdd(P.counts + r * P.n_query + __ldg(P.query.src + k), 1);
    }
    return n;
}

template <bool STRANDED, bool HAS_QUERY, int EXCL>
__global__ void __launch_bounds__(256) boot_count_kernel(const BootParams P) {
    const int lane = threadIdx.x & (kWarp - 1);
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x / kWarp);
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x / kWarp) + (threadIdx.x / kWarp);
    const int64_t n_items = P.n_iter * P.n_blocks;

    for (int64_t item = warp0; item < n_items; item += warps_total) {
        const int64_t r = item / P.n_blocks, b = item - r * P.n_blocks;
        const Item it = load_item(P, r, b);

        int32_t q_lo = 0, q_hi = 0, x_lo = 0, x_hi = 0;
        if (HAS_QUERY) {
            q_lo = __ldg(P.query.tile_lo + it.tile);
            q_hi = __ldg(P.query.tile_hi + it.tile);
        }
        if (EXCL == kExclTrim) {
            x_lo = __ldg(P.excl.tile_lo + it.tile);
            x_hi = __ldg(P.excl.tile_hi + it.tile);
        }
        int rows = 0, hits = 0;
        int overlaps = 0;
        for (int32_t i = it.lo + lane; i < it.hi; i += kWarp) {
            const int32_t st = __ldg(P.y.start + i), en = __ldg(P.y.end + i);
            int32_t s2, e2;
            hits += (P.permute ? (st >= it.bait_start) : (en >= it.bait_start)) ? 1 : 0;
            if (!transform(P, it, st, en, s2, e2)) continue;
            const int strand = (STRANDED && P.y.strand) ? __ldg(P.y.strand + i) : 0;
            if (EXCL == kExclTrim) {
                if (e2 < s2) continue;  // a zero-width survivor overlaps no gap
                for (int32_t k = x_lo; k < x_hi; ++k) {
                    const int32_t gs = __ldg(P.excl.start + k), ge = __ldg(P.excl.end + k);
                    if (s2 > ge || e2 < gs) continue;
                    ++rows;
                    if (HAS_QUERY) overlaps += count_piece<STRANDED>(P, r, q_lo, q_hi, max(s2, gs), min(e2, ge), strand);
                }
                continue;
            }
            if (EXCL == kExclDrop && excluded<STRANDED>(P.excl, it.tile, s2, e2, strand)) continue;
            ++rows;
            if (HAS_QUERY && e2 >= s2) overlaps += count_piece<STRANDED>(P, r, q_lo, q_hi, s2, e2, strand);
        }
        rows = __reduce_add_sync(kFull, rows);
        hits = __reduce_add_sync(kFull, hits);
        if (HAS_QUERY) overlaps = __reduce_add_sync(kFull, overlaps);
        if (lane == 0) {
            if (rows) atomicAdd(P.iter_nranges + r, (unsigned long long)rows);
            if (HAS_QUERY && overlaps) atomicAdd(P.iter_totals + r, (unsigned long long)overlaps);
            if (P.item_rows) P.item_rows[item] = rows;
            if (P.n_hits && hits) atomicAdd(P.n_hits, (unsigned long long)hits);
        }
    }
}

// ---- materialising path ---------------------------------------------------------------

// One CTA per iteration: exclusive scan of the rows-per-block of that iteration.
__global__ void __launch_bounds__(256) scan_iteration_kernel(const int32_t *__restrict__ item_rows, int64_t n_blocks,
                                                             int32_t *__restrict__ within_off) {
    __shared__ int warp_sums[8];
    __shared__ int carry;
    const int64_t r = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += blockDim.x) {
        const int64_t b = base + threadIdx.x;
        const int v = b < n_blocks ? item_rows[r * n_blocks + b] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(kFull, x, d);
            if (lane >= d) x += t;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        int before = carry;
        for (int w = 0; w < warp; ++w) before += warp_sums[w];
        if (b < n_blocks) within_off[r * n_blocks + b] = before + x - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + x;
        __syncthreads();
    }
}

struct RowsOut {
    int32_t *seq, *start, *end, *src, *block;
    const int64_t *iter_off;      // [n_iter + 1]
    const int32_t *within_off;    // [n_iter x n_blocks]
    const int32_t *item_rows;     // [n_iter x n_blocks] rows the count pass promised for the item
    int *overflow;                // set when a fill pass finds more rows than promised (never written past the promise)
};

// grid = (block groups, iterations); one warp per (iteration, block), 64 candidate ranges per
// loop trip (two independent loads per lane), ballot-compacted writes of the five columns.
template <bool STRANDED, int EXCL>
__global__ void __launch_bounds__(256) boot_fill_kernel(const BootParams P, const RowsOut O) {
    const int lane = threadIdx.x & (kWarp - 1);
    const unsigned below = (1u << lane) - 1u;
    const int64_t r = blockIdx.y;
    const int64_t row0 = __ldg(O.iter_off + r);
    for (int64_t b = (int64_t)blockIdx.x * (blockDim.x / kWarp) + (threadIdx.x / kWarp); b < P.n_blocks;
         b += (int64_t)gridDim.x * (blockDim.x / kWarp)) {
        const Item it = load_item(P, r, b);
        int64_t out = row0 + __ldg(O.within_off + r * P.n_blocks + b);
        const int64_t out_end = out + __ldg(O.item_rows + r * P.n_blocks + b);
        for (int32_t base = it.lo; base < it.hi; base += 2 * kWarp) {
            const int32_t i0 = base + lane, i1 = i0 + kWarp;
            const bool in0 = i0 < it.hi, in1 = i1 < it.hi;
            int32_t st0 = 0, en0 = 0, st1 = 0, en1 = 0;
            if (in0) { st0 = __ldg(P.y.start + i0); en0 = __ldg(P.y.end + i0); }
            if (in1) { st1 = __ldg(P.y.start + i1); en1 = __ldg(P.y.end + i1); }
            int32_t s0 = 0, e0 = 0, s1 = 0, e1 = 0;
            bool keep0 = in0 && transform(P, it, st0, en0, s0, e0);
            bool keep1 = in1 && transform(P, it, st1, en1, s1, e1);
            if (EXCL == kExclDrop) {
                if (keep0) keep0 = !excluded<STRANDED>(P.excl, it.tile, s0, e0, (STRANDED && P.y.strand) ? __ldg(P.y.strand + i0) : 0);
                if (keep1) keep1 = !excluded<STRANDED>(P.excl, it.tile, s1, e1, (STRANDED && P.y.strand) ? __ldg(P.y.strand + i1) : 0);
            }
            const unsigned m0 = __ballot_sync(kFull, keep0), m1 = __ballot_sync(kFull, keep1);
            if (out + __popc(m0) + __popc(m1) > out_end) {  // the two passes disagree: refuse to write
                if (lane == 0) atomicExch(O.overflow, 1);
                break;
            }
            if (keep0) {
                const int64_t pos = out + __popc(m0 & below);
                O.seq[pos] = it.tile_seq;
                O.start[pos] = s0;
                O.end[pos] = e0;
                O.src[pos] = P.y.src ? __ldg(P.y.src + i0) : i0;
                O.block[pos] = (int32_t)b;
            }
            out += __popc(m0);
            if (keep1) {
                const int64_t pos = out + __popc(m1 & below);
                O.seq[pos] = it.tile_seq;
                O.start[pos] = s1;
                O.end[pos] = e1;
                O.src[pos] = P.y.src ? __ldg(P.y.src + i1) : i1;
                O.block[pos] = (int32_t)b;
            }
            out += __popc(m1);
        }
    }
}

// excludeOption = "trim": a range becomes one row per gap it overlaps (its intersection with the
// gap), rows of a range in ascending gap order.  Lanes count their pieces, a warp scan places them.
__global__ void __launch_bounds__(256) boot_fill_trim_kernel(const BootParams P, const RowsOut O) {
    const int lane = threadIdx.x & (kWarp - 1);
    const int64_t r = blockIdx.y;
    const int64_t row0 = __ldg(O.iter_off + r);
    for (int64_t b = (int64_t)blockIdx.x * (blockDim.x / kWarp) + (threadIdx.x / kWarp); b < P.n_blocks;
         b += (int64_t)gridDim.x * (blockDim.x / kWarp)) {
        const Item it = load_item(P, r, b);
        const int32_t x_lo = __ldg(P.excl.tile_lo + it.tile), x_hi = __ldg(P.excl.tile_hi + it.tile);
        int64_t out = row0 + __ldg(O.within_off + r * P.n_blocks + b);
        const int64_t out_end = out + __ldg(O.item_rows + r * P.n_blocks + b);
        for (int32_t base = it.lo; base < it.hi; base += kWarp) {
            const int32_t i = base + lane;
            int32_t s2 = 0, e2 = -1;
            int n = 0;
            if (i < it.hi && transform(P, it, __ldg(P.y.start + i), __ldg(P.y.end + i), s2, e2) && e2 >= s2) {
                for (int32_t k = x_lo; k < x_hi; ++k) n += (s2 <= __ldg(P.excl.end + k) && e2 >= __ldg(P.excl.start + k));
            }
            int before = n;  // inclusive warp scan
#pragma unroll
            for (int d = 1; d < kWarp; d <<= 1) {
                const int t = __shfl_up_sync(kFull, before, d);
                if (lane >= d) before += t;
            }
            const int total = __shfl_sync(kFull, before, kWarp - 1);
            if (out + total > out_end) {
                if (lane == 0) atomicExch(O.overflow, 1);
                break;
            }
            if (n) {
                int64_t pos = out + (before - n);
                const int32_t src = P.y.src ? __ldg(P.y.src + i) : i;
                for (int32_t k = x_lo; k < x_hi; ++k) {
                    const int32_t gs = __ldg(P.excl.start + k), ge = __ldg(P.excl.end + k);
                    if (s2 > ge || e2 < gs) continue;
                    O.seq[pos] = it.tile_seq;
                    O.start[pos] = max(s2, gs);
                    O.end[pos] = min(e2, ge);
                    O.src[pos] = src;
                    O.block[pos] = (int32_t)b;
                    ++pos;
                }
            }
            out += total;
        }
    }
}

// ---- rank path: counting without streaming the ranges ------------------------------------
//
// For a bootstrap block sampled at [bs, be] on sequence c and moved by sh = tile_start - bs, the
// ranges that land on feature [gs, ge] of the tile's sequence (length L) are exactly
//     { i on c : start_i <= A  and  end_i >= Bv },   A = min(be, ge - sh, L - sh),  Bv = max(bs, gs - sh)
// (hit of the bait block, overlap after shift(), and start + sh <= L so trim() leaves a positive
// width; trim() never changes whether a positive-width range overlaps a feature inside [1, L]).
// When Bv <= A that set has  #{start <= A} - #{end < Bv}  members: two rank queries.  When A < Bv
// (the feature lies beside the tile, reached only by ranges hanging over the block edge) the
// members all contain [A, Bv]; they are enumerated backwards from #{start <= A} under the
// running-max-of-end bound.  Same arithmetic as the reference's findOverlaps + shift + trim +
// countOverlaps (R/bootUnsegmented.R:133-139, :175-191; vignettes/bootRanges.Rmd:497-503).

// Exclusion (excludeOption = "drop", R/bootRanges.R:108-109) stays on this path: the ranges
// dropped because they overlap exclude region X and that also land on feature g are those that
// overlap the pseudo-feature [max(xs, gs), min(xe, ge)] (when that is empty: those that span the gap),
// so they are counted by the same formula and SUBTRACTED.  Exclude regions are merged first
// (disjoint, sorted), so a range overlaps a run of consecutive regions and the union is
// sum_k |S_k| - sum_k |S_k and S_k+1|.  A feature therefore owns a list of signed windows per tile.

struct PairRec {               // one signed window to count for a feature on a tile (32 bytes)
    int32_t tile;              // block index b (row of the draw table)
    int32_t ws, we;            // window on the tile's sequence ([gs, ge] or an exclusion pseudo-feature)
    int32_t sign;              // +1 / -1
    int32_t tile_start, width, seq_len;  // copied from the block tables: saves three dependent loads
    int32_t cls;               // strand class of y the window counts (0 when y is unstranded)
};
static_assert(sizeof(PairRec) == 32, "PairRec is read as two int4");

struct FeatView {              // query features in visit order (grouped by list length)
    const int32_t *src;        // [n_query] index in the caller's order
    const int32_t *pair_off;   // [n_query + 1]
    const PairRec *pair_rec;
};

struct TileExcl {              // signed exclusion windows of each (tile, strand class), for the rows kernel
    const int32_t *off;        // [n_blocks * n_cls + 1], or nullptr without exclusion
    const int4 *rec;           // {ws, we, sign, 0}
};

// One (feature, tile, iteration) evaluation, split into stages so that the RPT evaluations a thread
// owns can be issued stage by stage: all table probes first, then all value reads, so that the
// dependent-load chain (draw -> sequence info -> table -> values) is paid once per thread, not per
// evaluation.  All coordinates are below 2^30 and widths at most 2^30: every sum fits int32.
struct PairProbe {
    int4 si;                   // seq_info of the bait sequence
    int32_t A, Bv;             // count = #{start <= A and end >= Bv}
    int32_t bs;                // bait start (the permute variants also need start >= bs)
    int32_t key_a, key_b;      // clamped search keys (A + 1 in starts, Bv in sorted ends)
    int32_t a0, an, b0, bn;    // bucket start and (capped) size in start[] and end_sorted[]
};

__device__ __forceinline__ void pair_keys(PairProbe &q, int2 draw, int32_t width, int32_t ts, int32_t seq_len,
                                          int32_t gs, int32_t ge, int rank_shift, int32_t end_shift, int32_t &ua, int32_t &ub) {
    const int32_t bs = draw.y;
    const int32_t sh = ts - bs;
    q.bs = bs;
    q.A = min(bs + (width - 1), min(ge, seq_len) - sh);
    q.Bv = max(bs, gs - sh);
    q.key_a = rank_clamp_key(q.A + 1, q.si);
    q.key_b = rank_clamp_key(q.Bv - end_shift, q.si);
    ua = q.si.w + (q.key_a >> rank_shift);
    ub = q.si.w + (q.key_b >> rank_shift);
}

// rank inside one bucket: first index with vals >= key (first three values already read; rare search)
__device__ __forceinline__ int32_t bucket_rank(const int32_t *__restrict__ vals, const uint32_t *__restrict__ table, int32_t u,
                                               const int4 si, int32_t a, int32_t n, int32_t key, int32_t v0, int32_t v1,
                                               int32_t v2) {
    if (v2 < key) {
        const int32_t b = n < kRankCountCap ? a + n : si.x + (int32_t)(__ldg(table + u + 1) >> kRankCountBits);
        return first_ge(vals, a + 3, b, key);
    }
    return a + (v0 < key) + (v1 < key) + (v2 < key);
}

// #{start <= a and end >= b} for a < b: such ranges contain [a, b]; walk back from #{start <= a}
// (index ia) under the running-max-of-end bound.  Usually zero or one step.
__device__ __forceinline__ int count_spanning(const RangesView &y, const int4 si, int32_t ia, int32_t b) {
    if (b > si.z) return 0;
    int n = 0;
    for (int32_t k = ia - 1; k >= si.x && __ldg(y.rpmax + k) >= b; --k) n += __ldg(y.rend + k) >= b;
    return n;
}

// PERMUTE: a range belongs to the ONE tile that holds its start (findOverlaps(select = "first"),
// R/bootUnsegmented.R:95, :155), so the members are {bs <= start <= A and end >= Bv}: the bootstrap
// count minus the ranges that start before the tile and reach Bv (they contain [bs - 1, Bv]).
template <bool PERMUTE>
__device__ __forceinline__ int pair_finish(const RangesView &y, const PairProbe &q, int32_t ia, int32_t ib) {
    if (q.A < 1 || (PERMUTE && q.A < q.bs)) return 0;  // permute: no start can lie in [bs, A]
    int n = q.Bv <= q.A ? ia - ib : count_spanning(y, q.si, ia, q.Bv);  // second form: the feature lies beside the tile
    if (PERMUTE && n > 0) n -= count_spanning(y, q.si, first_ge_ranked(y.rstart, y.rrank_start, y.rank_shift, q.si, q.bs), q.Bv);
    return n;
}

template <int RPT, bool PERMUTE = false>
__device__ __forceinline__ void count_pairs(const RangesView &y, const int2 (&d)[RPT], int cls, int32_t width, int32_t ts,
                                            int32_t seq_len, int32_t gs, int32_t ge, int (&cnt)[RPT]) {
    PairProbe q[RPT];
    int32_t ua[RPT], ub[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) q[j].si = __ldg(y.rseq_info + d[j].x * y.n_cls + cls);
#pragma unroll
    for (int j = 0; j < RPT; ++j) pair_keys(q[j], d[j], width, ts, seq_len, gs, ge, y.rank_shift, y.end_shift, ua[j], ub[j]);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const uint32_t ea = __ldg(y.rrank_start + ua[j]), eb = __ldg(y.rank_end + ub[j]);
        q[j].a0 = q[j].si.x + (int32_t)(ea >> kRankCountBits);
        q[j].an = (int32_t)(ea & kRankCountCap);
        q[j].b0 = q[j].si.x + (int32_t)(eb >> kRankCountBits);
        q[j].bn = (int32_t)(eb & kRankCountCap);
    }
    int32_t va[RPT][3], vb[RPT][3];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            va[j][k] = k < q[j].an ? __ldg(y.rstart + q[j].a0 + k) : INT32_MAX;
            vb[j][k] = k < q[j].bn ? __ldg(y.end_sorted + q[j].b0 + k) : INT32_MAX;
        }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int32_t ia = bucket_rank(y.rstart, y.rrank_start, ua[j], q[j].si, q[j].a0, q[j].an, q[j].key_a, va[j][0], va[j][1], va[j][2]);
        const int32_t ib = bucket_rank(y.end_sorted, y.rank_end, ub[j], q[j].si, q[j].b0, q[j].bn, q[j].key_b, vb[j][0], vb[j][1], vb[j][2]);
        cnt[j] += pair_finish<PERMUTE>(y, q[j], ia, ib);
    }
}

// Thread = (feature, RPT consecutive iterations); the block draws come from the table the
// block-rows kernel wrote.  Features are visited in F's order (grouped by tile count, so a warp's
// lanes loop the same number of times).  No atomics on the count matrix, no memset: every element
// is written exactly once.
template <int RPT, int MINB>
__global__ void __launch_bounds__(256, MINB) boot_rank_count_kernel(const BootParams P, const FeatView F) {
    __shared__ int warp_tot[8][RPT];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * RPT;
    const int n_live = (int)min((int64_t)RPT, P.n_iter - r0);
    int cnt[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) cnt[j] = 0;
    if (g < P.n_query) {
        const int32_t src = __ldg(F.src + g);
        const int32_t p1 = __ldg(F.pair_off + g + 1);
        for (int32_t p = __ldg(F.pair_off + g); p < p1; ++p) {
            const int4 ra = __ldg(reinterpret_cast<const int4 *>(F.pair_rec + p));      // tile, ws, we, sign
            const int4 rb = __ldg(reinterpret_cast<const int4 *>(F.pair_rec + p) + 1);  // tile_start, width, seq_len
            const int2 *__restrict__ tab = P.draw_tab + r0 * P.n_blocks + ra.x;
            int2 d[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) d[j] = __ldg(tab + (j < n_live ? j : 0) * P.n_blocks);  // tail: recount row 0, discarded
            int c[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) c[j] = 0;
            if (P.permute) count_pairs<RPT, true>(P.y, d, rb.w, rb.y, rb.x, rb.z, ra.y, ra.z, c);
            else count_pairs<RPT, false>(P.y, d, rb.w, rb.y, rb.x, rb.z, ra.y, ra.z, c);
#pragma unroll
            for (int j = 0; j < RPT; ++j) cnt[j] += ra.w * c[j];
        }
        if (P.counts) {
#pragma unroll
            for (int j = 0; j < RPT; ++j)
                if (j < n_live) P.counts[(r0 + j) * P.n_query + src] = cnt[j];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int s = __reduce_add_sync(kFull, cnt[j]);
        if (lane == 0) warp_tot[warp][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < n_live) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += (unsigned)warp_tot[w][threadIdx.x];
        if (s) atomicAdd(P.iter_totals + r0 + threadIdx.x, s);
    }
}

// ---- weighted sums on the rank path -----------------------------------------------------------
// sum of a numeric column of y over the ranges that land on a feature (vignettes/bootRanges.Rmd:532-540:
// summarize(sum(signal)) per (feature, iter)).  Same windows and ranks as the counts; instead of
// #{start <= A} - #{end < Bv} it takes the difference of two running sums of the weight, one along the
// start order and one along the sorted-end order (both restart at every virtual sequence).
struct WeightView {
    const double *w;           // weights in the rank path's (class-major) start order
    const double *ps;          // inclusive running sum of w inside each virtual sequence
    const double *pe;          // the same along the sorted-end order
};

__device__ __forceinline__ double run_sum(const double *__restrict__ p, int32_t first, int32_t idx) {
    return idx > first ? __ldg(p + idx - 1) : 0.0;  // sum of the first (idx - first) elements of the sequence
}
__device__ __forceinline__ double sum_spanning(const RangesView &y, const WeightView &W, const int4 si, int32_t ia, int32_t b) {
    if (b > si.z) return 0.0;
    double s = 0.0;
    for (int32_t k = ia - 1; k >= si.x && __ldg(y.rpmax + k) >= b; --k)
        if (__ldg(y.rend + k) >= b) s += __ldg(W.w + k);
    return s;
}

template <bool PERMUTE>
__device__ __forceinline__ double weighted_pair(const RangesView &y, const WeightView &W, int2 draw, int cls, int32_t width,
                                                int32_t ts, int32_t seq_len, int32_t ws, int32_t we) {
    PairProbe q;
    int32_t ua, ub;
    q.si = __ldg(y.rseq_info + draw.x * y.n_cls + cls);
    pair_keys(q, draw, width, ts, seq_len, ws, we, y.rank_shift, y.end_shift, ua, ub);
    if (q.A < 1 || (PERMUTE && q.A < q.bs)) return 0.0;
    const int32_t ia = first_ge_ranked(y.rstart, y.rrank_start, y.rank_shift, q.si, q.A + 1);
    double s;
    if (q.Bv <= q.A) {
        const int32_t ib = first_ge_ranked(y.end_sorted, y.rank_end, y.rank_shift, q.si, q.Bv - y.end_shift);
        s = run_sum(W.ps, q.si.x, ia) - run_sum(W.pe, q.si.x, ib);
    } else {
        s = sum_spanning(y, W, q.si, ia, q.Bv);
    }
    if (PERMUTE) s -= sum_spanning(y, W, q.si, first_ge_ranked(y.rstart, y.rrank_start, y.rank_shift, q.si, q.bs), q.Bv);
    return s;
}

// Thread = (feature, iteration), like boot_rank_count_kernel<1>; sums[r, g] written once, iter_sums by atomics.
__global__ void __launch_bounds__(256) boot_rank_wsum_kernel(const BootParams P, const FeatView F, const WeightView W,
                                                             double *__restrict__ sums, double *__restrict__ iter_sums) {
    __shared__ double warp_tot[8];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    double acc = 0.0;
    if (g < P.n_query) {
        const int32_t src = __ldg(F.src + g);
        const int32_t p1 = __ldg(F.pair_off + g + 1);
        for (int32_t p = __ldg(F.pair_off + g); p < p1; ++p) {
            const int4 ra = __ldg(reinterpret_cast<const int4 *>(F.pair_rec + p));
            const int4 rb = __ldg(reinterpret_cast<const int4 *>(F.pair_rec + p) + 1);
            const int2 d = __ldg(P.draw_tab + r * P.n_blocks + ra.x);
            const double v = P.permute ? weighted_pair<true>(P.y, W, d, rb.w, rb.y, rb.x, rb.z, ra.y, ra.z)
                                       : weighted_pair<false>(P.y, W, d, rb.w, rb.y, rb.x, rb.z, ra.y, ra.z);
            acc += ra.w * v;
        }
        if (sums) sums[r * P.n_query + src] = acc;
    }
    double t = acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += warp_tot[w];
        if (s != 0.0) atomicAdd(iter_sums + r, s);
    }
}

// gathers for the weight arrays: out[i] = in[order[i]] (order == nullptr: identity)
__global__ void gather_f64_kernel(const double *__restrict__ in, const int32_t *__restrict__ order, int64_t n, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[order ? order[i] : i];
}
// sort keys (virtual sequence << 32 | end) with the position as payload: the sorted-end order of the weights
__global__ void pack_vend_pairs_kernel(const unsigned *__restrict__ vkeys, const int32_t *__restrict__ seqid, const int32_t *__restrict__ end,
                                       int64_t n, unsigned long long *__restrict__ keys, int32_t *__restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned v = vkeys ? vkeys[i] : (unsigned)seqid[i];
        keys[i] = ((unsigned long long)v << 32) | (unsigned)end[i];
        idx[i] = (int32_t)i;
    }
}
__global__ void high32_kernel(const unsigned long long *__restrict__ keys, int64_t n, unsigned *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned)(keys[i] >> 32);
}
struct AddF64 {
    __host__ __device__ __forceinline__ double operator()(double a, double b) const { return a + b; }
};

// Per (iteration, block): draws the block (Philox) or reads the caller's draw, records it in the
// draw table for the count kernel, and counts by ranks the rows the reference's table would hold:
// hits of the bait block minus (where the variant filters width > 0, R/bootUnsegmented.R:189) the
// hits shifted wholly beyond the sequence end.  Requires tile_start <= seqlength for every tile
// (checked on the host).
template <int RPT>
__global__ void __launch_bounds__(256) boot_block_rows_kernel(const BootParams P, const TileExcl X) {
    __shared__ int warp_rows[8][RPT];
    __shared__ unsigned long long cta_hits;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * RPT;
    if (threadIdx.x == 0) cta_hits = 0;
    int rows[RPT];
    int hits_sum = 0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) rows[j] = 0;
    if (b < P.n_blocks) {
        const int32_t width = __ldg(P.bait_width + b);
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            if (r0 + j < P.n_iter) {
                const int64_t flat = (r0 + j) * P.n_blocks + b;
                int2 d;
                if (P.rng_mode == NRB200_RNG_PHILOX) {
                    draw_block(P, (uint64_t)(P.iter0 + r0 + j), b, d.x, d.y);
                } else {
                    d.x = __ldg(P.bait_seq + flat);
                    d.y = __ldg(P.bait_start + flat);
                }
                // bootstrap kinds: block b lands on tile b; permute kinds: on the tile the caller drew
                const int32_t t = P.dst_tile ? __ldg(P.dst_tile + flat) : (int32_t)b;
                const int32_t ts = __ldg(P.tile_start + t);
                const int32_t seq_len = (int32_t)__ldg(P.seqlen + __ldg(P.tile_seq + t));
                if (P.draw_tab) P.draw_tab[(r0 + j) * P.n_blocks + t] = d;  // the count kernel looks draws up by tile
                if (P.item_rec) {  // materialising: the fill pass streams canonical window [lo, hi) of this block
                    int32_t lo, hi;
                    locate_block(P.y, d.x, d.y, (int64_t)d.y + width - 1, lo, hi);
                    P.item_rec[flat] = make_int4(lo, hi, d.x, d.y);
                }
                int hits = 0;
                for (int cls = 0; cls < P.y.n_cls; ++cls) {  // one pass per strand class of y (one when y is unstranded)
                    const int4 si = __ldg(P.y.rseq_info + d.x * P.y.n_cls + cls);
                    const int32_t ie = idx_start_le(P.y, si, d.y + (width - 1));
                    // hits: overlap of the bait block (bootstrap) / start inside the tile (permute)
                    const int h1 = ie - (P.permute ? idx_start_le(P.y, si, d.y - 1) : idx_end_lt(P.y, si, d.y));
                    int dropped = 0;  // only a tile that overhangs its sequence end can push hits beyond it
                    if (P.drop_zero && ts + (width - 1) > seq_len) dropped = max(0, ie - idx_start_le(P.y, si, seq_len - (ts - d.y)));
                    hits += h1;
                    if (!P.rows_windows_only) rows[j] += h1 - dropped;
                    if (X.off) {  // minus the hits that overlap an exclude region after the move ("trim": plus the pieces)
                        const int32_t x1 = __ldg(X.off + t * P.y.n_cls + cls + 1);
                        for (int32_t k = __ldg(X.off + t * P.y.n_cls + cls); k < x1; ++k) {
                            const int4 w = __ldg(X.rec + k);
                            const int2 dd[1] = {d};
                            int c1[1] = {0};
                            if (P.permute) count_pairs<1, true>(P.y, dd, cls, width, ts, seq_len, w.x, w.y, c1);
                            else count_pairs<1, false>(P.y, dd, cls, width, ts, seq_len, w.x, w.y, c1);
                            rows[j] += w.z * c1[0];
                        }
                    }
                }
                hits_sum += hits;
                if (P.item_rows) P.item_rows[(r0 + j) * P.n_blocks + b] = rows[j];
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int s = __reduce_add_sync(kFull, rows[j]);
        if (lane == 0) warp_rows[warp][j] = s;
    }
    hits_sum = __reduce_add_sync(kFull, hits_sum);
    __syncthreads();
    if (lane == 0 && hits_sum) atomicAdd(&cta_hits, (unsigned long long)hits_sum);
    if (threadIdx.x < RPT && r0 + threadIdx.x < P.n_iter) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += (unsigned)warp_rows[w][threadIdx.x];
        if (s) atomicAdd(P.iter_nranges + r0 + threadIdx.x, s);
    }
    __syncthreads();
    if (threadIdx.x == 0 && P.n_hits && cta_hits) atomicAdd(P.n_hits, cta_hits);
}

// (seqid << 32 | end) sort keys and their unpacking: the sorted-ends array of the rank path
__global__ void pack_end_keys_kernel(const int32_t *__restrict__ seqid, const int32_t *__restrict__ end, int64_t n,
                                     unsigned long long *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((unsigned long long)(unsigned)seqid[i] << 32) | (unsigned)end[i];
}
__global__ void unpack_end_keys_kernel(const unsigned long long *__restrict__ keys, int64_t n, int32_t *__restrict__ end_sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) end_sorted[i] = (int32_t)(unsigned)keys[i];
}

// ---- index construction ---------------------------------------------------------------

// One pass over the caller's y: argument checks of nrb200_set_ranges, canonical-order test,
// strandedness, widest range, and per-sequence counts / largest end (warp-aggregated atomics).
struct InspectOut {
    int32_t err;       // bit mask: 1 seqid, 2 start < 1, 4 end < start, 8 end >= 2^30, 16 strand code
    int32_t unsorted;  // some neighbour pair is out of (seqid, start, end) order
    int32_t stranded;  // some strand is '+' or '-'
    int32_t wmax;      // widest range
    int32_t wmin;      // narrowest range (initialise to INT32_MAX)
};

__global__ void __launch_bounds__(256) inspect_ranges_kernel(const int32_t *__restrict__ seqid, const int32_t *__restrict__ start,
                                                             const int32_t *__restrict__ end, const int8_t *__restrict__ strand,
                                                             int64_t n, int n_seq, InspectOut *__restrict__ out,
                                                             unsigned *__restrict__ seq_count, int32_t *__restrict__ seq_max_end) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int err = 0, unsorted = 0, stranded = 0, wmax = 0, wmin = INT32_MAX;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += stride) {  // warp-uniform trip count
        const int64_t i = base + threadIdx.x;
        const bool live = i < n;
        int c = -1;
        int32_t e = 0;
        if (live) {
            c = seqid[i];
            const int32_t s = start[i];
            e = end[i];
            if (c < 0 || c >= n_seq) { err |= 1; c = -1; }
            if (s < 1) err |= 2;
            if (e < s) err |= 4;
            if (e >= (1 << 30)) err |= 8;
            if (strand) {
                const int t = strand[i];
                if (t < 0 || t > 2) err |= 16;
                stranded |= t != 0;
            }
            wmax = max(wmax, e - s + 1);
            wmin = min(wmin, e - s + 1);
            if (i > 0) {
                const int32_t pc = seqid[i - 1], ps = start[i - 1], pe = end[i - 1];
                if (c < pc || (c == pc && (s < ps || (s == ps && e < pe)))) unsorted = 1;
            }
        }
        const unsigned active = __ballot_sync(kFull, c >= 0);
        if (c >= 0) {
            const unsigned peers = __match_any_sync(active, c);
            const int32_t m = __reduce_max_sync(peers, e);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) {
                atomicAdd(seq_count + c, (unsigned)__popc(peers));
                atomicMax(seq_max_end + c, m);
            }
        }
    }
    err = __reduce_or_sync(kFull, err);
    unsorted = __reduce_or_sync(kFull, unsorted);
    stranded = __reduce_or_sync(kFull, stranded);
    wmax = __reduce_max_sync(kFull, wmax);
    wmin = __reduce_min_sync(kFull, wmin);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out->wmin, wmin);
        if (err) atomicOr(&out->err, err);
        if (unsorted) atomicOr(&out->unsorted, 1);
        if (stranded) atomicOr(&out->stranded, 1);
        atomicMax(&out->wmax, wmax);
    }
}

struct MaxOp {
    __host__ __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

// ---- device-side permutations (production mode of the permute variants) -------------------
// Image of sample(n) (R/bootUnsegmented.R:97, :156): tile b of local iteration r gets the sort key
// (r << 52) | (sequence << 40 -- within-chromosome variant only) | 40 Philox bits (stream 1); a stable
// radix sort groups each iteration (and sequence) and orders its tiles at random; block b moves to the
// tile whose index equals its position inside the iteration.
__global__ void permute_keys_kernel(uint64_t seed, int64_t iter0, int64_t n_iter, int64_t n_blocks, const int32_t *__restrict__ tile_seq,
                                    int within, unsigned long long *__restrict__ keys, int32_t *__restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_iter * n_blocks) return;
    const int64_t r = i / n_blocks, b = i - r * n_blocks;
    const Philox128 u = philox_draw(seed, (uint64_t)(iter0 + r), (uint32_t)b, 1u);
    const unsigned long long chrom = within ? (unsigned long long)tile_seq[b] : 0ull;
    keys[i] = ((unsigned long long)r << 52) | (chrom << 40) | (u.lo >> 24);
    vals[i] = (int32_t)b;
}
__global__ void permute_scatter_kernel(const int32_t *__restrict__ sorted_vals, int64_t n_iter, int64_t n_blocks,
                                       int32_t *__restrict__ dst_tile) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_iter * n_blocks) return;
    const int64_t r = p / n_blocks, q = p - r * n_blocks;
    dst_tile[r * n_blocks + sorted_vals[p]] = (int32_t)q;
}

// ---- device-side draws for proportionLength = FALSE (production mode) ----------------------
// Image of seg_bootstrap_noprop (R/bootRanges.R:256-331), three steps per batch of iterations:
// (1) B global draws: segment in proportion to width (states in order, segments in seg order inside a
// state), start round(U(start_j, start_j + (tiles_j - 1) L_b)); (2) per iteration, the draws of each state
// in draw order (stable compaction, one CTA per iteration); (3) tile k of state s takes the k-th draw of
// its state when there are enough, else draw number floor(U * rand_n) (the image of
// sample(rand_n, rearr_n, replace = TRUE)); a state without draws raises the error flag.
__global__ void noprop_draw_kernel(const SamplerView S, uint64_t seed, int64_t iter0, int64_t n_iter, int64_t n_blocks,
                                   int32_t *__restrict__ dseg, int32_t *__restrict__ dstart, int8_t *__restrict__ dstate) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_iter * n_blocks) return;
    const int64_t r = t / n_blocks, i = t - r * n_blocks;
    const Philox128 u = philox_draw(seed, (uint64_t)(iter0 + r), (uint32_t)i, 0u);
    int64_t total = 0;
    for (int st = 1; st <= S.n_states; ++st) total += S.state_len[st];
    int64_t pos = (int64_t)mul_hi_u64(u.lo, (uint64_t)total);
    int st = 1;
    while (st < S.n_states && pos >= S.state_len[st]) pos -= S.state_len[st++];
    const int64_t first = S.state_seg_off[st];
    const int n = (int)(S.state_seg_off[st + 1] - first);
    const int64_t j = first + pick_by_length(S.sg_cum + first, n, pos);
    dseg[t] = (int32_t)j;
    dstart[t] = (int32_t)rounded_uniform(u.hi, S.sg_start[j], (int64_t)(S.sg_tiles[j] - 1) * S.block_length);
    dstate[t] = (int8_t)st;
}

// one CTA per iteration: list[r][.] = draw indices grouped by state (ascending inside a state), cnt[r][s]
__global__ void __launch_bounds__(256) noprop_lists_kernel(int n_states, int64_t n_blocks, const int8_t *__restrict__ dstate,
                                                           int32_t *__restrict__ list, int32_t *__restrict__ cnt) {
    __shared__ int warp_sums[8];
    __shared__ int carry;
    const int64_t r = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int st = 1; st <= n_states; ++st) {
        const int begin = carry;
        __syncthreads();
        for (int64_t base = 0; base < n_blocks; base += blockDim.x) {
            const int64_t i = base + threadIdx.x;
            const int v = (i < n_blocks && dstate[r * n_blocks + i] == st) ? 1 : 0;
            int x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, x, d);
                if (lane >= d) x += t;
            }
            if (lane == 31) warp_sums[warp] = x;
            __syncthreads();
            int before = carry;
            for (int w = 0; w < warp; ++w) before += warp_sums[w];
            if (v) list[r * n_blocks + before + x - 1] = (int32_t)i;
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) carry = before + x;
            __syncthreads();
        }
        if (threadIdx.x == 0) cnt[r * (n_states + 1) + st] = carry - begin;
        __syncthreads();
    }
}

__global__ void noprop_assign_kernel(const SamplerView S, uint64_t seed, int64_t iter0, int64_t n_iter, int64_t n_blocks,
                                     const int32_t *__restrict__ dseg, const int32_t *__restrict__ dstart,
                                     const int32_t *__restrict__ list, const int32_t *__restrict__ cnt,
                                     int32_t *__restrict__ bait_seq, int32_t *__restrict__ bait_start, int *__restrict__ err) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_iter * n_blocks) return;
    const int64_t r = t / n_blocks, b = t - r * n_blocks;
    int st = 1;
    while (st < S.n_states && b >= S.state_block_off[st + 1]) ++st;
    const int64_t k = b - S.state_block_off[st], rearr_n = S.state_block_off[st + 1] - S.state_block_off[st];
    const int32_t *c = cnt + r * (S.n_states + 1);
    int64_t off = 0;
    for (int s2 = 1; s2 < st; ++s2) off += c[s2];
    const int64_t rand_n = c[st];
    if (rand_n == 0) {
        atomicExch(err, 1);
        bait_seq[t] = 0;
        bait_start[t] = 1;
        return;
    }
    int64_t m = k;
    if (rand_n < rearr_n) m = (int64_t)mul_hi_u64(philox_draw(seed, (uint64_t)(iter0 + r), (uint32_t)b, 2u).lo, (uint64_t)rand_n);
    const int32_t i = list[r * n_blocks + off + m];
    bait_seq[t] = S.sg_seq[dseg[r * n_blocks + i]];
    bait_start[t] = dstart[r * n_blocks + i];
}

// ---- strand-class grouping of y for the rank path -----------------------------------------
// virtual sequence of a range: seq * n_cls + class, class = strand class * n_width + (narrower than minoverlap)
__global__ void vseq_keys_kernel(const int32_t *__restrict__ seqid, const int8_t *__restrict__ strand,
                                 const int32_t *__restrict__ start, const int32_t *__restrict__ end, int64_t n, int n_cls,
                                 int n_width, int32_t minoverlap, unsigned *__restrict__ keys, int32_t *__restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned sc = strand ? (unsigned)strand[i] : 0u;
        const unsigned narrow = (n_width > 1 && end[i] - start[i] + 1 < minoverlap) ? 1u : 0u;
        keys[i] = (unsigned)seqid[i] * (unsigned)n_cls + sc * (unsigned)n_width + narrow;
        idx[i] = (int32_t)i;
    }
}
// class-major copies of start / end (order = stable sort of the canonical order on the virtual sequence)
__global__ void gather_by_order_kernel(const int32_t *__restrict__ order, const int32_t *__restrict__ start,
                                       const int32_t *__restrict__ end, int64_t n, int32_t *__restrict__ rstart,
                                       int32_t *__restrict__ rend) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int32_t o = order[i];
        rstart[i] = start[o];
        rend[i] = end[o];
    }
}
// voff[v] = first position of virtual sequence v in the sorted keys; seq_info of every virtual sequence
__global__ void vseq_info_kernel(const unsigned *__restrict__ sorted_keys, int64_t n, int n_vseq, int n_cls, const int32_t *__restrict__ max_pos,
                                 const int64_t *__restrict__ rank_off, int32_t *__restrict__ voff, int4 *__restrict__ info) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > n_vseq) return;
    auto lower = [&](unsigned key) {
        int64_t a = 0, b = n;
        while (a < b) {
            const int64_t m = (a + b) >> 1;
            if (sorted_keys[m] < key) a = m + 1; else b = m;
        }
        return (int32_t)a;
    };
    const int32_t lo = lower((unsigned)v);
    voff[v] = lo;
    if (v < n_vseq) info[v] = make_int4(lo, lower((unsigned)v + 1u), max_pos[v / n_cls], (int32_t)rank_off[v]);
}
__global__ void pack_vend_keys_kernel(const unsigned *__restrict__ vkeys, const int32_t *__restrict__ end, int64_t n,
                                      unsigned long long *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((unsigned long long)vkeys[i] << 32) | (unsigned)end[i];
}

// entry t of a packed rank table (see "direct-address rank tables")
__global__ void build_rank_table_kernel(const int32_t *__restrict__ values, const int32_t *__restrict__ seq_off,
                                        const int64_t *__restrict__ rank_off, int n_seq, int shift,
                                        uint32_t *__restrict__ table) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rank_off[n_seq]) return;
    int c = 0, hi = n_seq;  // sequence that owns entry t
    while (hi - c > 1) {
        const int m = (c + hi) >> 1;
        if (rank_off[m] <= t) c = m; else hi = m;
    }
    const int64_t bucket = t - rank_off[c];
    const int64_t key = bucket << shift, next = (bucket + 1) << shift;
    const int32_t a = seq_off[c], b = seq_off[c + 1];
    const int32_t first = key > INT32_MAX ? b : first_ge(values, a, b, (int32_t)key);
    const int32_t last = next > INT32_MAX ? b : first_ge(values, first, b, (int32_t)next);
    table[t] = ((uint32_t)(first - a) << kRankCountBits) | (uint32_t)min(last - first, kRankCountCap);
}

// countOverlaps(query, y) on the original ranges: one thread per y range.
template <bool STRANDED>
__global__ void count_observed_kernel(const RangesView y, int64_t n, const int32_t *__restrict__ y_seq,
                                      const int32_t *__restrict__ q_start, const int32_t *__restrict__ q_end,
                                      const int32_t *__restrict__ q_pmax, const int8_t *__restrict__ q_strand,
                                      const int32_t *__restrict__ q_src, const int32_t *__restrict__ q_seq_off,
                                      int32_t minoverlap, int32_t *__restrict__ counts, unsigned long long *__restrict__ total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int found = 0;
    if (i < n) {
        const int c = y_seq[i];
        const int32_t s = y.start[i], e = y.end[i];
        const int strand = (STRANDED && y.strand) ? y.strand[i] : 0;
        const int32_t a = q_seq_off[c];
        for (int32_t k = first_gt(q_start, a, q_seq_off[c + 1], e) - 1; k >= a && q_pmax[k] >= s; --k) {
            if (q_end[k] < s || (STRANDED && q_strand && !strands_match(strand, q_strand[k]))) continue;
            if (minoverlap > 0 && min(e, q_end[k]) - max(s, q_start[k]) + 1 < minoverlap) continue;  // width of the intersection
            ++found;
            atomicAdd(counts + q_src[k], 1);
        }
    }
    found = __reduce_add_sync(kFull, found);
    if ((threadIdx.x & 31) == 0 && found) atomicAdd(total, (unsigned long long)found);
}

// Per-feature moments over a batch of iterations: thread per feature, coalesced over g.
__global__ void accumulate_moments_kernel(const int32_t *__restrict__ counts, int64_t n_iter, int64_t n_query,
                                          const int32_t *__restrict__ observed, long long *__restrict__ sum,
                                          long long *__restrict__ sumsq, long long *__restrict__ n_ge) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_query) return;
    long long s = 0, q = 0, ge = 0;
    const int32_t obs = observed ? observed[g] : 0;
    for (int64_t r = 0; r < n_iter; ++r) {
        const long long v = counts[r * n_query + g];
        s += v;
        q += v * v;
        ge += (observed && v >= obs) ? 1 : 0;
    }
    sum[g] += s;
    sumsq[g] += q;
    n_ge[g] += ge;
}

}  // namespace nrb
