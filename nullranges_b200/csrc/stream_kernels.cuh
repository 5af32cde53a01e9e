// Streaming path: the fused count kernel (one warp per iteration x block) and the materialising scan / fill kernels.
// (part of kernels.cuh)
#pragma once
#include "device_views.cuh"

namespace nrb {

// ---- the fused kernel -----------------------------------------------------------------

struct Item {
    int bait_seq, tile;        // sampled-from sequence; destination tile index
    int32_t bait_start;
    int64_t bait_end;
    int tile_seq;
    int32_t shift;             // tile_start - bait_start       (block_shift, :179)
    int32_t seq_len;           // seqlengths[tile_seq]
    int32_t lo, hi;            // hit window in canonical y
    int32_t q_lo, q_hi, x_lo, x_hi;  // boot_stream_kernel only: the tile's query / exclude candidates
    int32_t interior;          // boot_stream_kernel only: no range of the hit window can leave [1, seq_len) after the shift
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
    if (e2 < s2 && !kZeroWidthStrictlyInside) return false;  // THE ZERO-WIDTH RULE (device_views.cuh)
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
            if (HAS_QUERY && (e2 >= s2 || kZeroWidthStrictlyInside)) overlaps += count_piece<STRANDED>(P, r, q_lo, q_hi, s2, e2, strand);
        }
        rows = __reduce_add_sync(kFull, rows);
        hits = __reduce_add_sync(kFull, hits);
        if (HAS_QUERY) overlaps = __reduce_add_sync(kFull, overlaps);
        if (lane == 0) {
            if (rows) atomicAdd(P.iter_nranges + r, (unsigned long long)rows);
            if (HAS_QUERY && overlaps) atomicAdd(P.iter_totals + r, (unsigned long long)overlaps);
            if (P.item_rows) P.item_rows[item] = rows;
            if (P.n_hits && hits) atomicAdd(P.n_hits + (r & (kHitSlots - 1)), (unsigned long long)hits);
        }
    }
}

// ---- the fused streaming kernel, tile-group form -------------------------------------------
//
// CTA = (iteration, group of up to kGroupTiles tiles that follow each other on one sequence).  The query features
// within reach of those tiles are ONE contiguous slice of the canonical query arrays; an elected thread brings the
// slice (start, end, running-max end, caller's position[, strand]) into shared memory with 1-D bulk asynchronous
// copies (cp.async.bulk + mbarrier: the copy engine works while the CTA draws its blocks and locates their hit
// windows), and the overlaps are tallied in a shared-memory histogram that leaves the CTA once, one global atomic per
// feature hit instead of one per overlap.  A warp takes a tile at a time and streams its hit window of y from the
// interleaved {start, end} array: 16 bytes per lane, 64 ranges per load instruction, the next trip's (or the next
// tile's first) load in flight while a trip is worked on; shift / trim / drop / exclude per range as in the
// reference (R/bootUnsegmented.R:175-191, R/bootRanges.R:106-116).  The 64 ranges of a load are neighbours on the
// genome, so the features they can reach are narrowed once per warp -- the window's two ends step forward over the
// sorted slice by one ballot each, from the warp's largest end and smallest start -- and every lane tests only those.
// The kernel is bound by instruction issue (profiles/README.md), so the loop is written for few instructions: tiles
// that provably stay inside their sequence skip trim() and count rows as hits (Item::interior), the shared arrays are
// addressed from one register, rows / hits leave a warp once, the total of overlaps is read off the histogram.
// A launch with few (group, iteration) pairs gives every group to gridDim.z CTAs that take its tiles in turn (run.cuh).
// Bootstrap kinds only (block b lands on tile b); the permute variants and slices that do not fit keep
// boot_count_kernel.
constexpr int kGroupTiles = 256;   // tiles per CTA at most (one thread draws each: <= the CTA's 256 threads).  The draw + hit-window
                                   // search is a ~4,000-cycle dependent chain before any range can be streamed: 16 tiles per warp
                                   // make that prologue a fifth of a CTA's life (32 tiles per CTA: nearly half, measured as barrier stalls)
constexpr int kStreamMinBlocks = 5; // resident CTAs per SM the stream kernel is compiled for (<= 51 registers: no spills); the plain variant
                                    // (no strands, no exclusion) fits 40 registers and takes 6
constexpr int kSePad = 2 * kWarp + 2; // sentinel elements after y's interleaved copy (a whole trip may be loaded from an even index <= n)
constexpr int kMaxSlice = 2048;    // features of a group's slice at most (shared memory: 21 bytes per feature)

struct TileGroup {
    int32_t first, count;   // tiles_sorted[first .. first + count): tile indices, ascending start on one sequence
    int32_t q_lo, q_hi;     // the group's slice of the canonical query, both multiples of 16 (16-byte aligned copies of every column)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
}

// Shared-memory accesses of the hot loop by 32-bit shared address: the base stays in one register instead of being rebuilt
// from the CTA's rank in its cluster at every use (what the compiler does under the register cap when it sees the pointers).
__device__ __forceinline__ int32_t lds_s32(uint32_t addr) {
    int32_t v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// n += on, as one predicated add (the compiler's own pattern is an add and a predicated move)
__device__ __forceinline__ void inc_if(unsigned &n, bool on) {
    asm("{\n"
        ".reg .pred q;\n"
        "setp.ne.s32 q, %1, 0;\n"
        "@q add.u32 %0, %0, 1;\n"
        "}\n"
        : "+r"(n)
        : "r"((int)on));
}
struct GroupView {
    const TileGroup *groups;
    const int32_t *tiles_sorted;
    const int2 *se;            // y as interleaved {start, end}, canonical order
    int slice_cap;             // features the shared-memory arrays hold (a multiple of 4)
};

// One tile of a group, worked by one warp: streams the tile's hit window of y, 64 ranges per trip, and tallies the overlaps
// with the features of the shared slice.  INTERIOR: the tile's window stays inside [1, seq_len) after the shift (Item::interior),
// so trim() and the fall-off-the-end rule cannot apply and every hit is a row.
template <bool STRANDED, bool HAS_QUERY, int EXCL, bool INTERIOR>
__device__ __forceinline__ void stream_tile(const BootParams &P, const GroupView &V, const Item &it, const int lane, const uint32_t a_start,
                                            const uint32_t cap4, const int32_t *s_start, const int32_t *s_end, int *s_hist,
                                            const int8_t *s_strand, unsigned &rows_out, unsigned &hits_lane, int4 &v, const int32_t next_base) {
    const int32_t q_lo = it.q_lo, q_hi = it.q_hi, x_lo = it.x_lo, x_hi = it.x_hi;
    // the features the warp's current ranges can reach, [w_lo, w_hi) of the shared slice: the ranges come in ascending
    // start order, so both ends only move forward (w_hi may run ahead: every lane tests exactly)
    int32_t w_lo = q_lo, w_hi = q_lo;
    unsigned rows = 0, hits = 0;
    // Ranges i0 and i0 + 1 per lane and trip.  v holds the lane's pair of the tile's first trip on entry (loaded while the warp
    // still worked on its previous tile) and the pair of the NEXT tile's first trip on exit; every load is unconditional: the
    // array carries kSePad sentinel elements at its end and the index tests below discard what lies outside the window.
    if ((it.lo & ~1) >= it.hi) {  // empty window
        v = __ldg(reinterpret_cast<const int4 *>(V.se + next_base + 2 * lane));
        rows_out = 0;
        return;
    }
    for (int32_t base = it.lo & ~1; base < it.hi; base += 2 * kWarp) {
        const int32_t i0 = base + 2 * lane;
        int32_t s2[2], e2[2];
        bool keep[2];
        int strand[2] = {0, 0};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int32_t i = i0 + u;
            const int32_t st = u ? v.z : v.x, en = u ? v.w : v.y;
            // bootstrap kinds: every range of the window starts at or before the block's end; a hit also ends at or after its start.
            // Only the even-aligned first slot of the first trip can lie before the window.
            const bool hit = (u ? true : i >= it.lo) && i < it.hi && en >= it.bait_start;
            inc_if(hits, hit);
            s2[u] = st + it.shift;
            e2[u] = en + it.shift;
            keep[u] = hit;
            if (!INTERIOR) {
                if (s2[u] > it.seq_len) {       // wholly beyond the sequence end -> [L + 1, L]
                    if (P.drop_zero) keep[u] = false;
                    s2[u] = it.seq_len + 1;
                    e2[u] = it.seq_len;
                } else {
                    s2[u] = max(s2[u], 1);
                    e2[u] = min(e2[u], it.seq_len);
                }
            }
            if (STRANDED && keep[u] && P.y.strand) strand[u] = __ldg(P.y.strand + i);
            if (EXCL == kExclDrop && keep[u] && x_hi > x_lo && (e2[u] >= s2[u] || kZeroWidthStrictlyInside)) {  // THE ZERO-WIDTH RULE (device_views.cuh)
                for (int32_t k = x_lo; k < x_hi; ++k) {  // the tile's exclude candidates (few tiles have any)
                    if (s2[u] <= __ldg(P.excl.end + k) && e2[u] >= __ldg(P.excl.start + k) &&
                        (!STRANDED || !P.excl.strand || strands_match(strand[u], __ldg(P.excl.strand + k)))) {
                        keep[u] = false;
                        break;
                    }
                }
            }
            if (EXCL == kExclTrim && keep[u] && e2[u] < s2[u]) keep[u] = false;  // a zero-width survivor overlaps no gap
        }
        // the pair is consumed: the next trip's load (the next tile's first, after the last trip) goes out now and has the rest
        // of this trip to arrive
        v = __ldg(reinterpret_cast<const int4 *>(V.se + (base + 2 * kWarp < it.hi ? i0 + 2 * kWarp : next_base + 2 * lane)));
        if (EXCL == kExclTrim) {
            // each range is cut to the gaps it overlaps: one row (and one piece to count) per overlap
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!keep[u]) continue;
                for (int32_t k = x_lo; k < x_hi; ++k) {
                    const int32_t gs = __ldg(P.excl.start + k), ge = __ldg(P.excl.end + k);
                    if (s2[u] > ge || e2[u] < gs) continue;
                    ++rows;
                    if (HAS_QUERY) {
                        const int32_t ps = max(s2[u], gs), pe = min(e2[u], ge);
                        for (int32_t q = q_lo; q < q_hi; ++q) {
                            const int32_t qs = s_start[q], qe = s_end[q];
                            if (ps > qe || pe < qs) continue;
                            if (STRANDED && P.query.strand && !strands_match(strand[u], s_strand[q])) continue;
                            if (P.minoverlap > 0 && min(pe, qe) - max(ps, qs) + 1 < P.minoverlap) continue;
                            atomicAdd(&s_hist[q], 1);
                        }
                    }
                }
            }
            continue;
        }
        if (!INTERIOR) {  // (interior tiles: every hit is a row)
            inc_if(rows, keep[0]);
            inc_if(rows, keep[1]);
        }
        if (HAS_QUERY && q_hi > q_lo) {
            const bool cnt0 = keep[0] && (e2[0] >= s2[0] || kZeroWidthStrictlyInside), cnt1 = keep[1] && (e2[1] >= s2[1] || kZeroWidthStrictlyInside);
            const int32_t my_min = min(cnt0 ? s2[0] : INT32_MAX, cnt1 ? s2[1] : INT32_MAX);
            const int32_t my_max = max(cnt0 ? e2[0] : INT32_MIN, cnt1 ? e2[1] : INT32_MIN);
            const int32_t w_min = __reduce_min_sync(kFull, my_min), w_max = __reduce_max_sync(kFull, my_max);
            if (w_max != INT32_MIN) {
                // both ends of the window advance 32 candidates at a time: starts and running-max ends ascend along the slice,
                // so the lanes that pass are a prefix and their number is the step
                for (;;) {  // features starting at or before the largest end
                    const int32_t q = w_hi + lane;
                    const unsigned m = __ballot_sync(kFull, q < q_hi && lds_s32(a_start + 4u * (uint32_t)q) <= w_max);
                    w_hi += __popc(m);
                    if (m != kFull) break;
                }
                for (;;) {  // ... whose running-max end reaches the smallest start
                    const int32_t q = w_lo + lane;
                    const unsigned m = __ballot_sync(kFull, q < w_hi && lds_s32(a_start + 2u * cap4 + 4u * (uint32_t)q) < w_min);
                    w_lo += __popc(m);
                    if (m != kFull) break;
                }
                const uint32_t a_stop = a_start + 4u * (uint32_t)w_hi;
#pragma unroll 1
                for (uint32_t a = a_start + 4u * (uint32_t)w_lo; a < a_stop; a += 4u) {
                    const int32_t qs = lds_s32(a), qe = lds_s32(a + cap4);  // broadcast reads
                    bool h0 = cnt0 && s2[0] <= qe && e2[0] >= qs, h1 = cnt1 && s2[1] <= qe && e2[1] >= qs;
                    if (STRANDED && P.query.strand) {
                        const int qst = s_strand[(a - a_start) >> 2];
                        h0 = h0 && strands_match(strand[0], qst);
                        h1 = h1 && strands_match(strand[1], qst);
                    }
                    if (P.minoverlap > 0) {
                        h0 = h0 && min(e2[0], qe) - max(s2[0], qs) + 1 >= P.minoverlap;
                        h1 = h1 && min(e2[1], qe) - max(s2[1], qs) + 1 >= P.minoverlap;
                    }
                    // one add per warp and feature, whatever the density of y (with every lane adding for itself a dense y --
                    // thousands of ranges per feature -- turned each trip into 64 serialised same-address atomics: 4x slower
                    // at 5e7 ranges)
                    const int n = __reduce_add_sync(kFull, (h0 ? 1 : 0) + (h1 ? 1 : 0));
                    if (n && lane == 0) red_shared_add(a + 4u * cap4, n);
                }
            }
        }
    }
    hits_lane += hits;
    rows_out = INTERIOR ? hits : rows;
}

template <bool STRANDED, bool HAS_QUERY, int EXCL>
__global__ void __launch_bounds__(256, (!STRANDED && EXCL == kExclNone) ? kStreamMinBlocks + 1 : kStreamMinBlocks) boot_stream_kernel(const BootParams P, const GroupView V) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned long long s_tally[3];  // rows, hits, overlaps of the CTA's tiles
    __shared__ Item s_item[kGroupTiles];       // the group's blocks: drawn and located once, by one lane each
    if (threadIdx.x < 3) s_tally[threadIdx.x] = 0;
    const int cap = V.slice_cap;
    int32_t *s_start = reinterpret_cast<int32_t *>(smem_raw);
    int32_t *s_end = s_start + cap, *s_pmax = s_end + cap, *s_src = s_pmax + cap;
    int *s_hist = s_src + cap;
    int8_t *s_strand = reinterpret_cast<int8_t *>(s_hist + cap);
    uint32_t a_start = smem_u32(smem_raw);     // shared address of s_start; s_end, s_pmax, s_hist follow at 1, 2, 4 times cap4
    asm volatile("" : "+r"(a_start));           // opaque to the compiler: kept, not rematerialised
    const uint32_t cap4 = 4u * (uint32_t)cap;
    const int lane = threadIdx.x & (kWarp - 1), warp = threadIdx.x / kWarp;
    const TileGroup grp = V.groups[blockIdx.x];
    const int64_t r = blockIdx.y;
    const int n_slice = grp.q_hi - grp.q_lo;             // a multiple of 16 (the arrays carry that much slack at their ends)
    const bool stage = HAS_QUERY && n_slice > 0;
    if (stage) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)n_slice * 4u;
            const bool with_strand = STRANDED && P.query.strand;
            mbar_expect_tx(&bar, 4u * bytes + (with_strand ? (uint32_t)n_slice : 0u));
            bulk_copy_g2s(s_start, P.query.start + grp.q_lo, bytes, &bar);
            bulk_copy_g2s(s_end, P.query.end + grp.q_lo, bytes, &bar);
            bulk_copy_g2s(s_pmax, P.query.pmax + grp.q_lo, bytes, &bar);
            bulk_copy_g2s(s_src, P.query.src + grp.q_lo, bytes, &bar);
            if (with_strand) bulk_copy_g2s(s_strand, P.query.strand + grp.q_lo, (uint32_t)n_slice, &bar);
        }
        for (int k = threadIdx.x; k < n_slice; k += blockDim.x) s_hist[k] = 0;  // generic-proxy writes to a region the copies do not touch
    }
    // While the copy engine works: thread j draws block j of the group and locates its hit window
    // (Philox + pick + two rank probes: ~275 instructions that 32 lanes would otherwise repeat for every tile).
    // (gridDim.z > 1, small launches: this CTA works the group's tiles z, z + gridDim.z, ...: n_tiles of them)
    const int n_tiles = (grp.count - (int)blockIdx.z + (int)gridDim.z - 1) / (int)gridDim.z;
    if ((int)threadIdx.x < n_tiles) {
        Item it = load_item(P, r, __ldg(V.tiles_sorted + grp.first + blockIdx.z + threadIdx.x * gridDim.z));
        it.q_lo = it.q_hi = it.x_lo = it.x_hi = 0;
        if (HAS_QUERY) {
            it.q_lo = __ldg(P.query.tile_lo + it.tile) - grp.q_lo;
            it.q_hi = __ldg(P.query.tile_hi + it.tile) - grp.q_lo;
        }
        if (EXCL != kExclNone) {
            it.x_lo = __ldg(P.excl.tile_lo + it.tile);
            it.x_hi = __ldg(P.excl.tile_hi + it.tile);
        }
        // Most tiles sit well inside their sequence: when the window's smallest start and (a bound on) its largest end stay inside
        // after the shift, trim() is the identity for every range of the tile and nothing can fall off the sequence end.
        it.interior = 0;
        if (EXCL == kExclNone && it.hi > it.lo) {
            const int64_t s_min = (int64_t)__ldg(&V.se[it.lo].x) + it.shift, e_max = (int64_t)__ldg(P.y.pmax + it.hi - 1) + it.shift;
            it.interior = (s_min >= 1 && e_max < it.seq_len) ? 1 : 0;
        }
        s_item[threadIdx.x] = it;
    }
    if (stage) mbar_wait(&bar, 0);  // the slice has landed (every thread observes the phase itself)
    __syncthreads();                // ... the histogram is zeroed and the items are in place
    // rows and hits are tallied per lane over all the warp's tiles (a lane sees at most 1/32 of every hit window, so 32 bits hold
    // them) and leave the warp once; the overlaps are read off the histogram when it leaves the CTA.
    unsigned rows_lane = 0, hits_lane = 0;
    int4 v = make_int4(0, 0, 0, 0);
    if (warp < n_tiles) v = __ldg(reinterpret_cast<const int4 *>(V.se + (s_item[warp].lo & ~1) + 2 * lane));
    for (int j = warp; j < n_tiles; j += blockDim.x / kWarp) {
        const Item it = s_item[j];
        const int j_next = j + blockDim.x / kWarp;
        const int32_t next_base = (j_next < n_tiles ? s_item[j_next].lo : it.lo) & ~1;  // (after the warp's last tile: a load nobody reads)
        unsigned rows = 0;
        bool done = false;
        if constexpr (EXCL == kExclNone) {
            if (it.interior) {
                stream_tile<STRANDED, HAS_QUERY, EXCL, true>(P, V, it, lane, a_start, cap4, s_start, s_end, s_hist, s_strand, rows, hits_lane, v, next_base);
                done = true;
            }
        }
        if (!done) stream_tile<STRANDED, HAS_QUERY, EXCL, false>(P, V, it, lane, a_start, cap4, s_start, s_end, s_hist, s_strand, rows, hits_lane, v, next_base);
        rows_lane += rows;
        if (P.item_rows) {
            const unsigned rows_tile = __reduce_add_sync(kFull, rows);
            if (lane == 0) P.item_rows[r * P.n_blocks + it.tile] = (int32_t)rows_tile;
        }
    }
    // The per-iteration tallies leave the CTA once: all CTAs of an iteration add to the same three words.
    {
        unsigned long long rows_warp = rows_lane, hits_warp = hits_lane;
#pragma unroll
        for (int d = kWarp / 2; d > 0; d >>= 1) {
            rows_warp += __shfl_xor_sync(kFull, rows_warp, d);
            hits_warp += __shfl_xor_sync(kFull, hits_warp, d);
        }
        if (lane == 0) {
            if (rows_warp) atomicAdd(&s_tally[0], rows_warp);
            if (hits_warp) atomicAdd(&s_tally[1], hits_warp);
        }
    }
    __syncthreads();  // ... and the histogram is complete
    if (HAS_QUERY && stage) {
        unsigned long long overlaps = 0;
        const bool store = P.counts != nullptr;
        for (int k = threadIdx.x; k < n_slice; k += blockDim.x) {
            const int n = s_hist[k];
            if (n) {
                overlaps += (unsigned)n;
                if (store) atomicAdd(P.counts + r * P.n_query + s_src[k], n);
            }
        }
#pragma unroll
        for (int d = kWarp / 2; d > 0; d >>= 1) overlaps += __shfl_xor_sync(kFull, overlaps, d);
        if (lane == 0 && overlaps) atomicAdd(&s_tally[2], overlaps);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s_tally[0]) atomicAdd(P.iter_nranges + r, s_tally[0]);
        if (P.n_hits && s_tally[1]) atomicAdd(P.n_hits + (r & (kHitSlots - 1)), s_tally[1]);
        if (HAS_QUERY && s_tally[2]) atomicAdd(P.iter_totals + r, s_tally[2]);
    }
}

// y as interleaved {start, end} pairs (one 16-byte load = two ranges) followed by kSePad sentinel elements: boot_stream_kernel
// loads whole trips (2 * kWarp ranges from an even index) without looking at the window's end first
__global__ void interleave_ranges_kernel(const int32_t *__restrict__ start, const int32_t *__restrict__ end, int64_t n, int2 *__restrict__ se) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) se[i] = make_int2(start[i], end[i]);
    else if (i < n + kSePad) se[i] = make_int2(INT32_MAX, INT32_MIN);
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

}  // namespace nrb
