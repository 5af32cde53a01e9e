// Rank path: counting (and weighted sums) by rank differences, and the per-block rows kernel.
// (part of kernels.cuh)
#pragma once
#include "device_views.cuh"
#include "stream_kernels.cuh"

namespace nrb {

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

// One signed window to count for a feature on a tile, 16 bytes (half a sector: the count kernel is bound by the
// sectors it pulls through the L2, so the record carries only what is per window):
//   {(tile << 4) | (cls << 1) | negative, ws, we, tile_start}
// tile = block index b (row of the draw table), [ws, we] = window on the tile's sequence ([gs, ge] or an exclusion
// pseudo-feature), cls = class of y the window counts (0 when y has one class), negative = the window is subtracted.
// The tile's bait width and sequence length come from tile_info[tile] (8 bytes per tile, read beside the draw -- the
// same tile for most lanes of a warp, so one broadcast).
constexpr int kRecTileShift = 4;
__device__ __forceinline__ int rec_tile(int32_t x) { return x >> kRecTileShift; }
__device__ __forceinline__ int rec_cls(int32_t x) { return (x >> 1) & 7; }
__device__ __forceinline__ int rec_sign(int32_t x) { return 1 - 2 * (x & 1); }

struct FeatView {              // query features in visit order (grouped by list length)
    const int32_t *src;        // [n_query] index in the caller's order
    const int32_t *pair_off;   // [n_query + 1]
    const int4 *rec;           // [n_pairs]
};

struct TileExcl {              // signed exclusion windows of each (tile, strand class), for the rows kernel
    const int32_t *off;        // [n_blocks * n_cls + 1], or nullptr without exclusion
    const int4 *rec;           // {ws, we, sign, 0}
};

// One (feature window, tile, iteration) evaluation.  All coordinates are below 2^30 and widths at most 2^30: every
// sum fits int32.
struct PairProbe {
    int4 si;                   // seq_info of the bait sequence
    int32_t A, Bv;             // count = #{start <= A and end >= Bv}
    int32_t bs;                // bait start (the permute variants also need start >= bs)
};

__device__ __forceinline__ void pair_keys(PairProbe &q, int2 draw, int32_t width, int32_t ts, int32_t seq_len, int32_t gs, int32_t ge) {
    const int32_t bs = draw.y;
    const int32_t sh = ts - bs;
    q.bs = bs;
    q.A = min(bs + (width - 1), min(ge, seq_len) - sh);
    q.Bv = max(bs, gs - sh);
}

// #{start <= a and end >= b} for a < b: such ranges contain [a, b]; walk back from #{start <= a}
// (index ia) under the running-max-of-end bound.  Usually zero or one step.
__device__ __forceinline__ int count_spanning(const RangesView &y, const int4 si, int32_t ia, int32_t b) {
    if (b > si.z) return 0;
    int n = 0;
    for (int32_t k = ia - 1; k >= si.x && __ldg(y.rpmax + k) >= b; --k) n += __ldg(y.rend + k) >= b;
    return n;
}

// The counts of RPT evaluations that share a window (RPT consecutive iterations), stage by stage so that the
// independent loads of a stage are in flight together: sequence infos, then both table entries of every evaluation
// (two 8-byte loads, one sector each), then the ranks out of the entries themselves.
// PERMUTE: a range belongs to the ONE tile that holds its start (findOverlaps(select = "first"),
// R/bootUnsegmented.R:95, :155), so the members are {bs <= start <= A and end >= Bv}: the bootstrap
// count minus the ranges that start before the tile and reach Bv (they contain [bs - 1, Bv]).
template <int RPT, bool PERMUTE>
__device__ __forceinline__ void count_pairs(const RangesView &y, const int2 (&d)[RPT], int cls, int32_t width, int32_t ts,
                                            int32_t seq_len, int32_t gs, int32_t ge, int (&cnt)[RPT]) {
    PairProbe q[RPT];
    int32_t key_a[RPT], key_b[RPT], ua[RPT], ub[RPT];
    RankEntry ea[RPT], eb[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) q[j].si = __ldg(y.rseq_info + d[j].x * y.n_cls + cls);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        pair_keys(q[j], d[j], width, ts, seq_len, gs, ge);
        key_a[j] = rank_clamp_key(q[j].A + 1, q[j].si);
        key_b[j] = rank_clamp_key(q[j].Bv - y.end_shift, q[j].si);
        ua[j] = q[j].si.w + (key_a[j] >> y.rank_shift);
        ub[j] = q[j].si.w + (key_b[j] >> y.rank_shift);
        ea[j] = __ldg(y.rrank_start + ua[j]);
        eb[j] = __ldg(y.rank_end + ub[j]);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        if (q[j].A < 1 || (PERMUTE && q[j].A < q[j].bs)) continue;  // permute: no start can lie in [bs, A]
        const int32_t ia = rank_from_entry(y.rstart, y.rrank_start, y.rank_shift, q[j].si, key_a[j], ua[j], ea[j]);
        int n;
        if (q[j].Bv <= q[j].A) n = ia - rank_from_entry(y.end_sorted, y.rank_end, y.rank_shift, q[j].si, key_b[j], ub[j], eb[j]);
        else n = count_spanning(y, q[j].si, ia, q[j].Bv);  // the feature lies beside the tile
        if (PERMUTE && n > 0)
            n -= count_spanning(y, q[j].si, first_ge_ranked(y.rstart, y.rrank_start, y.rank_shift, q[j].si, q[j].bs), q[j].Bv);
        cnt[j] += n;
    }
}

template <bool PERMUTE>
__device__ __forceinline__ int count_pair(const RangesView &y, int2 d, int cls, int32_t width, int32_t ts, int32_t seq_len,
                                          int32_t gs, int32_t ge) {
    const int2 dd[1] = {d};
    int c[1] = {0};
    count_pairs<1, PERMUTE>(y, dd, cls, width, ts, seq_len, gs, ge, c);
    return c[0];
}

constexpr int kCountThreads = 128;  // CTA size of the count kernels: 128 beat 64 / 256 / 512 in round 1's sweep

// The count kernel.  Thread = feature, staying on it for `ipt` consecutive iterations: the window record, the
// tile's bait width / sequence length and the list bounds are read once per thread instead of once per iteration,
// and the next iteration's draw is requested while the current one is evaluated, so that per iteration the chain
// is draw (prefetched) -> sequence info (an L1 hit) -> the two table entries.  Per-iteration totals are summed in
// shared memory and leave the CTA once.  With MOMENTS the counts are not stored at all: every thread keeps
// sum, sum of squares and #{count >= observed} of its feature in registers and adds them to the per-feature
// vectors when its iterations are done -- the [R x G] matrix never exists (BASELINE config 4: R = 10^4, G = 2 * 10^5).
struct MomentsOut {
    const int32_t *observed;       // [n_query] caller's order, or nullptr
    long long *sum, *sumsq, *n_ge; // [n_query] caller's order
};
constexpr int kLoopMaxIpt = 64;
template <bool MOMENTS, int MINB>
__global__ void __launch_bounds__(kCountThreads, MINB) boot_rank_loop_kernel(const BootParams P, const FeatView F, int ipt, const MomentsOut M) {
    __shared__ int s_tot[kCountThreads / 32][kLoopMaxIpt];  // per-iteration totals, one row per warp (plain adds by lane 0)
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * ipt;
    const int n_it = (int)min((int64_t)ipt, P.n_iter - r0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = lane; k < n_it; k += 32) s_tot[warp][k] = 0;
    __syncwarp();
    int32_t src = 0, p0 = 0, p1 = 0;
    if (g < P.n_query) {
        src = __ldg(F.src + g);
        p0 = __ldg(F.pair_off + g);
        p1 = __ldg(F.pair_off + g + 1);
    }
    const bool live = g < P.n_query, any = p1 > p0;
    int4 rec0 = make_int4(0, 0, 0, 0);
    int2 ti0 = make_int2(0, 0), d_next = make_int2(0, 0);
    const int2 *__restrict__ tab = P.draw_tab + r0 * P.n_blocks;
    const int2 *__restrict__ dp = tab;  // the first window's draw of the current iteration; advances by one row per iteration
    if (any) {
        rec0 = __ldg(F.rec + p0);
        ti0 = __ldg(P.tile_info + rec_tile(rec0.x));
        dp += rec_tile(rec0.x);
        d_next = __ldg(dp);
    }
    long long m_sum = 0, m_sq = 0;
    int m_ge = 0;
    const int32_t obs = (MOMENTS && live && M.observed) ? __ldg(M.observed + src) : 0;
#pragma unroll 1
    for (int k = 0; k < n_it; ++k) {
        int cnt = 0;
        if (any) {
            const int2 d = d_next;
            dp += P.n_blocks;
            if (k + 1 < n_it) d_next = __ldg(dp);
            const int c0 = P.permute ? count_pair<true>(P.y, d, rec_cls(rec0.x), ti0.x, rec0.w, ti0.y, rec0.y, rec0.z)
                                     : count_pair<false>(P.y, d, rec_cls(rec0.x), ti0.x, rec0.w, ti0.y, rec0.y, rec0.z);
            cnt = rec_sign(rec0.x) * c0;
            for (int32_t p = p0 + 1; p < p1; ++p) {  // further windows of the feature (other tiles, exclusion windows): L1 hits after the first iteration
                const int4 rec = __ldg(F.rec + p);
                const int tile = rec_tile(rec.x);
                const int2 ti = __ldg(P.tile_info + tile);
                const int2 dd = __ldg(tab + (int64_t)k * P.n_blocks + tile);
                const int c = P.permute ? count_pair<true>(P.y, dd, rec_cls(rec.x), ti.x, rec.w, ti.y, rec.y, rec.z)
                                        : count_pair<false>(P.y, dd, rec_cls(rec.x), ti.x, rec.w, ti.y, rec.y, rec.z);
                cnt += rec_sign(rec.x) * c;
            }
        }
        if (live) {
            if (MOMENTS) {
                m_sum += cnt;
                m_sq += (long long)cnt * cnt;
                m_ge += (M.observed && cnt >= obs) ? 1 : 0;
            } else if (P.counts16) {
                __stcs(P.counts16 + (r0 + k) * P.n_query + src, (unsigned short)min(cnt, 65535));
                if (cnt > 65535) atomicOr(P.overflow16, 1);
            } else if (P.n_peers > 0) {  // the tile goes straight into every member's matrix: the all-gather IS the kernel's store
                const int64_t at = (r0 + k) * P.n_query + src;
                for (int w = 0; w < P.n_peers; ++w) __stcs(P.peer_counts[w] + at, cnt);
            } else if (P.counts) {
                __stcs(P.counts + (r0 + k) * P.n_query + src, cnt);
            }
        }
        const int s = __reduce_add_sync(kFull, cnt);
        if (lane == 0) s_tot[warp][k] += s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_it; k += kCountThreads) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kCountThreads / 32; ++w) t += (unsigned)s_tot[w][k];
        if (t) atomicAdd(P.iter_totals + r0 + k, t);
    }
    if (MOMENTS && live) {
        if (m_sum) atomicAdd(reinterpret_cast<unsigned long long *>(M.sum + src), (unsigned long long)m_sum);
        if (m_sq) atomicAdd(reinterpret_cast<unsigned long long *>(M.sumsq + src), (unsigned long long)m_sq);
        if (m_ge) atomicAdd(reinterpret_cast<unsigned long long *>(M.n_ge + src), (unsigned long long)m_ge);
    }
}

// Device group, gathered result: this GPU's slices of the per-iteration vectors into every member's vectors.
struct PeerVectors {
    int n;
    unsigned long long *nranges[kMaxPeers], *totals[kMaxPeers];  // already offset to this GPU's first iteration
};
__global__ void broadcast_iteration_vectors_kernel(const unsigned long long *__restrict__ nranges, const unsigned long long *__restrict__ totals,
                                                   int64_t n, const PeerVectors V) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long a = nranges[i], b = totals[i];
        for (int w = 0; w < V.n; ++w) {
            V.nranges[w][i] = a;
            V.totals[w][i] = b;
        }
    }
}

// Device group, gathered result: this GPU's tile of the count matrix ([n x G] int32, already in its own copy of the
// matrix) into the same place of every other member's copy -- 16-byte stores over peer-mapped pointers (NVLink), fully
// coalesced, every element read once and written W - 1 times.  (Scattered 4-byte peer stores straight from the count
// kernel were measured first: 190 GB/s per GPU; this form moves the same bytes as full 128-byte lines.)
struct PeerTiles {
    int n;                       // the OTHER members
    int32_t *dst[kMaxPeers];     // their matrices, offset to this GPU's first row
};
__global__ void __launch_bounds__(256) push_tile_kernel(const int32_t *__restrict__ src, int64_t n_elems, const PeerTiles T) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // head up to 16-byte alignment (the tiles sit at the same offset in every member's matrix), vector body, tail
    const int64_t head = min(n_elems, (int64_t)((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) / 4);
    const int64_t n_vec = (n_elems - head) / 4;
    const int4 *__restrict__ vsrc = reinterpret_cast<const int4 *>(src + head);
    for (int64_t i = tid; i < n_vec; i += stride) {
        const int4 v = __ldcs(vsrc + i);
        for (int w = 0; w < T.n; ++w) __stcs(reinterpret_cast<int4 *>(T.dst[w] + head) + i, v);
    }
    for (int64_t i = tid; i < head; i += stride)
        for (int w = 0; w < T.n; ++w) T.dst[w][i] = src[i];
    for (int64_t i = head + 4 * n_vec + tid; i < n_elems; i += stride)
        for (int w = 0; w < T.n; ++w) T.dst[w][i] = src[i];
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
    q.si = __ldg(y.rseq_info + draw.x * y.n_cls + cls);
    pair_keys(q, draw, width, ts, seq_len, ws, we);
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

// Thread = (feature, iteration); sums[r, g] written once, iter_sums by atomics.
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
            const int4 rec = __ldg(F.rec + p);
            const int tile = rec_tile(rec.x);
            const int2 ti = __ldg(P.tile_info + tile);
            const int2 d = __ldg(P.draw_tab + r * P.n_blocks + tile);
            const double v = P.permute ? weighted_pair<true>(P.y, W, d, rec_cls(rec.x), ti.x, rec.w, ti.y, rec.y, rec.z)
                                       : weighted_pair<false>(P.y, W, d, rec_cls(rec.x), ti.x, rec.w, ti.y, rec.y, rec.z);
            acc += rec_sign(rec.x) * v;
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
// SIMPLE: one class of y, bootstrap kinds, no exclusion, no table wanted -- the common fused-counts call; the general
// branches are compiled out.
// Sum over the warp of non-negative per-lane counts (each below 2^31) without wrapping: a few blocks over very many
// ranges can push 32 lanes past 2^32, so the halves are reduced separately.
__device__ __forceinline__ unsigned long long warp_sum_u64(int v) {
    const unsigned lo = __reduce_add_sync(kFull, (unsigned)v & 0xffffu), hi = __reduce_add_sync(kFull, (unsigned)v >> 16);
    return ((unsigned long long)hi << 16) + lo;
}

constexpr int kRowsThreads = 256;  // 128 measured the same
template <bool SIMPLE>
__global__ void __launch_bounds__(kRowsThreads, 6) boot_block_rows_kernel(const BootParams P, const TileExcl X, int ipc) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
  constexpr int RPT = 1;
#pragma unroll 1
  for (int it = 0; it < ipc; ++it) {  // a CTA keeps its 256 blocks for `ipc` iterations (fewer, longer-lived CTAs)
    const int64_t r0 = ((int64_t)blockIdx.y * ipc + it) * RPT;
    if (r0 >= P.n_iter) break;
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
                const int32_t t = (!SIMPLE && P.dst_tile) ? __ldg(P.dst_tile + flat) : (int32_t)b;
                const int32_t ts = __ldg(P.tile_start + t);
                const int32_t seq_len = (int32_t)__ldg(P.seqlen + __ldg(P.tile_seq + t));
                if (P.draw_tab) P.draw_tab[(r0 + j) * P.n_blocks + t] = d;  // the count kernel looks draws up by tile
                if (!SIMPLE && P.item_rec) {  // materialising: the fill pass streams canonical window [lo, hi) of this block
                    int32_t lo, hi;
                    locate_block(P.y, d.x, d.y, (int64_t)d.y + width - 1, lo, hi);
                    P.item_rec[flat] = make_int4(lo, hi, d.x, d.y);
                }
                int hits = 0;
                const int n_cls = SIMPLE ? 1 : P.y.n_cls;
                for (int cls = 0; cls < n_cls; ++cls) {  // one pass per strand class of y (one when y is unstranded)
                    const int4 si = __ldg(P.y.rseq_info + d.x * n_cls + cls);
                    const int32_t ie = idx_start_le(P.y, si, d.y + (width - 1));
                    // hits: overlap of the bait block (bootstrap) / start inside the tile (permute)
                    const int h1 = ie - ((!SIMPLE && P.permute) ? idx_start_le(P.y, si, d.y - 1) : idx_end_lt(P.y, si, d.y));
                    int dropped = 0;  // only a tile that overhangs its sequence end can push hits beyond it
                    if (P.drop_zero && ts + (width - 1) > seq_len) dropped = max(0, ie - idx_start_le(P.y, si, seq_len - (ts - d.y)));
                    hits += h1;
                    if (SIMPLE || !P.rows_windows_only) rows[j] += h1 - dropped;
                    if (!SIMPLE && X.off) {  // minus the hits that overlap an exclude region after the move ("trim": plus the pieces)
                        const int32_t x1 = __ldg(X.off + t * n_cls + cls + 1);
                        for (int32_t k = __ldg(X.off + t * n_cls + cls); k < x1; ++k) {
                            const int4 w = __ldg(X.rec + k);
                            rows[j] += w.z * (P.permute ? count_pair<true>(P.y, d, cls, width, ts, seq_len, w.x, w.y)
                                                        : count_pair<false>(P.y, d, cls, width, ts, seq_len, w.x, w.y));
                        }
                    }
                }
                hits_sum += hits;
                if (!SIMPLE && P.item_rows) P.item_rows[(r0 + j) * P.n_blocks + b] = rows[j];
            }
        }
    }
    // every warp adds its share straight to the per-iteration tallies (spread over the iterations' addresses): no
    // shared-memory stage, no barrier, a warp retires as soon as its 32 blocks are done
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const unsigned long long s = warp_sum_u64(rows[j]);
        if (lane == 0 && s && r0 + j < P.n_iter) atomicAdd(P.iter_nranges + r0 + j, s);
    }
    const unsigned long long hs = warp_sum_u64(hits_sum);
    if (lane == 0 && P.n_hits && hs) atomicAdd(P.n_hits + (r0 & (kHitSlots - 1)), hs);
  }
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

}  // namespace nrb
