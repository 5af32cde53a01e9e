// Index construction, device-side draws (permutations, proportionLength = FALSE), observed counts and moments.
// (part of kernels.cuh)
#pragma once
#include "device_views.cuh"

namespace nrb {

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

constexpr int kInspectSeqs = 1024;  // per-sequence tallies are kept in shared memory up to this many sequences

// Every CTA takes one contiguous piece of y (a sorted y then touches one or two sequences per CTA) and tallies the
// per-sequence counts / largest ends in shared memory; global atomics only when the CTA is done.
__global__ void __launch_bounds__(256) inspect_ranges_kernel(const int32_t *__restrict__ seqid, const int32_t *__restrict__ start,
                                                             const int32_t *__restrict__ end, const int8_t *__restrict__ strand,
                                                             int64_t n, int n_seq, InspectOut *__restrict__ out,
                                                             unsigned *__restrict__ seq_count, int32_t *__restrict__ seq_max_end) {
    __shared__ unsigned s_count[kInspectSeqs];
    __shared__ int32_t s_max[kInspectSeqs];
    const bool local = n_seq <= kInspectSeqs;
    if (local) {
        for (int c = threadIdx.x; c < n_seq; c += blockDim.x) {
            s_count[c] = 0;
            s_max[c] = INT32_MIN;
        }
        __syncthreads();
    }
    unsigned *cnt = local ? s_count : seq_count;
    int32_t *mx = local ? s_max : seq_max_end;
    const int64_t piece = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * piece, hi = min(n, lo + piece);
    int err = 0, unsorted = 0, stranded = 0, wmax = 0, wmin = INT32_MAX;
    for (int64_t base = lo; base < hi; base += blockDim.x) {  // CTA-uniform trip count
        const int64_t i = base + threadIdx.x;
        const bool live = i < hi;
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
                atomicAdd(cnt + c, (unsigned)__popc(peers));
                atomicMax(mx + c, m);
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
    if (local) {
        __syncthreads();
        for (int c = threadIdx.x; c < n_seq; c += blockDim.x)
            if (s_count[c]) {
                atomicAdd(seq_count + c, s_count[c]);
                atomicMax(seq_max_end + c, s_max[c]);
            }
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

// ---- canonical order of an unsorted y, on the device ------------------------------------------
// (seqid, start, end) order = stable sort on end, then stable sort on (seqid << 32 | start)
__global__ void canon_keys_end_kernel(const int32_t *__restrict__ end, int64_t n, unsigned *__restrict__ keys, int32_t *__restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = (unsigned)end[i];
        idx[i] = (int32_t)i;
    }
}
__global__ void canon_keys_seq_start_kernel(const int32_t *__restrict__ order, const int32_t *__restrict__ seqid,
                                            const int32_t *__restrict__ start, int64_t n, unsigned long long *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int32_t o = order[i];
        keys[i] = ((unsigned long long)(unsigned)seqid[o] << 32) | (unsigned)start[o];
    }
}
__global__ void gather_ranges_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ seqid, const int32_t *__restrict__ start,
                                     const int32_t *__restrict__ end, int64_t n, int32_t *__restrict__ c_seqid,
                                     int32_t *__restrict__ c_start, int32_t *__restrict__ c_end) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int32_t o = src[i];
        c_seqid[i] = seqid[o];
        c_start[i] = start[o];
        c_end[i] = end[o];
    }
}
__global__ void gather_i8_kernel(const int32_t *__restrict__ src, const int8_t *__restrict__ in, int64_t n, int8_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[src[i]];
}

// Packed rank table (see "direct-address rank tables").  A thread fills kTableRun consecutive entries: one binary
// search for the first, then it walks the values forward bucket by bucket (an entry's end is the next entry's
// start) and packs the bucket's first values into the entry; the CTA's entries go out through shared memory so
// that the stores are contiguous.
constexpr int kTableRun = 8;
__global__ void __launch_bounds__(256) build_rank_table_kernel(const int32_t *__restrict__ values, const int32_t *__restrict__ seq_off,
                                                               const int64_t *__restrict__ rank_off, int n_seq, int shift,
                                                               RankEntry *__restrict__ table) {
    __shared__ RankEntry buf[256 * kTableRun + 256 / 2];  // skewed: one pad entry per two threads' runs
    const int64_t total = rank_off[n_seq];
    const int64_t cta_first = (int64_t)blockIdx.x * (256 * kTableRun);
    const int64_t t0 = cta_first + (int64_t)threadIdx.x * kTableRun;
    if (t0 < total) {
        int c = 0, hi = n_seq;  // sequence that owns entry t0
        while (hi - c > 1) {
            const int m = (c + hi) >> 1;
            if (rank_off[m] <= t0) c = m; else hi = m;
        }
        int64_t seq_last = rank_off[c + 1];
        int32_t a = seq_off[c], b = seq_off[c + 1];
        int64_t bucket = t0 - rank_off[c];
        int32_t p = (bucket << shift) > INT32_MAX ? b : first_ge(values, a, b, (int32_t)(bucket << shift));
        const int out0 = threadIdx.x * kTableRun + threadIdx.x / 2;
        for (int j = 0; j < kTableRun && t0 + j < total; ++j, ++bucket) {
            if (t0 + j >= seq_last) {  // the run crosses into the next sequence (every sequence has entries)
                ++c;
                seq_last = rank_off[c + 1];
                a = seq_off[c];
                b = seq_off[c + 1];
                bucket = 0;
                p = a;
            }
            const int64_t next = (bucket + 1) << shift;
            int32_t q = p;
            if (next > INT32_MAX) {
                q = b;
            } else {
                int steps = 0;
                while (q < b && __ldg(values + q) < (int32_t)next) {
                    ++q;
                    if (++steps == 16) {  // a crowded bucket: finish with a binary search
                        q = first_ge(values, q, b, (int32_t)next);
                        break;
                    }
                }
            }
            buf[out0 + j] = rank_entry_pack(values, a, p, q, bucket << shift, shift);
            p = q;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256 * kTableRun; i += 256)
        if (cta_first + i < total) table[cta_first + i] = buf[i + (i / kTableRun) / 2];
}
inline unsigned rank_table_grid(int64_t entries) { return (unsigned)((entries + 256 * kTableRun - 1) / (256 * kTableRun)); }

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

// Ends in ascending order inside each sequence WITHOUT a sort, for y in canonical (start) order whose widest range is
// short against the spacing of the ranges: end_i = start_i + w_i - 1 is then nearly sorted already.  Rank of end_i =
// #{start_j <= end_i - wmax} (all of those end before it) + the (end, index)-smaller ones among the few ranges that start
// in (end_i - wmax, end_i] -- two probes of the start table and a short scan; every range writes its end to its rank.
__global__ void end_rank_scatter_kernel(const int32_t *__restrict__ seqid, const int32_t *__restrict__ start,
                                        const int32_t *__restrict__ end, const int4 *__restrict__ seq_info,
                                        const RankEntry *__restrict__ rank_start, int rank_shift, int64_t n, int32_t wmax,
                                        int32_t *__restrict__ end_sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 si = __ldg(seq_info + seqid[i]);
    const int32_t e = end[i];
    const int32_t lo = first_ge_ranked(start, rank_start, rank_shift, si, e - wmax + 1);
    const int32_t hi = first_ge_ranked(start, rank_start, rank_shift, si, e + 1);
    int32_t rank = lo;
    for (int32_t j = lo; j < hi; ++j) {
        const int32_t ej = __ldg(end + j);
        rank += (ej < e || (ej == e && j < (int32_t)i)) ? 1 : 0;
    }
    end_sorted[rank] = e;
}

// ---- calibration: what the L2 delivers to scattered 8-byte loads ------------------------------------------------
// The rank-path count kernel is bound by 32-byte sectors pulled through the L2 at random addresses (two table entries
// per evaluation).  This kernel measures that ceiling on the device at hand: every thread issues kCalibLoads
// independent 8-byte loads per round at hashed, sector-distinct positions of an L2-resident buffer.
constexpr int kCalibLoads = 8;
__global__ void __launch_bounds__(256) l2_sector_rate_kernel(const uint2 *__restrict__ buf, uint32_t n_mask, int rounds,
                                                             unsigned long long *__restrict__ sink) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        uint2 v[kCalibLoads];
#pragma unroll
        for (int j = 0; j < kCalibLoads; ++j) {
            x = x * 1664525u + 1013904223u;
            v[j] = __ldg(buf + (((x >> 4) & n_mask) << 2));  // one 8-byte word per 32-byte sector
        }
#pragma unroll
        for (int j = 0; j < kCalibLoads; ++j) acc += v[j].x ^ v[j].y;
        x ^= acc & 1u;  // the next round's addresses depend on this round's data: rounds do not overlap freely
    }
    if (acc == 0x9e3779b9u) atomicAdd(sink, 1ull);
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

// The same for the weighted sums (double): sequential over the batch's iterations, so the result does not depend on
// the launch shape.
__global__ void accumulate_moments_f64_kernel(const double *__restrict__ sums, int64_t n_iter, int64_t n_query,
                                              const double *__restrict__ observed, double *__restrict__ sum,
                                              double *__restrict__ sumsq, long long *__restrict__ n_ge) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_query) return;
    double s = 0.0, q = 0.0;
    long long ge = 0;
    const double obs = observed ? observed[g] : 0.0;
    for (int64_t r = 0; r < n_iter; ++r) {
        const double v = sums[r * n_query + g];
        s += v;
        q += v * v;
        ge += (observed && v >= obs) ? 1 : 0;
    }
    sum[g] += s;
    sumsq[g] += q;
    n_ge[g] += ge;
}

}  // namespace nrb
