// Philox4x32-10 counter RNG (Salmon, Moraes, Dror, Shaw 2011) and the integer-only block
// draws of production mode.  One call yields 128 bits: 64 pick WHERE the block is sampled
// from (sequence / segment, proportional to length), 64 pick its start.
//
// The draws are the integer image of the reference's
//   sample(names(L_s), n, replace=TRUE, prob=L_s)                   R/bootUnsegmented.R:122
//   round(runif(n, 1, (n_per_chrom - 1) * L_b + 1))                 R/bootUnsegmented.R:60,124
//   sample(length(seg_s), n, TRUE, prob = width(seg_s)); round(runif(n, rmin, rmax))
//                                                                   R/bootRanges.R:202-209
// No floating point is involved, so host and device agree bit for bit.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define NRB_HD __host__ __device__ __forceinline__
#else
#define NRB_HD inline
#endif

namespace nrb {

NRB_HD uint64_t mul_hi_u64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

NRB_HD uint32_t mul_hi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

struct Philox128 {
    uint64_t lo, hi;
};

// counter = (element, iteration low, iteration high, stream), key = 64-bit seed
NRB_HD Philox128 philox_draw(uint64_t seed, uint64_t iter, uint32_t element, uint32_t stream) {
    uint32_t c0 = element, c1 = (uint32_t)iter, c2 = (uint32_t)(iter >> 32), c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t h0 = mul_hi_u32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mul_hi_u32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return {((uint64_t)c1 << 32) | c0, ((uint64_t)c3 << 32) | c2};
}

NRB_HD uint64_t mul_hi_u64_any(uint64_t a, uint64_t b);

// round(runif(1, lo, lo + span)) on integers: endpoints carry half the weight of interior
// points, exactly like rounding a continuous uniform.
NRB_HD int64_t rounded_uniform(uint64_t bits, int64_t lo, int64_t span) {
    if (span <= 0) return lo;
    const uint64_t v = mul_hi_u64_any(bits, (uint64_t)(2 * span));
    return lo + (int64_t)((v + 1) >> 1);
}

// Largest i in [0, n) with cum[i] <= pos (cum is non-decreasing, cum[0] = 0).
NRB_HD int pick_by_length(const int64_t *cum, int n, int64_t pos) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] <= pos)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// The same answer through a coarse table: tab[k] = pick_by_length(cum, n, k << shift) for the kPickBuckets buckets that
// cover [0, total); a draw then starts at its bucket's entry and walks forward (0 or 1 steps when the elements are wider
// than a bucket, which chromosomes and segmentation states are), falling back to the search after a few steps.
constexpr int kPickBuckets = 1024;
NRB_HD int pick_by_length_tab(const int64_t *cum, int n, int64_t pos, const int32_t *tab, int shift) {
    int i = tab[pos >> shift];
    for (int step = 0; i + 1 < n && cum[i + 1] <= pos; ++step) {
        ++i;
        if (step == 4) return i + pick_by_length(cum + i, n - i, pos);
    }
    return i;
}

// floor(a * b / 2^64) for b < 2^32 (two 32 x 32 -> 64 multiplies instead of the full 64 x 64 high product); exact.
NRB_HD uint64_t mul_hi_u64_small(uint64_t a, uint64_t b) {
    const uint64_t lo = (a & 0xffffffffu) * b, hi = (a >> 32) * b;  // b < 2^32: no overflow
    return (hi + (lo >> 32)) >> 32;
}
NRB_HD uint64_t mul_hi_u64_any(uint64_t a, uint64_t b) { return (b >> 32) ? mul_hi_u64(a, b) : mul_hi_u64_small(a, b); }

}  // namespace nrb
