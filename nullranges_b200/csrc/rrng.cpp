// Base-R random number stream for hosts that are not R (the Python mirror of bootRanges()).
//
// In the R package the wrapper calls R's own sample()/runif() exactly where the reference
// does (R/bootUnsegmented.R:60,97,122,124,156; R/bootRanges.R:202,209,263,267,296), so none
// of this is linked into the R shim.  A non-R host needs the identical stream to reproduce a
// `set.seed(s); bootRanges(...)` run, hence this restatement of R 4.x defaults:
// RNGkind("Mersenne-Twister", "Inversion", "Rejection").
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/nrb200.h"

struct nrb200_rrng {
    static constexpr int N = 624, M = 397;
    uint32_t mt[N];
    int pos;

    explicit nrb200_rrng(uint32_t seed) {
        // set.seed(): scramble 50 times, discard one value (the slot that holds mti), then
        // fill the 624-word state; mti = N forces a twist before the first output.
        for (int i = 0; i < 50; ++i) seed = seed * 69069u + 1u;
        seed = seed * 69069u + 1u;
        for (auto &w : mt) {
            seed = seed * 69069u + 1u;
            w = seed;
        }
        pos = N;
    }

    void twist() {
        auto mix = [](uint32_t hi, uint32_t lo) {
            uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
            return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        };
        for (int k = 0; k < N; ++k) mt[k] = mt[(k + M) % N] ^ mix(mt[k], mt[(k + 1) % N]);
        pos = 0;
    }

    uint32_t bits32() {
        if (pos >= N) twist();
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    // unif_rand(): value in the open interval (0, 1)
    double unif() {
        constexpr double kScale = 2.3283064365386963e-10;  // 2^-32
        constexpr double kEps = 2.328306437080797e-10;     // 1 / (2^32 - 1)
        double x = bits32() * kScale;
        if (x <= 0.0) return 0.5 * kEps;
        if (1.0 - x <= 0.0) return 1.0 - 0.5 * kEps;
        return x;
    }

    // R_unif_index(dn) under sample.kind = "Rejection"
    double index_below(double dn) {
        if (dn <= 0) return 0.0;
        const int nbits = (int)std::ceil(std::log2(dn));
        for (;;) {
            int64_t v = 0;
            for (int got = 0; got <= nbits; got += 16)
                v = v * 65536 + (int)std::floor(unif() * 65536);
            if (nbits < 64) v &= ((int64_t)1 << nbits) - 1;
            if ((double)v < dn) return (double)v;
        }
    }
};

namespace {

// revsort() of R's sort.c: heapsort into descending order, carrying the index vector.  The
// order of tied probabilities is a property of this exact sift-down, so it is kept.
void heap_sort_desc(std::vector<double> &key, std::vector<int32_t> &idx) {
    const int n = (int)key.size();
    if (n < 2) return;
    double *a = key.data() - 1;  // 1-based views
    int32_t *ib = idx.data() - 1;
    int l = n / 2 + 1, ir = n;
    while (true) {
        double ra;
        int32_t ii;
        if (l > 1) {
            --l;
            ra = a[l];
            ii = ib[l];
        } else {
            ra = a[ir];
            ii = ib[ir];
            a[ir] = a[1];
            ib[ir] = ib[1];
            if (--ir == 1) {
                a[1] = ra;
                ib[1] = ii;
                return;
            }
        }
        int i = l, j = 2 * l;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) ++j;
            if (ra > a[j]) {
                a[i] = a[j];
                ib[i] = ib[j];
                i = j;
                j *= 2;
            } else {
                break;
            }
        }
        a[i] = ra;
        ib[i] = ii;
    }
}

// Walker alias tables + draws (walker_ProbSampleReplace), used when > 200 categories matter.
void sample_alias(nrb200_rrng &g, const std::vector<double> &p, int64_t size, int32_t *out) {
    const int n = (int)p.size();
    std::vector<double> q(n);
    std::vector<int32_t> alias(n, 0), HL(n);
    int32_t *H = HL.data() - 1, *L = HL.data() + n;
    for (int i = 0; i < n; ++i) {
        q[i] = p[i] * n;
        if (q[i] < 1.0)
            *++H = i;
        else
            *--L = i;
    }
    if (H >= HL.data() && L < HL.data() + n) {
        for (int k = 0; k < n - 1; ++k) {
            const int i = HL[k], j = *L;
            alias[i] = j;
            q[j] += q[i] - 1;
            if (q[j] < 1.0) ++L;
            if (L >= HL.data() + n) break;
        }
    }
    for (int i = 0; i < n; ++i) q[i] += i;
    for (int64_t s = 0; s < size; ++s) {
        const double rU = g.unif() * n;
        const int k = (int)rU;
        out[s] = (rU < q[k]) ? k + 1 : alias[k] + 1;
    }
}

// ProbSampleReplace: inversion by linear scan over descending cumulative probabilities.
void sample_inversion(nrb200_rrng &g, std::vector<double> p, int64_t size, int32_t *out) {
    const int n = (int)p.size();
    std::vector<int32_t> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i + 1;
    heap_sort_desc(p, perm);
    for (int i = 1; i < n; ++i) p[i] += p[i - 1];
    for (int64_t s = 0; s < size; ++s) {
        const double rU = g.unif();
        int j = 0;
        while (j < n - 1 && !(rU <= p[j])) ++j;
        out[s] = perm[j];
    }
}

}  // namespace

extern "C" {

nrb200_rrng *nrb200_rrng_create(uint32_t seed) { return new nrb200_rrng(seed); }
void nrb200_rrng_destroy(nrb200_rrng *g) { delete g; }
double nrb200_rrng_unif_rand(nrb200_rrng *g) { return g->unif(); }
double nrb200_r_round(double x) { return std::nearbyint(x); }

int nrb200_rrng_runif(nrb200_rrng *g, int64_t n, const double *lo, int64_t n_lo, const double *hi,
                      int64_t n_hi, double *out) {
    if (!g || n < 0 || (n > 0 && (n_lo <= 0 || n_hi <= 0 || !lo || !hi || !out)))
        return NRB200_ERR_INVALID;
    for (int64_t i = 0; i < n; ++i) {
        const double a = lo[i % n_lo], b = hi[i % n_hi];
        if (!(std::isfinite(a) && std::isfinite(b)) || b < a) return NRB200_ERR_INVALID;
        out[i] = (a == b) ? a : a + (b - a) * g->unif();
    }
    return NRB200_OK;
}

int nrb200_rrng_sample(nrb200_rrng *g, int32_t n, int64_t size, int32_t replace, const double *prob,
                       int32_t *out) {
    if (!g || n < 1 || size < 0 || (size > 0 && !out)) return NRB200_ERR_INVALID;
    if (prob) {
        if (!replace) return NRB200_ERR_UNSUPPORTED;
        std::vector<double> p(prob, prob + n);
        double total = 0;
        int positive = 0;
        for (double v : p) {
            if (!std::isfinite(v) || v < 0) return NRB200_ERR_INVALID;
            if (v > 0) {
                ++positive;
                total += v;
            }
        }
        if (!positive) return NRB200_ERR_INVALID;
        int heavy = 0;
        for (double &v : p) {
            v /= total;
            if (n * v > 0.1) ++heavy;
        }
        if (heavy > 200)
            sample_alias(*g, p, size, out);
        else
            sample_inversion(*g, std::move(p), size, out);
        return NRB200_OK;
    }
    if (replace || size < 2) {
        for (int64_t i = 0; i < size; ++i) out[i] = (int32_t)g->index_below((double)n) + 1;
        return NRB200_OK;
    }
    if (size > n) return NRB200_ERR_INVALID;
    std::vector<int32_t> pool(n);
    for (int i = 0; i < n; ++i) pool[i] = i;
    int remaining = n;
    for (int64_t i = 0; i < size; ++i) {
        const int j = (int)g->index_below((double)remaining);
        out[i] = pool[j] + 1;
        pool[j] = pool[--remaining];
    }
    return NRB200_OK;
}

}  // extern "C"
