/*
 * nr_oracle.c -- CPU ORACLE for the nullranges bootRanges hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (nullranges_b200/, libnrb200.so) never links, imports or calls it.
 *
 * What it is: a plain-C, line-by-line restatement of the reference algorithm
 *   R/bootRanges.R:72-331        (front-end loop, segmented samplers, exclude "drop" / "trim")
 *   R/bootUnsegmented.R:1-191    (unsegmented samplers, shift_and_swap_chrom)
 * plus the semantics of the third-party calls those files make, none of which is
 * vendored under /root/reference (IRanges / GenomicRanges / S4Vectors, Bioconductor 3.17,
 * and base R 4.3's RNG: set.seed, unif_rand, runif, sample, round).  Their published
 * behaviour is restated here; every function cites the reference call site it serves.
 *
 * PARITY STATUS: "parity unpinned" at the value level.  R and Bioconductor are not
 * installed in the build image, and the reference's own tests (tests/testthat/test_boot.R,
 * test_trim_exclude.R) pin structural properties only, never a coordinate or a count.
 * What IS pinned: the R RNG stream (set.seed/runif/sample known answers from real R,
 * tests/test_oracle_golden.py + tests/golden/), Philox4x32-10 (Random123 known answers), the worked
 * examples of SURVEY 8c, the structural properties of the reference tests
 * (tests/test_reference_properties.py), and an independent numpy brute force of the whole iteration
 * (tests/test_oracle_bruteforce.py).
 *
 * Conventions: coordinates are R integers (int32), 1-based closed intervals.  chrom ids
 * are 0-based seqlevel indices.  strand codes: 0 '*', 1 '+', 2 '-'.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

static __thread char orc_errbuf[512];
ORC_API const char *orc_last_error(void) { return orc_errbuf; }
static int orc_fail(const char *msg) {
    snprintf(orc_errbuf, sizeof orc_errbuf, "%s", msg);
    return -1;
}

/* ======================================================================================
 * 1. Base-R RNG (R 4.3 defaults: Mersenne-Twister / Inversion / Rejection)
 *    serves: sample() at R/bootUnsegmented.R:97,122,156, R/bootRanges.R:202,263,296
 *            runif()  at R/bootUnsegmented.R:60,124,     R/bootRanges.R:209,267
 * ====================================================================================== */
#define MT_N 624
#define MT_M 397

typedef struct orc_rng {
    uint32_t mt[MT_N];
    int mti;
} orc_rng;

/* set.seed(seed): 50 rounds of LCG scrambling, then 625 LCG outputs; the first is
 * replaced by mti = 624 so the state is regenerated on the first draw. */
ORC_API orc_rng *orc_rng_new(uint32_t seed) {
    orc_rng *g = (orc_rng *)malloc(sizeof *g);
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    seed = 69069u * seed + 1u; /* i_seed[0], overwritten by mti */
    for (int j = 0; j < MT_N; j++) {
        seed = 69069u * seed + 1u;
        g->mt[j] = seed;
    }
    g->mti = MT_N;
    return g;
}
ORC_API void orc_rng_free(orc_rng *g) { free(g); }

static uint32_t mt_next(orc_rng *g) {
    static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
    if (g->mti >= MT_N) {
        int kk;
        uint32_t y;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            y = (g->mt[kk] & 0x80000000u) | (g->mt[kk + 1] & 0x7fffffffu);
            g->mt[kk] = g->mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < MT_N - 1; kk++) {
            y = (g->mt[kk] & 0x80000000u) | (g->mt[kk + 1] & 0x7fffffffu);
            g->mt[kk] = g->mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (g->mt[MT_N - 1] & 0x80000000u) | (g->mt[0] & 0x7fffffffu);
        g->mt[MT_N - 1] = g->mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
        g->mti = 0;
    }
    uint32_t y = g->mt[g->mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* unif_rand(): MT output scaled into [0,1), then pushed strictly inside (0,1). */
ORC_API double orc_unif_rand(orc_rng *g) {
    const double i2_32m1 = 2.328306437080797e-10;
    double x = (double)mt_next(g) * 2.3283064365386963e-10;
    if (x <= 0.0) return 0.5 * i2_32m1;
    if (1.0 - x <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return x;
}

/* runif(n, a, b) with recycled vector bounds; a == b returns a and consumes NO draw. */
ORC_API void orc_runif(orc_rng *g, int64_t n, const double *a, int64_t na, const double *b,
                       int64_t nb, double *out) {
    for (int64_t i = 0; i < n; i++) {
        double ai = a[i % na], bi = b[i % nb];
        if (ai == bi)
            out[i] = ai;
        else
            out[i] = ai + (bi - ai) * orc_unif_rand(g);
    }
}

/* round(x): digits = 0 is IEEE round-half-to-even (R's private_rint). */
ORC_API double orc_round(double x) { return nearbyint(x); }

/* R_unif_index(dn), sample.kind = "Rejection": draw ceil(log2(dn)) bits in 16-bit chunks,
 * reject values >= dn. */
static double r_unif_index(orc_rng *g, double dn) {
    if (dn <= 0) return 0.0;
    int bits = (int)ceil(log2(dn));
    double dv;
    do {
        int64_t v = 0;
        for (int n = 0; n <= bits; n += 16) {
            int v1 = (int)floor(orc_unif_rand(g) * 65536);
            v = 65536 * v + v1;
        }
        if (bits < 64) v &= (((int64_t)1) << bits) - 1;
        dv = (double)v;
    } while (dn <= dv);
    return dv;
}

/* sample.int(n, size, replace = TRUE) without prob: 1-based results. */
ORC_API void orc_sample_replace(orc_rng *g, double dn, int64_t size, int32_t *out) {
    for (int64_t i = 0; i < size; i++) out[i] = (int32_t)(r_unif_index(g, dn) + 1);
}

/* sample.int(n, k, replace = FALSE): partial Fisher-Yates with swap-delete.
 * (k < 2 takes the replace branch in do_sample; the draw count is the same.) */
ORC_API void orc_sample_noreplace(orc_rng *g, int32_t n, int32_t k, int32_t *out) {
    if (k < 2) {
        for (int i = 0; i < k; i++) out[i] = (int32_t)(r_unif_index(g, (double)n) + 1);
        return;
    }
    int32_t *x = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) x[i] = i;
    int m = n;
    for (int i = 0; i < k; i++) {
        int j = (int)r_unif_index(g, (double)m);
        out[i] = x[j] + 1;
        x[j] = x[--m];
    }
    free(x);
}

/* R's revsort(): heapsort of a[] into DESCENDING order carrying ib[] alongside.
 * Tie order is whatever this particular heapsort produces, so it is restated exactly. */
static void r_revsort(double *a, int32_t *ib, int n) {
    if (n <= 1) return;
    a--;
    ib--;
    int l = (n >> 1) + 1, ir = n;
    for (;;) {
        double ra;
        int32_t ii;
        if (l > 1) {
            l--;
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
        int i = l, j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) ++j;
            if (ra > a[j]) {
                a[i] = a[j];
                ib[i] = ib[j];
                i = j;
                j += j;
            } else
                j = ir + 1;
        }
        a[i] = ra;
        ib[i] = ii;
    }
}

/* sample.int(n, size, replace = TRUE, prob = p): FixupProb normalisation, then
 * ProbSampleReplace (<= 200 categories with n*p > 0.1) or Walker's alias method. */
ORC_API int orc_sample_prob(orc_rng *g, int32_t n, const double *prob, int64_t size,
                            int32_t *out) {
    if (n <= 0) return orc_fail("sample: incorrect number of probabilities");
    double *p = (double *)malloc(sizeof(double) * (size_t)n);
    double sum = 0.0;
    int npos = 0;
    for (int i = 0; i < n; i++) {
        if (!isfinite(prob[i]) || prob[i] < 0) {
            free(p);
            return orc_fail("sample: NA or negative in probability vector");
        }
        if (prob[i] > 0) {
            npos++;
            sum += prob[i];
        }
    }
    if (npos == 0) {
        free(p);
        return orc_fail("sample: too few positive probabilities");
    }
    for (int i = 0; i < n; i++) p[i] = prob[i] / sum;
    int nc = 0;
    for (int i = 0; i < n; i++)
        if (n * p[i] > 0.1) nc++;
    if (nc > 200) {
        /* Walker alias */
        int32_t *HL = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        int32_t *a = (int32_t *)calloc((size_t)n, sizeof(int32_t));
        double *q = (double *)malloc(sizeof(double) * (size_t)n);
        int32_t *H = HL - 1, *L = HL + n;
        for (int i = 0; i < n; i++) {
            q[i] = p[i] * n;
            if (q[i] < 1.0)
                *++H = i;
            else
                *--L = i;
        }
        if (H >= HL && L < HL + n) {
            for (int k = 0; k < n - 1; k++) {
                int i = HL[k];
                int j = *L;
                a[i] = j;
                q[j] += q[i] - 1;
                if (q[j] < 1.0) L++;
                if (L >= HL + n) break;
            }
        }
        for (int i = 0; i < n; i++) q[i] += i;
        for (int64_t i = 0; i < size; i++) {
            double rU = orc_unif_rand(g) * n;
            int k = (int)rU;
            out[i] = (rU < q[k]) ? k + 1 : a[k] + 1;
        }
        free(HL);
        free(a);
        free(q);
    } else {
        int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        for (int i = 0; i < n; i++) perm[i] = i + 1;
        r_revsort(p, perm, n);
        for (int i = 1; i < n; i++) p[i] += p[i - 1];
        int nm1 = n - 1;
        for (int64_t i = 0; i < size; i++) {
            double rU = orc_unif_rand(g);
            int j;
            for (j = 0; j < nm1; j++)
                if (rU <= p[j]) break;
            out[i] = perm[j];
        }
        free(perm);
    }
    free(p);
    return 0;
}

/* ======================================================================================
 * 2. Philox4x32-10 (Salmon et al. 2011) -- the production-mode counter RNG.  The CUDA
 *    engine draws with the same function, so production mode is bit-comparable too.
 * ====================================================================================== */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
}
ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        if (r > 0) {
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        philox_round(c, k);
    }
    memcpy(out, c, sizeof c);
}
static inline uint64_t mulhi64(uint64_t a, uint64_t b) {
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
}
/* Two 64-bit words for (iteration, element, stream) under a 64-bit seed. */
static inline void philox_draw(uint64_t seed, uint64_t iter, uint32_t elem, uint32_t stream,
                               uint64_t *u0, uint64_t *u1) {
    uint32_t ctr[4] = {elem, (uint32_t)iter, (uint32_t)(iter >> 32), stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    *u0 = ((uint64_t)o[1] << 32) | o[0];
    *u1 = ((uint64_t)o[3] << 32) | o[2];
}
/* Integer image of round(runif(1, lo, lo + M)): the two endpoints get half weight,
 * interior points full weight (R/bootUnsegmented.R:60,124; R/bootRanges.R:209,267). */
static inline int64_t philox_round_unif(uint64_t u, int64_t lo, int64_t M) {
    if (M <= 0) return lo;
    uint64_t v = mulhi64(u, (uint64_t)(2 * M));
    return lo + (int64_t)((v + 1) >> 1);
}

/* ======================================================================================
 * 3. Sampling plan: block tables for each sampler of the reference.
 * ====================================================================================== */
enum {
    ORC_UNSEG_ACROSS = 0,  /* R/bootUnsegmented.R:116-143 */
    ORC_UNSEG_WITHIN = 1,  /* R/bootUnsegmented.R:50-75   */
    ORC_SEG_PROP = 2,      /* R/bootRanges.R:179-253      */
    ORC_SEG_NOPROP = 3,    /* R/bootRanges.R:256-331      */
    ORC_PERMUTE_WITHIN = 4,/* R/bootUnsegmented.R:83-109  */
    ORC_PERMUTE_ACROSS = 5 /* R/bootUnsegmented.R:150-165 */
};

typedef struct orc_plan {
    int kind, C, S, nseg;
    int64_t L_b, B, L_g;
    int64_t *seqlen;      /* [C] */
    int64_t *n_per_chrom; /* [C]  ceiling(L_c / L_b) */
    int32_t *seg_chrom, *seg_start, *seg_end, *seg_state; /* [nseg] */
    /* destination tiles (iteration-invariant) */
    int32_t *tile_chrom, *tile_start; /* [B] */
    int32_t *bait_width;              /* [B] width of the bait block paired with tile b */
    int32_t *tile_state;              /* [B] 1-based state (seg kinds) or 0 */
    /* seg_prop: per state */
    int64_t *L_bs;         /* [S+1] block length per state (index 1..S) */
    int64_t *state_off;    /* [S+2] block offset of each state */
    int64_t *state_seg_off;/* [S+2] offset into state_segs */
    int32_t *state_segs;   /* segment ids grouped by state, seg order */
    int64_t *seg_nblk;     /* [nseg] tiles per segment */
} orc_plan;

ORC_API void orc_plan_free(orc_plan *p) {
    if (!p) return;
    free(p->seqlen); free(p->n_per_chrom);
    free(p->seg_chrom); free(p->seg_start); free(p->seg_end); free(p->seg_state);
    free(p->tile_chrom); free(p->tile_start); free(p->bait_width); free(p->tile_state);
    free(p->L_bs); free(p->state_off); free(p->state_seg_off); free(p->state_segs);
    free(p->seg_nblk);
    free(p);
}

static int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

static void *dupmem(const void *src, size_t bytes) {
    void *d = malloc(bytes ? bytes : 1);
    if (bytes) memcpy(d, src, bytes);
    return d;
}

ORC_API orc_plan *orc_plan_new(int kind, int C, const int64_t *seqlen, int64_t L_b, int nseg,
                               const int32_t *seg_chrom, const int32_t *seg_start,
                               const int32_t *seg_end, const int32_t *seg_state) {
    if (C <= 0 || L_b <= 0) { orc_fail("plan: need C > 0 and blockLength > 0"); return NULL; }
    orc_plan *p = (orc_plan *)calloc(1, sizeof *p);
    p->kind = kind; p->C = C; p->L_b = L_b;
    p->seqlen = (int64_t *)dupmem(seqlen, sizeof(int64_t) * (size_t)C);
    p->n_per_chrom = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    p->L_g = 0;
    for (int c = 0; c < C; c++) {
        if (seqlen[c] <= 0) { orc_fail("plan: seqlengths must be known and positive"); orc_plan_free(p); return NULL; }
        p->L_g += seqlen[c];
        p->n_per_chrom[c] = ceil_div(seqlen[c], L_b);
    }
    int unseg = (kind == ORC_UNSEG_ACROSS || kind == ORC_UNSEG_WITHIN ||
                 kind == ORC_PERMUTE_WITHIN || kind == ORC_PERMUTE_ACROSS);
    if (unseg) {
        /* stopifnot(all(chrom_lens >= L_b))  R/bootUnsegmented.R:8 */
        for (int c = 0; c < C; c++)
            if (seqlen[c] < L_b) { orc_fail("all(chrom_lens >= L_b) is not TRUE"); orc_plan_free(p); return NULL; }
        int64_t B = 0;
        for (int c = 0; c < C; c++) B += p->n_per_chrom[c];
        p->B = B;
        p->tile_chrom = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
        p->tile_start = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
        p->bait_width = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
        p->tile_state = (int32_t *)calloc((size_t)B, sizeof(int32_t));
        int64_t b = 0;
        /* tileGenome(cut.last.tile.in.chrom=TRUE) (:129) and successiveIRanges (:63,:93)
         * have the same tile STARTS: 1 + k * L_b per chromosome in seqlevel order. */
        for (int c = 0; c < C; c++)
            for (int64_t k = 0; k < p->n_per_chrom[c]; k++, b++) {
                p->tile_chrom[b] = c;
                p->tile_start[b] = (int32_t)(1 + k * L_b);
                p->bait_width[b] = (int32_t)L_b;
            }
        return p;
    }
    /* segmented kinds */
    if (nseg <= 0) { orc_fail("plan: segmentation is empty"); orc_plan_free(p); return NULL; }
    p->nseg = nseg;
    p->seg_chrom = (int32_t *)dupmem(seg_chrom, sizeof(int32_t) * (size_t)nseg);
    p->seg_start = (int32_t *)dupmem(seg_start, sizeof(int32_t) * (size_t)nseg);
    p->seg_end = (int32_t *)dupmem(seg_end, sizeof(int32_t) * (size_t)nseg);
    p->seg_state = (int32_t *)dupmem(seg_state, sizeof(int32_t) * (size_t)nseg);
    int S = 0;
    for (int j = 0; j < nseg; j++) {
        if (seg_state[j] < 1) { orc_fail("plan: seg$state must be >= 1"); orc_plan_free(p); return NULL; }
        if (seg_chrom[j] < 0 || seg_chrom[j] >= C) { orc_fail("plan: seg seqname not in seqlevels(y)"); orc_plan_free(p); return NULL; }
        if (seg_end[j] < seg_start[j]) { orc_fail("plan: zero-width segment"); orc_plan_free(p); return NULL; }
        if (seg_state[j] > S) S = seg_state[j];
    }
    p->S = S; /* esses <- seq_len(max(state))  R/bootRanges.R:146 */
    p->L_bs = (int64_t *)calloc((size_t)S + 2, sizeof(int64_t));
    p->state_off = (int64_t *)calloc((size_t)S + 2, sizeof(int64_t));
    p->state_seg_off = (int64_t *)calloc((size_t)S + 2, sizeof(int64_t));
    p->state_segs = (int32_t *)malloc(sizeof(int32_t) * (size_t)nseg);
    p->seg_nblk = (int64_t *)calloc((size_t)nseg, sizeof(int64_t));
    int64_t off = 0;
    for (int s = 1; s <= S; s++) {
        p->state_seg_off[s] = off;
        for (int j = 0; j < nseg; j++)
            if (seg_state[j] == s) p->state_segs[off++] = j;
    }
    p->state_seg_off[S + 1] = off;
    for (int s = 1; s <= S; s++) {
        if (p->state_seg_off[s + 1] == p->state_seg_off[s]) {
            orc_fail("plan: a state in 1..max(state) has no segments (reference errors in sample())");
            orc_plan_free(p); return NULL;
        }
        if (kind == ORC_SEG_PROP) {
            /* L_b_per_state <- round(L_b * L_s / L_g)   R/bootRanges.R:181-189 */
            int64_t L_s = 0;
            for (int64_t t = p->state_seg_off[s]; t < p->state_seg_off[s + 1]; t++) {
                int j = p->state_segs[t];
                L_s += (int64_t)seg_end[j] - seg_start[j] + 1;
            }
            double v = (double)L_b * (double)L_s / (double)p->L_g;
            p->L_bs[s] = (int64_t)orc_round(v);
            if (p->L_bs[s] <= 0) { orc_fail("plan: per-state block length rounds to 0"); orc_plan_free(p); return NULL; }
        } else {
            p->L_bs[s] = L_b;
        }
    }
    /* tiles: seq(from=start_j, to=end_j, by=L_bs) per segment (:213-215, :269-271),
     * count = ceiling(width_j / L_bs) (:198, :259); states concatenated in order. */
    int64_t B = 0;
    for (int j = 0; j < nseg; j++) {
        int64_t w = (int64_t)seg_end[j] - seg_start[j] + 1;
        p->seg_nblk[j] = ceil_div(w, p->L_bs[seg_state[j]]);
        B += p->seg_nblk[j];
    }
    p->B = B;
    p->tile_chrom = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    p->tile_start = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    p->bait_width = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    p->tile_state = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    int64_t b = 0;
    for (int s = 1; s <= S; s++) {
        p->state_off[s] = b;
        for (int64_t t = p->state_seg_off[s]; t < p->state_seg_off[s + 1]; t++) {
            int j = p->state_segs[t];
            for (int64_t k = 0; k < p->seg_nblk[j]; k++, b++) {
                p->tile_chrom[b] = seg_chrom[j];
                p->tile_start[b] = (int32_t)(seg_start[j] + k * p->L_bs[s]);
                p->bait_width[b] = (int32_t)p->L_bs[s]; /* rep(L_b_per_state, nblocks_per_state) :240-243; L_b at :321 */
                p->tile_state[b] = s;
            }
        }
    }
    p->state_off[S + 1] = b;
    return p;
}

ORC_API int64_t orc_plan_nblocks(const orc_plan *p) { return p->B; }
ORC_API int orc_plan_nstates(const orc_plan *p) { return p->S; }
ORC_API void orc_plan_tiles(const orc_plan *p, int32_t *tile_chrom, int32_t *tile_start,
                            int32_t *bait_width) {
    memcpy(tile_chrom, p->tile_chrom, sizeof(int32_t) * (size_t)p->B);
    memcpy(tile_start, p->tile_start, sizeof(int32_t) * (size_t)p->B);
    memcpy(bait_width, p->bait_width, sizeof(int32_t) * (size_t)p->B);
}
ORC_API int64_t orc_plan_state_block_length(const orc_plan *p, int s) {
    return (s >= 1 && s <= p->S) ? p->L_bs[s] : -1;
}

/* --------------------------------------------------------------------------------------
 * One iteration of block draws with the R RNG, consuming the stream in the reference's
 * order.  Outputs, all [B]:
 *   bait_chrom/bait_start : where block b is sampled FROM
 *   dst_tile              : index of the tile block b is moved TO (b itself for bootstrap)
 * -------------------------------------------------------------------------------------- */
ORC_API int orc_plan_draw_R(const orc_plan *p, orc_rng *g, int32_t *bait_chrom,
                            int32_t *bait_start, int32_t *dst_tile) {
    const int64_t B = p->B;
    for (int64_t b = 0; b < B; b++) dst_tile[b] = (int32_t)b;
    switch (p->kind) {
    case ORC_UNSEG_ACROSS: {
        /* random_chr <- sample(names(L_s), n, replace=TRUE, prob=L_s)              :122
         * random_start <- round(runif(n, 1, (n_per_chrom[random_chr]-1)*L_b + 1))  :124 */
        double *prob = (double *)malloc(sizeof(double) * (size_t)p->C);
        for (int c = 0; c < p->C; c++) prob[c] = (double)p->seqlen[c];
        int rc = orc_sample_prob(g, p->C, prob, B, bait_chrom);
        free(prob);
        if (rc) return rc;
        for (int64_t b = 0; b < B; b++) {
            int c = bait_chrom[b] - 1;
            bait_chrom[b] = c;
            double a = 1.0, hi = (double)(p->n_per_chrom[c] - 1) * (double)p->L_b + 1.0, u;
            orc_runif(g, 1, &a, 1, &hi, 1, &u);
            bait_start[b] = (int32_t)orc_round(u);
        }
        return 0;
    }
    case ORC_UNSEG_WITHIN: {
        /* per seqlevel: round(runif(n, 1, (n-1)*L_b + 1))                          :58-60 */
        int64_t b = 0;
        for (int c = 0; c < p->C; c++) {
            double a = 1.0, hi = (double)(p->n_per_chrom[c] - 1) * (double)p->L_b + 1.0, u;
            for (int64_t k = 0; k < p->n_per_chrom[c]; k++, b++) {
                orc_runif(g, 1, &a, 1, &hi, 1, &u);
                bait_chrom[b] = c;
                bait_start[b] = (int32_t)orc_round(u);
            }
        }
        return 0;
    }
    case ORC_SEG_PROP: {
        /* lapply over states, each: sample segments by width, then runif starts   :192-224 */
        for (int s = 1; s <= p->S; s++) {
            int64_t n = p->state_off[s + 1] - p->state_off[s];
            int32_t ns = (int32_t)(p->state_seg_off[s + 1] - p->state_seg_off[s]);
            const int32_t *segs = p->state_segs + p->state_seg_off[s];
            double *prob = (double *)malloc(sizeof(double) * (size_t)ns);
            for (int t = 0; t < ns; t++)
                prob[t] = (double)((int64_t)p->seg_end[segs[t]] - p->seg_start[segs[t]] + 1);
            int32_t *rr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
            int rc = orc_sample_prob(g, ns, prob, n, rr);
            free(prob);
            if (rc) { free(rr); return rc; }
            for (int64_t i = 0; i < n; i++) {
                int j = segs[rr[i] - 1];
                double rmin = (double)p->seg_start[j];
                double rmax = (double)(p->seg_nblk[j] - 1) * (double)p->L_bs[s] + rmin, u;
                orc_runif(g, 1, &rmin, 1, &rmax, 1, &u);
                int64_t b = p->state_off[s] + i;
                bait_chrom[b] = p->seg_chrom[j];
                bait_start[b] = (int32_t)orc_round(u);
            }
            free(rr);
        }
        return 0;
    }
    case ORC_SEG_NOPROP: {
        /* one global draw over all segments (:263-267), then per state keep the first
         * rearr_n draws or up-sample with sample(rand_n, rearr_n, replace=TRUE) (:292-301) */
        int64_t n = B;
        double *prob = (double *)malloc(sizeof(double) * (size_t)p->nseg);
        for (int j = 0; j < p->nseg; j++)
            prob[j] = (double)((int64_t)p->seg_end[j] - p->seg_start[j] + 1);
        int32_t *rs = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        int32_t *rstart = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        int rc = orc_sample_prob(g, p->nseg, prob, n, rs);
        free(prob);
        if (rc) { free(rs); free(rstart); return rc; }
        for (int64_t i = 0; i < n; i++) {
            int j = rs[i] - 1;
            double rmin = (double)p->seg_start[j];
            double rmax = (double)(p->seg_nblk[j] - 1) * (double)p->L_b + rmin, u;
            orc_runif(g, 1, &rmin, 1, &rmax, 1, &u);
            rstart[i] = (int32_t)orc_round(u);
        }
        int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        for (int s = 1; s <= p->S; s++) {
            int64_t rand_n = 0;
            for (int64_t i = 0; i < n; i++)
                if (p->seg_state[rs[i] - 1] == s) idx[rand_n++] = (int32_t)i;
            int64_t rearr_n = p->state_off[s + 1] - p->state_off[s];
            if (rand_n < rearr_n) {
                if (rand_n == 0) {
                    free(rs); free(rstart); free(idx);
                    return orc_fail("seg_bootstrap_noprop: a state drew no segments "
                                    "(reference fails: length(random) != length(rearranged))");
                }
                int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * (size_t)rearr_n);
                orc_sample_replace(g, (double)rand_n, rearr_n, ri);
                for (int64_t k = 0; k < rearr_n; k++) {
                    int64_t i = idx[ri[k] - 1], b = p->state_off[s] + k;
                    bait_chrom[b] = p->seg_chrom[rs[i] - 1];
                    bait_start[b] = rstart[i];
                }
                free(ri);
            } else {
                for (int64_t k = 0; k < rearr_n; k++) {
                    int64_t i = idx[k], b = p->state_off[s] + k;
                    bait_chrom[b] = p->seg_chrom[rs[i] - 1];
                    bait_start[b] = rstart[i];
                }
            }
        }
        free(rs); free(rstart); free(idx);
        return 0;
    }
    case ORC_PERMUTE_WITHIN: {
        /* per seqlevel: perm <- sample(n); block k moves to tile perm[k]          :97-101 */
        int64_t off = 0;
        for (int c = 0; c < p->C; c++) {
            int32_t n = (int32_t)p->n_per_chrom[c];
            int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
            orc_sample_noreplace(g, n, n, perm);
            for (int k = 0; k < n; k++) {
                bait_chrom[off + k] = c;
                bait_start[off + k] = p->tile_start[off + k];
                dst_tile[off + k] = (int32_t)(off + perm[k] - 1);
            }
            free(perm);
            off += n;
        }
        return 0;
    }
    case ORC_PERMUTE_ACROSS: {
        /* perm <- sample(length(blocks)); rearranged_blocks <- blocks[perm]        :156-157 */
        int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
        orc_sample_noreplace(g, (int32_t)B, (int32_t)B, perm);
        for (int64_t b = 0; b < B; b++) {
            bait_chrom[b] = p->tile_chrom[b];
            bait_start[b] = p->tile_start[b];
            dst_tile[b] = perm[b] - 1;
        }
        free(perm);
        return 0;
    }
    }
    return orc_fail("plan: unknown kind");
}

/* Production-mode draws (counter-based, integer only).  Distribution-equivalent to the
 * R-RNG draws above; identical arithmetic runs on the GPU (csrc/sampler.cuh). */
ORC_API int orc_plan_draw_philox(const orc_plan *p, uint64_t seed, uint64_t iter,
                                 int32_t *bait_chrom, int32_t *bait_start, int32_t *dst_tile) {
    const int64_t B = p->B;
    for (int64_t b = 0; b < B; b++) dst_tile[b] = (int32_t)b;
    uint64_t u0, u1;
    switch (p->kind) {
    case ORC_UNSEG_ACROSS:
        for (int64_t b = 0; b < B; b++) {
            philox_draw(seed, iter, (uint32_t)b, 0, &u0, &u1);
            int64_t pos = (int64_t)mulhi64(u0, (uint64_t)p->L_g);
            int c = 0;
            int64_t cum = 0;
            while (c < p->C - 1 && pos >= cum + p->seqlen[c]) cum += p->seqlen[c++];
            bait_chrom[b] = c;
            bait_start[b] = (int32_t)philox_round_unif(u1, 1, (p->n_per_chrom[c] - 1) * p->L_b);
        }
        return 0;
    case ORC_UNSEG_WITHIN:
        for (int64_t b = 0; b < B; b++) {
            philox_draw(seed, iter, (uint32_t)b, 0, &u0, &u1);
            int c = p->tile_chrom[b];
            bait_chrom[b] = c;
            bait_start[b] = (int32_t)philox_round_unif(u1, 1, (p->n_per_chrom[c] - 1) * p->L_b);
        }
        return 0;
    case ORC_SEG_PROP:
        for (int s = 1; s <= p->S; s++) {
            const int32_t *segs = p->state_segs + p->state_seg_off[s];
            int64_t ns = p->state_seg_off[s + 1] - p->state_seg_off[s];
            int64_t L_s = 0;
            for (int64_t t = 0; t < ns; t++)
                L_s += (int64_t)p->seg_end[segs[t]] - p->seg_start[segs[t]] + 1;
            for (int64_t b = p->state_off[s]; b < p->state_off[s + 1]; b++) {
                philox_draw(seed, iter, (uint32_t)b, 0, &u0, &u1);
                int64_t pos = (int64_t)mulhi64(u0, (uint64_t)L_s);
                int64_t t = 0, cum = 0;
                for (;;) {
                    int64_t w = (int64_t)p->seg_end[segs[t]] - p->seg_start[segs[t]] + 1;
                    if (t == ns - 1 || pos < cum + w) break;
                    cum += w;
                    t++;
                }
                int j = segs[t];
                bait_chrom[b] = p->seg_chrom[j];
                bait_start[b] = (int32_t)philox_round_unif(u1, p->seg_start[j],
                                                            (p->seg_nblk[j] - 1) * p->L_bs[s]);
            }
        }
        return 0;
    case ORC_SEG_NOPROP: {
        /* Production-mode image of seg_bootstrap_noprop (R/bootRanges.R:256-331).  Draw i of the B global
         * draws (counter i, stream 0) picks a segment in proportion to its width -- segments taken state by
         * state, in seg order inside a state -- and a start round(U(start_j, start_j + (nblk_j - 1) L_b)).
         * State s then owns the draws that fell on its segments, in draw order; its k-th tile takes the k-th
         * of them when there are enough (:297-301), else draw number floor(U * rand_n) with U from counter
         * (tile, stream 2) -- the image of sample(rand_n, rearr_n, replace = TRUE) (:295-296).  A state
         * without any draw is an error, as in the reference (:303). */
        int64_t L_tot = 0;
        for (int j = 0; j < p->nseg; j++) L_tot += (int64_t)p->seg_end[j] - p->seg_start[j] + 1;
        int32_t *d_seg = (int32_t *)malloc(sizeof(int32_t) * (size_t)(B ? B : 1));
        int32_t *d_start = (int32_t *)malloc(sizeof(int32_t) * (size_t)(B ? B : 1));
        for (int64_t i = 0; i < B; i++) {
            philox_draw(seed, iter, (uint32_t)i, 0, &u0, &u1);
            int64_t pos = (int64_t)mulhi64(u0, (uint64_t)L_tot), cum = 0;
            int64_t t = 0;
            for (;; t++) { /* walk state_segs: state-major, seg order inside a state */
                int j = p->state_segs[t];
                int64_t w = (int64_t)p->seg_end[j] - p->seg_start[j] + 1;
                if (t == p->nseg - 1 || pos < cum + w) break;
                cum += w;
            }
            int j = p->state_segs[t];
            d_seg[i] = j;
            d_start[i] = (int32_t)philox_round_unif(u1, p->seg_start[j], (p->seg_nblk[j] - 1) * p->L_b);
        }
        int32_t *list = (int32_t *)malloc(sizeof(int32_t) * (size_t)(B ? B : 1));
        int rc = 0;
        for (int s = 1; s <= p->S && !rc; s++) {
            int64_t rand_n = 0;
            for (int64_t i = 0; i < B; i++)
                if (p->seg_state[d_seg[i]] == s) list[rand_n++] = (int32_t)i;
            int64_t rearr_n = p->state_off[s + 1] - p->state_off[s];
            if (rand_n == 0 && rearr_n > 0) { rc = orc_fail("seg_bootstrap_noprop: a state drew no segments (reference fails: length(random) != length(rearranged))"); break; }
            for (int64_t k = 0; k < rearr_n; k++) {
                int64_t b = p->state_off[s] + k, m = k;
                if (rand_n < rearr_n) {
                    philox_draw(seed, iter, (uint32_t)b, 2, &u0, &u1);
                    m = (int64_t)mulhi64(u0, (uint64_t)rand_n);
                }
                int32_t i = list[m];
                bait_chrom[b] = p->seg_chrom[d_seg[i]];
                bait_start[b] = d_start[i];
            }
        }
        free(d_seg); free(d_start); free(list);
        return rc;
    }
    case ORC_PERMUTE_WITHIN:
    case ORC_PERMUTE_ACROSS: {
        /* Production-mode image of sample(n) (R/bootUnsegmented.R:97, :156): every tile gets a 40-bit
         * Philox key (stream 1); tiles are ranked by (sequence -- within-chromosome variant only --, key),
         * ties by tile index; block b moves to the tile of its rank.  A uniformly random permutation
         * up to key ties (probability ~ B^2 / 2^41 per iteration). */
        typedef struct { uint64_t key; int32_t b; } pk;
        pk *v = (pk *)malloc(sizeof(pk) * (size_t)(B ? B : 1));
        for (int64_t b = 0; b < B; b++) {
            philox_draw(seed, iter, (uint32_t)b, 1, &u0, &u1);
            uint64_t chrom = p->kind == ORC_PERMUTE_WITHIN ? (uint64_t)p->tile_chrom[b] : 0;
            v[b].key = (chrom << 40) | (u0 >> 24);
            v[b].b = (int32_t)b;
            bait_chrom[b] = p->tile_chrom[b];
            bait_start[b] = p->tile_start[b];
        }
        /* stable insertion-free sort: merge sort by key, ties keep ascending b */
        pk *tmp = (pk *)malloc(sizeof(pk) * (size_t)(B ? B : 1));
        for (int64_t w = 1; w < B; w *= 2) {
            for (int64_t lo = 0; lo < B; lo += 2 * w) {
                int64_t mid = lo + w < B ? lo + w : B, hi = lo + 2 * w < B ? lo + 2 * w : B;
                int64_t i = lo, j = mid, k = lo;
                while (i < mid && j < hi) tmp[k++] = (v[j].key < v[i].key) ? v[j++] : v[i++];
                while (i < mid) tmp[k++] = v[i++];
                while (j < hi) tmp[k++] = v[j++];
            }
            pk *t = v; v = tmp; tmp = t;
        }
        for (int64_t q = 0; q < B; q++) dst_tile[v[q].b] = (int32_t)q;
        free(v); free(tmp);
        return 0;
    }
    default:
        return orc_fail("philox draws: implemented for unsegmented (across/within) and "
                        "segmented proportionLength=TRUE bootstrap");
    }
}

/* ======================================================================================
 * 4. Interval algebra: findOverlaps / shift / trim / drop / overlapsAny / countOverlaps
 * ====================================================================================== */
typedef struct orc_index {
    int64_t n;
    int C;
    const int32_t *chrom, *start, *end;
    const int8_t *strand; /* may be NULL = all '*' */
    int64_t *off;         /* [C+1] */
    int32_t *ord;         /* ids grouped by chrom, sorted by (start, id) */
    int32_t *pmax;        /* running max of end along ord within chrom */
    int32_t *wmax;        /* [C] largest width */
    int ord_is_identity;
    int64_t slack;        /* as a countOverlaps() query: maxgap + 1 (0 = plain overlap, maxgap = -1) */
    int64_t minoverlap;   /* as a countOverlaps() query: smallest intersection width that counts (0 = any) */
} orc_index;

static const orc_index *g_sort_ix;
static int cmp_ord(const void *a, const void *b) {
    int32_t i = *(const int32_t *)a, j = *(const int32_t *)b;
    const orc_index *ix = g_sort_ix;
    if (ix->chrom[i] != ix->chrom[j]) return ix->chrom[i] < ix->chrom[j] ? -1 : 1;
    if (ix->start[i] != ix->start[j]) return ix->start[i] < ix->start[j] ? -1 : 1;
    return i < j ? -1 : (i > j);
}

ORC_API void orc_index_free(orc_index *ix) {
    if (!ix) return;
    free(ix->off); free(ix->ord); free(ix->pmax); free(ix->wmax); free(ix);
}

/* The caller keeps chrom/start/end/strand alive for the life of the index. */
ORC_API orc_index *orc_index_new(int64_t n, int C, const int32_t *chrom, const int32_t *start,
                                 const int32_t *end, const int8_t *strand) {
    for (int64_t i = 0; i < n; i++) {
        if (chrom[i] < 0 || chrom[i] >= C) { orc_fail("index: seqname id out of range"); return NULL; }
        if (end[i] < start[i]) { orc_fail("index: zero-width or negative-width range"); return NULL; }
    }
    orc_index *ix = (orc_index *)calloc(1, sizeof *ix);
    ix->n = n; ix->C = C; ix->chrom = chrom; ix->start = start; ix->end = end; ix->strand = strand;
    ix->off = (int64_t *)calloc((size_t)C + 1, sizeof(int64_t));
    ix->ord = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    ix->pmax = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    ix->wmax = (int32_t *)calloc((size_t)C, sizeof(int32_t));
    int sorted = 1;
    for (int64_t i = 0; i < n; i++) {
        ix->ord[i] = (int32_t)i;
        ix->off[chrom[i] + 1]++;
        int32_t w = end[i] - start[i] + 1;
        if (w > ix->wmax[chrom[i]]) ix->wmax[chrom[i]] = w;
        if (i > 0 && (chrom[i] < chrom[i - 1] || (chrom[i] == chrom[i - 1] && start[i] < start[i - 1])))
            sorted = 0;
    }
    for (int c = 0; c < C; c++) ix->off[c + 1] += ix->off[c];
    if (!sorted) {
        g_sort_ix = ix;
        qsort(ix->ord, (size_t)n, sizeof(int32_t), cmp_ord);
    }
    ix->ord_is_identity = sorted;
    for (int c = 0; c < C; c++) {
        int32_t m = INT32_MIN;
        for (int64_t k = ix->off[c]; k < ix->off[c + 1]; k++) {
            int32_t e = end[ix->ord[k]];
            if (e > m) m = e;
            ix->pmax[k] = m;
        }
    }
    return ix;
}

/* countOverlaps(x, boots, maxgap = m) (vignettes/bootRanges.Rmd:532-540 uses maxgap = 1e3): two ranges
 * count as overlapping when the gap between them is <= m (adjacent ranges have gap 0; overlapping
 * ranges -1), i.e. start_q <= end_s + m + 1 and end_q >= start_s - m - 1. */
ORC_API void orc_index_set_maxgap(orc_index *ix, int64_t maxgap) { ix->slack = maxgap < -1 ? 0 : maxgap + 1; }
/* countOverlaps(x, y, minoverlap = k): the intersection must be at least k wide
 * (R/segmentDensity.R:77 counts ranges per genome tile with minoverlap = 8). */
ORC_API void orc_index_set_minoverlap(orc_index *ix, int64_t k) { ix->minoverlap = k < 0 ? 0 : k; }

/* #{k in [lo,hi) : start[ord[k]] <= x} + lo */
static int64_t upper_start(const orc_index *ix, int64_t lo, int64_t hi, int64_t x) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)ix->start[ix->ord[mid]] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static inline int strand_compatible(int a, int b) { return a == 0 || b == 0 || a == b; }

/* ---- THE ZERO-WIDTH RULE (one named constant; the kernels carry the same one, device_views.cuh) --------------
 * The within-chromosome variants keep ranges that trim() shrank to width 0 at [L + 1, L] (no width filter at
 * R/bootUnsegmented.R:73, :107).  Whether such a range overlaps an exclude / query range is decided inside IRanges,
 * which is not in the reference tree and cannot run here (INTEGRATION.md "R-check items"):
 *   0  a zero-width range overlaps nothing                                        (what this build ships)
 *   1  the plain overlap test start_q <= end_s && end_q >= start_s also for width 0, i.e. a zero-width range is a
 *      hit of every range that strictly contains its position (IRanges' documented behaviour for maxgap = -1,
 *      minoverlap = 0); only ranges reaching past the sequence end can tell the two apart.
 * Build with -DORC_ZERO_WIDTH_RULE=1 (and the engine with -DNRB_ZERO_WIDTH_RULE=1) to flip it. */
#ifndef ORC_ZERO_WIDTH_RULE
#define ORC_ZERO_WIDTH_RULE 0
#endif
ORC_API int orc_zero_width_rule(void) { return ORC_ZERO_WIDTH_RULE; }
static inline int zero_width_skips(int64_t s, int64_t e) { return e < s && !ORC_ZERO_WIDTH_RULE; }

/* overlapsAny(range, index) with findOverlaps(type="any") semantics on positive-width
 * ranges: same seqname, start_q <= end_s, end_q >= start_s, compatible strands.
 * A zero-width query range overlaps nothing that lies inside the sequence bounds. */
static int overlaps_any(const orc_index *ix, int c, int64_t s, int64_t e, int strand) {
    if (!ix || zero_width_skips(s, e)) return 0;
    int64_t lo = ix->off[c], hi = upper_start(ix, ix->off[c], ix->off[c + 1], e);
    for (int64_t k = hi - 1; k >= lo; k--) {
        if ((int64_t)ix->pmax[k] < s) break;
        int32_t id = ix->ord[k];
        if ((int64_t)ix->end[id] >= s &&
            strand_compatible(strand, ix->strand ? ix->strand[id] : 0))
            return 1;
    }
    return 0;
}

/* countOverlaps(query, y') accumulated one y' range at a time: counts[g]++ for every
 * query feature g overlapping [s,e] on chrom c (vignettes/bootRanges.Rmd:482-483, 497-503).
 * Returns the number of overlapped features. */
static int64_t count_into(const orc_index *q, int c, int64_t s, int64_t e, int strand,
                          int32_t *counts) {
    if (!q || zero_width_skips(s, e)) return 0;
    if (e < s && q->minoverlap > 0) return 0; /* an empty intersection is never minoverlap wide */
    int64_t n = 0;
    const int64_t m1 = q->slack;
    int64_t lo = q->off[c], hi = upper_start(q, q->off[c], q->off[c + 1], e + m1);
    for (int64_t k = hi - 1; k >= lo; k--) {
        if ((int64_t)q->pmax[k] < s - m1) break;
        int32_t id = q->ord[k];
        if ((int64_t)q->end[id] >= s - m1 &&
            strand_compatible(strand, q->strand ? q->strand[id] : 0)) {
            if (q->minoverlap > 0) {
                int64_t lo2 = s > q->start[id] ? s : q->start[id], hi2 = e < q->end[id] ? e : q->end[id];
                if (hi2 - lo2 + 1 < q->minoverlap) continue;
            }
            n++;
            if (counts) counts[id]++;
        }
    }
    return n;
}

typedef struct {
    int32_t *chrom, *start, *end, *src, *block;
    int64_t n, cap; /* cap < 0: count only */
} orc_rows;

#define ORC_EXCLUDE_DROP 0 /* excl = the exclude ranges; overlapping ranges are removed */
#define ORC_EXCLUDE_TRIM 1 /* excl = gaps(exclude); ranges are cut down to the gaps      */

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return x < y ? -1 : (x > y);
}

/* One bootstrap/permutation iteration over prepared block draws.
 *   findOverlaps(random_blocks, y) + gather + block label   R/bootUnsegmented.R:133-137,
 *                                                           :67-71, R/bootRanges.R:166-170
 *   shift_and_swap_chrom                                    R/bootUnsegmented.R:175-191
 *   exclude "drop"                                          R/bootRanges.R:108-109
 *   fused countOverlaps                                     vignette :482-503
 * Rows come out in ascending (block, y-index) order (SURVEY 8a ordering note).
 * Bootstrap kinds: a y range is a hit of block b when it overlaps the bait block
 * (strand ignored: the bait is '*').  Permute kinds: y belongs to the ONE tile that
 * contains its start (findOverlaps(select="first"), :95, :155). */
static int64_t one_iteration(const orc_plan *p, const orc_index *y, const orc_index *excl, int excl_mode,
                             const orc_index *query, const int32_t *bait_chrom,
                             const int32_t *bait_start, const int32_t *dst_tile,
                             orc_rows *rows, int32_t *counts, int64_t *total_overlaps,
                             int32_t **scratch, int64_t *scratch_cap) {
    const int permute = (p->kind == ORC_PERMUTE_WITHIN || p->kind == ORC_PERMUTE_ACROSS);
    /* within-chrom variants keep zero-width survivors (no width filter at :73, :107) */
    const int drop_zero = !(p->kind == ORC_UNSEG_WITHIN || p->kind == ORC_PERMUTE_WITHIN);
    int64_t nrows = 0, tot = 0;
    for (int64_t b = 0; b < p->B; b++) {
        int c = bait_chrom[b];
        int64_t s = bait_start[b], e = s + p->bait_width[b] - 1;
        int d = dst_tile[b], tc = p->tile_chrom[d];
        int64_t shift = (int64_t)p->tile_start[d] - s; /* block_shift  :179 */
        int64_t Lc = p->seqlen[tc];
        int64_t lo = y->off[c], hi = upper_start(y, y->off[c], y->off[c + 1], e);
        int64_t nh = 0;
        for (int64_t k = hi - 1; k >= lo; k--) {
            if ((int64_t)y->pmax[k] < s) break;
            int32_t id = y->ord[k];
            int hit = permute ? ((int64_t)y->start[id] >= s) : ((int64_t)y->end[id] >= s);
            if (!hit) continue;
            if (nh == *scratch_cap) {
                *scratch_cap = *scratch_cap ? *scratch_cap * 2 : 1024;
                *scratch = (int32_t *)realloc(*scratch, sizeof(int32_t) * (size_t)*scratch_cap);
            }
            (*scratch)[nh++] = id;
        }
        if (y->ord_is_identity) { /* collected descending -> reverse */
            for (int64_t i = 0, j = nh - 1; i < j; i++, j--) {
                int32_t t = (*scratch)[i]; (*scratch)[i] = (*scratch)[j]; (*scratch)[j] = t;
            }
        } else {
            qsort(*scratch, (size_t)nh, sizeof(int32_t), cmp_i32);
        }
        for (int64_t h = 0; h < nh; h++) {
            int32_t id = (*scratch)[h];
            int64_t s2 = y->start[id] + shift, e2 = y->end[id] + shift; /* shift()  :185 */
            /* trim() = restrict(start=1, end=seqlength, keep.all.ranges=TRUE)  :188 */
            if (s2 > Lc) { s2 = Lc + 1; e2 = Lc; }
            else if (e2 < 1) { s2 = 1; e2 = 0; }
            else { if (s2 < 1) s2 = 1; if (e2 > Lc) e2 = Lc; }
            if (drop_zero && e2 < s2) continue; /* y_prime[width(y_prime) > 0]  :189 */
            int st = y->strand ? y->strand[id] : 0;
            if (excl && excl_mode == ORC_EXCLUDE_TRIM) {
                /* excludeOption = "trim"  (R/bootRanges.R:110-115): excl holds
                 * gaps(exclude, end = seqlengths(y)) restricted to strand "*"; the range is replaced by
                 * its intersections with the gaps it overlaps (join_overlap_intersect: one row per
                 * (range, gap) hit, gaps ascending).  Gaps are disjoint and sorted, so the overlapped
                 * ones are consecutive.  A zero-width range: see THE ZERO-WIDTH RULE. */
                if (zero_width_skips(s2, e2)) continue;
                int64_t glo = excl->off[tc], ghi = upper_start(excl, glo, excl->off[tc + 1], e2);
                int64_t k0 = ghi;
                while (k0 > glo && (int64_t)excl->end[excl->ord[k0 - 1]] >= s2) k0--;
                for (int64_t k = k0; k < ghi; k++) {
                    int32_t gid = excl->ord[k];
                    int64_t ps = s2 > excl->start[gid] ? s2 : excl->start[gid];
                    int64_t pe = e2 < excl->end[gid] ? e2 : excl->end[gid];
                    if (rows && rows->cap >= 0 && rows->n < rows->cap) {
                        int64_t r = rows->n;
                        rows->chrom[r] = tc; rows->start[r] = (int32_t)ps; rows->end[r] = (int32_t)pe;
                        rows->src[r] = id; rows->block[r] = (int32_t)b;
                    }
                    if (rows) rows->n++;
                    nrows++;
                    if (query) tot += count_into(query, tc, ps, pe, st, counts);
                }
                continue;
            }
            if (excl && overlaps_any(excl, tc, s2, e2, st)) continue; /* bootRanges.R:109 */
            if (rows && rows->cap >= 0 && rows->n < rows->cap) {
                int64_t r = rows->n;
                rows->chrom[r] = tc; rows->start[r] = (int32_t)s2; rows->end[r] = (int32_t)e2;
                rows->src[r] = id; rows->block[r] = (int32_t)b;
            }
            if (rows) rows->n++;
            nrows++;
            if (query) tot += count_into(query, tc, s2, e2, st, counts);
        }
    }
    if (total_overlaps) *total_overlaps = tot;
    return nrows;
}

/* Rows of ONE iteration.  Call with cap = 0 to size, then again to fill.  Returns the
 * row count (or -1). */
ORC_API int64_t orc_iteration_rows(const orc_plan *p, const orc_index *y, const orc_index *excl, int excl_mode,
                                   const int32_t *bait_chrom, const int32_t *bait_start,
                                   const int32_t *dst_tile, int64_t cap, int32_t *out_chrom,
                                   int32_t *out_start, int32_t *out_end, int32_t *out_src,
                                   int32_t *out_block) {
    orc_rows rows = {out_chrom, out_start, out_end, out_src, out_block, 0, cap};
    int32_t *scratch = NULL;
    int64_t scap = 0;
    one_iteration(p, y, excl, excl_mode, NULL, bait_chrom, bait_start, dst_tile, &rows, NULL, NULL,
                  &scratch, &scap);
    free(scratch);
    return rows.n;
}

/* countOverlaps(query, rows) for an explicit row table (used to cross-check the fused
 * path against the two-step reference workflow). strand may be NULL. */
ORC_API int64_t orc_count_overlaps(const orc_index *query, int64_t n, const int32_t *chrom,
                                   const int32_t *start, const int32_t *end,
                                   const int8_t *strand, int32_t *counts) {
    int64_t tot = 0;
    for (int64_t i = 0; i < n; i++)
        tot += count_into(query, chrom[i], start[i], end[i], strand ? strand[i] : 0, counts);
    return tot;
}

/* Full fused path over R iterations.
 *   rng_mode 0: block draws supplied by the caller, [R x B] each (validation mode: they
 *               were drawn with the R RNG in the reference's order);
 *   rng_mode 1: Philox draws for global iterations iter0 .. iter0+R-1.
 * Outputs: iter_nranges[R] (rows per iteration), iter_totals[R] (sum of overlaps),
 *          feature_counts [R x G] or NULL.  nthreads > 1 spreads iterations over pthreads
 *          (the reference itself is single-threaded R; threads are a courtesy to the
 *          CPU baseline, iterations are independent). */
typedef struct {
    const orc_plan *p; const orc_index *y, *excl, *query; int excl_mode;
    int64_t R; int rng_mode; uint64_t seed, iter0;
    const int32_t *bait_chrom, *bait_start, *dst_tile;
    int64_t *iter_nranges, *iter_totals; int32_t *feature_counts;
    int64_t next; int err; pthread_mutex_t mu;
    char errmsg[256];
} orc_job;

static void *orc_worker(void *arg) {
    orc_job *j = (orc_job *)arg;
    const orc_plan *p = j->p;
    const int64_t B = p->B, G = j->query ? j->query->n : 0;
    int32_t *scratch = NULL, *bc = NULL, *bs = NULL, *dt = NULL;
    int64_t scap = 0;
    if (j->rng_mode == 1) {
        bc = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
        bs = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
        dt = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    }
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t r = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (r >= j->R) break;
        const int32_t *pc, *ps, *pd;
        if (j->rng_mode == 1) {
            if (orc_plan_draw_philox(p, j->seed, j->iter0 + (uint64_t)r, bc, bs, dt)) {
                pthread_mutex_lock(&j->mu);  /* the message lives in this thread's buffer: hand it to the caller */
                if (!j->err) snprintf(j->errmsg, sizeof j->errmsg, "%.250s", orc_errbuf);
                j->err = 1;
                pthread_mutex_unlock(&j->mu);
                continue;
            }
            pc = bc; ps = bs; pd = dt;
        } else {
            pc = j->bait_chrom + r * B; ps = j->bait_start + r * B; pd = j->dst_tile + r * B;
        }
        int32_t *cnt = j->feature_counts ? j->feature_counts + r * G : NULL;
        if (cnt) memset(cnt, 0, sizeof(int32_t) * (size_t)G);
        int64_t tot = 0;
        int64_t nr = one_iteration(p, j->y, j->excl, j->excl_mode, j->query, pc, ps, pd, NULL, cnt, &tot,
                                   &scratch, &scap);
        if (j->iter_nranges) j->iter_nranges[r] = nr;
        if (j->iter_totals) j->iter_totals[r] = tot;
    }
    free(scratch); free(bc); free(bs); free(dt);
    return NULL;
}

ORC_API int orc_boot_counts(const orc_plan *p, const orc_index *y, const orc_index *excl, int excl_mode,
                            const orc_index *query, int64_t R, int rng_mode, uint64_t seed,
                            uint64_t iter0, const int32_t *bait_chrom,
                            const int32_t *bait_start, const int32_t *dst_tile,
                            int64_t *iter_nranges, int64_t *iter_totals,
                            int32_t *feature_counts, int nthreads) {
    orc_job j = {p, y, excl, query, excl_mode, R, rng_mode, seed, iter0, bait_chrom, bait_start, dst_tile,
                 iter_nranges, iter_totals, feature_counts, 0, 0, PTHREAD_MUTEX_INITIALIZER, {0}};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) {
        orc_worker(&j);
    } else {
        pthread_t th[256];
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, orc_worker, &j);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    if (j.err) return orc_fail(j.errmsg);
    return 0;
}

ORC_API int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
