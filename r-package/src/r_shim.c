/*
 * r_shim.c -- .Call entry points binding libnrb200 (include/nrb200.h) into R.
 *
 * The reference package is pure R (no src/, no useDynLib); this shim is the thin C layer the
 * north star asks for.  It is compiled only where R exists (R CMD INSTALL r-package, with
 * PKG_LIBS = -lnrb200); the build image has no R, so the same ABI is exercised from Python
 * (nullranges_b200/_lib.py) and this file is kept mechanical: borrow the INTEGER()/REAL() pointers,
 * call one nrb200_* function, copy the status out.  R owns every input; nothing is written to them.
 *
 * Error convention: nrb200_* never longjmp and never throw.  On a non-zero status the shim copies
 * the message, leaves the handle valid (the external pointer's finalizer frees the GPU resources), and
 * only then calls Rf_error().
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "nrb200.h"

static void check(nrb200_handle *h, int rc) {
    if (rc != NRB200_OK) {
        char msg[512];
        strncpy(msg, nrb200_last_error(h), sizeof msg - 1);
        msg[sizeof msg - 1] = 0;
        Rf_error("libnrb200 (%d): %s", rc, msg);
    }
}

static void handle_finalizer(SEXP ptr) {
    nrb200_handle *h = (nrb200_handle *)R_ExternalPtrAddr(ptr);
    if (h) nrb200_destroy(h);
    R_ClearExternalPtr(ptr);
}

static nrb200_handle *H(SEXP ptr) {
    nrb200_handle *h = (nrb200_handle *)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("nrb200 handle was destroyed");
    return h;
}

/* NA_integer_ is INT_MIN: reject before it reaches the device */
static const int *ints(SEXP x, const char *what) {
    const int *p = INTEGER(x);
    for (R_xlen_t i = 0, n = XLENGTH(x); i < n; i++)
        if (p[i] == NA_INTEGER) Rf_error("%s contains NA", what);
    return p;
}

static int8_t *strand_codes(SEXP x) { /* integer 0/1/2 -> int8 (R_alloc: freed by R at .Call exit) */
    if (x == R_NilValue) return NULL;
    R_xlen_t n = XLENGTH(x);
    int8_t *out = (int8_t *)R_alloc(n ? n : 1, 1);
    const int *p = ints(x, "strand");
    for (R_xlen_t i = 0; i < n; i++) out[i] = (int8_t)p[i];
    return out;
}

/* devices: integer vector of CUDA ordinals; more than one = a device group (the iterations of every run are sharded
 * over those GPUs from this one R thread, see nrb200_config in nrb200.h) */
SEXP nrb200_R_create(SEXP devices) {
    const R_xlen_t W = XLENGTH(devices);
    if (W < 1) Rf_error("nrb200: no device given");
    const int *dev = ints(devices, "devices");
    nrb200_config cfg = {dev[0], 0, NULL, (int32_t)W, dev};
    nrb200_handle *h = NULL;
    check(NULL, nrb200_create(&cfg, &h));
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* seqlengths: double (may exceed int range in principle), seqid 0-based */
SEXP nrb200_R_set_ranges(SEXP ptr, SEXP seqlengths, SEXP seqid, SEXP start, SEXP end, SEXP strand) {
    R_xlen_t C = XLENGTH(seqlengths);
    int64_t *sl = (int64_t *)R_alloc(C ? C : 1, sizeof(int64_t));
    for (R_xlen_t c = 0; c < C; c++) sl[c] = ISNA(REAL(seqlengths)[c]) ? -1 : (int64_t)REAL(seqlengths)[c];
    check(H(ptr), nrb200_set_ranges(H(ptr), XLENGTH(start), (int32_t)C, sl, ints(seqid, "seqnames"), ints(start, "start"),
                                    ints(end, "end"), strand_codes(strand)));
    int32_t sorted = 0, stranded = 0, wmax = 0;
    check(H(ptr), nrb200_ranges_info(H(ptr), &sorted, &stranded, &wmax));
    return Rf_ScalarLogical(sorted);
}

SEXP nrb200_R_set_query(SEXP ptr, SEXP seqid, SEXP start, SEXP end, SEXP strand, SEXP maxgap, SEXP minoverlap) {
    check(H(ptr), nrb200_set_query_ex(H(ptr), XLENGTH(start), ints(seqid, "seqnames(x)"), ints(start, "start(x)"),
                                      ints(end, "end(x)"), strand_codes(strand), Rf_asInteger(maxgap), Rf_asInteger(minoverlap)));
    return R_NilValue;
}

/* option: 0 = excludeOption "drop", 1 = "trim" */
SEXP nrb200_R_set_exclude(SEXP ptr, SEXP seqid, SEXP start, SEXP end, SEXP strand, SEXP option) {
    check(H(ptr), nrb200_set_exclude(H(ptr), XLENGTH(start), ints(seqid, "seqnames(exclude)"), ints(start, "start(exclude)"),
                                     ints(end, "end(exclude)"), strand_codes(strand)));
    check(H(ptr), nrb200_set_exclude_option(H(ptr), Rf_asInteger(option)));
    return R_NilValue;
}

/* returns list(tile_seqid, tile_start, bait_width) so that R can draw the blocks itself */
SEXP nrb200_R_set_plan(SEXP ptr, SEXP kind, SEXP block_length, SEXP seg_seqid, SEXP seg_start, SEXP seg_end, SEXP seg_state) {
    nrb200_plan_spec spec = {Rf_asInteger(kind), (int64_t)Rf_asReal(block_length), 0, NULL, NULL, NULL, NULL};
    if (seg_start != R_NilValue) {
        spec.n_seg = XLENGTH(seg_start);
        spec.seg_seqid = ints(seg_seqid, "seqnames(seg)");
        spec.seg_start = ints(seg_start, "start(seg)");
        spec.seg_end = ints(seg_end, "end(seg)");
        spec.seg_state = ints(seg_state, "seg$state");
    }
    int64_t B = 0;
    check(H(ptr), nrb200_set_plan(H(ptr), &spec, &B));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    for (int k = 0; k < 3; k++) SET_VECTOR_ELT(out, k, Rf_allocVector(INTSXP, B));
    check(H(ptr), nrb200_get_tiles(H(ptr), INTEGER(VECTOR_ELT(out, 0)), INTEGER(VECTOR_ELT(out, 1)), INTEGER(VECTOR_ELT(out, 2))));
    UNPROTECT(1);
    return out;
}

/* Long runs go to the engine in batches of iterations with R_CheckUserInterrupt() in between (it longjmps out on
 * Ctrl-C: nothing but R-managed memory and the still-valid handle is live at that point).  Iterations are
 * independent and Philox is keyed by the global iteration, so batching does not change any result. */
#define NRB_R_BATCH 4096
static nrb200_draws batch_of(const nrb200_draws *d, int64_t r0, int64_t n, int64_t n_blocks) {
    nrb200_draws b = *d;
    b.n_iter = n;
    b.iter0 = d->iter0 + r0;
    if (d->rng_mode == NRB200_RNG_HOST) {
        b.bait_seqid = d->bait_seqid + r0 * n_blocks;
        b.bait_start = d->bait_start + r0 * n_blocks;
        b.dst_tile = d->dst_tile ? d->dst_tile + r0 * n_blocks : NULL;
    }
    return b;
}
static int64_t blocks_of(const nrb200_draws *d, SEXP bait_start) { /* n_blocks of host-mode draws (0 when unused) */
    return (d->rng_mode == NRB200_RNG_HOST && d->n_iter > 0) ? (int64_t)(XLENGTH(bait_start) / d->n_iter) : 0;
}

static void fill_draws(nrb200_draws *d, SEXP n_iter, SEXP iter0, SEXP seed, SEXP bait_seqid, SEXP bait_start, SEXP dst_tile) {
    memset(d, 0, sizeof *d);
    d->n_iter = (int64_t)Rf_asReal(n_iter);
    d->iter0 = (int64_t)Rf_asReal(iter0);
    if (bait_start == R_NilValue) { /* production mode: Philox on the device */
        d->rng_mode = NRB200_RNG_PHILOX;
        d->seed = (uint64_t)Rf_asReal(seed);
    } else {                        /* validation mode: R drew the blocks under set.seed() */
        d->rng_mode = NRB200_RNG_HOST;
        d->bait_seqid = ints(bait_seqid, "bait_seqid");
        d->bait_start = ints(bait_start, "bait_start");
        d->dst_tile = dst_tile == R_NilValue ? NULL : ints(dst_tile, "dst_tile");
    }
}

/* fused path: list(iter_nranges = double[R], iter_totals = double[R], counts = integer matrix [G x R] or NULL).
 * Totals can exceed 2^31 - 1, so they come back as doubles (exact below 2^53).  The matrix is
 * [n_iter x n_query] row-major in C = [n_query x n_iter] column-major in R: counts[g, r]. */
SEXP nrb200_R_run_counts(SEXP ptr, SEXP n_iter, SEXP iter0, SEXP seed, SEXP bait_seqid, SEXP bait_start, SEXP dst_tile,
                         SEXP n_query, SEXP want_matrix) {
    nrb200_draws d;
    fill_draws(&d, n_iter, iter0, seed, bait_seqid, bait_start, dst_tile);
    const R_xlen_t R = (R_xlen_t)d.n_iter, G = (R_xlen_t)Rf_asReal(n_query);
    int64_t *nr = (int64_t *)R_alloc(R ? R : 1, 8), *tot = (int64_t *)R_alloc(R ? R : 1, 8);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP mat = R_NilValue;
    if (Rf_asLogical(want_matrix) && G > 0) {
        mat = Rf_allocMatrix(INTSXP, (int)G, (int)R);
        SET_VECTOR_ELT(out, 2, mat);
    }
    const int64_t B = blocks_of(&d, bait_start);
    for (int64_t r0 = 0; r0 < (int64_t)R || r0 == 0; r0 += NRB_R_BATCH) {
        const int64_t n = (int64_t)R - r0 < NRB_R_BATCH ? (int64_t)R - r0 : NRB_R_BATCH;
        const nrb200_draws b = batch_of(&d, r0, n, B);
        check(H(ptr), nrb200_run_counts(H(ptr), &b, nr + r0, tot + r0, mat == R_NilValue ? NULL : INTEGER(mat) + r0 * (int64_t)G, NULL));
        if (r0 + NRB_R_BATCH < (int64_t)R) R_CheckUserInterrupt();
    }
    SEXP a = Rf_allocVector(REALSXP, R);
    SET_VECTOR_ELT(out, 0, a); /* reachable from `out` before the next allocation can trigger a collection */
    SEXP b = Rf_allocVector(REALSXP, R);
    SET_VECTOR_ELT(out, 1, b);
    for (R_xlen_t r = 0; r < R; r++) {
        REAL(a)[r] = (double)nr[r];
        REAL(b)[r] = (double)tot[r];
    }
    UNPROTECT(1);
    return out;
}

/* weighted statistic: list(iter_sums = double[R], sums = double matrix [G x R] or NULL).  weights: double vector
 * along y (NA rejected). */
SEXP nrb200_R_run_weighted(SEXP ptr, SEXP weights, SEXP n_iter, SEXP iter0, SEXP seed, SEXP bait_seqid, SEXP bait_start,
                           SEXP dst_tile, SEXP n_query, SEXP want_matrix) {
    for (R_xlen_t i = 0, n = XLENGTH(weights); i < n; i++)
        if (ISNA(REAL(weights)[i])) Rf_error("weights contain NA");
    check(H(ptr), nrb200_set_weights(H(ptr), REAL(weights)));
    nrb200_draws d;
    fill_draws(&d, n_iter, iter0, seed, bait_seqid, bait_start, dst_tile);
    const R_xlen_t R = (R_xlen_t)d.n_iter, G = (R_xlen_t)Rf_asReal(n_query);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP it = Rf_allocVector(REALSXP, R);
    SET_VECTOR_ELT(out, 0, it);
    SEXP mat = R_NilValue;
    if (Rf_asLogical(want_matrix) && G > 0) {
        mat = Rf_allocMatrix(REALSXP, (int)G, (int)R);
        SET_VECTOR_ELT(out, 1, mat);
    }
    const int64_t B = blocks_of(&d, bait_start);
    for (int64_t r0 = 0; r0 < (int64_t)R || r0 == 0; r0 += NRB_R_BATCH) {
        const int64_t n = (int64_t)R - r0 < NRB_R_BATCH ? (int64_t)R - r0 : NRB_R_BATCH;
        const nrb200_draws b = batch_of(&d, r0, n, B);
        check(H(ptr), nrb200_run_weighted(H(ptr), &b, REAL(it) + r0, mat == R_NilValue ? NULL : REAL(mat) + r0 * (int64_t)G, NULL));
        if (r0 + NRB_R_BATCH < (int64_t)R) R_CheckUserInterrupt();
    }
    UNPROTECT(1);
    return out;
}

/* per-feature moments of the weighted sums, no [G x R] matrix: list(iter_sums = double[R], sum = double[G],
 * sumsq = double[G], n_ge = double[G] (counts of iterations; exact in a double)).  observed: double[G] or NULL. */
SEXP nrb200_R_run_weighted_moments(SEXP ptr, SEXP weights, SEXP observed, SEXP n_iter, SEXP iter0, SEXP seed, SEXP bait_seqid,
                                   SEXP bait_start, SEXP dst_tile, SEXP n_query) {
    for (R_xlen_t i = 0, n = XLENGTH(weights); i < n; i++)
        if (ISNA(REAL(weights)[i])) Rf_error("weights contain NA");
    check(H(ptr), nrb200_set_weights(H(ptr), REAL(weights)));
    nrb200_draws d;
    fill_draws(&d, n_iter, iter0, seed, bait_seqid, bait_start, dst_tile);
    const R_xlen_t R = (R_xlen_t)d.n_iter, G = (R_xlen_t)Rf_asReal(n_query);
    if (observed != R_NilValue && XLENGTH(observed) != G) Rf_error("observed must have one value per feature");
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SEXP it = Rf_allocVector(REALSXP, R);
    SET_VECTOR_ELT(out, 0, it);
    SEXP sum = Rf_allocVector(REALSXP, G);
    SET_VECTOR_ELT(out, 1, sum);
    SEXP sumsq = Rf_allocVector(REALSXP, G);
    SET_VECTOR_ELT(out, 2, sumsq);
    SEXP nge = Rf_allocVector(REALSXP, G);
    SET_VECTOR_ELT(out, 3, nge);
    int64_t *ge = (int64_t *)R_alloc(G ? G : 1, sizeof(int64_t));
    check(H(ptr), nrb200_run_weighted_moments(H(ptr), &d, observed == R_NilValue ? NULL : REAL(observed), REAL(it), REAL(sum),
                                              REAL(sumsq), ge, NULL));
    for (R_xlen_t g = 0; g < G; g++) REAL(nge)[g] = (double)ge[g];
    UNPROTECT(1);
    return out;
}

/* materialising path: list(lens = double[R], seqid, start, end, src_idx (1-based), block (1-based)) */
SEXP nrb200_R_run_materialize(SEXP ptr, SEXP n_iter, SEXP iter0, SEXP seed, SEXP bait_seqid, SEXP bait_start, SEXP dst_tile) {
    nrb200_draws d;
    fill_draws(&d, n_iter, iter0, seed, bait_seqid, bait_start, dst_tile);
    const R_xlen_t R = (R_xlen_t)d.n_iter;
    int64_t *offs = (int64_t *)R_alloc(R + 1, 8);
    check(H(ptr), nrb200_run_materialize(H(ptr), &d, offs, NULL));
    const int64_t n = offs[R];
    if (n > INT32_MAX) Rf_error("bootRanges: %lld rows exceed R's GRanges limit; use bootCounts() (fused path)", (long long)n);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 6));
    SEXP lens = Rf_allocVector(REALSXP, R);
    SET_VECTOR_ELT(out, 0, lens);
    for (R_xlen_t r = 0; r < R; r++) REAL(lens)[r] = (double)(offs[r + 1] - offs[r]);
    for (int k = 1; k < 6; k++) SET_VECTOR_ELT(out, k, Rf_allocVector(INTSXP, (R_xlen_t)n));
    check(H(ptr), nrb200_fetch_rows(H(ptr), 0, n, INTEGER(VECTOR_ELT(out, 1)), INTEGER(VECTOR_ELT(out, 2)),
                                    INTEGER(VECTOR_ELT(out, 3)), INTEGER(VECTOR_ELT(out, 4)), INTEGER(VECTOR_ELT(out, 5))));
    int *src = INTEGER(VECTOR_ELT(out, 4)), *blk = INTEGER(VECTOR_ELT(out, 5));
    for (int64_t i = 0; i < n; i++) { src[i] += 1; blk[i] += 1; }
    UNPROTECT(1);
    return out;
}

/* per-feature moments of the overlap counts, no [G x R] matrix (bootMoments): list(iter_nranges = double[R],
 * iter_totals = double[R], sum = double[G], sumsq = double[G], n_ge = double[G]).  observed: integer[G] or NULL.
 * Accumulates batch by batch on the device(s); R_CheckUserInterrupt() between batches. */
SEXP nrb200_R_run_moments(SEXP ptr, SEXP observed, SEXP n_iter, SEXP iter0, SEXP seed, SEXP bait_seqid, SEXP bait_start, SEXP dst_tile,
                          SEXP n_query) {
    nrb200_draws d;
    fill_draws(&d, n_iter, iter0, seed, bait_seqid, bait_start, dst_tile);
    const R_xlen_t R = (R_xlen_t)d.n_iter, G = (R_xlen_t)Rf_asReal(n_query);
    if (observed != R_NilValue && XLENGTH(observed) != G) Rf_error("observed must have one value per feature");
    const int *obs = observed == R_NilValue ? NULL : ints(observed, "observed");
    int64_t *nr = (int64_t *)R_alloc(R ? R : 1, 8), *tot = (int64_t *)R_alloc(R ? R : 1, 8);
    int64_t *m = (int64_t *)R_alloc(G ? 3 * G : 1, 8);
    check(H(ptr), nrb200_moments_reset(H(ptr)));
    const int64_t B = blocks_of(&d, bait_start);
    for (int64_t r0 = 0; r0 < (int64_t)R; r0 += NRB_R_BATCH) {
        const int64_t n = (int64_t)R - r0 < NRB_R_BATCH ? (int64_t)R - r0 : NRB_R_BATCH;
        const nrb200_draws b = batch_of(&d, r0, n, B);
        check(H(ptr), nrb200_run_moments(H(ptr), &b, 0, obs, nr + r0, tot + r0, NULL));
        if (r0 + NRB_R_BATCH < (int64_t)R) R_CheckUserInterrupt();
    }
    check(H(ptr), nrb200_moments_fetch(H(ptr), m, m + G, m + 2 * G));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    for (int k = 0; k < 5; k++) SET_VECTOR_ELT(out, k, Rf_allocVector(REALSXP, k < 2 ? R : G));
    for (R_xlen_t r = 0; r < R; r++) {
        REAL(VECTOR_ELT(out, 0))[r] = (double)nr[r];
        REAL(VECTOR_ELT(out, 1))[r] = (double)tot[r];
    }
    for (int k = 0; k < 3; k++)
        for (R_xlen_t g = 0; g < G; g++) REAL(VECTOR_ELT(out, 2 + k))[g] = (double)m[k * G + g];
    UNPROTECT(1);
    return out;
}

/* countOverlaps(x, y) on the original ranges (the observed statistic): integer[G] */
SEXP nrb200_R_count_observed(SEXP ptr, SEXP n_query, SEXP minoverlap) {
    const R_xlen_t G = (R_xlen_t)Rf_asReal(n_query);
    SEXP out = PROTECT(Rf_allocVector(INTSXP, G));
    int64_t total = 0;
    check(H(ptr), nrb200_count_observed_ex(H(ptr), Rf_asInteger(minoverlap), INTEGER(out), &total));
    UNPROTECT(1);
    return out;
}

/* permute variants drawn on the device: the tile (1-based) every block of iterations 1..n_iter moved to, as an
 * integer matrix [n_blocks x n_iter] -- what blockStart needs */
SEXP nrb200_R_fetch_dst_tiles(SEXP ptr, SEXP n_iter, SEXP n_blocks) {
    const int R = Rf_asInteger(n_iter), B = Rf_asInteger(n_blocks);
    SEXP out = PROTECT(Rf_allocMatrix(INTSXP, B, R));
    check(H(ptr), nrb200_fetch_dst_tiles(H(ptr), 0, R, INTEGER(out)));
    int *p = INTEGER(out);
    for (int64_t i = 0, n = (int64_t)R * B; i < n; i++) p[i] += 1;
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"nrb200_R_create", (DL_FUNC)&nrb200_R_create, 1},
    {"nrb200_R_set_ranges", (DL_FUNC)&nrb200_R_set_ranges, 6},
    {"nrb200_R_set_query", (DL_FUNC)&nrb200_R_set_query, 7},
    {"nrb200_R_set_exclude", (DL_FUNC)&nrb200_R_set_exclude, 6},
    {"nrb200_R_set_plan", (DL_FUNC)&nrb200_R_set_plan, 7},
    {"nrb200_R_run_counts", (DL_FUNC)&nrb200_R_run_counts, 9},
    {"nrb200_R_run_weighted", (DL_FUNC)&nrb200_R_run_weighted, 10},
    {"nrb200_R_run_weighted_moments", (DL_FUNC)&nrb200_R_run_weighted_moments, 10},
    {"nrb200_R_run_materialize", (DL_FUNC)&nrb200_R_run_materialize, 7},
    {"nrb200_R_run_moments", (DL_FUNC)&nrb200_R_run_moments, 9},
    {"nrb200_R_count_observed", (DL_FUNC)&nrb200_R_count_observed, 3},
    {"nrb200_R_fetch_dst_tiles", (DL_FUNC)&nrb200_R_fetch_dst_tiles, 3},
    {NULL, NULL, 0}};

void R_init_nullrangesB200(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
