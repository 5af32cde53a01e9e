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

SEXP nrb200_R_create(SEXP device) {
    nrb200_config cfg = {Rf_asInteger(device), 0, NULL};
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
    check(H(ptr), nrb200_run_counts(H(ptr), &d, nr, tot, mat == R_NilValue ? NULL : INTEGER(mat), NULL));
    SEXP a = Rf_allocVector(REALSXP, R), b = Rf_allocVector(REALSXP, R);
    SET_VECTOR_ELT(out, 0, a);
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
    check(H(ptr), nrb200_run_weighted(H(ptr), &d, REAL(it), mat == R_NilValue ? NULL : REAL(mat), NULL));
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
    {NULL, NULL, 0}};

void R_init_nullrangesB200(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
