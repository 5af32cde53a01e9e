/*
 * nrb200.h -- C ABI of libnrb200, the B200 (sm_100a) engine behind nullranges::bootRanges().
 *
 * The reference (nullranges 1.7.10) is pure R: it has NO foreign-function boundary
 * (NAMESPACE has no useDynLib; R/ has no .Call).  This header therefore DEFINES the
 * boundary an R `.Call` shim (r-package/src/r_shim.c) binds; each entry
 * point names the reference code it replaces.  The same ABI is driven from Python
 * (nullranges_b200/_lib.py, ctypes) because R is not installed in the build image.
 *
 * Conventions
 *   - plain C types only; every pointer is caller-owned HOST memory unless a name ends in
 *     `_dev`; the library never keeps a host pointer after a call returns;
 *   - coordinates are R `integer` (int32), 1-based closed intervals [start, end];
 *   - seqname ids are 0-based indices into seqlevels(y); strand codes 0 '*', 1 '+', 2 '-'
 *     (strand pointers may be NULL = all '*');
 *   - every function returns NRB200_OK (0) or a negative code; the text is in
 *     nrb200_last_error().  Nothing is thrown across the boundary, so an R shim can free
 *     its resources and only then call Rf_error();
 *   - a handle is bound to one GPU, or to a device group (nrb200_config.n_devices: the iterations of a run are
 *     sharded over the GPUs of the box from ONE calling thread), and is not thread-safe (R is single-threaded).
 */
#ifndef NRB200_H
#define NRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define NRB200_OK 0
#define NRB200_ERR_INVALID (-1)     /* bad argument (mirrors a stopifnot() of the reference) */
#define NRB200_ERR_CUDA (-2)        /* CUDA runtime failure / no device */
#define NRB200_ERR_STATE (-3)       /* call order: ranges / plan not set */
#define NRB200_ERR_UNSUPPORTED (-4) /* valid in the reference, not (yet) on the device */
#define NRB200_ERR_NOMEM (-5)

typedef struct nrb200_handle nrb200_handle;

/* config flags */
#define NRB200_FLAG_FORCE_STREAM 1u /* always stream the sampled ranges (boot_stream_kernel), even where
                                       the rank path (counting by rank differences) applies; for tests
                                       and for measuring the streaming kernel */

typedef struct nrb200_config {
    int32_t device;     /* CUDA ordinal (ignored when n_devices > 0) */
    uint32_t flags;     /* NRB200_FLAG_* */
    void *stream;       /* cudaStream_t to launch on, or NULL = a private stream (must be NULL for a device group) */
    /* Device group: n_devices > 1 ordinals in `devices` (n_devices == 1: that one device; 0: `device`).  The handle
     * then shards the bootstrap iterations of every run over those GPUs -- replicate(R, ...) of R/bootRanges.R:90 is
     * R independent draws -- and delivers ONE result into the caller's arrays, as the reference does
     * (R/bootRanges.R:120-122): inputs go to every GPU from the caller's single host copy, the window lists are built
     * once, GPU k runs iterations [ceil(k R / W), ceil((k + 1) R / W)) and copies its rows of the count matrix
     * straight into the caller's matrix.  Philox draws are keyed by the global iteration: results are bit-identical
     * for every W.  Every entry point below works on a group handle except the *_resident pair (device pointers of
     * several GPUs) -- NRB200_ERR_UNSUPPORTED; nrb200_run_materialize / nrb200_count_observed run on the first
     * device. */
    int32_t n_devices;
    const int32_t *devices;
} nrb200_config;

int nrb200_create(const nrb200_config *cfg, nrb200_handle **out);
/* Number of GPUs behind the handle (1 unless it is a device group). */
int nrb200_n_devices(const nrb200_handle *h);
void nrb200_destroy(nrb200_handle *h);
/* h may be NULL: error of the last failed nrb200_create() on this thread. */
const char *nrb200_last_error(const nrb200_handle *h);
const char *nrb200_version(void);
/* How this build treats the zero-width ranges the within-chromosome variants keep at [L + 1, L] (no width filter at
 * R/bootUnsegmented.R:73, :107) when they meet exclude / query ranges: 0 = they overlap nothing, 1 = IRanges' plain
 * overlap test (a hit of every range strictly containing the position).  A compile-time constant
 * (NRB_ZERO_WIDTH_RULE, csrc/device_views.cuh); INTEGRATION.md lists the R one-liner that settles it. */
int nrb200_zero_width_rule(void);

/* ---- inputs ------------------------------------------------------------------------ */

/* y of bootRanges(y, ...) plus seqlengths(y) (R/bootRanges.R:80-81: all must be known).
 * Any order is accepted; the engine canonicalises to (seqname, start, end) and remembers
 * the permutation, so src_idx in results always refers to the caller's order. */
int nrb200_set_ranges(nrb200_handle *h, int64_t n, int32_t n_seq, const int64_t *seqlengths,
                      const int32_t *seqid, const int32_t *start, const int32_t *end,
                      const int8_t *strand);

/* The same in two halves, for callers with host-side work to do in between.  _begin checks seqlengths, then only
 * QUEUES the upload of y, its inspection on the device and the copy of the verdict; _end waits for them, reports
 * what nrb200_set_ranges() would have reported and queues the index build.  Between the two, only
 * nrb200_set_query*(), nrb200_set_exclude*() and nrb200_set_plan() may be called (they need seqlengths, not y), and
 * the arrays given to _begin must stay valid and unchanged until _end returns.  The front end uses this to prepare
 * the query and the plan while y is on the wire (R/bootRanges.R:80-88 has no such ordering constraint: its checks
 * are independent of each other). */
int nrb200_set_ranges_begin(nrb200_handle *h, int64_t n, int32_t n_seq, const int64_t *seqlengths,
                            const int32_t *seqid, const int32_t *start, const int32_t *end,
                            const int8_t *strand);
int nrb200_set_ranges_end(nrb200_handle *h);

/* What set_ranges found: is_sorted = the caller's order was already (seqname, start, end) -- for
 * an unstranded y that is the all(y == sort(y)) check of R/bootUnsegmented.R:31, so the wrapper
 * need not sort again; is_stranded = some strand is '+' or '-'; max_width = widest range. */
int nrb200_ranges_info(const nrb200_handle *h, int32_t *is_sorted, int32_t *is_stranded, int32_t *max_width);

/* x of countOverlaps(x, boots) / join_overlap_inner(x, boots)
 * (vignettes/bootRanges.Rmd:482-483, 497-503, 749).  n = 0 clears. */
int nrb200_set_query(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start,
                     const int32_t *end, const int8_t *strand);

/* The same with countOverlaps(x, boots, maxgap = , minoverlap = ) / join_overlap_inner(x, boots, maxgap = ...)
 * (vignettes/bootRanges.Rmd:532-540): maxgap -- ranges separated by at most that many positions count as
 * overlapping (adjacent ranges: gap 0); minoverlap -- the intersection must be at least that wide.  The two
 * cannot be combined (as in GenomicRanges).  maxgap = -1, minoverlap = 0 is nrb200_set_query(). */
int nrb200_set_query_ex(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start,
                        const int32_t *end, const int8_t *strand, int32_t maxgap, int32_t minoverlap);

/* exclude of bootRanges(..., exclude, excludeOption = "drop") (R/bootRanges.R:106-109).
 * n = 0 clears. */
int nrb200_set_exclude(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start,
                       const int32_t *end, const int8_t *strand);

/* excludeOption of bootRanges() (R/bootRanges.R:82, 106-116).  DROP (default): bootstrapped ranges that
 * overlap an exclude range are removed.  TRIM: they are cut down to gaps(exclude, end = seqlengths(y)) --
 * one output row per (range, gap) overlap; needs all(strand(exclude) == "*") (:84).  Call after
 * nrb200_set_exclude(); setting a stranded exclude afterwards fails at the next run. */
#define NRB200_EXCLUDE_DROP 0
#define NRB200_EXCLUDE_TRIM 1
int nrb200_set_exclude_option(nrb200_handle *h, int32_t option);

/* Which sampler of the reference runs (dispatch at R/bootRanges.R:91-104 and
 * R/bootUnsegmented.R:15-40). */
enum {
    NRB200_UNSEG_ACROSS = 0,   /* unseg_bootstrap_across_chrom  R/bootUnsegmented.R:116-143 */
    NRB200_UNSEG_WITHIN = 1,   /* unseg_bootstrap_within_chrom  R/bootUnsegmented.R:50-75   */
    NRB200_SEG_PROP = 2,       /* seg_bootstrap_prop            R/bootRanges.R:179-253      */
    NRB200_SEG_NOPROP = 3,     /* seg_bootstrap_noprop          R/bootRanges.R:256-331      */
    NRB200_PERMUTE_WITHIN = 4, /* unseg_permute_within_chrom    R/bootUnsegmented.R:83-109  */
    NRB200_PERMUTE_ACROSS = 5  /* unseg_permute_across_chrom    R/bootUnsegmented.R:150-165 */
};

typedef struct nrb200_plan_spec {
    int32_t kind;            /* one of the enum above */
    int64_t block_length;    /* blockLength (L_b) */
    /* seg of bootRanges(..., seg): ranges(seg), seqnames(seg), seg$state (>= 1). */
    int64_t n_seg;
    const int32_t *seg_seqid;
    const int32_t *seg_start;
    const int32_t *seg_end;
    const int32_t *seg_state;
} nrb200_plan_spec;

/* Builds the iteration-invariant block tables: the tiles blocks are moved TO
 * (tileGenome / successiveIRanges / seq(by = L_bs)) and the bait width of each block.
 * Requires nrb200_set_ranges() first (it needs seqlengths).  Block order: unsegmented =
 * seqlevel-major; segmented = state 1..S, then segments of that state in `seg` order,
 * then tiles left to right -- the order the reference concatenates them in. */
int nrb200_set_plan(nrb200_handle *h, const nrb200_plan_spec *spec, int64_t *n_blocks);

/* Copies out the block tables ([n_blocks] each; any pointer may be NULL). */
int nrb200_get_tiles(const nrb200_handle *h, int32_t *tile_seqid, int32_t *tile_start,
                     int32_t *bait_width);

/* ---- block draws ------------------------------------------------------------------- */

#define NRB200_RNG_HOST 0   /* validation mode: the caller drew the blocks (R: sample()/runif()
                               under set.seed, in the order of R/bootUnsegmented.R:122-124,
                               :60, R/bootRanges.R:202-209, :263-267, :296) */
#define NRB200_RNG_PHILOX 1 /* production mode: Philox4x32-10 on the device, counter =
                               (global iteration, block), key = seed.  Bootstrap samplers: the block's
                               source; permute variants: a random permutation of the tiles per iteration
                               (40-bit keys, device radix sort); proportionLength = FALSE: global draws,
                               per-state compaction and resampling, all on the device. */

typedef struct nrb200_draws {
    int64_t n_iter;            /* R */
    int64_t iter0;             /* global index of the first iteration (multi-GPU shards) */
    int32_t rng_mode;
    uint64_t seed;             /* NRB200_RNG_PHILOX */
    /* NRB200_RNG_HOST, [n_iter x n_blocks] row-major: */
    const int32_t *bait_seqid; /* seqname id the block is sampled FROM */
    const int32_t *bait_start; /* its start */
    const int32_t *dst_tile;   /* permute kinds: tile index the block moves TO; else NULL */
} nrb200_draws;

typedef struct nrb200_stats {
    double kernel_ms;      /* device time of the bootstrap kernels (CUDA events) */
    double total_ms;       /* device time incl. uploads, memsets and result copies */
    int64_t n_launches;    /* kernels launched by the call */
    int64_t n_hits;        /* sum over iterations of ranges sampled (H of the byte model) */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    double rows_kernel_ms;  /* rank path: boot_block_rows_kernel (draws + rows per block) */
    double count_kernel_ms; /* rank path: boot_rank_loop_kernel; streaming path: boot_stream_kernel (boot_count_kernel for permute) */
    int64_t n_windows;      /* rank path: signed windows evaluated per iteration (0 on the streaming path) */
} nrb200_stats;

/* ---- fused path: bootstrap -> countOverlaps, the R x N table is never built -------- */

/* Replaces, per iteration, the body of replicate() at R/bootRanges.R:90-118
 * (sampler + findOverlaps + shift_and_swap_chrom + exclude drop) AND the user-side
 * countOverlaps / group_by(x_id, iter) of vignettes/bootRanges.Rmd:497-503.
 * Outputs (host, any may be NULL):
 *   iter_nranges   [n_iter]          rows the reference's BootRanges would hold for iter r
 *   iter_totals    [n_iter]          sum over features of the overlap counts
 *   feature_counts [n_iter x n_query] countOverlaps(x, boots[iter == r]) in the caller's
 *                                     query order */
int nrb200_run_counts(nrb200_handle *h, const nrb200_draws *draws, int64_t *iter_nranges,
                      int64_t *iter_totals, int32_t *feature_counts, nrb200_stats *stats);

/* The same with the count matrix delivered as uint16 -- half the bytes over PCIe and into host memory, for callers
 * whose counts fit (overlaps per feature and iteration: cfg 2's mean is 15).  A count above 65535 is stored as 65535
 * and *overflow is set: call nrb200_run_counts() then.  Rank path only (NRB200_ERR_UNSUPPORTED otherwise).  An R
 * caller keeps nrb200_run_counts(): R's integer is 32-bit. */
int nrb200_run_counts_u16(nrb200_handle *h, const nrb200_draws *draws, int64_t *iter_nranges, int64_t *iter_totals,
                          uint16_t *feature_counts, int32_t *overflow, nrb200_stats *stats);

/* Same computation, results left in device memory owned by the handle (valid until the next
 * run or destroy).  want_feature_counts = 0 skips the [n_iter x n_query] matrix. */
int nrb200_run_counts_resident(nrb200_handle *h, const nrb200_draws *draws,
                               int32_t want_feature_counts, nrb200_stats *stats);
int nrb200_resident_pointers(const nrb200_handle *h, void **iter_nranges_dev,
                             void **iter_totals_dev, void **feature_counts_dev);

/* Device group, device-resident consumers: the same computation, and EVERY GPU of the group ends up with the whole
 * result in its own memory -- iter_nranges / iter_totals [n_iter] (uint64) and feature_counts [n_iter x n_query] (int32)
 * -- for work that continues on the GPUs.  This is SURVEY 8e's all-gather of the [R/W x G] tiles without a collective
 * library or a second process: GPU k computes the rows of its iterations into its own copy of the matrix and pushes
 * them into the copies of the other members with coalesced 16-byte stores over peer-mapped pointers (NVLink; peer
 * access is required), all GPUs at once.  On a single-GPU handle it is nrb200_run_counts_resident().  nrb200_gathered_pointers: member = 0 .. nrb200_n_devices(h) - 1; the pointers live on
 * *device and stay valid until the group's next gathered run. */
int nrb200_run_counts_gathered(nrb200_handle *h, const nrb200_draws *draws, nrb200_stats *stats);
int nrb200_gathered_pointers(const nrb200_handle *h, int32_t member, int32_t *device, void **iter_nranges_dev,
                             void **iter_totals_dev, void **feature_counts_dev);

/* Per-feature bootstrap moments without the n_iter x n_query matrix (BASELINE config 4):
 * sum_r c[r,g], sum_r c[r,g]^2, and #{r : c[r,g] >= observed[g]} (observed may be NULL).
 * Accumulates over calls until nrb200_moments_reset(). Iterations are processed in batches
 * of `batch_iter` (0 = pick). */
int nrb200_moments_reset(nrb200_handle *h);
int nrb200_run_moments(nrb200_handle *h, const nrb200_draws *draws, int64_t batch_iter,
                       const int32_t *observed, int64_t *iter_nranges, int64_t *iter_totals,
                       nrb200_stats *stats);
int nrb200_moments_fetch(nrb200_handle *h, int64_t *feat_sum, int64_t *feat_sumsq,
                         int64_t *feat_n_ge_observed);

/* ---- weighted statistic: sum of a numeric column of y over the ranges that land on each feature -----
 * (vignettes/bootRanges.Rmd:532-540: join_overlap_inner(...) %>% group_by(x_id, iter) %>% summarize(sum(signal))).
 * weights: [n of nrb200_set_ranges], in the caller's order; NULL clears.  Call after nrb200_set_ranges().
 * nrb200_run_weighted: iter_sums [n_iter] = sum over features, feature_sums [n_iter x n_query] (may be NULL).
 * Floating point (double): sums are formed from running sums along two orders of y, so they agree with a
 * plain summation to rounding (relative 1e-9 in the tests), not bit for bit. */
int nrb200_set_weights(nrb200_handle *h, const double *weights);
int nrb200_run_weighted(nrb200_handle *h, const nrb200_draws *draws, double *iter_sums, double *feature_sums,
                        nrb200_stats *stats);
/* The same without the [n_iter x n_query] matrix: per-feature moments of the weighted sums over the iterations of
 * this call -- sum_r s[r,g], sum_r s[r,g]^2 (double) and #{r : s[r,g] >= observed[g]} (observed may be NULL) --
 * enough for the z-scores / empirical p-values of vignettes/bootRanges.Rmd:540-560 at any number of iterations.
 * Every output pointer may be NULL.  Accumulate over calls on the host if wanted. */
int nrb200_run_weighted_moments(nrb200_handle *h, const nrb200_draws *draws, const double *observed,
                                double *iter_sums, double *feat_sum, double *feat_sumsq,
                                int64_t *feat_n_ge_observed, nrb200_stats *stats);

/* countOverlaps(x, y) on the ORIGINAL ranges (the observed statistic,
 * vignettes/bootRanges.Rmd:482-483). */
int nrb200_count_observed(nrb200_handle *h, int32_t *feature_counts, int64_t *total);
/* The same with countOverlaps(x, y, minoverlap = minoverlap): only intersections at least that wide count.
 * With x = the tiles of the genome this is the counting step of segmentDensity()
 * (R/segmentDensity.R:77: countOverlaps(query_accept, x, minoverlap = 8)).  Not combinable with maxgap. */
int nrb200_count_observed_ex(nrb200_handle *h, int32_t minoverlap, int32_t *feature_counts, int64_t *total);

/* ---- materialising path: the literal BootRanges table ------------------------------ */

/* Phase 1: runs the bootstrap, keeps the rows on the device, returns iter_offsets
 * [n_iter + 1] (rows of iteration r are [iter_offsets[r], iter_offsets[r+1]) -- this is
 * `lens`/`iter` of R/bootRanges.R:120-122).  Rows are ordered (iter, block, canonical y
 * order). */
int nrb200_run_materialize(nrb200_handle *h, const nrb200_draws *draws, int64_t *iter_offsets,
                           nrb200_stats *stats);
/* Phase 2: copies rows [row0, row0 + n) to the caller.  src_idx is the 0-based index into
 * the caller's y (gather strand(y)/mcols(y) with it); block is the 0-based block index
 * (the reference's internal `block` column, R/bootRanges.R:126). Pointers may be NULL. */
int nrb200_fetch_rows(nrb200_handle *h, int64_t row0, int64_t n, int32_t *seqid,
                      int32_t *start, int32_t *end, int32_t *src_idx, int32_t *block);
/* Permute variants with device draws: the tile each block of iterations [r0, r0 + n) of the last run moved TO
 * ([n x n_blocks], the dst_tile a host-mode caller would have passed in) -- what `blockStart` needs.  */
int nrb200_fetch_dst_tiles(nrb200_handle *h, int64_t r0, int64_t n, int32_t *dst_tile);

/* ---- diagnostics ----------------------------------------------------------------------- */
/* Measures, on the handle's GPU, the rate at which the L2 serves scattered 8-byte loads that each touch their own
 * 32-byte sector of an L2-resident buffer of `buffer_bytes` (a power of two is used; 0 = 32 MiB): the ceiling of the
 * rank-path count kernel, whose evaluations are two such loads each (kernels).  *gbytes_per_s = sectors
 * per second x 32 bytes / 1e9.  bench.py reports the count kernel against this number. */
int nrb200_measure_l2_sector_rate(nrb200_handle *h, int64_t buffer_bytes, double *gbytes_per_s);

/* ---- base-R RNG for hosts that are not R ------------------------------------------- */
/* In R the wrapper draws blocks with R's own sample()/runif().  A non-R host (the Python
 * mirror) needs the same stream under set.seed(): Mersenne-Twister / Inversion / Rejection,
 * R >= 3.6 defaults. */
typedef struct nrb200_rrng nrb200_rrng;
nrb200_rrng *nrb200_rrng_create(uint32_t seed);                 /* set.seed(seed) */
void nrb200_rrng_destroy(nrb200_rrng *g);
double nrb200_rrng_unif_rand(nrb200_rrng *g);
/* runif(n, min, max) with recycled bounds; min == max consumes no draw. */
int nrb200_rrng_runif(nrb200_rrng *g, int64_t n, const double *min, int64_t n_min,
                      const double *max, int64_t n_max, double *out);
/* sample.int(n, size, replace, prob): 1-based results. prob may be NULL. Weighted sampling
 * without replacement is not used by the reference path and returns NRB200_ERR_UNSUPPORTED. */
int nrb200_rrng_sample(nrb200_rrng *g, int32_t n, int64_t size, int32_t replace,
                       const double *prob, int32_t *out);
double nrb200_r_round(double x); /* round(x): half to even */
/* rep(first:(first + n_runs - 1), diff(offsets)) as int32: the dense image of the `iter` column
 * (Rle(factor(rep(seq_len(R), lens))), R/bootRanges.R:122) for hosts without run-length vectors; R builds the Rle
 * from iter_offsets directly and never needs it.  out: [offsets[n_runs] - offsets[0]]. */
void nrb200_expand_runs(const int64_t *offsets, int64_t n_runs, int32_t first, int32_t *out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NRB200_H */
