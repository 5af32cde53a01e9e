#!/usr/bin/env Rscript
# Pins the repo's parity claims against the REAL reference, on any machine that has R (>= 4.3) with Bioconductor's
# nullranges, GenomicRanges, plyranges and jsonlite:
#
#     Rscript tests/golden/make_golden.R            # writes tests/golden/r_reference_golden.json
#     python -m pytest tests/test_r_golden.py       # oracle (CPU) and product (GPU, -m gpu) against that file
#
# The build image has no R, so the file this script writes is not committed by the builder; until somebody runs it,
# value-level parity rests on the C oracle (DESIGN.md "parity unpinned").  Every case below runs the reference's own
# fixtures (tests/testthat/test_boot.R:7-16, :80-82; tests/testthat/test_trim_exclude.R:4-9; R/bootRanges.R:65-69)
# through nullranges::bootRanges under set.seed and dumps EVERY output row (iteration, seqname, start, end, strand,
# index of the source range in y, in the reference's own row order) plus countOverlaps(x, boots[iter == r]) against a
# small query set, and the two facts about IRanges the oracle could not verify (INTEGRATION.md, "R-check items").
suppressPackageStartupMessages({
  library(nullranges); library(GenomicRanges); library(jsonlite)
})
out_path <- file.path(dirname(sub("--file=", "", grep("--file=", commandArgs(), value = TRUE)[1])), "r_reference_golden.json")

strand_code <- function(x) unname(c("*" = 0L, "+" = 1L, "-" = 2L)[as.character(strand(x))])
gr_list <- function(x, lv) {
  if (is.null(x)) return(NULL)
  l <- list(seqid0 = match(as.character(seqnames(x)), lv) - 1L, start = start(x), end = end(x), strand = strand_code(x))
  if (!is.null(x$state)) l$state <- as.integer(x$state)
  l
}

run_case <- function(name, source, y, seed, blockLength, R, query, ...) {
  lv <- seqlevels(y)
  y$src <- seq_along(y)                      # carried through as a metadata column: the source range of every row
  set.seed(seed)
  br <- bootRanges(y, blockLength = blockLength, R = R, ...)
  it <- as.integer(br$iter)
  counts <- t(vapply(seq_len(R), function(r) countOverlaps(query, br[it == r]), integer(length(query))))
  args <- list(...)
  list(name = name, source = source, seed = seed, block_length = blockLength, R = R,
       type = if (is.null(args$type)) "bootstrap" else args$type,
       within_chrom = isTRUE(args$withinChrom),
       proportion_length = if (is.null(args$proportionLength)) TRUE else args$proportionLength,
       exclude_option = if (is.null(args$excludeOption)) "drop" else args$excludeOption,
       seqlevels = lv, seqlengths = as.numeric(seqlengths(y)),
       y = gr_list(y, lv), seg = gr_list(args$seg, lv), exclude = gr_list(args$exclude, lv), query = gr_list(query, lv),
       rows = list(iter = it, seqid0 = match(as.character(seqnames(br)), lv) - 1L, start = start(br), end = end(br),
                   strand = strand_code(br), src_1based = br$src),
       rows_per_iteration = as.integer(table(factor(it, levels = seq_len(R)))),
       counts = counts)
}

# ---- fixtures of the reference's own tests
seq_nms <- rep(c("chr1", "chr2", "chr3"), c(4, 5, 2))
gr <- GRanges(seq_nms, IRanges(start = c(1, 101, 121, 201, 101, 201, 216, 231, 401, 1, 101),
                               width = c(20, 5, 5, 30, 20, 5, 5, 5, 30, 80, 40)),
              seqlengths = c(chr1 = 300, chr2 = 450, chr3 = 200))
seg <- GRanges(c("chr1", "chr1", "chr2", "chr2", "chr3"), IRanges(c(1, 151, 1, 226, 1), c(150, 300, 225, 450, 200)),
               state = c(1, 2, 1, 2, 1), seqlengths = seqlengths(gr))
q3 <- GRanges(c("chr1", "chr1", "chr2", "chr2", "chr3", "chr3"), IRanges(c(1, 120, 90, 300, 1, 150), c(60, 260, 240, 450, 90, 200)),
              seqlengths = seqlengths(gr))
ex3 <- GRanges(c("chr1", "chr2"), IRanges(c(40, 200), c(60, 260)), seqlengths = seqlengths(gr))
y_trim <- sort(GRanges("chr1", IRanges(1 + 2 * 0:45, width = 10), strand = rep(c("+", "-"), length = 46), seqlengths = c(chr1 = 100)))
ex_trim <- GRanges("chr1", IRanges(36, 45), seqlengths = c(chr1 = 100))
q_trim <- GRanges("chr1", IRanges(c(1, 30, 50, 80), c(25, 60, 70, 100)), strand = c("*", "+", "-", "*"), seqlengths = c(chr1 = 100))
gr1 <- GRanges("chr1", IRanges(0:4 * 10 + 1, width = 5), seqlengths = c(chr1 = 50))
q1 <- GRanges("chr1", IRanges(c(1, 20, 44), c(12, 30, 50)), seqlengths = c(chr1 = 50))

cases <- list(
  run_case("roxygen_example", "R/bootRanges.R:65-69", gr1, 1, 10, 1, q1),
  run_case("boot_across", "tests/testthat/test_boot.R:67", gr, 1, 100, 5, q3, type = "bootstrap", withinChrom = FALSE),
  run_case("boot_within", "tests/testthat/test_boot.R:41", gr, 1, 100, 5, q3, type = "bootstrap", withinChrom = TRUE),
  run_case("permute_within", "tests/testthat/test_boot.R:27", gr, 1, 100, 5, q3, type = "permute", withinChrom = TRUE),
  run_case("permute_across", "tests/testthat/test_boot.R:54", gr, 1, 100, 5, q3, type = "permute", withinChrom = FALSE),
  run_case("segmented_prop", "tests/testthat/test_boot.R:80-86", gr, 1, 100, 5, q3, seg = seg),
  run_case("segmented_noprop", "tests/testthat/test_boot.R:80-82 with proportionLength = FALSE", gr, 3, 100, 5, q3, seg = seg, proportionLength = FALSE),
  run_case("exclude_drop", "tests/testthat/test_boot.R:7-16 + exclude", gr, 2, 100, 5, q3, exclude = ex3, excludeOption = "drop"),
  run_case("exclude_trim", "tests/testthat/test_trim_exclude.R:4-11", y_trim, 1, 20, 3, q_trim, exclude = ex_trim, excludeOption = "trim")
)

# ---- the two facts about IRanges the C oracle takes on trust (INTEGRATION.md, "R-check items")
zw <- GRanges("chr1", IRanges(101, width = 0), seqlengths = c(chr1 = 100))           # what trim() leaves at [L + 1, L]
past <- suppressWarnings(GRanges("chr1", IRanges(90, 110), seqlengths = c(chr1 = 100)))   # reaches past the sequence end
inside <- GRanges("chr1", IRanges(90, 100), seqlengths = c(chr1 = 100))
blocks <- IRanges(c(1, 6), width = 10)
yy <- IRanges(c(5, 1, 9, 3, 12), width = 4)
fo <- findOverlaps(blocks, yy)
facts <- list(
  zero_width_hits_range_past_end = as.integer(countOverlaps(past, zw)),     # 0 => zero-width rule 0, 1 => rule 1
  zero_width_hits_range_inside = as.integer(countOverlaps(inside, zw)),     # expected 0 under both rules
  zero_width_overlapsAny_past_end = as.integer(overlapsAny(zw, past)),
  within_block_hit_order = lapply(seq_along(blocks), function(b) subjectHits(fo)[queryHits(fo) == b])  # ascending y index?
)

writeLines(toJSON(list(provenance = list(R = R.version.string, nullranges = as.character(packageVersion("nullranges")),
                                         GenomicRanges = as.character(packageVersion("GenomicRanges")),
                                         IRanges = as.character(packageVersion("IRanges")), date = as.character(Sys.Date())),
                       facts = facts, cases = cases),
                  auto_unbox = TRUE, null = "null", digits = NA, pretty = FALSE), out_path)
cat("wrote", out_path, "\n")
