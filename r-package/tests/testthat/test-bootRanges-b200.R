# Replays the reference's own tests (nullranges tests/testthat/test_boot.R, test_trim_exclude.R) against
# nullrangesB200::bootRanges, and -- the value-level pin the reference's tests do not make -- compares every output
# row with nullranges::bootRanges under the same set.seed().  Needs a B200 (the engine has no CPU path); skipped
# where libnrb200 cannot create a handle.
library(GenomicRanges)

has_gpu <- function() !inherits(try(nullrangesB200:::.handle(0L), silent = TRUE), "try-error")

seq_nms <- rep(c("chr1", "chr2", "chr3"), c(4, 5, 2))
gr <- GRanges(seqnames = seq_nms,
              IRanges(start = c(1, 101, 121, 201, 101, 201, 216, 231, 401, 1, 101),
                      width = c(20, 5, 5, 30, 20, 5, 5, 5, 30, 80, 40)),
              seqlengths = c(chr1 = 300, chr2 = 450, chr3 = 200),
              chrom = as.integer(factor(seq_nms)))
seg <- GRanges(c("chr1", "chr1", "chr2", "chr2", "chr3"), IRanges(c(1, 151, 1, 226, 1), c(150, 300, 225, 450, 200)),
               state = c(1, 2, 1, 2, 1))
blockLength <- 100
R <- 5

same_table <- function(a, b) {           # same rows, in the same order, same iter / mcols
  expect_equal(as.character(seqnames(a)), as.character(seqnames(b)))
  expect_equal(start(a), start(b)); expect_equal(end(a), end(b))
  expect_equal(as.character(strand(a)), as.character(strand(b)))
  expect_equal(as.integer(a$iter), as.integer(b$iter))
  expect_equal(a$chrom, b$chrom)
}

test_that("the reference's properties hold (test_boot.R:27-96)", {
  skip_if_not(has_gpu())
  set.seed(1); p <- nullrangesB200::bootRanges(gr, blockLength, R, type = "permute", withinChrom = TRUE)
  tally <- table(as.integer(p$iter), p$chrom)
  expect_true(all(t(tally) == as.integer(table(gr$chrom))))                       # :27-35
  set.seed(1); b <- nullrangesB200::bootRanges(gr, blockLength, R, type = "bootstrap", withinChrom = TRUE)
  origin <- tapply(b$chrom, list(as.integer(b$iter), as.character(seqnames(b))), function(v) length(unique(v)))
  expect_true(all(origin[!is.na(origin)] == 1))                                   # :41-48
  set.seed(1); p <- nullrangesB200::bootRanges(gr, blockLength, R, type = "permute", withinChrom = FALSE)
  expect_equal(as.integer(table(as.integer(p$iter))), rep(length(gr), R))         # :54-61
  set.seed(1); b <- nullrangesB200::bootRanges(gr, blockLength, R, type = "bootstrap", withinChrom = FALSE)
  expect_true(!all(table(as.integer(b$iter)) == length(gr)))                      # :67-74
})

test_that("every row equals nullranges::bootRanges under the same seed", {
  skip_if_not(has_gpu())
  skip_if_not_installed("nullranges")
  for (args in list(list(type = "bootstrap", withinChrom = FALSE), list(type = "bootstrap", withinChrom = TRUE),
                    list(type = "permute", withinChrom = FALSE), list(type = "permute", withinChrom = TRUE),
                    list(seg = seg), list(seg = seg, proportionLength = FALSE))) {
    set.seed(1); want <- do.call(nullranges::bootRanges, c(list(gr, blockLength, R), args))
    set.seed(1); got <- do.call(nullrangesB200::bootRanges, c(list(gr, blockLength, R), args))
    same_table(got, want)
  }
})

test_that("trim excludeOption works as expected (test_trim_exclude.R:4-13) and equals the reference", {
  skip_if_not(has_gpu())
  y <- sort(GRanges("chr1", IRanges(1 + 2 * 0:45, width = 10), strand = rep(c("+", "-"), length = 46), seqlengths = c(chr1 = 100)))
  exclude <- GRanges("chr1", IRanges(36, 45), seqlengths = c(chr1 = 100))
  y$block <- 1
  set.seed(1); br <- nullrangesB200::bootRanges(y, blockLength = 20, exclude = exclude, excludeOption = "trim")
  expect_true(!any(overlapsAny(br, exclude)))
  skip_if_not_installed("nullranges")
  set.seed(1); want <- nullranges::bootRanges(y, blockLength = 20, exclude = exclude, excludeOption = "trim")
  expect_equal(start(br), start(want)); expect_equal(end(br), end(want))
})

test_that("fused counts equal countOverlaps on the reference's table; N GPUs equal one", {
  skip_if_not(has_gpu())
  skip_if_not_installed("nullranges")
  x <- GRanges(c("chr1", "chr2", "chr3"), IRanges(c(1, 90, 1), c(260, 240, 90)), seqlengths = seqlengths(gr))
  set.seed(1); want <- nullranges::bootRanges(gr, blockLength, R)
  set.seed(1); got <- nullrangesB200::bootCounts(gr, x, blockLength, R, rng = "R")
  ref <- vapply(seq_len(R), function(r) countOverlaps(x, want[as.integer(want$iter) == r]), integer(length(x)))
  expect_equal(got$counts, ref)
  a <- nullrangesB200::bootCounts(gr, x, blockLength, 40, rng = "philox", seed = 7)
  b <- try(nullrangesB200::bootCounts(gr, x, blockLength, 40, rng = "philox", seed = 7, gpus = 2L), silent = TRUE)
  if (!inherits(b, "try-error")) expect_equal(a$counts, b$counts)
})
