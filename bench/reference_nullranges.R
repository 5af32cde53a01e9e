#!/usr/bin/env Rscript
# The TRUE reference arm: nullranges::bootRanges + GenomicRanges::countOverlaps per iteration on the
# synthetic inputs written by bench/export_inputs.py.  Not runnable in the build image (no R there);
# run it wherever R >= 4.3 with Bioconductor 3.17 (nullranges, GenomicRanges, plyranges) exists:
#   python bench/export_inputs.py 2 /tmp/cfg2 && Rscript bench/reference_nullranges.R /tmp/cfg2 20 5e5
# Prints iterations/s and writes per-iteration totals (totals.tsv) for a cross-check against
# nullranges_b200 run with rng = "R" under the same seed.
suppressPackageStartupMessages({ library(nullranges); library(GenomicRanges) })
args <- commandArgs(trailingOnly = TRUE)
dir <- args[1]; R <- as.integer(args[2]); L_b <- as.numeric(args[3])
sl <- read.delim(file.path(dir, "seqlengths.tsv"), header = FALSE)
seqlens <- setNames(sl$V2, sl$V1)
rd <- function(name) {
  d <- read.delim(file.path(dir, paste0(name, ".tsv")))
  gr <- GRanges(factor(d$seqnames, levels = names(seqlens)), IRanges(d$start, d$end), seqlengths = seqlens)
  if ("state" %in% names(d)) gr$state <- d$state
  gr
}
y <- rd("y"); x <- rd("x")
seg <- if (file.exists(file.path(dir, "seg.tsv"))) rd("seg") else NULL
exclude <- if (file.exists(file.path(dir, "exclude.tsv"))) rd("exclude") else NULL
set.seed(2024)
t <- system.time({
  boots <- bootRanges(y, blockLength = L_b, R = R, seg = seg, exclude = exclude)
  per_iter <- split(boots, boots$iter)
  counts <- vapply(per_iter, function(b) countOverlaps(x, b), integer(length(x)))
})
cat(sprintf("reference: %d iterations in %.2f s = %.3f iter/s (single R process)\n", R, t[["elapsed"]], R / t[["elapsed"]]))
write.table(data.frame(iter = seq_len(R), total = colSums(counts), nranges = lengths(per_iter)),
            file.path(dir, "totals.tsv"), sep = "\t", quote = FALSE, row.names = FALSE)
