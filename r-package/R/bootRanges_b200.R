# bootRanges() / bootCounts() served by libnrb200 (B200).  Same arguments as nullranges::bootRanges
# (man/bootRanges.Rd); new arguments default to the reference's behaviour.
#
# rng = "R": block draws come from R's own sample()/runif() under the caller's set.seed(), consumed in
# the order of the reference (R/bootUnsegmented.R:60,122-124; R/bootRanges.R:202-209,263-267,296), so
# the result equals nullranges::bootRanges() run under the same seed.  rng = "philox": draws happen on
# the GPU (counter-based, reproducible from `seed`, independent of how iterations are split).
#
# This file cannot be executed in the build image (no R); its logic is mirrored one to one by
# nullranges_b200/bootranges.py + samplers.py, which the test-suite runs against the oracle.

.nrb_env <- new.env()
# gpus: NULL (one GPU: `device`), N (the first N GPUs of the box) or a vector of CUDA ordinals.  With more than one
# GPU the handle is a device group: every run shards its R iterations over the GPUs (replicate() at
# R/bootRanges.R:90 is R independent draws) and returns ONE result, bit-identical for every number of GPUs.
.devices <- function(device = 0L, gpus = NULL) {
  if (is.null(gpus)) return(as.integer(device))
  if (length(gpus) == 1L) seq_len(as.integer(gpus)) - 1L else as.integer(gpus)
}
.handle <- function(device = 0L, gpus = NULL) {
  dev <- .devices(device, gpus)
  key <- paste0("h", paste(dev, collapse = "_"))
  if (is.null(.nrb_env[[key]])) .nrb_env[[key]] <- .Call("nrb200_R_create", dev)
  .nrb_env[[key]]
}
.strand_code <- function(x) {
  s <- as.character(GenomicRanges::strand(x))
  code <- c("*" = 0L, "+" = 1L, "-" = 2L)[s]
  if (all(code == 0L)) NULL else unname(code)
}
.seqid <- function(x, levels) match(as.character(GenomicRanges::seqnames(x)), levels) - 1L
.kind <- function(seg, proportionLength, type, withinChrom) {
  if (!is.null(seg)) return(if (proportionLength) 2L else 3L)       # type / withinChrom ignored (bootRanges.R:91-98)
  if (type == "bootstrap") (if (withinChrom) 1L else 0L) else (if (withinChrom) 4L else 5L)
}

# One iteration of block draws on R's RNG -> list(seqid (0-based), start, dst (0-based tile or NULL)).
.draw_blocks <- function(kind, L_s, L_b, tiles, seg) {
  n_c <- ceiling(L_s / L_b)
  if (kind == 0L) {                                   # across chromosomes
    pick <- sample(length(L_s), sum(n_c), replace = TRUE, prob = L_s)
    st <- round(runif(sum(n_c), 1, (n_c[pick] - 1) * L_b + 1))
    return(list(seqid = pick - 1L, start = as.integer(st), dst = NULL))
  }
  if (kind == 1L) {                                   # within chromosome: one runif() per seqlevel
    st <- lapply(seq_along(L_s), function(c) round(runif(n_c[c], 1, (n_c[c] - 1) * L_b + 1)))
    return(list(seqid = rep(seq_along(L_s) - 1L, n_c), start = as.integer(unlist(st)), dst = NULL))
  }
  if (kind == 4L) {                                   # permute within: tile k of a chromosome -> perm[k]
    off <- c(0, cumsum(n_c))
    dst <- unlist(lapply(seq_along(L_s), function(c) off[c] + sample(n_c[c]) - 1L))
    return(list(seqid = tiles[[1]], start = tiles[[2]], dst = as.integer(dst)))
  }
  if (kind == 5L) {                                   # permute across
    return(list(seqid = tiles[[1]], start = tiles[[2]], dst = sample(length(tiles[[1]])) - 1L))
  }
  w <- seg$end - seg$start + 1
  states <- seq_len(max(seg$state))
  if (kind == 2L) {                                   # segmented, block length proportional to state size
    L_bs <- round(L_b * vapply(states, function(s) sum(w[seg$state == s]), 0) / sum(L_s))
    out <- lapply(states, function(s) {
      i <- which(seg$state == s)
      nb <- ceiling(w[i] / L_bs[s])
      pick <- sample(length(i), sum(nb), replace = TRUE, prob = w[i])
      lo <- seg$start[i][pick]
      list(seg$seqid[i][pick], round(runif(sum(nb), lo, (nb[pick] - 1) * L_bs[s] + lo)))
    })
    return(list(seqid = unlist(lapply(out, `[[`, 1)), start = as.integer(unlist(lapply(out, `[[`, 2))), dst = NULL))
  }
  nb <- ceiling(w / L_b)                              # kind 3: fixed block length, resampled per state
  pick <- sample(length(w), sum(nb), replace = TRUE, prob = w)
  lo <- seg$start[pick]
  st <- round(runif(sum(nb), lo, (nb[pick] - 1) * L_b + lo))
  out <- lapply(states, function(s) {
    idx <- which(seg$state[pick] == s)
    want <- sum(nb[seg$state == s])
    stopifnot(length(idx) > 0)
    sel <- if (length(idx) < want) sample(length(idx), want, replace = TRUE) else seq_len(want)
    list(seg$seqid[pick[idx]][sel], st[idx][sel])
  })
  list(seqid = unlist(lapply(out, `[[`, 1)), start = as.integer(unlist(lapply(out, `[[`, 2))), dst = NULL)
}

.setup <- function(y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom, device, gpus = NULL) {
  chr_lens <- GenomeInfoDb::seqlengths(y)
  stopifnot(all(!is.na(chr_lens)))
  if (excludeOption == "trim") stopifnot(all(GenomicRanges::strand(exclude) == "*"))   # R/bootRanges.R:83-85
  lv <- GenomeInfoDb::seqlevels(y)
  h <- .handle(device, gpus)
  kind <- .kind(seg, proportionLength, type, withinChrom)
  sorted <- .Call("nrb200_R_set_ranges", h, as.numeric(chr_lens), .seqid(y, lv), GenomicRanges::start(y),
                  GenomicRanges::end(y), .strand_code(y))
  if (kind %in% c(0L, 5L)) {
    ok <- if (is.null(.strand_code(y))) sorted else all(y == GenomicRanges::sort(y))
    stopifnot("all(y == GenomicRanges::sort(y)) is not TRUE" = ok)
  }
  if (is.null(exclude)) .Call("nrb200_R_set_exclude", h, integer(), integer(), integer(), NULL, 0L)
  else {
    exclude <- exclude[as.character(GenomicRanges::seqnames(exclude)) %in% lv]
    .Call("nrb200_R_set_exclude", h, .seqid(exclude, lv), GenomicRanges::start(exclude), GenomicRanges::end(exclude),
          .strand_code(exclude), if (excludeOption == "trim") 1L else 0L)
  }
  segl <- NULL
  if (!is.null(seg)) {
    stopifnot("state" %in% names(S4Vectors::mcols(seg)))
    segl <- list(seqid = .seqid(seg, lv), start = GenomicRanges::start(seg), end = GenomicRanges::end(seg),
                 state = as.integer(seg$state))
  }
  tiles <- .Call("nrb200_R_set_plan", h, kind, as.numeric(blockLength), segl$seqid, segl$start, segl$end, segl$state)
  list(h = h, kind = kind, L_s = as.numeric(chr_lens), tiles = tiles, seg = segl, lv = lv, sorted = sorted)
}

.draws <- function(s, R, blockLength, rng) {
  if (rng == "philox") return(list(NULL, NULL, NULL))
  d <- lapply(seq_len(R), function(i) .draw_blocks(s$kind, s$L_s, blockLength, s$tiles, s$seg))
  list(unlist(lapply(d, `[[`, "seqid")), unlist(lapply(d, `[[`, "start")),
       if (is.null(d[[1]]$dst)) NULL else unlist(lapply(d, `[[`, "dst")))
}

bootRanges <- function(y, blockLength, R = 1, seg = NULL, exclude = NULL, excludeOption = c("drop", "trim"),
                       proportionLength = TRUE, type = c("bootstrap", "permute"), withinChrom = FALSE,
                       storeBlockLength = FALSE, rng = c("R", "philox"), seed = 0, iter0 = 0, storeBlockStart = FALSE,
                       device = 0L, gpus = NULL) {
  excludeOption <- match.arg(excludeOption); type <- match.arg(type); rng <- match.arg(rng)
  s <- .setup(y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom, device, gpus)
  d <- .draws(s, R, blockLength, rng)
  res <- .Call("nrb200_R_run_materialize", s$h, as.numeric(R), as.numeric(iter0), as.numeric(seed), d[[1]], d[[2]], d[[3]])
  if (!s$sorted) {   # row order of the reference follows the order y was given in (see bootranges.py)
    it <- rep(seq_len(R), res[[1]])
    second <- if (s$kind %in% c(4L, 5L)) .seqid(y, s$lv)[res[[5]]] else res[[6]]
    o <- order(it, second, res[[5]], res[[3]])
    for (k in 2:6) res[[k]] <- res[[k]][o]
  }
  src <- res[[5]]
  sn <- res[[2]] + 1L                       # integer codes straight into a factor: no per-row strings
  attr(sn, "levels") <- s$lv
  class(sn) <- "factor"
  br <- GenomicRanges::GRanges(S4Vectors::Rle(sn),
                               IRanges::IRanges(res[[3]], res[[4]]), strand = GenomicRanges::strand(y)[src],
                               seqinfo = GenomeInfoDb::seqinfo(y))
  mc <- S4Vectors::mcols(y)[src, , drop = FALSE]
  mc$block <- NULL
  S4Vectors::mcols(br) <- mc
  # Rle(factor(rep(seq_len(R), lens))) of R/bootRanges.R:122, built from (values, run lengths): no per-row work
  S4Vectors::mcols(br)$iter <- S4Vectors::Rle(factor(seq_len(R), levels = seq_len(R)), res[[1]])
  if (storeBlockLength) S4Vectors::mcols(br)$blockLength <- S4Vectors::Rle(as.integer(blockLength), length(br))
  if (storeBlockStart) {
    # start of the tile the row's block LANDED on (the reference drops its `block` column, R/bootRanges.R:126; this
    # column is the free extra SURVEY fact 5 describes).  Bootstrap kinds: block b lands on tile b.  Permute kinds:
    # on the tile it was drawn for -- R's own draw (d[[3]], 0-based, iteration-major) or the device's permutation.
    dst <- res[[6]]
    if (s$kind %in% c(4L, 5L)) {
      B <- length(s$tiles[[1]])
      perm <- if (rng == "R") matrix(d[[3]] + 1L, nrow = B) else .Call("nrb200_R_fetch_dst_tiles", s$h, as.integer(R), as.integer(B))
      dst <- perm[cbind(res[[6]], rep(seq_len(R), res[[1]]))]
    }
    S4Vectors::mcols(br)$blockStart <- s$tiles[[2]][dst]
  }
  new("BootRanges", br)
}

# countOverlaps(x, bootRanges(y, ...), maxgap = maxgap) per iteration without building the R x N table.
# Returns list(iter_totals, iter_nranges, counts = integer matrix [length(x), R]).
bootCounts <- function(y, x, blockLength, R = 1, seg = NULL, exclude = NULL, excludeOption = c("drop", "trim"),
                       proportionLength = TRUE, type = c("bootstrap", "permute"), withinChrom = FALSE,
                       rng = c("philox", "R"), seed = 0, iter0 = 0, perFeature = TRUE, maxgap = -1L, minoverlap = 0L, weights = NULL,
                       device = 0L, gpus = NULL) {
  excludeOption <- match.arg(excludeOption); type <- match.arg(type); rng <- match.arg(rng)
  s <- .setup(y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom, device, gpus)
  keep <- which(as.character(GenomicRanges::seqnames(x)) %in% s$lv)
  xk <- x[keep]
  .Call("nrb200_R_set_query", s$h, .seqid(xk, s$lv), GenomicRanges::start(xk), GenomicRanges::end(xk), .strand_code(xk),
        as.integer(maxgap), as.integer(minoverlap))
  d <- .draws(s, R, blockLength, rng)
  res <- .Call("nrb200_R_run_counts", s$h, as.numeric(R), as.numeric(iter0), as.numeric(seed), d[[1]], d[[2]], d[[3]],
               as.numeric(length(xk)), perFeature)
  counts <- res[[3]]
  if (!is.null(counts) && length(keep) != length(x)) {
    full <- matrix(0L, length(x), R); full[keep, ] <- counts; counts <- full
  }
  out <- list(iter_nranges = res[[1]], iter_totals = res[[2]], counts = counts)
  if (!is.null(weights)) {   # sum of a numeric column of y over the overlapping ranges (vignette: summarize(sum(signal)))
    w <- as.numeric(S4Vectors::mcols(y)[[weights]])
    ws <- .Call("nrb200_R_run_weighted", s$h, w, as.numeric(R), as.numeric(iter0), as.numeric(seed), d[[1]], d[[2]], d[[3]],
                as.numeric(length(xk)), perFeature)
    sums <- ws[[2]]
    if (!is.null(sums) && length(keep) != length(x)) {
      full <- matrix(0, length(x), R); full[keep, ] <- sums; sums <- full
    }
    out$iter_sums <- ws[[1]]; out$sums <- sums
  }
  out
}

# Per-feature mean / sd / z / empirical p of the weighted statistic over R iterations, reduced on the GPU (no G x R
# matrix): the summary vignettes/bootRanges.Rmd:532-560 computes from the per-iteration table.  observed: the statistic
# on y itself (numeric, one value per feature of x), or NULL for mean and sd only.
bootWeightedMoments <- function(y, x, blockLength, weights, R = 1, observed = NULL, seg = NULL, exclude = NULL,
                                excludeOption = c("drop", "trim"), proportionLength = TRUE, type = c("bootstrap", "permute"),
                                withinChrom = FALSE, rng = c("philox", "R"), seed = 0, iter0 = 0, maxgap = -1L, minoverlap = 0L,
                                device = 0L, gpus = NULL) {
  excludeOption <- match.arg(excludeOption); type <- match.arg(type); rng <- match.arg(rng)
  s <- .setup(y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom, device, gpus)
  stopifnot("every seqlevel of x must be a seqlevel of y" = all(as.character(GenomicRanges::seqnames(x)) %in% s$lv))
  .Call("nrb200_R_set_query", s$h, .seqid(x, s$lv), GenomicRanges::start(x), GenomicRanges::end(x), .strand_code(x),
        as.integer(maxgap), as.integer(minoverlap))
  d <- .draws(s, R, blockLength, rng)
  w <- as.numeric(S4Vectors::mcols(y)[[weights]])
  m <- .Call("nrb200_R_run_weighted_moments", s$h, w, if (is.null(observed)) NULL else as.numeric(observed), as.numeric(R),
             as.numeric(iter0), as.numeric(seed), d[[1]], d[[2]], d[[3]], as.numeric(length(x)))
  mean <- m[[2]] / R
  sd <- sqrt(pmax(m[[3]] - R * mean^2, 0) / max(R - 1, 1))
  out <- list(iter_sums = m[[1]], mean = mean, sd = sd)
  if (!is.null(observed)) {
    out$z <- (observed - mean) / sd
    out$p_ge <- (m[[4]] + 1) / (R + 1)
  }
  out
}

# Per-feature mean / sd / z / empirical p of countOverlaps(x, boots) over R iterations, reduced on the GPU(s): no
# G x R matrix, so R = 10^4 against 2 * 10^5 features (BASELINE config 4) is fine.  observed defaults to
# countOverlaps(x, y) on the original ranges (vignettes/bootRanges.Rmd:482-483).
bootMoments <- function(y, x, blockLength, R = 1, observed = NULL, seg = NULL, exclude = NULL, excludeOption = c("drop", "trim"),
                        proportionLength = TRUE, type = c("bootstrap", "permute"), withinChrom = FALSE, rng = c("philox", "R"),
                        seed = 0, iter0 = 0, maxgap = -1L, minoverlap = 0L, device = 0L, gpus = NULL) {
  excludeOption <- match.arg(excludeOption); type <- match.arg(type); rng <- match.arg(rng)
  s <- .setup(y, blockLength, seg, exclude, excludeOption, proportionLength, type, withinChrom, device, gpus)
  stopifnot("every seqlevel of x must be a seqlevel of y" = all(as.character(GenomicRanges::seqnames(x)) %in% s$lv))
  .Call("nrb200_R_set_query", s$h, .seqid(x, s$lv), GenomicRanges::start(x), GenomicRanges::end(x), .strand_code(x),
        as.integer(maxgap), as.integer(minoverlap))
  if (is.null(observed)) observed <- .Call("nrb200_R_count_observed", s$h, as.numeric(length(x)), as.integer(minoverlap))
  d <- .draws(s, R, blockLength, rng)
  m <- .Call("nrb200_R_run_moments", s$h, as.integer(observed), as.numeric(R), as.numeric(iter0), as.numeric(seed), d[[1]], d[[2]],
             d[[3]], as.numeric(length(x)))
  mean <- m[[3]] / R
  sd <- sqrt(pmax(m[[4]] - R * mean^2, 0) / max(R - 1, 1))
  list(iter_nranges = m[[1]], iter_totals = m[[2]], mean = mean, sd = sd, observed = observed, z = (observed - mean) / sd,
       p_ge = (m[[5]] + 1) / (R + 1))
}
