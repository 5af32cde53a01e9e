// Device side of libnrb200, in parts:
//   device_views.cuh    views of the engine's buffers, searches, direct-address rank tables, Philox block draws
//   rank_kernels.cuh    RANK PATH (default): per (iteration, block) rows kernel + per (feature, iteration) count /
//                       weighted-sum kernels that count by rank differences and never stream the sampled ranges
//   stream_kernels.cuh  STREAMING PATH: one warp per (iteration, block) streams the hits, shift / trim / drop /
//                       exclude per range (fused count kernel; scan + fill kernels of the materialised table)
//   index_kernels.cuh   index construction, device-side permutations and proportionLength = FALSE draws,
//                       observed counts, moments
//
// Reference code this replaces, per iteration: R/bootUnsegmented.R:116-143 / :50-75 / :83-109 / :150-165,
// R/bootRanges.R:141-331 (samplers + findOverlaps + gather), R/bootUnsegmented.R:175-191
// (shift_and_swap_chrom), R/bootRanges.R:106-116 (exclude drop / trim) and the user-side
// countOverlaps of vignettes/bootRanges.Rmd:497-503, 532-540.
#pragma once
#include "device_views.cuh"
#include "stream_kernels.cuh"
#include "rank_kernels.cuh"
#include "index_kernels.cuh"
