// Iteration-invariant block tables of one bootRanges() sampler, built on the host once per
// call.  The reference rebuilds the same information inside every replicate() iteration
// (tileGenome at R/bootUnsegmented.R:129, successiveIRanges at :63, seq(by = L_bs) at
// R/bootRanges.R:213-215, :269-271).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/nrb200.h"

namespace nrb {

struct Plan {
    int kind = -1;
    int64_t block_length = 0;
    int n_seq = 0;
    int n_states = 0;
    int64_t genome_length = 0;  // L_g = sum(seqlengths(y))
    std::vector<int64_t> seqlen;
    std::vector<int64_t> tiles_per_seq;  // ceiling(L_c / L_b)
    std::vector<int64_t> seq_cum;        // [n_seq + 1] cumulative seqlengths

    // Block b is moved TO tile b (bootstrap kinds).  bait_width[b] is the width of the block
    // sampled for it (L_b, or the per-state L_bs of R/bootRanges.R:189).
    std::vector<int32_t> tile_seq, tile_start, bait_width;

    // Segmented kinds: tables the device sampler needs, grouped by state (1-based).
    std::vector<int64_t> state_block_off;  // [S + 2] first block of each state
    std::vector<int64_t> state_seg_off;    // [S + 2] first entry of each state in sg_*
    std::vector<int64_t> state_len;        // [S + 2] L_s: total width of the state
    std::vector<int64_t> state_block_len;  // [S + 2] L_bs
    std::vector<int64_t> sg_cum;           // width of the state's segments BEFORE this one
    std::vector<int32_t> sg_seq, sg_start, sg_tiles;

    bool permute() const { return kind == NRB200_PERMUTE_WITHIN || kind == NRB200_PERMUTE_ACROSS; }
    bool segmented() const { return kind == NRB200_SEG_PROP || kind == NRB200_SEG_NOPROP; }
    // Within-chromosome variants keep zero-width survivors: there is no width filter after
    // trim() at R/bootUnsegmented.R:73 and :107, unlike shift_and_swap_chrom (:189).
    bool drop_zero_width() const { return !(kind == NRB200_UNSEG_WITHIN || kind == NRB200_PERMUTE_WITHIN); }
    int64_t n_blocks() const { return (int64_t)tile_seq.size(); }
};

// Returns an empty string on success, else the message of the reference check that failed.
std::string build_plan(const nrb200_plan_spec &spec, const std::vector<int64_t> &seqlengths, Plan &out);

}  // namespace nrb
