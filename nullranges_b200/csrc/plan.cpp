#include "plan.h"

#include <cmath>

namespace nrb {

static int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

std::string build_plan(const nrb200_plan_spec &spec, const std::vector<int64_t> &seqlengths, Plan &out) {
    Plan p;
    p.kind = spec.kind;
    p.block_length = spec.block_length;
    p.n_seq = (int)seqlengths.size();
    p.seqlen = seqlengths;
    if (p.n_seq == 0) return "seqlengths(y) is empty";
    if (spec.block_length <= 0) return "blockLength must be a positive integer";
    if (spec.block_length > (int64_t)1 << 30) return "blockLength above 2^30 is not supported";
    if (spec.kind < NRB200_UNSEG_ACROSS || spec.kind > NRB200_PERMUTE_ACROSS) return "unknown sampler kind";

    p.seq_cum.assign(p.n_seq + 1, 0);
    p.tiles_per_seq.resize(p.n_seq);
    for (int c = 0; c < p.n_seq; ++c) {
        p.seq_cum[c + 1] = p.seq_cum[c] + seqlengths[c];
        p.tiles_per_seq[c] = ceil_div(seqlengths[c], spec.block_length);
    }
    p.genome_length = p.seq_cum[p.n_seq];
    // the block tables are materialised: refuse absurd block counts before allocating them
    constexpr int64_t kMaxBlocks = (int64_t)1 << 28;
    auto too_many = [&](int64_t n) { return n > kMaxBlocks; };

    if (!p.segmented()) {
        // stopifnot(all(chrom_lens >= L_b))                       R/bootUnsegmented.R:8
        for (int64_t len : seqlengths)
            if (len < spec.block_length) return "all(chrom_lens >= L_b) is not TRUE";
        int64_t total_tiles = 0;
        for (int64_t t : p.tiles_per_seq) total_tiles += t;
        if (too_many(total_tiles)) return "too many blocks per iteration (blockLength too small for this genome)";
        // Tile k of sequence c starts at 1 + k * L_b, sequences in seqlevel order.
        for (int c = 0; c < p.n_seq; ++c)
            for (int64_t k = 0; k < p.tiles_per_seq[c]; ++k) {
                p.tile_seq.push_back(c);
                p.tile_start.push_back((int32_t)(1 + k * spec.block_length));
                p.bait_width.push_back((int32_t)spec.block_length);
            }
        out = std::move(p);
        return "";
    }

    // ---- segmented ----
    if (spec.n_seg <= 0 || !spec.seg_seqid || !spec.seg_start || !spec.seg_end || !spec.seg_state)
        return "seg must be a non-empty GRanges with a 'state' column";
    int S = 0;
    for (int64_t j = 0; j < spec.n_seg; ++j) {
        if (spec.seg_state[j] < 1) return "seg$state must be integers >= 1";
        if (spec.seg_seqid[j] < 0 || spec.seg_seqid[j] >= p.n_seq) return "seqnames(seg) must be among seqlevels(y)";
        if (spec.seg_start[j] < 1 || spec.seg_end[j] < spec.seg_start[j]) return "segments must have start >= 1 and positive width";
        S = std::max(S, spec.seg_state[j]);
    }
    p.n_states = S;  // esses <- seq_len(max(state))                R/bootRanges.R:146
    p.state_block_off.assign(S + 2, 0);
    p.state_seg_off.assign(S + 2, 0);
    p.state_len.assign(S + 2, 0);
    p.state_block_len.assign(S + 2, 0);

    for (int s = 1; s <= S; ++s) {
        p.state_seg_off[s] = (int64_t)p.sg_seq.size();
        p.state_block_off[s] = p.n_blocks();
        // seg_s <- seg[state == s]: segments of the state in the order given
        std::vector<int64_t> members;
        int64_t total = 0;
        for (int64_t j = 0; j < spec.n_seg; ++j)
            if (spec.seg_state[j] == s) {
                members.push_back(j);
                total += (int64_t)spec.seg_end[j] - spec.seg_start[j] + 1;
            }
        if (members.empty()) return "every state in 1..max(seg$state) needs at least one segment";
        p.state_len[s] = total;
        int64_t Lbs = spec.block_length;
        if (spec.kind == NRB200_SEG_PROP) {
            // L_b_per_state <- round(L_b * L_s / L_g)               R/bootRanges.R:189
            Lbs = (int64_t)std::nearbyint((double)spec.block_length * (double)total / (double)p.genome_length);
            if (Lbs <= 0) return "a state's scaled block length round(L_b * L_s / L_g) is 0";
        }
        p.state_block_len[s] = Lbs;
        if (too_many(p.n_blocks() + ceil_div(total, Lbs))) return "too many blocks per iteration (blockLength too small for this genome)";
        int64_t before = 0;
        for (int64_t j : members) {
            const int64_t width = (int64_t)spec.seg_end[j] - spec.seg_start[j] + 1;
            const int64_t tiles = ceil_div(width, Lbs);  // ceiling(width / L_bs)  :198, :259
            p.sg_cum.push_back(before);
            p.sg_seq.push_back(spec.seg_seqid[j]);
            p.sg_start.push_back(spec.seg_start[j]);
            p.sg_tiles.push_back((int32_t)tiles);
            before += width;
            for (int64_t k = 0; k < tiles; ++k) {  // seq(from = start, to = end, by = L_bs)
                p.tile_seq.push_back(spec.seg_seqid[j]);
                p.tile_start.push_back((int32_t)(spec.seg_start[j] + k * Lbs));
                p.bait_width.push_back((int32_t)Lbs);
            }
        }
    }
    p.state_seg_off[S + 1] = (int64_t)p.sg_seq.size();
    p.state_block_off[S + 1] = p.n_blocks();
    out = std::move(p);
    return "";
}

}  // namespace nrb
