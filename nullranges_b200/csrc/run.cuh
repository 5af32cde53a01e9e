// libnrb200: dispatch, block draws, kernel launches and the core of the counting entry points.
// (textually included by engine.cu)
namespace {

// The rank path covers every sampler (bootstrap and permute, with or without exclusion, any strands);
// only a segmentation reaching beyond its sequence falls back to streaming the hits (boot_stream_kernel; boot_count_kernel for permute).
bool rank_path_ok(const nrb200_handle *h) {
    if (h->flags & NRB200_FLAG_FORCE_STREAM) return false;
    if (kZeroWidthStrictlyInside && h->have_plan && !h->plan.drop_zero_width()) {
        // zero-width rule 1 (device_views.cuh): the window lists encode rule 0; the rules differ only when an exclude or
        // query range reaches past the end of its sequence -- those inputs take the streaming kernels
        for (const RangeSet *rs : {&h->query, &h->excl})
            for (int c = 0; rs->n > 0 && c < h->n_seq; ++c)
                if ((int64_t)rs->max_end[c] > h->seqlen[c]) return false;
    }
    return h->tiles_in_bounds;
}

// Everything a run needs that hangs on the plan, the query, the exclude set and -- of y -- only its widest range and
// its classes: the window lists of the rank path, the per-tile slices of the streaming kernels.
int prepare_lists(nrb200_handle *h) {
    int rc;
    NvtxRange nvtx("nrb200: window lists");
    Trace tr("prepare");
    h->tiles_in_bounds = true;
    for (int64_t t = 0; t < h->plan.n_blocks(); ++t)
        if (h->plan.tile_start[t] > h->seqlen[h->plan.tile_seq[t]]) h->tiles_in_bounds = false;
    if (trim_mode(h) && h->excl.stranded) return fail(h, NRB200_ERR_INVALID, "all(strand(exclude) == \"*\") is not TRUE");
    if (trim_mode(h) && (rc = build_gaps(h))) return rc;
    const bool rank = rank_path_ok(h);
    if (rank) {
        if ((rc = build_pairs(h))) return rc;  // window lists of the features and of the tiles
        tr.lap("build_pairs");
    }
    // the streaming kernels (count without the rank path, and the fill pass) test exclusion per range
    if (h->excl.n && (rc = build_slices(h, active_excl(h), h->x_tile_lo, h->x_tile_hi))) return rc;
    if (!rank) {
        if (h->query.n && (rc = range_set_to_device(h, h->query))) return rc;
        if (h->query.n && (rc = build_slices(h, h->query, h->q_tile_lo, h->q_tile_hi))) return rc;
        if ((rc = build_tile_groups(h))) return rc;
    }
    return NRB200_OK;
}

int prepare(nrb200_handle *h) {
    if (!h->have_y) return fail(h, NRB200_ERR_STATE, h->y_pending ? "nrb200_set_ranges_end() has not been called" : "nrb200_set_ranges() has not been called");
    if (!h->have_plan) return fail(h, NRB200_ERR_STATE, "nrb200_set_plan() has not been called");
    if (h->slices_dirty) {
        int rc;
        if (h->index_minoverlap != h->query_minoverlap && (rc = build_rank_tables(h))) return rc;  // the width classes changed
        if ((rc = prepare_lists(h))) return rc;
        h->slices_dirty = false;
    }
    if (!rank_path_ok(h)) {
        if (int rc = ensure_interleaved(h)) return rc;
        return ensure_pmax_table(h);
    }
    return NRB200_OK;
}

int stage_draws(nrb200_handle *h, const nrb200_draws *d) {
    if (!d || d->n_iter < 0) return fail(h, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
    const Plan &p = h->plan;
    const int64_t B = p.n_blocks();
    if (d->rng_mode == NRB200_RNG_PHILOX) {
        if (p.kind == NRB200_SEG_NOPROP) {
            // proportionLength = FALSE: the draws are made on the device into the tables host mode would upload
            const int64_t R = d->n_iter, items = R * B;
            const int S = p.n_states;
            if (S > 127) return fail(h, NRB200_ERR_UNSUPPORTED, "device draws for proportionLength=FALSE support at most 127 states");
            for (DevMem *m : {&h->d_bait_seq, &h->d_bait_start, &h->np_seg, &h->np_start, &h->np_list})
                NRB_CUDA(h, m->reserve((size_t)std::max<int64_t>(items, 1) * 4));
            NRB_CUDA(h, h->np_state.reserve((size_t)std::max<int64_t>(items, 1)));
            NRB_CUDA(h, h->np_cnt.reserve((size_t)std::max<int64_t>(R, 1) * (S + 1) * 4 + 4));
            if (items == 0) return NRB200_OK;
            int *err = h->np_cnt.as<int>() + R * (S + 1);
            NRB_CUDA(h, cudaMemsetAsync(err, 0, 4, h->stream));
            const unsigned grid = (unsigned)((items + 255) / 256);
            noprop_draw_kernel<<<grid, 256, 0, h->stream>>>(h->sampler_view, d->seed, d->iter0, R, B, h->np_seg.as<int32_t>(),
                                                             h->np_start.as<int32_t>(), h->np_state.as<int8_t>());
            noprop_lists_kernel<<<(unsigned)R, 256, 0, h->stream>>>(S, B, h->np_state.as<int8_t>(), h->np_list.as<int32_t>(),
                                                                    h->np_cnt.as<int32_t>());
            noprop_assign_kernel<<<grid, 256, 0, h->stream>>>(h->sampler_view, d->seed, d->iter0, R, B, h->np_seg.as<int32_t>(),
                                                               h->np_start.as<int32_t>(), h->np_list.as<int32_t>(), h->np_cnt.as<int32_t>(),
                                                               h->d_bait_seq.as<int32_t>(), h->d_bait_start.as<int32_t>(), err);
            NRB_CUDA(h, cudaGetLastError());
            int flag = 0;
            NRB_CUDA(h, cudaMemcpyAsync(&flag, err, 4, cudaMemcpyDeviceToHost, h->stream));
            if (int rc_sync = sync_stream(h)) return rc_sync;
            // the reference stops here too: length(rearranged_blocks_start) == length(random_blocks_start) fails (R/bootRanges.R:303)
            if (flag) return fail(h, NRB200_ERR_INVALID, "seg_bootstrap_noprop: a state drew no segments in some iteration");
            return NRB200_OK;
        }
        if (!p.permute()) return NRB200_OK;
        // permute variants: one random permutation of the tiles per iteration, made on the device
        if (h->n_seq > 4096) return fail(h, NRB200_ERR_UNSUPPORTED, "device permutations support at most 4096 sequences");
        const int64_t R = d->n_iter;
        NRB_CUDA(h, h->d_dst_tile.reserve((size_t)std::max<int64_t>(R * B, 1) * 4));
        h->dst_tile_items = R * B;
        const int64_t kPermBatch = 4096;  // 12 bits of the sort key
        const int64_t cap = std::min(R, kPermBatch) * B;
        NRB_CUDA(h, h->sort_keys_a.reserve((size_t)std::max<int64_t>(cap, 1) * 8));
        NRB_CUDA(h, h->sort_keys_b.reserve((size_t)std::max<int64_t>(cap, 1) * 8));
        NRB_CUDA(h, h->perm_vals_a.reserve((size_t)std::max<int64_t>(cap, 1) * 4));
        NRB_CUDA(h, h->perm_vals_b.reserve((size_t)std::max<int64_t>(cap, 1) * 4));
        for (int64_t r0 = 0; r0 < R; r0 += kPermBatch) {
            const int64_t n = std::min(kPermBatch, R - r0), items = n * B;
            if (items == 0) continue;
            const unsigned grid = (unsigned)((items + 255) / 256);
            permute_keys_kernel<<<grid, 256, 0, h->stream>>>(d->seed, d->iter0 + r0, n, B, h->tile_seq.as<int32_t>(),
                                                              p.kind == NRB200_PERMUTE_WITHIN, h->sort_keys_a.as<unsigned long long>(),
                                                              h->perm_vals_a.as<int32_t>());
            size_t tmp_bytes = 0;
            NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->sort_keys_a.as<unsigned long long>(),
                                                        h->sort_keys_b.as<unsigned long long>(), h->perm_vals_a.as<int32_t>(),
                                                        h->perm_vals_b.as<int32_t>(), items, 0, 64, h->stream));
            NRB_CUDA(h, h->sort_tmp.reserve(tmp_bytes));
            NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp_bytes, h->sort_keys_a.as<unsigned long long>(),
                                                        h->sort_keys_b.as<unsigned long long>(), h->perm_vals_a.as<int32_t>(),
                                                        h->perm_vals_b.as<int32_t>(), items, 0, 64, h->stream));
            permute_scatter_kernel<<<grid, 256, 0, h->stream>>>(h->perm_vals_b.as<int32_t>(), n, B, h->d_dst_tile.as<int32_t>() + r0 * B);
            NRB_CUDA(h, cudaGetLastError());
        }
        return NRB200_OK;
    }
    if (d->rng_mode != NRB200_RNG_HOST) return fail(h, NRB200_ERR_INVALID, "draws: unknown rng_mode");
    const int64_t n = d->n_iter * B;
    if (n > 0 && (!d->bait_seqid || !d->bait_start)) return fail(h, NRB200_ERR_INVALID, "draws: bait_seqid/bait_start are required in host mode");
    if (p.permute() && n > 0 && !d->dst_tile) return fail(h, NRB200_ERR_INVALID, "draws: permute kinds need dst_tile");
    for (int64_t i = 0; i < n; ++i) {
        if (d->bait_seqid[i] < 0 || d->bait_seqid[i] >= h->n_seq) return fail(h, NRB200_ERR_INVALID, "draws: bait_seqid out of range");
        if (d->bait_start[i] < 1 || d->bait_start[i] >= (1 << 30)) return fail(h, NRB200_ERR_INVALID, "draws: bait_start out of range");
        if (d->dst_tile && (d->dst_tile[i] < 0 || d->dst_tile[i] >= B)) return fail(h, NRB200_ERR_INVALID, "draws: dst_tile out of range");
    }
    if (p.permute() && n > 0) {  // sample(n) (R/bootUnsegmented.R:97, :156): every tile is the destination of exactly one block
        std::vector<uint8_t> seen((size_t)B);
        for (int64_t r = 0; r < d->n_iter; ++r) {
            std::fill(seen.begin(), seen.end(), 0);
            for (int64_t b = 0; b < B; ++b) {
                uint8_t &s = seen[d->dst_tile[r * B + b]];
                if (s) return fail(h, NRB200_ERR_INVALID, "draws: dst_tile must be a permutation of the tiles in every iteration");
                s = 1;
            }
        }
    }
    int rc;
    if ((rc = upload(h, h->d_bait_seq, d->bait_seqid, (size_t)n))) return rc;
    if ((rc = upload(h, h->d_bait_start, d->bait_start, (size_t)n))) return rc;
    if (d->dst_tile && (rc = upload(h, h->d_dst_tile, d->dst_tile, (size_t)n))) return rc;
    h->dst_tile_items = d->dst_tile ? n : 0;
    return NRB200_OK;
}

BootParams make_params(nrb200_handle *h, const nrb200_draws *d) {
    BootParams P{};
    const RangeSet &y = h->y;
    RangesView Y{};
    Y.start = y.start.as<int32_t>();
    Y.end = y.end.as<int32_t>();
    Y.pmax = y.pmax.as<int32_t>();
    Y.strand = y.stranded ? y.strand.as<int8_t>() : nullptr;
    Y.src = y.sorted_input ? nullptr : y.src.as<int32_t>();
    Y.seq_off = y.seq_off.as<int32_t>();
    Y.seq_info = h->seq_info_dev.as<int4>();
    Y.rank_start = h->rank_start.as<RankEntry>();
    Y.rank_pmax = h->rank_pmax.as<RankEntry>();
    Y.rank_shift = h->rank_shift;
    Y.n_cls = h->n_cls;
    Y.end_sorted = h->end_sorted.as<int32_t>();
    Y.rank_end = h->rank_end.as<RankEntry>();
    if (h->n_cls == 1) {  // unstranded: the rank path reads the canonical arrays
        Y.rstart = Y.start;
        Y.rend = Y.end;
        Y.rpmax = Y.pmax;
        Y.rseq_info = Y.seq_info;
        Y.rrank_start = Y.rank_start;
    } else {
        Y.rstart = h->r_start.as<int32_t>();
        Y.rend = h->r_end.as<int32_t>();
        Y.rpmax = h->r_pmax.as<int32_t>();
        Y.rseq_info = h->rseq_info_dev.as<int4>();
        Y.rrank_start = h->rrank_start.as<RankEntry>();
    }
    Y.end_shift = h->end_shift;
    if (h->ends_alias_starts) {
        Y.end_sorted = Y.rstart;
        Y.rank_end = Y.rrank_start;
    }
    P.y = Y;
    const RangeSet &q = h->query, &x = active_excl(h);
    P.query = SliceView{q.start.as<int32_t>(), q.end.as<int32_t>(), q.pmax.as<int32_t>(), q.stranded ? q.strand.as<int8_t>() : nullptr,
                        q.src.as<int32_t>(), h->q_tile_lo.as<int32_t>(), h->q_tile_hi.as<int32_t>()};
    P.excl = SliceView{x.start.as<int32_t>(), x.end.as<int32_t>(), x.pmax.as<int32_t>(), x.stranded ? x.strand.as<int8_t>() : nullptr,
                       x.src.as<int32_t>(), h->x_tile_lo.as<int32_t>(), h->x_tile_hi.as<int32_t>()};
    P.rows_windows_only = trim_mode(h);
    P.minoverlap = h->query_minoverlap;
    P.sampler = h->sampler_view;
    P.n_blocks = h->plan.n_blocks();
    P.tile_seq = h->tile_seq.as<int32_t>();
    P.tile_start = h->tile_start.as<int32_t>();
    P.bait_width = h->bait_width.as<int32_t>();
    P.tile_info = h->tile_info.as<int2>();
    P.seqlen = h->seqlen_dev.as<int64_t>();
    P.permute = h->plan.permute();
    P.drop_zero = h->plan.drop_zero_width();
    // proportionLength = FALSE in production mode: stage_draws() filled the host-mode tables on the device
    const bool noprop_dev = d->rng_mode == NRB200_RNG_PHILOX && h->plan.kind == NRB200_SEG_NOPROP;
    P.rng_mode = noprop_dev ? NRB200_RNG_HOST : d->rng_mode;
    P.seed = d->seed;
    P.iter0 = d->iter0;
    if (d->rng_mode == NRB200_RNG_HOST || noprop_dev) {
        P.bait_seq = h->d_bait_seq.as<int32_t>();
        P.bait_start = h->d_bait_start.as<int32_t>();
        P.dst_tile = (!noprop_dev && d->dst_tile) ? h->d_dst_tile.as<int32_t>() : nullptr;
    } else if (h->plan.permute()) {
        P.dst_tile = h->d_dst_tile.as<int32_t>();  // made by stage_draws() on the device
    }
    P.n_iter = d->n_iter;
    P.n_query = q.n;
    return P;
}

template <bool ST, bool HQ>
void launch_stream_t(nrb200_handle *h, const BootParams &P, int excl) {
    const GroupView V{h->groups_dev.as<TileGroup>(), h->tiles_sorted_dev.as<int32_t>(), h->y_se.as<int2>(), h->group_slice_cap};
    // Few iterations (or few groups) would leave most SMs without a CTA: a group is then worked by `split` CTAs, CTA z taking
    // the group's tiles z, z + split, ... (each stages the group's query slice for itself and flushes its own histogram).
    const int64_t ctas = (int64_t)h->n_groups * P.n_iter, want = (int64_t)h->sm_count * kStreamMinBlocks * 2;
    const unsigned split = (unsigned)std::max<int64_t>(1, std::min<int64_t>(kGroupTiles / 32, (want + ctas - 1) / std::max<int64_t>(ctas, 1)));
    const dim3 grid((unsigned)h->n_groups, (unsigned)P.n_iter, split);
    const size_t smem = ((size_t)h->group_slice_cap * 21 + 15) & ~(size_t)15;
    // dynamic slice (up to 43 KB) + the static item table (14 KB) can pass the 48 KB a kernel gets without asking
    auto launch = [&](auto kernel) {
        if (smem > ((size_t)32 << 10)) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<grid, 256, smem, h->stream>>>(P, V);
    };
    if (excl == kExclNone) launch(boot_stream_kernel<ST, HQ, kExclNone>);
    else if (excl == kExclDrop) launch(boot_stream_kernel<ST, HQ, kExclDrop>);
    else launch(boot_stream_kernel<ST, HQ, kExclTrim>);
}

template <bool ST, bool HQ>
void launch_count_t(nrb200_handle *h, const BootParams &P, unsigned grid, int excl) {
    if (excl == kExclNone) boot_count_kernel<ST, HQ, kExclNone><<<grid, 256, 0, h->stream>>>(P);
    else if (excl == kExclDrop) boot_count_kernel<ST, HQ, kExclDrop><<<grid, 256, 0, h->stream>>>(P);
    else boot_count_kernel<ST, HQ, kExclTrim><<<grid, 256, 0, h->stream>>>(P);
}

constexpr int64_t kMaxBatch = 2048;  // iterations per launch at most (gridDim.y and scratch stay small); NRB200_MAX_BATCH lowers it
constexpr int kChunkEvents = 64;
constexpr size_t kMaxStagingBytes = (size_t)2 << 30;  // pinned staging for pageable result buffers, at most

// Returns the number of kernels launched.
int launch_rank(nrb200_handle *h, const BootParams &P, bool has_query) {
    if (P.n_iter == 0) return 0;
    int launches = 0;
    if (P.n_blocks > 0) {
        const TileExcl X{h->excl.n > 0 ? h->tile_x_off.as<int32_t>() : nullptr, h->tile_x_rec.as<int4>()};  // drop: minus, trim: plus
        // iterations a CTA loops over: up to NRB200_ROWS_IPC, as long as the grid still holds a few CTAs per resident slot
        const int64_t ctas_x = (P.n_blocks + kRowsThreads - 1) / kRowsThreads;
        const int ipc = (int)std::max<int64_t>(1, std::min<int64_t>(h->tune_rows_ipc, P.n_iter * ctas_x / ((int64_t)h->sm_count * 12)));
        const dim3 rows_grid((unsigned)ctas_x, (unsigned)((P.n_iter + ipc - 1) / ipc));
        const bool simple = P.y.n_cls == 1 && !X.off && !P.item_rec && !P.item_rows && !P.permute && !P.dst_tile && !P.rows_windows_only;
        if (simple) boot_block_rows_kernel<true><<<rows_grid, kRowsThreads, 0, h->stream>>>(P, X, ipc);
        else boot_block_rows_kernel<false><<<rows_grid, kRowsThreads, 0, h->stream>>>(P, X, ipc);
        ++launches;
    }
    if (h->split_ev) cudaEventRecord(h->split_ev, h->stream);  // stats: end of the rows kernel
    if (has_query) {
        const FeatView F{h->f_src.as<int32_t>(), h->pair_off.as<int32_t>(), h->pair_rec.as<int4>()};
        // a thread stays on its feature for up to NRB200_IPT iterations, as long as the grid still holds a few CTAs per
        // resident slot (small runs: one iteration per thread)
        const unsigned gx = (unsigned)((P.n_query + kCountThreads - 1) / kCountThreads);
        const int ipt = (int)std::max<int64_t>(1, std::min<int64_t>(h->tune_ipt, (int64_t)gx * P.n_iter / ((int64_t)h->sm_count * 16 * 3)));
        const dim3 lgrid(gx, (unsigned)((P.n_iter + ipt - 1) / ipt));
        if (h->tune_loop_minb >= 16) boot_rank_loop_kernel<false, 16><<<lgrid, kCountThreads, 0, h->stream>>>(P, F, ipt, MomentsOut{});
        else boot_rank_loop_kernel<false, 12><<<lgrid, kCountThreads, 0, h->stream>>>(P, F, ipt, MomentsOut{});
        ++launches;
    }
    return launches;
}

int launch_count(nrb200_handle *h, const BootParams &P, bool has_query) {
    if (rank_path_ok(h)) return launch_rank(h, P, has_query);
    const int64_t items = P.n_iter * P.n_blocks;
    if (items == 0) return 0;
    const int64_t ctas = (items + 7) / 8;
    const unsigned grid = (unsigned)std::min<int64_t>(ctas, (int64_t)h->sm_count * 64);
    const bool st = h->y.stranded && ((has_query && h->query.stranded) || (h->excl.n && active_excl(h).stranded));
    const int excl = excl_kind(h);
    if (h->groups_ok && !P.permute && !P.dst_tile && !(h->tune_stream_old)) {  // tile-group form (bootstrap kinds)
        if (st && has_query) launch_stream_t<true, true>(h, P, excl);
        else if (st) launch_stream_t<true, false>(h, P, excl);
        else if (has_query) launch_stream_t<false, true>(h, P, excl);
        else launch_stream_t<false, false>(h, P, excl);
        return 1;
    }
    if (st && has_query) launch_count_t<true, true>(h, P, grid, excl);
    else if (st) launch_count_t<true, false>(h, P, grid, excl);
    else if (has_query) launch_count_t<false, true>(h, P, grid, excl);
    else launch_count_t<false, false>(h, P, grid, excl);
    return 1;
}

// P restricted to local iterations [r0, r0 + n).
BootParams slice_params(const BootParams &P, int64_t r0, int64_t n) {
    BootParams Q = P;
    Q.n_iter = n;
    Q.iter0 = P.iter0 + r0;
    if (P.bait_seq) Q.bait_seq = P.bait_seq + r0 * P.n_blocks;
    if (P.bait_start) Q.bait_start = P.bait_start + r0 * P.n_blocks;
    if (P.dst_tile) Q.dst_tile = P.dst_tile + r0 * P.n_blocks;
    if (P.iter_nranges) Q.iter_nranges = P.iter_nranges + r0;
    if (P.iter_totals) Q.iter_totals = P.iter_totals + r0;
    if (P.counts) Q.counts = P.counts + r0 * P.n_query;
    if (P.counts16) Q.counts16 = P.counts16 + r0 * P.n_query;
    for (int w = 0; w < P.n_peers; ++w) Q.peer_counts[w] = P.peer_counts[w] + r0 * P.n_query;
    if (P.item_rows) Q.item_rows = P.item_rows + r0 * P.n_blocks;
    if (P.item_rec) Q.item_rec = P.item_rec + r0 * P.n_blocks;
    return Q;
}

// uint16 -> int32 with streaming stores, AVX-512 / AVX2 / SSE2 chosen at run time: host_simd.cpp
// Core of every counting entry point: results stay on the device.  With host_counts set, the
// iterations run in chunks and each chunk of the [R x G] matrix starts its way to the host
// (copy stream) while the next chunk computes.
//
// Wire format of a host-bound matrix (rank path): uint16.  Counts of overlaps per (feature, iteration) are small
// (cfg 2: mean 15, max a few hundred), so the kernel stores them as uint16 -- half the bytes through L2, HBM and
// PCIe -- into a pinned staging buffer, and the host widens every chunk into the caller's int32 matrix (R `integer`)
// with all its threads while later chunks still compute / travel.  A count above 65535 raises the chunk's overflow
// flag; such a chunk is computed again with int32 output.  NRB200_WIRE16=0 turns the format off.
int run_counts_core(nrb200_handle *h, const nrb200_draws *d, bool want_counts, int32_t *host_counts,
                    nrb200_stats *stats) {
    int rc;
    if ((rc = prepare(h))) return rc;
    NvtxRange nvtx("nrb200: run (chunks)");
    h->timeline.mark("lists ready", h->stream);
    NRB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if ((rc = stage_draws(h, d))) return rc;
    const int64_t R = d->n_iter, G = h->query.n;
    const bool has_query = G > 0;
    want_counts = want_counts && has_query;
    NRB_CUDA(h, h->d_iter_nranges.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->d_iter_totals.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_nranges.p, 0, R * 8, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_totals.p, 0, R * 8, h->stream));
    if ((rc = hits_reset(h))) return rc;
    uint16_t *host16 = h->out16;  // nrb200_run_counts_u16: the caller takes the matrix as uint16 (saturating), no widening
    h->out16 = nullptr;
    h->out16_overflow = 0;
    if (host16 && !rank_path_ok(h)) return fail(h, NRB200_ERR_UNSUPPORTED, "uint16 counts need the rank path (segmentation within seqlengths, no NRB200_FLAG_FORCE_STREAM)");
    const bool pipelined = (host_counts || host16) && want_counts && R > 0;
    bool pageable = false;
    if (pipelined) {
        cudaPointerAttributes attr{};
        pageable = cudaPointerGetAttributes(&attr, host16 ? (const void *)host16 : (const void *)host_counts) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();  // an unregistered pointer may leave a sticky-free error behind on older drivers
    }
    const size_t cells = (size_t)R * (size_t)std::max<int64_t>(G, 0);
    // tune_wire16: 0 never, 1 always, 2 (default) where it pays.  A pageable destination always does (the matrix has to be
    // moved by the host anyway).  A pinned one could take the int32 matrix by DMA without any host work, but half the bytes
    // on the wire and a widening pass by the host threads is faster as long as the host has the bandwidth to spare: 1.65
    // against 2.14 ms per 1000 iterations on one GPU, 1.96 against 2.25 ms per 2000 on two; from four GPUs on the host's
    // memory is the bottleneck (device groups) and plain DMA wins (8 GPUs: 7.8 against 10 ms per 8000).
    bool wire16 = pipelined && rank_path_ok(h) && cells * 2 <= kMaxStagingBytes &&
                  (h->tune_wire16 == 1 || (h->tune_wire16 == 2 && (pageable || h->group_w <= 2)));
    if (host16) wire16 = pipelined;
    const bool stage16 = wire16 && !(host16 && !pageable);  // a pinned uint16 destination takes the chunks by DMA as they are
    if (host16 && pageable && cells * 2 > kMaxStagingBytes) return fail(h, NRB200_ERR_UNSUPPORTED, "uint16 counts into pageable memory: the matrix exceeds the staging limit; use a pinned buffer or fewer iterations per call");
    if (stage16 && h->stage_cap < cells * 2) {
        if (h->stage_buf) cudaFreeHost(h->stage_buf);
        h->stage_buf = nullptr;
        h->stage_cap = 0;
        if (cudaMallocHost((void **)&h->stage_buf, cells * 2) == cudaSuccess) h->stage_cap = cells * 2;
        else {
            cudaGetLastError();
            if (host16) return fail(h, NRB200_ERR_NOMEM, "no pinned staging memory for the uint16 matrix");
            wire16 = false;
        }
    }
    int32_t *counts = nullptr;
    uint16_t *counts16 = nullptr;
    int *flags_dev = nullptr;
    if (wire16) {
        NRB_CUDA(h, h->d_counts16.reserve(((cells * 2 + 15) & ~(size_t)15) + kChunkEvents * sizeof(int)));  // matrix, then the flags
        counts16 = h->d_counts16.as<uint16_t>();
        flags_dev = reinterpret_cast<int *>(h->d_counts16.as<char>() + ((cells * 2 + 15) & ~(size_t)15));
        NRB_CUDA(h, cudaMemsetAsync(flags_dev, 0, kChunkEvents * sizeof(int), h->stream));
        if (!h->flags_host) NRB_CUDA(h, cudaMallocHost((void **)&h->flags_host, kChunkEvents * sizeof(int)));
        std::memset(h->flags_host, 0, kChunkEvents * sizeof(int));
    } else if (want_counts && !host_counts && h->peer_out.n > 0) {
        if (!rank_path_ok(h)) return fail(h, NRB200_ERR_UNSUPPORTED, "the gathered result of a device group needs the rank path");
        // gathered run of a device group: the kernel stores into every member's matrix, nothing local
    } else if (want_counts) {
        NRB_CUDA(h, h->d_counts.reserve(cells * sizeof(int32_t)));
        counts = h->d_counts.as<int32_t>();
        // the rank path writes every element; the streaming path accumulates with atomics
        if (!rank_path_ok(h)) NRB_CUDA(h, cudaMemsetAsync(counts, 0, cells * sizeof(int32_t), h->stream));
    }
    BootParams P = make_params(h, d);
    P.iter_nranges = h->d_iter_nranges.as<unsigned long long>();
    P.iter_totals = h->d_iter_totals.as<unsigned long long>();
    P.counts = counts;
    P.counts16 = counts16;
    if (want_counts && !host_counts && !counts && !counts16) {
        P.n_peers = h->peer_out.n;
        for (int w = 0; w < P.n_peers; ++w) P.peer_counts[w] = h->peer_out.counts[w];
    }
    P.n_hits = stats ? h->d_n_hits.as<unsigned long long>() : nullptr;  // the H statistic is tallied only when it is asked for
    NRB_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    int launches = 0;
    // Iterations go in batches: bounded draw-table scratch on the rank path and, when the matrix
    // is wanted on the host, chunk k copies out (copy stream) while chunk k + 1 computes.
    int64_t chunk = h->tune_max_batch;
    if (pipelined) {
        if (!h->copy_stream) NRB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        chunk = std::min<int64_t>(h->tune_max_batch, std::max<int64_t>(((R + h->tune_chunks - 1) / h->tune_chunks + 7) / 8 * 8, 64));
    }
    if (pipelined) chunk = std::max(chunk, (R + kChunkEvents - 1) / kChunkEvents);  // one event / flag slot per chunk in flight
    if (rank_path_ok(h) && has_query) {
        NRB_CUDA(h, h->draw_tab.reserve((size_t)std::min(chunk, std::max<int64_t>(R, 1)) * h->plan.n_blocks() * sizeof(int2)));
        P.draw_tab = h->draw_tab.as<int2>();
    }
    // A pageable int32 destination (R's own vectors are): cudaMemcpyAsync would stage it synchronously at a few GB/s.
    // Copy chunk by chunk into a pinned staging buffer of the handle instead and move each chunk on with a
    // parallel memcpy as soon as its copy event fires, while later chunks still compute / travel.
    int32_t *staging = nullptr;
    std::vector<std::pair<int64_t, int64_t>> staged;  // (first iteration, count) of every chunk, in order
    if (pipelined && !wire16) {
        const size_t bytes = cells * sizeof(int32_t);
        if (pageable) advise_huge_pages(host_counts, bytes);
        if (pageable && bytes <= kMaxStagingBytes) {
            if (h->stage_cap < bytes) {
                if (h->stage_buf) cudaFreeHost(h->stage_buf);
                h->stage_buf = nullptr;
                h->stage_cap = 0;
                if (cudaMallocHost((void **)&h->stage_buf, bytes) == cudaSuccess) h->stage_cap = bytes;
                else cudaGetLastError();
            }
            if (h->stage_cap >= bytes) staging = static_cast<int32_t *>(h->stage_buf);
        }
    }
    if (wire16 && pageable) advise_huge_pages(host16 ? (void *)host16 : (void *)host_counts, cells * (host16 ? 2 : 4));
    uint16_t *staging16 = stage16 ? static_cast<uint16_t *>(h->stage_buf) : (wire16 ? host16 : nullptr);
    int k = 0;
    const bool split = stats && rank_path_ok(h);
    for (int64_t r0 = 0; r0 < R; r0 += chunk, ++k) {
        const int64_t n = std::min(chunk, R - r0);
        if (split) {
            while (h->batch_ev.size() < (size_t)3 * (k + 1)) {
                cudaEvent_t e;
                NRB_CUDA(h, cudaEventCreate(&e));
                h->batch_ev.push_back(e);
            }
            NRB_CUDA(h, cudaEventRecord(h->batch_ev[3 * k], h->stream));
            h->split_ev = h->batch_ev[3 * k + 1];
        }
        BootParams Q = slice_params(P, r0, n);
        if (wire16) Q.overflow16 = flags_dev + (k % kChunkEvents);
        launches += launch_count(h, Q, has_query);
        h->split_ev = nullptr;
        if (split) NRB_CUDA(h, cudaEventRecord(h->batch_ev[3 * k + 2], h->stream));
        NRB_CUDA(h, cudaGetLastError());
        h->timeline.mark("chunk computed", h->stream);
        if (pipelined) {
            cudaEvent_t &ev = h->chunk_ev[k % kChunkEvents];
            if (!ev) NRB_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            NRB_CUDA(h, cudaEventRecord(ev, h->stream));
            NRB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, ev, 0));
            if (wire16) {
                NRB_CUDA(h, cudaMemcpyAsync(staging16 + r0 * G, counts16 + r0 * G, (size_t)n * G * sizeof(uint16_t), cudaMemcpyDeviceToHost,
                                            h->copy_stream));
                NRB_CUDA(h, cudaMemcpyAsync(h->flags_host + (k % kChunkEvents), flags_dev + (k % kChunkEvents), sizeof(int),
                                            cudaMemcpyDeviceToHost, h->copy_stream));
            } else {
                NRB_CUDA(h, cudaMemcpyAsync((staging ? staging : host_counts) + r0 * G, counts + r0 * G, (size_t)n * G * sizeof(int32_t),
                                            cudaMemcpyDeviceToHost, h->copy_stream));
            }
            h->timeline.mark("chunk copied out", h->copy_stream);
            if (staging || stage16) {
                cudaEvent_t &done = h->copied_ev[k % kChunkEvents];
                if (!done) NRB_CUDA(h, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
                NRB_CUDA(h, cudaEventRecord(done, h->copy_stream));
                staged.emplace_back(r0, n);
            }
        }
    }
    NRB_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    // Chunks that landed in pinned staging memory are handed to the caller by ONE team of host threads that stays up
    // for the whole drain: thread 0 waits for chunk c's copy event and publishes it, every thread then converts
    // (uint16 -> int32) or copies its slice of the chunk and moves on to the next -- no fork / join per chunk.
    std::vector<size_t> redo;  // wire16: chunks whose overflow flag is up
    if (!staged.empty()) {
        NvtxRange nvtx_drain("nrb200: drain staged chunks");
        std::atomic<int64_t> landed{-1};
        std::atomic<int> drain_err{0};
        const int64_t n_chunks = (int64_t)staged.size();
#pragma omp parallel num_threads(host_threads(h))
        {
            const int tid = omp_get_thread_num(), nth = omp_get_num_threads();
            if (pageable && nth > 1 && tid > 0) {
                // A freshly allocated destination (every R call's result is) faults its pages in on first touch, and with
                // transparent huge pages (numpy and glibc ask for them for a matrix this size) one fault zeroes 2 MB: ~0.15 ms
                // during which every other thread that needs the same 2 MB waits.  So the threads take whole 2 MB blocks of the
                // destination in turn, lowest addresses (first chunks) first, NOW, while the first chunks still compute and
                // travel, instead of inside the conversion loop.  (Thread 0 goes straight to the copy events.)
                char *base = host16 ? reinterpret_cast<char *>(host16) : reinterpret_cast<char *>(host_counts);
                const size_t total = cells * (host16 ? 2 : 4);
                constexpr uintptr_t kBlock = (uintptr_t)2 << 20;
                const auto t_pop = std::chrono::steady_clock::now();
                const uintptr_t first = (uintptr_t)base & ~(kBlock - 1), end = (uintptr_t)base + total;
                for (uintptr_t blk = first + (uintptr_t)(tid - 1) * kBlock; blk < end; blk += (uintptr_t)(nth - 1) * kBlock) {
                    char *lo = reinterpret_cast<char *>(std::max(blk, (uintptr_t)base)), *hi = reinterpret_cast<char *>(std::min(blk + kBlock, end));
                    if (populate_pages(lo, (size_t)(hi - lo))) continue;  // one call, no trap per page
                    for (volatile char *q = lo; q < hi; q += 4096) *q = 0;  // (older kernels: one store per 4 KB page)
                    *(volatile char *)(hi - 1) = 0;
                }
                if (Trace::on() && tid == 1)
                    fprintf(stderr, "[nrb200] drain/populate (thread 1 of %d) %.3f ms\n", nth,
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pop).count());
            }
            for (int64_t c = 0; c < n_chunks; ++c) {
                if (tid == 0) {
                    if (cudaEventSynchronize(h->copied_ev[c % kChunkEvents]) != cudaSuccess) drain_err.store(1);
                    landed.store(c, std::memory_order_release);
                } else {
                    while (landed.load(std::memory_order_acquire) < c) GroupSpin::pause();
                }
                if (drain_err.load()) continue;
                const int64_t r0 = staged[c].first, n = staged[c].second;
                if (wire16 && !host16 && h->flags_host[c % kChunkEvents]) continue;  // redone in int32 below
                const size_t elems = (size_t)n * G;
                const size_t per = ((elems + nth - 1) / nth + 15) & ~(size_t)15;
                const size_t lo = std::min(elems, (size_t)tid * per), hi = std::min(elems, lo + per);
                if (hi <= lo) continue;
                if (host16) std::memcpy(host16 + r0 * G + lo, staging16 + r0 * G + lo, (hi - lo) * sizeof(uint16_t));
                else if (wire16) widen_u16_to_i32(staging16 + r0 * G + lo, host_counts + r0 * G + lo, hi - lo);
                else std::memcpy(host_counts + r0 * G + lo, staging + r0 * G + lo, (hi - lo) * sizeof(int32_t));
            }
        }
        if (drain_err.load()) return fail(h, NRB200_ERR_CUDA, "cudaEventSynchronize failed while draining the count matrix");
        if (wire16 && !host16)
            for (int64_t c = 0; c < n_chunks; ++c)
                if (h->flags_host[c % kChunkEvents]) redo.push_back((size_t)c);
    }
    if (host16) {  // saturated counts are reported, not recomputed: the caller asked for uint16
        NRB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
        for (int c = 0; c < kChunkEvents; ++c) h->out16_overflow |= h->flags_host[c];
    }
    if (!redo.empty()) {
        // some count exceeds 65535: those chunks again, with int32 output (the per-iteration vectors of the second pass go
        // to scratch: the first pass already holds them)
        int64_t n_max = 0;
        for (size_t c : redo) n_max = std::max(n_max, staged[c].second);
        NRB_CUDA(h, h->d_counts.reserve((size_t)n_max * G * sizeof(int32_t)));
        NRB_CUDA(h, h->redo_scratch.reserve((size_t)n_max * 16));
        for (size_t c : redo) {
            const int64_t r0 = staged[c].first, n = staged[c].second;
            NRB_CUDA(h, cudaMemsetAsync(h->redo_scratch.p, 0, (size_t)n * 16, h->stream));
            BootParams Q = slice_params(P, r0, n);
            Q.counts16 = nullptr;
            Q.overflow16 = nullptr;
            Q.counts = h->d_counts.as<int32_t>();
            Q.iter_nranges = h->redo_scratch.as<unsigned long long>();
            Q.iter_totals = h->redo_scratch.as<unsigned long long>() + n;
            Q.n_hits = nullptr;
            launches += launch_count(h, Q, has_query);
            NRB_CUDA(h, cudaGetLastError());
            if ((rc = download(h, host_counts + r0 * G, h->d_counts.p, (size_t)n * G * sizeof(int32_t)))) return rc;
        }
    }
    h->res_has_counts = want_counts && !wire16;
    h->last_wire16 = wire16;
    if (stats) {
        NRB_CUDA(h, cudaEventSynchronize(h->ev[2]));
        float k_ms = 0, t_ms = 0;
        cudaEventElapsedTime(&k_ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&t_ms, h->ev[0], h->ev[2]);
        unsigned long long hit_words[kHitSlots + 1];
        if ((rc = hits_fetch(h, hit_words))) return rc;
        if (int rc_sync = sync_stream(h)) return rc_sync;
        const unsigned long long hits = hits_total(hit_words);
        std::memset(stats, 0, sizeof *stats);
        stats->kernel_ms = k_ms;
        stats->total_ms = t_ms;
        stats->n_launches = launches;
        stats->n_hits = (int64_t)hits;
        if (split) {
            stats->n_windows = has_query ? h->n_pairs : 0;
            for (int b = 0; b < k; ++b) {
                float a_ms = 0, b_ms = 0;
                cudaEventElapsedTime(&a_ms, h->batch_ev[3 * b], h->batch_ev[3 * b + 1]);
                cudaEventElapsedTime(&b_ms, h->batch_ev[3 * b + 1], h->batch_ev[3 * b + 2]);
                stats->rows_kernel_ms += a_ms;
                stats->count_kernel_ms += b_ms;
            }
        } else {
            stats->count_kernel_ms = k_ms;  // streaming path: one kernel does both
        }
        stats->h2d_bytes = d->rng_mode == NRB200_RNG_HOST ? R * h->plan.n_blocks() * (d->dst_tile ? 12 : 8) : 0;
    }
    return NRB200_OK;
}

int fetch_u64_as_i64(nrb200_handle *h, const DevMem &src, int64_t n, int64_t *dst) {
    if (!dst || n == 0) return NRB200_OK;
    NRB_CUDA(h, cudaMemcpyAsync(dst, src.p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    return NRB200_OK;
}

}  // namespace
