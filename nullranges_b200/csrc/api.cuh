// libnrb200: the extern "C" entry points of include/nrb200.h.  (textually included by engine.cu)
extern "C" {

const char *nrb200_version(void) { return "nrb200 0.2.0 (sm_100a)"; }

int nrb200_zero_width_rule(void) { return kZeroWidthStrictlyInside ? 1 : 0; }

const char *nrb200_last_error(const nrb200_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int nrb200_create(const nrb200_config *cfg, nrb200_handle **out) {
    if (!out) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: out is NULL");
    *out = nullptr;
    if (cfg && cfg->n_devices > 1) return group_create(cfg, out);
    if (cfg && cfg->n_devices < 0) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: n_devices must be >= 0");
    if (cfg && cfg->n_devices == 1 && !cfg->devices) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: n_devices = 1 needs the device list");
    const int dev = cfg ? (cfg->n_devices == 1 ? cfg->devices[0] : cfg->device) : 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, NRB200_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                                 " (libnrb200 has no CPU path)");
    if (dev < 0 || dev >= count) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: device ordinal out of range");
    NRB_CUDA(nullptr, cudaSetDevice(dev));
    nrb200_handle *h = new nrb200_handle();
    h->device = dev;
    h->flags = cfg ? cfg->flags : 0;
    if (const char *e = getenv("NRB200_IPT")) h->tune_ipt = std::min(std::max(atoi(e), 1), kLoopMaxIpt);
    if (const char *e = getenv("NRB200_LOOP_MINB")) h->tune_loop_minb = atoi(e);
    if (const char *e = getenv("NRB200_STREAM_OLD")) h->tune_stream_old = atoi(e) != 0;
    if (const char *e = getenv("NRB200_END_RANK")) h->tune_end_rank = atoi(e) != 0;
    if (const char *e = getenv("NRB200_ROWS_IPC")) h->tune_rows_ipc = std::min(std::max(atoi(e), 1), 64);
    if (const char *e = getenv("NRB200_TABLE_MULT")) h->tune_table_mult = atof(e);
    if (const char *e = getenv("NRB200_TABLE_CAP")) h->tune_table_cap = atoll(e);
    if (const char *e = getenv("NRB200_WIRE16")) h->tune_wire16 = std::min(std::max(atoi(e), 0), 2);
    if (const char *e = getenv("NRB200_CHUNKS")) h->tune_chunks = std::min(std::max(atoi(e), 1), 64);
    if (const char *e = getenv("NRB200_MAX_BATCH")) h->tune_max_batch = std::min<int64_t>(std::max<int64_t>(atoll(e), 8), 2048);
    if (cfg && cfg->stream) {
        h->stream = (cudaStream_t)cfg->stream;
    } else {
        e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete h;
            return fail(nullptr, NRB200_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
        }
        h->own_stream = true;
    }
    for (auto &ev : h->ev) cudaEventCreate(&ev);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
    *out = h;
    return NRB200_OK;
}

int nrb200_n_devices(const nrb200_handle *h) { return !h ? 0 : (is_group(h) ? group_size(h) : 1); }

void nrb200_destroy(nrb200_handle *h) {
    if (!h) return;
    if (is_group(h)) return group_destroy(h);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto &ev : h->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->chunk_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->copied_ev)
        if (ev) cudaEventDestroy(ev);
    if (h->stage_buf) cudaFreeHost(h->stage_buf);
    if (h->stage_in) cudaFreeHost(h->stage_in);
    if (h->m_stage) cudaFreeHost(h->m_stage);
    if (h->y_ev) cudaEventDestroy(h->y_ev);
    if (h->inspect_host) cudaFreeHost(h->inspect_host);
    if (h->flags_host) cudaFreeHost(h->flags_host);
    for (auto &ev : h->batch_ev) cudaEventDestroy(ev);
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamDestroy(h->copy_stream);
    }
    if (h->arena) cudaFreeHost(h->arena);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

int nrb200_set_ranges_begin(nrb200_handle *h, int64_t n, int32_t n_seq, const int64_t *seqlengths, const int32_t *seqid,
                            const int32_t *start, const int32_t *end, const int8_t *strand) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_set_ranges_begin(h, n, n_seq, seqlengths, seqid, start, end, strand);
    NRB_CUDA(h, cudaSetDevice(h->device));
    NvtxRange nvtx("nrb200: upload y + inspect");
    h->y_pending = false;
    if (n_seq <= 0 || !seqlengths) return fail(h, NRB200_ERR_INVALID, "seqlengths(y) must be known");
    for (int c = 0; c < n_seq; ++c) {
        // stopifnot(all(!is.na(chr_lens)))                          R/bootRanges.R:81
        if (seqlengths[c] <= 0) return fail(h, NRB200_ERR_INVALID, "all(!is.na(seqlengths(y))) is not TRUE");
        if (seqlengths[c] >= (1 << 30)) return fail(h, NRB200_ERR_UNSUPPORTED, "seqlengths must be below 2^30");
    }
    h->have_y = h->have_plan = false;
    h->query.n = h->excl.n = 0;  // seqlevels may have changed
    h->query_minoverlap = 0;
    h->query_widen = 0;
    h->have_weights = false;      // weights belong to the ranges they were set for
    h->moments_ready = false;     // accumulated moments belong to the query they were taken against
    h->n_seq = n_seq;
    h->seqlen.assign(seqlengths, seqlengths + n_seq);
    int rc;
    if ((rc = upload(h, h->seqlen_dev, seqlengths, (size_t)n_seq))) return rc;
    if ((rc = load_y_begin(h, n, seqid, start, end, strand))) return rc;
    h->y_pending = true;
    return NRB200_OK;
}

int nrb200_set_ranges_end(nrb200_handle *h) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_all(h, [](int, nrb200_handle *m) { return nrb200_set_ranges_end(m); });
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (!h->y_pending) return fail(h, NRB200_ERR_STATE, "nrb200_set_ranges_begin() has not been called");
    NvtxRange nvtx("nrb200: verdict + index build");
    h->y_pending = false;
    Trace tr("set_ranges_end");
    h->slices_dirty = true;
    int rc = load_y_end(h);
    if (rc < 0) return rc;
    tr.lap("load_y_end");
    if (rc == 1 && (rc = load_range_set(h, h->y, "y", h->pend.n, h->pend.seqid, h->pend.start, h->pend.end, h->pend.strand, true)))
        return rc;
    h->pend = {};
    h->timeline.mark("verdict read", h->stream);
    if ((rc = build_rank_tables(h))) return rc;
    h->timeline.mark("index built", h->stream);
    tr.lap("build_rank_tables");
    // the large host mirror of y is not needed again
    std::vector<int32_t>().swap(h->y.h_start);
    std::vector<int32_t>().swap(h->y.h_end);
    std::vector<int32_t>().swap(h->y.h_pmax);
    std::vector<int32_t>().swap(h->y.h_src);
    std::vector<int8_t>().swap(h->y.h_strand);
    h->have_y = true;
    h->slices_dirty = true;
    return NRB200_OK;
}

int nrb200_set_ranges(nrb200_handle *h, int64_t n, int32_t n_seq, const int64_t *seqlengths, const int32_t *seqid,
                      const int32_t *start, const int32_t *end, const int8_t *strand) {
    if (int rc = nrb200_set_ranges_begin(h, n, n_seq, seqlengths, seqid, start, end, strand)) return rc;
    return nrb200_set_ranges_end(h);
}

int nrb200_ranges_info(const nrb200_handle *h, int32_t *is_sorted, int32_t *is_stranded, int32_t *max_width) {
    if (is_group(h)) h = h->members[0];
    if (!h || !h->have_y) return NRB200_ERR_STATE;
    if (is_sorted) *is_sorted = h->y.sorted_input;
    if (is_stranded) *is_stranded = h->y.stranded;
    if (max_width) *max_width = h->y.wmax;
    return NRB200_OK;
}

static int set_side(nrb200_handle *h, RangeSet &rs, const char *what, int64_t n, const int32_t *seqid,
                    const int32_t *start, const int32_t *end, const int8_t *strand, int32_t widen = 0) {
    if (!h) return NRB200_ERR_INVALID;
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (!h->have_y && !h->y_pending) return fail(h, NRB200_ERR_STATE, "call nrb200_set_ranges() first (it defines the seqlevels)");
    h->slices_dirty = true;
    if (&rs == &h->query) h->moments_ready = false;  // a new query: sums of the old one must not be mixed in (nor its G trusted)
    if (n == 0) {
        rs.n = 0;
        return NRB200_OK;
    }
    return load_range_set(h, rs, what, n, seqid, start, end, strand, false, &rs == &h->query, widen);
}

int nrb200_set_query(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start, const int32_t *end,
                     const int8_t *strand) {
    return nrb200_set_query_ex(h, n, seqid, start, end, strand, -1, 0);
}

// maxgap: two ranges count as overlapping when at most `maxgap` positions separate them
// (start_q <= end_s + maxgap + 1 and end_q >= start_s - maxgap - 1).  That is plain overlap with every
// feature widened by maxgap + 1 on both sides, which is how the query is stored; nothing downstream
// changes.  (Bootstrapped ranges lie inside [1, seqlength], so the widening cannot create overlaps
// that trim() would have removed.)
// minoverlap: the intersection must be at least that wide.  On the rank path this is a condition on the window's
// end points plus "the range itself is at least that wide": y is indexed in width classes (build_rank_tables).
int nrb200_set_query_ex(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start, const int32_t *end,
                        const int8_t *strand, int32_t maxgap, int32_t minoverlap) {
    if (!h) return NRB200_ERR_INVALID;
    if (maxgap < -1) return fail(h, NRB200_ERR_INVALID, "'maxgap' must be >= -1");
    if (maxgap >= (1 << 29)) return fail(h, NRB200_ERR_UNSUPPORTED, "'maxgap' must be below 2^29");
    if (minoverlap < 0) return fail(h, NRB200_ERR_INVALID, "'minoverlap' must be >= 0");
    if (minoverlap > 0 && maxgap != -1) return fail(h, NRB200_ERR_INVALID, "'maxgap' and 'minoverlap' cannot both be set");
    if (is_group(h)) return group_all(h, [&](int, nrb200_handle *m) { return nrb200_set_query_ex(m, n, seqid, start, end, strand, maxgap, minoverlap); });
    const int rc = set_side(h, h->query, "query", n, seqid, start, end, strand, maxgap + 1);
    if (rc == NRB200_OK) {
        h->query_widen = maxgap + 1;
        h->query_minoverlap = n > 0 ? minoverlap : 0;
    }
    return rc;
}

int nrb200_set_exclude(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start, const int32_t *end,
                       const int8_t *strand) {
    if (is_group(h)) return group_all(h, [&](int, nrb200_handle *m) { return nrb200_set_exclude(m, n, seqid, start, end, strand); });
    if (!h) return NRB200_ERR_INVALID;
    return set_side(h, h->excl, "exclude", n, seqid, start, end, strand);
}

int nrb200_set_exclude_option(nrb200_handle *h, int32_t option) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_all(h, [&](int, nrb200_handle *m) { return nrb200_set_exclude_option(m, option); });
    if (option != NRB200_EXCLUDE_DROP && option != NRB200_EXCLUDE_TRIM)
        return fail(h, NRB200_ERR_INVALID, "'excludeOption' should be one of 'drop', 'trim'");
    // stopifnot(all(strand(exclude) == "*"))                       R/bootRanges.R:84
    if (option == NRB200_EXCLUDE_TRIM && h->excl.n > 0 && h->excl.stranded)
        return fail(h, NRB200_ERR_INVALID, "all(strand(exclude) == \"*\") is not TRUE");
    h->excl_option = option;
    h->slices_dirty = true;
    return NRB200_OK;
}

int nrb200_set_plan(nrb200_handle *h, const nrb200_plan_spec *spec, int64_t *n_blocks) {
    if (!h || !spec) return NRB200_ERR_INVALID;
    if (is_group(h)) {
        const int rc = group_all(h, [&](int, nrb200_handle *m) { return nrb200_set_plan(m, spec, nullptr); });
        if (rc == NRB200_OK && n_blocks) *n_blocks = h->members[0]->plan.n_blocks();
        return rc;
    }
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (!h->have_y && !h->y_pending) return fail(h, NRB200_ERR_STATE, "call nrb200_set_ranges() first (the plan needs seqlengths)");
    NvtxRange nvtx("nrb200: plan");
    Plan p;
    const std::string msg = build_plan(*spec, h->seqlen, p);
    if (!msg.empty()) return fail(h, NRB200_ERR_INVALID, msg);
    if (p.n_blocks() >= ((int64_t)1 << (31 - kRecTileShift))) return fail(h, NRB200_ERR_UNSUPPORTED, "too many blocks per iteration");
    h->plan = std::move(p);
    const Plan &pl = h->plan;
    const size_t B = (size_t)pl.n_blocks();
    int rc;
    if ((rc = upload(h, h->tile_seq, pl.tile_seq.data(), B))) return rc;
    if ((rc = upload(h, h->tile_start, pl.tile_start.data(), B))) return rc;
    if ((rc = upload(h, h->bait_width, pl.bait_width.data(), B))) return rc;
    {
        std::vector<int32_t> info(2 * B);  // {bait width, length of the tile's sequence} per tile, one 8-byte load
        for (size_t t = 0; t < B; ++t) {
            info[2 * t] = pl.bait_width[t];
            info[2 * t + 1] = (int32_t)h->seqlen[pl.tile_seq[t]];
        }
        if ((rc = upload(h, h->tile_info, info.data(), info.size()))) return rc;
    }
    // sampler tables packed into one int64 and one int32 buffer
    std::vector<int64_t> i64;
    std::vector<int32_t> i32;
    auto put64 = [&](const std::vector<int64_t> &v) { size_t o = i64.size(); i64.insert(i64.end(), v.begin(), v.end()); return o; };
    auto put32 = [&](const std::vector<int32_t> &v) { size_t o = i32.size(); i32.insert(i32.end(), v.begin(), v.end()); return o; };
    const size_t o_cum = put64(pl.seq_cum), o_tps = put64(pl.tiles_per_seq), o_sbo = put64(pl.state_block_off),
                 o_sso = put64(pl.state_seg_off), o_sl = put64(pl.state_len), o_sbl = put64(pl.state_block_len),
                 o_sgc = put64(pl.sg_cum);
    const size_t o_sq = put32(pl.sg_seq), o_ss = put32(pl.sg_start), o_st = put32(pl.sg_tiles);
    // coarse pick tables (philox.cuh, pick_by_length_tab): [0] over the sequences, [st] over the segments of state st
    std::vector<int32_t> pick_tab((size_t)(pl.n_states + 1) * kPickBuckets, 0), pick_shift(pl.n_states + 1, 0);
    auto make_pick = [&](int slot, const int64_t *cum, int n, int64_t total) {
        int shift = 0;
        while ((total >> shift) >= kPickBuckets) ++shift;
        pick_shift[slot] = shift;
        int i = 0;
        for (int k = 0; k < kPickBuckets; ++k) {
            const int64_t at = (int64_t)k << shift;
            while (i + 1 < n && cum[i + 1] <= at) ++i;
            pick_tab[(size_t)slot * kPickBuckets + k] = i;
        }
    };
    make_pick(0, pl.seq_cum.data(), pl.n_seq, pl.genome_length);
    for (int st = 1; st <= pl.n_states; ++st) {
        const int64_t first = pl.state_seg_off[st];
        make_pick(st, pl.sg_cum.data() + first, (int)(pl.state_seg_off[st + 1] - first), pl.state_len[st]);
    }
    const size_t o_pt = put32(pick_tab), o_ps = put32(pick_shift);
    if ((rc = upload(h, h->sampler_i64, i64.data(), i64.size()))) return rc;
    if ((rc = upload(h, h->sampler_i32, i32.data(), i32.size()))) return rc;
    const int64_t *b64 = h->sampler_i64.as<int64_t>();
    const int32_t *b32 = h->sampler_i32.as<int32_t>();
    h->sampler_view = SamplerView{pl.kind, pl.n_seq, pl.n_states, pl.block_length, pl.genome_length,
                                  b64 + o_cum, b64 + o_tps, b64 + o_sbo, b64 + o_sso, b64 + o_sl, b64 + o_sbl,
                                  b64 + o_sgc, b32 + o_sq, b32 + o_ss, b32 + o_st, b32 + o_pt, b32 + o_ps};
    {   // tiles of each sequence by start, for the window lists (build_pairs)
        auto &T = h->tiles_by_seq;
        const int C = h->n_seq;
        T.id.assign(C, {});
        T.ts.assign(C, {});
        T.te.assign(C, {});
        T.te_max.assign(C, {});
        for (int64_t t = 0; t < pl.n_blocks(); ++t) T.id[pl.tile_seq[t]].push_back((int32_t)t);
        for (int c = 0; c < C; ++c) {
            auto &v = T.id[c];
            bool sorted = true;
            for (size_t j = 1; j < v.size(); ++j) sorted &= pl.tile_start[v[j - 1]] <= pl.tile_start[v[j]];
            if (!sorted)
                std::sort(v.begin(), v.end(), [&](int32_t a, int32_t b) {
                    return pl.tile_start[a] != pl.tile_start[b] ? pl.tile_start[a] < pl.tile_start[b] : a < b;
                });
            T.ts[c].resize(v.size());
            T.te[c].resize(v.size());
            T.te_max[c].resize(v.size());
            int64_t m = INT64_MIN;
            for (size_t j = 0; j < v.size(); ++j) {
                T.ts[c][j] = pl.tile_start[v[j]];
                T.te[c][j] = T.ts[c][j] + pl.bait_width[v[j]] - 1;
                m = std::max(m, T.te[c][j]);
                T.te_max[c][j] = m;
            }
        }
    }
    h->have_plan = true;
    h->slices_dirty = true;
    if (n_blocks) *n_blocks = pl.n_blocks();
    return NRB200_OK;
}

int nrb200_get_tiles(const nrb200_handle *h, int32_t *tile_seqid, int32_t *tile_start, int32_t *bait_width) {
    if (is_group(h)) h = h->members[0];
    if (!h || !h->have_plan) return NRB200_ERR_STATE;
    const size_t bytes = (size_t)h->plan.n_blocks() * sizeof(int32_t);
    if (tile_seqid) std::memcpy(tile_seqid, h->plan.tile_seq.data(), bytes);
    if (tile_start) std::memcpy(tile_start, h->plan.tile_start.data(), bytes);
    if (bait_width) std::memcpy(bait_width, h->plan.bait_width.data(), bytes);
    return NRB200_OK;
}

int nrb200_run_counts_resident(nrb200_handle *h, const nrb200_draws *draws, int32_t want_feature_counts,
                               nrb200_stats *stats) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return fail(h, NRB200_ERR_UNSUPPORTED, "device-resident results are per GPU: use one handle per device (or nrb200_run_counts)");
    NRB_CUDA(h, cudaSetDevice(h->device));
    return run_counts_core(h, draws, want_feature_counts != 0, nullptr, stats);
}

int nrb200_resident_pointers(const nrb200_handle *h, void **iter_nranges_dev, void **iter_totals_dev,
                             void **feature_counts_dev) {
    if (!h || is_group(h)) return NRB200_ERR_INVALID;
    if (iter_nranges_dev) *iter_nranges_dev = h->d_iter_nranges.p;
    if (iter_totals_dev) *iter_totals_dev = h->d_iter_totals.p;
    if (feature_counts_dev) *feature_counts_dev = h->res_has_counts ? h->d_counts.p : nullptr;
    return NRB200_OK;
}

int nrb200_run_counts_gathered(nrb200_handle *h, const nrb200_draws *draws, nrb200_stats *stats) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_run_counts_gathered(h, draws, stats);
    if (int rc = nrb200_run_counts_resident(h, draws, 1, stats)) return rc;  // one GPU: gathered is resident ...
    return sync_stream(h);                                                    // ... and complete on return, as for a group
}

int nrb200_gathered_pointers(const nrb200_handle *h, int32_t member, int32_t *device, void **iter_nranges_dev, void **iter_totals_dev,
                             void **feature_counts_dev) {
    if (!h) return NRB200_ERR_INVALID;
    if (!is_group(h)) {
        if (member != 0) return NRB200_ERR_INVALID;
        if (device) *device = h->device;
        return nrb200_resident_pointers(h, iter_nranges_dev, iter_totals_dev, feature_counts_dev);
    }
    if (member < 0 || member >= group_size(h)) return NRB200_ERR_INVALID;
    const nrb200_handle *m = h->members[member];
    if (device) *device = m->device;
    if (iter_nranges_dev) *iter_nranges_dev = m->gat_nranges.p;
    if (iter_totals_dev) *iter_totals_dev = m->gat_totals.p;
    if (feature_counts_dev) *feature_counts_dev = m->gat_counts.p;
    return NRB200_OK;
}

int nrb200_run_counts(nrb200_handle *h, const nrb200_draws *draws, int64_t *iter_nranges, int64_t *iter_totals,
                      int32_t *feature_counts, nrb200_stats *stats) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_run_counts(h, draws, iter_nranges, iter_totals, feature_counts, stats);
    NRB_CUDA(h, cudaSetDevice(h->device));
    int rc = run_counts_core(h, draws, feature_counts != nullptr, feature_counts, stats);
    if (rc) return rc;
    const int64_t R = draws->n_iter, G = h->query.n;
    if ((rc = fetch_u64_as_i64(h, h->d_iter_nranges, R, iter_nranges))) return rc;
    if ((rc = fetch_u64_as_i64(h, h->d_iter_totals, R, iter_totals))) return rc;
    if (int rc_sync = sync_stream(h)) return rc_sync;
    if (h->copy_stream) NRB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    h->timeline.mark("run_counts returns", h->stream);
    h->timeline.dump();
    if (stats) stats->d2h_bytes = (iter_nranges ? R * 8 : 0) + (iter_totals ? R * 8 : 0) + (feature_counts ? R * G * (h->last_wire16 ? 2 : 4) : 0);
    return NRB200_OK;
}

int nrb200_run_counts_u16(nrb200_handle *h, const nrb200_draws *draws, int64_t *iter_nranges, int64_t *iter_totals,
                          uint16_t *feature_counts, int32_t *overflow, nrb200_stats *stats) {
    if (!h || !feature_counts) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_run_counts_u16(h, draws, iter_nranges, iter_totals, feature_counts, overflow, stats);
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (h->query.n == 0) return fail(h, NRB200_ERR_STATE, "nrb200_set_query() has not been called");
    h->out16 = feature_counts;
    int rc = run_counts_core(h, draws, true, nullptr, stats);
    h->out16 = nullptr;
    if (rc) return rc;
    const int64_t R = draws->n_iter, G = h->query.n;
    if ((rc = fetch_u64_as_i64(h, h->d_iter_nranges, R, iter_nranges))) return rc;
    if ((rc = fetch_u64_as_i64(h, h->d_iter_totals, R, iter_totals))) return rc;
    if (int rc_sync = sync_stream(h)) return rc_sync;
    if (h->copy_stream) NRB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    if (overflow) *overflow = h->out16_overflow;
    if (stats) stats->d2h_bytes = (iter_nranges ? R * 8 : 0) + (iter_totals ? R * 8 : 0) + R * G * 2;
    return NRB200_OK;
}

int nrb200_moments_reset(nrb200_handle *h) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_all(h, [](int, nrb200_handle *m) { return nrb200_moments_reset(m); });
    NRB_CUDA(h, cudaSetDevice(h->device));
    const int64_t G = h->query.n;
    if (G == 0) return fail(h, NRB200_ERR_STATE, "moments need a query (nrb200_set_query)");
    NRB_CUDA(h, h->m_sum.reserve(G * 8));
    NRB_CUDA(h, h->m_sumsq.reserve(G * 8));
    NRB_CUDA(h, h->m_nge.reserve(G * 8));
    NRB_CUDA(h, cudaMemsetAsync(h->m_sum.p, 0, G * 8, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->m_sumsq.p, 0, G * 8, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->m_nge.p, 0, G * 8, h->stream));
    h->moments_ready = true;
    h->moments_G = G;
    return NRB200_OK;
}

int nrb200_run_moments(nrb200_handle *h, const nrb200_draws *draws, int64_t batch_iter, const int32_t *observed,
                       int64_t *iter_nranges, int64_t *iter_totals, nrb200_stats *stats) {
    if (!h || !draws) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_run_moments(h, draws, batch_iter, observed, iter_nranges, iter_totals, stats);
    NRB_CUDA(h, cudaSetDevice(h->device));
    const int64_t G = h->query.n;
    if (G == 0) return fail(h, NRB200_ERR_STATE, "moments need a query (nrb200_set_query)");
    int rc;
    if ((!h->moments_ready || h->moments_G != G) && (rc = nrb200_moments_reset(h))) return rc;
    if (observed && (rc = upload(h, h->m_observed, observed, (size_t)G))) return rc;
    if ((rc = prepare(h))) return rc;
    if (rank_path_ok(h)) {
        // Fused: the count kernel keeps sum / sum of squares / #{>= observed} of every feature in registers over a run
        // of iterations and never writes the [R x G] matrix; batches (bounded draw-table scratch) follow each other on
        // the stream without a host synchronisation; one wait at the end.
        const int64_t R = draws->n_iter, Bk = h->plan.n_blocks();
        if (R < 0) return fail(h, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
        NRB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
        if ((rc = stage_draws(h, draws))) return rc;
        const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(batch_iter > 0 ? batch_iter : h->tune_max_batch, h->tune_max_batch));
        NRB_CUDA(h, h->d_iter_nranges.reserve(std::max<int64_t>(R, 1) * 8));
        NRB_CUDA(h, h->d_iter_totals.reserve(std::max<int64_t>(R, 1) * 8));
        NRB_CUDA(h, cudaMemsetAsync(h->d_iter_nranges.p, 0, R * 8, h->stream));
        NRB_CUDA(h, cudaMemsetAsync(h->d_iter_totals.p, 0, R * 8, h->stream));
        if ((rc = hits_reset(h))) return rc;
        NRB_CUDA(h, h->draw_tab.reserve((size_t)std::min(batch, std::max<int64_t>(R, 1)) * Bk * sizeof(int2)));
        BootParams P = make_params(h, draws);
        P.iter_nranges = h->d_iter_nranges.as<unsigned long long>();
        P.iter_totals = h->d_iter_totals.as<unsigned long long>();
        P.n_hits = stats ? h->d_n_hits.as<unsigned long long>() : nullptr;
        P.draw_tab = h->draw_tab.as<int2>();
        const FeatView F{h->f_src.as<int32_t>(), h->pair_off.as<int32_t>(), h->pair_rec.as<int4>()};
        const MomentsOut M{observed ? h->m_observed.as<int32_t>() : nullptr, h->m_sum.as<long long>(), h->m_sumsq.as<long long>(),
                           h->m_nge.as<long long>()};
        NRB_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
        int launches = 0;
        for (int64_t r0 = 0; r0 < R; r0 += batch) {
            const BootParams Q = slice_params(P, r0, std::min(batch, R - r0));
            launches += launch_rank(h, Q, false);  // draws + rows per iteration
            const unsigned gx = (unsigned)((G + kCountThreads - 1) / kCountThreads);
            const int ipt = (int)std::max<int64_t>(1, std::min<int64_t>(kLoopMaxIpt, (int64_t)gx * Q.n_iter / ((int64_t)h->sm_count * 16 * 3)));
            boot_rank_loop_kernel<true, 10><<<dim3(gx, (unsigned)((Q.n_iter + ipt - 1) / ipt)), kCountThreads, 0, h->stream>>>(Q, F, ipt, M);
            ++launches;
            NRB_CUDA(h, cudaGetLastError());
        }
        NRB_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
        if ((rc = fetch_u64_as_i64(h, h->d_iter_nranges, R, iter_nranges))) return rc;
        if ((rc = fetch_u64_as_i64(h, h->d_iter_totals, R, iter_totals))) return rc;
        unsigned long long hit_words[kHitSlots + 1] = {};
        if (stats && (rc = hits_fetch(h, hit_words))) return rc;
        if (int rc_sync = sync_stream(h)) return rc_sync;
        if (stats) {
            float k_ms = 0, t_ms = 0;
            cudaEventElapsedTime(&k_ms, h->ev[1], h->ev[2]);
            cudaEventElapsedTime(&t_ms, h->ev[0], h->ev[2]);
            std::memset(stats, 0, sizeof *stats);
            stats->kernel_ms = k_ms;
            stats->total_ms = t_ms;
            stats->n_launches = launches;
            stats->n_hits = (int64_t)hits_total(hit_words);
            stats->n_windows = h->n_pairs;
            stats->d2h_bytes = (iter_nranges ? R * 8 : 0) + (iter_totals ? R * 8 : 0);
            stats->h2d_bytes = draws->rng_mode == NRB200_RNG_HOST ? R * Bk * (draws->dst_tile ? 12 : 8) : 0;
        }
        return NRB200_OK;
    }
    // streaming path: the matrix of a batch of iterations, then its reduction
    if (batch_iter <= 0) batch_iter = std::max<int64_t>(1, std::min<int64_t>(draws->n_iter, ((int64_t)1 << 28) / std::max<int64_t>(G, 1)));
    const int64_t B = h->plan.n_blocks();
    nrb200_stats acc{};
    for (int64_t r0 = 0; r0 < draws->n_iter; r0 += batch_iter) {
        nrb200_draws d = *draws;
        d.n_iter = std::min(batch_iter, draws->n_iter - r0);
        d.iter0 = draws->iter0 + r0;
        if (draws->rng_mode == NRB200_RNG_HOST) {
            d.bait_seqid = draws->bait_seqid + r0 * B;
            d.bait_start = draws->bait_start + r0 * B;
            d.dst_tile = draws->dst_tile ? draws->dst_tile + r0 * B : nullptr;
        }
        nrb200_stats st{};
        if ((rc = run_counts_core(h, &d, true, nullptr, &st))) return rc;
        const int threads = 256;
        accumulate_moments_kernel<<<(unsigned)((G + threads - 1) / threads), threads, 0, h->stream>>>(
            h->d_counts.as<int32_t>(), d.n_iter, G, observed ? h->m_observed.as<int32_t>() : nullptr,
            h->m_sum.as<long long>(), h->m_sumsq.as<long long>(), h->m_nge.as<long long>());
        NRB_CUDA(h, cudaGetLastError());
        if ((rc = fetch_u64_as_i64(h, h->d_iter_nranges, d.n_iter, iter_nranges ? iter_nranges + r0 : nullptr))) return rc;
        if ((rc = fetch_u64_as_i64(h, h->d_iter_totals, d.n_iter, iter_totals ? iter_totals + r0 : nullptr))) return rc;
        if (int rc_sync = sync_stream(h)) return rc_sync;
        acc.kernel_ms += st.kernel_ms;
        acc.rows_kernel_ms += st.rows_kernel_ms;
        acc.count_kernel_ms += st.count_kernel_ms;
        acc.total_ms += st.total_ms;
        acc.n_launches += st.n_launches + 1;
        acc.n_hits += st.n_hits;
        acc.h2d_bytes += st.h2d_bytes;
    }
    if (stats) *stats = acc;
    return NRB200_OK;
}

int nrb200_moments_fetch(nrb200_handle *h, int64_t *feat_sum, int64_t *feat_sumsq, int64_t *feat_n_ge_observed) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_moments_fetch(h, feat_sum, feat_sumsq, feat_n_ge_observed);
    if (!h->moments_ready || h->moments_G != h->query.n) return fail(h, NRB200_ERR_STATE, "no moments accumulated for this query");
    NRB_CUDA(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->moments_G * 8;
    if (feat_sum) NRB_CUDA(h, cudaMemcpyAsync(feat_sum, h->m_sum.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (feat_sumsq) NRB_CUDA(h, cudaMemcpyAsync(feat_sumsq, h->m_sumsq.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (feat_n_ge_observed) NRB_CUDA(h, cudaMemcpyAsync(feat_n_ge_observed, h->m_nge.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (int rc_sync = sync_stream(h)) return rc_sync;
    return NRB200_OK;
}

int nrb200_set_weights(nrb200_handle *h, const double *weights) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_all(h, [&](int, nrb200_handle *m) { return nrb200_set_weights(m, weights); });
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (!h->have_y) return fail(h, NRB200_ERR_STATE, "call nrb200_set_ranges() first");
    if (!weights) {
        h->have_weights = false;
        return NRB200_OK;
    }
    const int64_t n = h->y.n;
    int rc;
    if ((rc = upload_direct(h, h->w_in, weights, (size_t)n))) return rc;
    NRB_CUDA(h, h->w_c.reserve((size_t)std::max<int64_t>(n, 1) * 8));
    if (n > 0) {
        // caller's order -> canonical order (identity when the input was already canonical; else y.src, on the device)
        gather_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->w_in.as<double>(), h->y.sorted_input ? nullptr : h->y.src.as<int32_t>(),
                                                                              n, h->w_c.as<double>());
        NRB_CUDA(h, cudaGetLastError());
    }
    if ((rc = sync_stream(h))) return rc;  // the caller's array is no longer read
    h->have_weights = true;
    h->weights_dirty = true;
    return NRB200_OK;
}

struct WeightedMoments {      // per-feature moments of the weighted sums over the iterations of one call (all optional)
    const double *observed;
    double *sum, *sumsq;
    int64_t *n_ge;
};

static int run_weighted_core(nrb200_handle *h, const nrb200_draws *draws, double *iter_sums, double *feature_sums,
                             const WeightedMoments *mom, nrb200_stats *stats) {
    if (!h || !draws) return NRB200_ERR_INVALID;
    NRB_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = prepare(h))) return rc;
    if (!h->have_weights) return fail(h, NRB200_ERR_STATE, "nrb200_set_weights() has not been called");
    if (h->query.n == 0) return fail(h, NRB200_ERR_STATE, "nrb200_set_query() has not been called");
    if (!rank_path_ok(h))
        return fail(h, NRB200_ERR_UNSUPPORTED, "weighted sums run on the rank path only (segmentation within seqlengths, no NRB200_FLAG_FORCE_STREAM)");
    if ((rc = build_weight_index(h))) return rc;
    NRB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if ((rc = stage_draws(h, draws))) return rc;
    const int64_t R = draws->n_iter, G = h->query.n, B = h->plan.n_blocks();
    const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(h->tune_max_batch, ((int64_t)1 << 27) / std::max<int64_t>(G, 1)));
    NRB_CUDA(h, h->d_iter_nranges.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->ws_iter.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->ws_sums.reserve((size_t)std::min(batch, std::max<int64_t>(R, 1)) * G * 8));
    NRB_CUDA(h, h->draw_tab.reserve((size_t)std::min(batch, std::max<int64_t>(R, 1)) * B * sizeof(int2)));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_nranges.p, 0, R * 8, h->stream));
    if ((rc = hits_reset(h))) return rc;
    NRB_CUDA(h, cudaMemsetAsync(h->ws_iter.p, 0, R * 8, h->stream));
    if (mom) {
        for (DevMem *m : {&h->wm_sum, &h->wm_sumsq, &h->wm_nge}) {
            NRB_CUDA(h, m->reserve((size_t)G * 8));
            NRB_CUDA(h, cudaMemsetAsync(m->p, 0, (size_t)G * 8, h->stream));
        }
        if (mom->observed && (rc = upload(h, h->wm_observed, mom->observed, (size_t)G))) return rc;
    }
    BootParams P = make_params(h, draws);
    P.iter_nranges = h->d_iter_nranges.as<unsigned long long>();
    P.n_hits = h->d_n_hits.as<unsigned long long>();
    P.draw_tab = h->draw_tab.as<int2>();
    const FeatView F{h->f_src.as<int32_t>(), h->pair_off.as<int32_t>(), h->pair_rec.as<int4>()};
    const WeightView W{h->n_cls > 1 ? h->w_r.as<double>() : h->w_c.as<double>(), h->w_ps.as<double>(),
                       h->ends_alias_starts ? h->w_ps.as<double>() : h->w_pe.as<double>()};
    NRB_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    int launches = 0;
    for (int64_t r0 = 0; r0 < R; r0 += batch) {
        const int64_t n = std::min(batch, R - r0);
        const BootParams Q = slice_params(P, r0, n);
        launches += launch_rank(h, Q, false);  // rows kernel: draws the blocks, fills the draw table
        boot_rank_wsum_kernel<<<dim3((unsigned)((G + 255) / 256), (unsigned)n), 256, 0, h->stream>>>(Q, F, W, h->ws_sums.as<double>(),
                                                                                                     h->ws_iter.as<double>() + r0);
        ++launches;
        NRB_CUDA(h, cudaGetLastError());
        if (mom) {
            accumulate_moments_f64_kernel<<<(unsigned)((G + 255) / 256), 256, 0, h->stream>>>(
                h->ws_sums.as<double>(), n, G, mom->observed ? h->wm_observed.as<double>() : nullptr, h->wm_sum.as<double>(),
                h->wm_sumsq.as<double>(), h->wm_nge.as<long long>());
            ++launches;
            NRB_CUDA(h, cudaGetLastError());
        }
        if (feature_sums)
            NRB_CUDA(h, cudaMemcpyAsync(feature_sums + r0 * G, h->ws_sums.p, (size_t)n * G * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    NRB_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    if (iter_sums && R > 0) NRB_CUDA(h, cudaMemcpyAsync(iter_sums, h->ws_iter.p, R * 8, cudaMemcpyDeviceToHost, h->stream));
    if (mom) {
        if (mom->sum) NRB_CUDA(h, cudaMemcpyAsync(mom->sum, h->wm_sum.p, (size_t)G * 8, cudaMemcpyDeviceToHost, h->stream));
        if (mom->sumsq) NRB_CUDA(h, cudaMemcpyAsync(mom->sumsq, h->wm_sumsq.p, (size_t)G * 8, cudaMemcpyDeviceToHost, h->stream));
        if (mom->n_ge) NRB_CUDA(h, cudaMemcpyAsync(mom->n_ge, h->wm_nge.p, (size_t)G * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    if ((rc = sync_stream(h))) return rc;
    if (stats) {
        float k_ms = 0, t_ms = 0;
        cudaEventElapsedTime(&k_ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&t_ms, h->ev[0], h->ev[2]);
        std::memset(stats, 0, sizeof *stats);
        stats->kernel_ms = k_ms;
        stats->total_ms = t_ms;
        stats->n_launches = launches;
        stats->n_windows = h->n_pairs;
        stats->d2h_bytes = (feature_sums ? R * G * 8 : 0) + (iter_sums ? R * 8 : 0);
    }
    return NRB200_OK;
}

int nrb200_run_weighted(nrb200_handle *h, const nrb200_draws *draws, double *iter_sums, double *feature_sums, nrb200_stats *stats) {
    if (is_group(h)) return group_run_weighted(h, draws, iter_sums, feature_sums, stats);
    return run_weighted_core(h, draws, iter_sums, feature_sums, nullptr, stats);
}

int nrb200_run_weighted_moments(nrb200_handle *h, const nrb200_draws *draws, const double *observed, double *iter_sums,
                                double *feat_sum, double *feat_sumsq, int64_t *feat_n_ge_observed, nrb200_stats *stats) {
    if (is_group(h)) return group_run_weighted_moments(h, draws, observed, iter_sums, feat_sum, feat_sumsq, feat_n_ge_observed, stats);
    const WeightedMoments mom{observed, feat_sum, feat_sumsq, feat_n_ge_observed};
    return run_weighted_core(h, draws, iter_sums, nullptr, &mom, stats);
}

int nrb200_count_observed(nrb200_handle *h, int32_t *feature_counts, int64_t *total) {
    return nrb200_count_observed_ex(h, 0, feature_counts, total);
}

int nrb200_count_observed_ex(nrb200_handle *h, int32_t minoverlap, int32_t *feature_counts, int64_t *total) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_lead(h, [&](nrb200_handle *m) { return nrb200_count_observed_ex(m, minoverlap, feature_counts, total); });
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (minoverlap < 0) return fail(h, NRB200_ERR_INVALID, "'minoverlap' must be >= 0");
    if (minoverlap > 0 && h->query_widen > 0) return fail(h, NRB200_ERR_INVALID, "'maxgap' and 'minoverlap' cannot both be set");
    if (!h->have_y) return fail(h, NRB200_ERR_STATE, "nrb200_set_ranges() has not been called");
    if (h->query.n == 0) return fail(h, NRB200_ERR_STATE, "nrb200_set_query() has not been called");
    {
        int rc0 = range_set_to_device(h, h->query);
        if (rc0) return rc0;
    }
    const RangeSet &q = h->query;
    DevMem &counts = h->obs_counts, &tot = h->obs_total;  // handle-owned: no cudaMalloc / cudaFree per call
    NRB_CUDA(h, counts.reserve(q.n * 4));
    NRB_CUDA(h, tot.reserve(8));
    NRB_CUDA(h, cudaMemsetAsync(counts.p, 0, q.n * 4, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(tot.p, 0, 8, h->stream));
    const RangeSet &y = h->y;
    RangesView yv{};
    yv.start = y.start.as<int32_t>();
    yv.end = y.end.as<int32_t>();
    yv.strand = y.stranded ? y.strand.as<int8_t>() : nullptr;
    const int threads = 256;
    const unsigned grid = (unsigned)((y.n + threads - 1) / threads);
    if (y.n > 0) {
        if (y.stranded && q.stranded)
            count_observed_kernel<true><<<grid, threads, 0, h->stream>>>(yv, y.n, y.seqid.as<int32_t>(), q.start.as<int32_t>(),
                q.end.as<int32_t>(), q.pmax.as<int32_t>(), q.strand.as<int8_t>(), q.src.as<int32_t>(), q.seq_off.as<int32_t>(),
                minoverlap, counts.as<int32_t>(), tot.as<unsigned long long>());
        else
            count_observed_kernel<false><<<grid, threads, 0, h->stream>>>(yv, y.n, y.seqid.as<int32_t>(), q.start.as<int32_t>(),
                q.end.as<int32_t>(), q.pmax.as<int32_t>(), nullptr, q.src.as<int32_t>(), q.seq_off.as<int32_t>(),
                minoverlap, counts.as<int32_t>(), tot.as<unsigned long long>());
        NRB_CUDA(h, cudaGetLastError());
    }
    unsigned long long t = 0;
    if (feature_counts) NRB_CUDA(h, cudaMemcpyAsync(feature_counts, counts.p, q.n * 4, cudaMemcpyDeviceToHost, h->stream));
    NRB_CUDA(h, cudaMemcpyAsync(&t, tot.p, 8, cudaMemcpyDeviceToHost, h->stream));
    if (int rc_sync = sync_stream(h)) return rc_sync;
    if (total) *total = (int64_t)t;
    return NRB200_OK;
}

int nrb200_run_materialize(nrb200_handle *h, const nrb200_draws *draws, int64_t *iter_offsets, nrb200_stats *stats) {
    if (!h || !draws) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_lead(h, [&](nrb200_handle *m) { return nrb200_run_materialize(m, draws, iter_offsets, stats); });
    NRB_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = prepare(h))) return rc;
    if ((rc = ensure_pmax_table(h))) return rc;  // the fill kernel locates the hit windows
    NRB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if ((rc = stage_draws(h, draws))) return rc;
    const int64_t R = draws->n_iter, B = h->plan.n_blocks(), items = R * B;
    NRB_CUDA(h, h->d_iter_nranges.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->item_rows.reserve(std::max<int64_t>(items, 1) * 4));
    NRB_CUDA(h, h->within_off.reserve(std::max<int64_t>(items, 1) * 4));
    NRB_CUDA(h, h->iter_off_dev.reserve((R + 1) * 8));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_nranges.p, 0, R * 8, h->stream));
    if ((rc = hits_reset(h))) return rc;  // partial hit sums, then the fill-overflow flag
    BootParams P = make_params(h, draws);
    P.n_query = 0;
    P.iter_nranges = h->d_iter_nranges.as<unsigned long long>();
    P.iter_totals = nullptr;
    P.item_rows = h->item_rows.as<int32_t>();
    P.n_hits = h->d_n_hits.as<unsigned long long>();
    NRB_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    int count_launches = 0;
    const bool rank_rows = rank_path_ok(h);
    if (rank_rows && items > 0) {  // the rows kernel leaves each block's draw and hit window for the fill kernel
        NRB_CUDA(h, h->item_rec.reserve((size_t)items * sizeof(int4)));
        P.item_rec = h->item_rec.as<int4>();
    }
    for (int64_t r0 = 0; r0 < R; r0 += kMaxBatch) {
        BootParams Q = slice_params(P, r0, std::min(kMaxBatch, R - r0));
        count_launches += launch_count(h, Q, false);
        NRB_CUDA(h, cudaGetLastError());
    }
    if (R > 0)
        scan_iteration_kernel<<<(unsigned)R, 256, 0, h->stream>>>(h->item_rows.as<int32_t>(), B, h->within_off.as<int32_t>());
    NRB_CUDA(h, cudaGetLastError());
    std::vector<int64_t> per_iter(R), offs(R + 1, 0);
    if ((rc = fetch_u64_as_i64(h, h->d_iter_nranges, R, per_iter.data()))) return rc;
    if (int rc_sync = sync_stream(h)) return rc_sync;
    for (int64_t r = 0; r < R; ++r) {
        // positions inside an iteration are int32 (within_off): an iteration of 2^31 rows or more cannot be laid out
        if (per_iter[r] > (int64_t)INT32_MAX) return fail(h, NRB200_ERR_UNSUPPORTED, "an iteration holds more than 2^31 - 1 rows");
        offs[r + 1] = offs[r] + per_iter[r];
    }
    const int64_t rows = offs[R];
    if ((rc = upload(h, h->iter_off_dev, offs.data(), (size_t)R + 1))) return rc;
    for (DevMem *m : {&h->row_seq, &h->row_start, &h->row_end, &h->row_src, &h->row_block})
        NRB_CUDA(h, m->reserve(std::max<int64_t>(rows, 1) * 4));
    int fill_launches = 0;
    if (items > 0) {
        const bool st = h->y.stranded && h->excl.n && active_excl(h).stranded;
        const int excl = excl_kind(h);
        const int64_t kFillBatch = 32768;  // gridDim.y
        for (int64_t r0 = 0; r0 < R; r0 += kFillBatch, ++fill_launches) {
            BootParams Q = slice_params(P, r0, std::min(kFillBatch, R - r0));
            const RowsOut O{h->row_seq.as<int32_t>(), h->row_start.as<int32_t>(), h->row_end.as<int32_t>(), h->row_src.as<int32_t>(),
                            h->row_block.as<int32_t>(), h->iter_off_dev.as<int64_t>() + r0, h->within_off.as<int32_t>() + r0 * B,
                            h->item_rows.as<int32_t>() + r0 * B, reinterpret_cast<int *>(h->d_n_hits.as<unsigned long long>() + kHitSlots)};
            const dim3 grid((unsigned)std::min<int64_t>((B + 7) / 8, 65535), (unsigned)Q.n_iter);
            if (excl == kExclTrim) boot_fill_trim_kernel<<<grid, 256, 0, h->stream>>>(Q, O);
            else if (st) boot_fill_kernel<true, kExclDrop><<<grid, 256, 0, h->stream>>>(Q, O);
            else if (excl == kExclDrop) boot_fill_kernel<false, kExclDrop><<<grid, 256, 0, h->stream>>>(Q, O);
            else boot_fill_kernel<false, kExclNone><<<grid, 256, 0, h->stream>>>(Q, O);
            NRB_CUDA(h, cudaGetLastError());
        }
    }
    NRB_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    unsigned long long hit_words[kHitSlots + 1] = {};
    if ((rc = hits_fetch(h, hit_words))) return rc;
    if (int rc_sync = sync_stream(h)) return rc_sync;
    if (hit_words[kHitSlots] != 0) {
        h->n_rows = 0;
        return fail(h, NRB200_ERR_STATE, "internal error: the count and fill passes of the table disagree (no row was written out of place)");
    }
    h->n_rows = rows;
    if (iter_offsets) std::memcpy(iter_offsets, offs.data(), (size_t)(R + 1) * 8);
    if (stats) {
        float k_ms = 0, t_ms = 0;
        cudaEventElapsedTime(&k_ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&t_ms, h->ev[0], h->ev[2]);
        const unsigned long long hits = hits_total(hit_words);
        std::memset(stats, 0, sizeof *stats);
        stats->kernel_ms = k_ms;
        stats->total_ms = t_ms;
        stats->n_launches = items > 0 ? count_launches + 1 + fill_launches : 0;
        stats->n_hits = (int64_t)hits;
    }
    return NRB200_OK;
}

int nrb200_fetch_rows(nrb200_handle *h, int64_t row0, int64_t n, int32_t *seqid, int32_t *start, int32_t *end,
                      int32_t *src_idx, int32_t *block) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_lead(h, [&](nrb200_handle *m) { return nrb200_fetch_rows(m, row0, n, seqid, start, end, src_idx, block); });
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (row0 < 0 || n < 0 || row0 + n > h->n_rows) return fail(h, NRB200_ERR_INVALID, "fetch_rows: range outside the materialised table");
    struct { int32_t *dst; DevMem *src; } cols[] = {{seqid, &h->row_seq}, {start, &h->row_start}, {end, &h->row_end},
                                                     {src_idx, &h->row_src}, {block, &h->row_block}};
    if (int rc_sync = sync_stream(h)) return rc_sync;
    for (auto &c : cols)
        if (c.dst && n)
            if (int rc = download(h, c.dst, c.src->as<int32_t>() + row0, (size_t)n * 4)) return rc;
    return NRB200_OK;
}

int nrb200_fetch_dst_tiles(nrb200_handle *h, int64_t r0, int64_t n, int32_t *dst_tile) {
    if (!h) return NRB200_ERR_INVALID;
    if (is_group(h)) return group_lead(h, [&](nrb200_handle *m) { return nrb200_fetch_dst_tiles(m, r0, n, dst_tile); });
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (!h->have_plan || !h->plan.permute()) return fail(h, NRB200_ERR_STATE, "fetch_dst_tiles: the last plan is not a permute variant");
    const int64_t B = h->plan.n_blocks();
    if (r0 < 0 || n < 0 || (size_t)(r0 + n) * B * 4 > h->d_dst_tile.cap || (size_t)(r0 + n) * B > (size_t)h->dst_tile_items)
        return fail(h, NRB200_ERR_INVALID, "fetch_dst_tiles: iterations outside the last run");
    if (n == 0 || !dst_tile) return NRB200_OK;
    NRB_CUDA(h, cudaMemcpyAsync(dst_tile, h->d_dst_tile.as<int32_t>() + r0 * B, (size_t)n * B * 4, cudaMemcpyDeviceToHost, h->stream));
    return sync_stream(h);
}

int nrb200_measure_l2_sector_rate(nrb200_handle *h, int64_t buffer_bytes, double *gbytes_per_s) {
    if (!h || !gbytes_per_s) return NRB200_ERR_INVALID;
    if (is_group(h)) h = h->members[0];
    NRB_CUDA(h, cudaSetDevice(h->device));
    size_t bytes = (size_t)32 << 20;
    if (buffer_bytes > 0) {
        bytes = 1 << 20;
        while (bytes * 2 <= (size_t)buffer_bytes && bytes < ((size_t)1 << 30)) bytes *= 2;
    }
    DevMem buf, sink;
    NRB_CUDA(h, buf.reserve(bytes));
    NRB_CUDA(h, sink.reserve(8));
    NRB_CUDA(h, cudaMemsetAsync(buf.p, 1, bytes, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(sink.p, 0, 8, h->stream));
    const uint32_t n_mask = (uint32_t)(bytes / 32 - 1);
    const int rounds = 64;
    const unsigned grid = (unsigned)h->sm_count * 8 * 4;  // 8 resident CTAs of 256 threads per SM, four waves
    cudaEvent_t a, b;
    NRB_CUDA(h, cudaEventCreate(&a));
    NRB_CUDA(h, cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // the first pass also pulls the buffer into the L2
        cudaEventRecord(a, h->stream);
        l2_sector_rate_kernel<<<grid, 256, 0, h->stream>>>(buf.as<uint2>(), n_mask, rounds, sink.as<unsigned long long>());
        cudaEventRecord(b, h->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    NRB_CUDA(h, cudaGetLastError());
    const double sectors = (double)grid * 256.0 * rounds * kCalibLoads;
    *gbytes_per_s = sectors * 32.0 / (best * 1e-3) / 1e9;
    return NRB200_OK;
}

void nrb200_expand_runs(const int64_t *offsets, int64_t n_runs, int32_t first, int32_t *out) {
    if (!offsets || !out || n_runs <= 0) return;
    const int64_t base = offsets[0], total = offsets[n_runs] - base;
    const int64_t piece = (int64_t)1 << 18;  // elements per task
    const int64_t pieces = (total + piece - 1) / piece;
#pragma omp parallel for schedule(static) if (total >= ((int64_t)1 << 20))
    for (int64_t p = 0; p < pieces; ++p) {
        const int64_t lo = base + p * piece, hi = std::min(lo + piece, base + total);
        int64_t r = std::upper_bound(offsets, offsets + n_runs + 1, lo) - offsets - 1;  // run that holds element lo
        for (int64_t i = lo; i < hi;) {
            while (offsets[r + 1] <= i) ++r;
            const int64_t stop = std::min(hi, offsets[r + 1]);
            std::fill(out + (i - base), out + (stop - base), (int32_t)(first + r));
            i = stop;
        }
    }
}

}  // extern "C"
