// libnrb200: loading y / query / exclude, canonical order, rank tables, weight running sums.
// (textually included by engine.cu after handle.h)
namespace {

// Canonicalise one set of ranges on the host and mirror it to the device.
int range_set_to_device(nrb200_handle *h, RangeSet &rs) {
    if (rs.on_device) return NRB200_OK;
    int rc;
    const size_t n = (size_t)rs.n;
    // 16 elements of slack behind every column: boot_stream_kernel copies 16-element-aligned slices of them
    for (DevMem *m : {&rs.start, &rs.end, &rs.pmax, &rs.src}) NRB_CUDA(h, m->reserve((n + 16) * 4));
    if (rs.stranded) NRB_CUDA(h, rs.strand.reserve(n + 16));
    if ((rc = upload(h, rs.start, rs.h_start.data(), n))) return rc;
    if ((rc = upload(h, rs.end, rs.h_end.data(), n))) return rc;
    if ((rc = upload(h, rs.pmax, rs.h_pmax.data(), n))) return rc;
    if ((rc = upload(h, rs.src, rs.h_src.data(), n))) return rc;
    if ((rc = upload(h, rs.seq_off, rs.h_seq_off.data(), rs.h_seq_off.size()))) return rc;
    if (rs.stranded && (rc = upload(h, rs.strand, rs.h_strand.data(), n))) return rc;
    rs.on_device = true;
    return NRB200_OK;
}

int load_range_set(nrb200_handle *h, RangeSet &rs, const char *what, int64_t n, const int32_t *seqid,
                   const int32_t *start, const int32_t *end, const int8_t *strand, bool keep_seqid_dev,
                   bool defer_device = false, int32_t widen = 0) {
    if (n < 0 || n > (int64_t)INT32_MAX - 64) return fail(h, NRB200_ERR_INVALID, std::string(what) + ": length out of range");
    if (n > 0 && (!seqid || !start || !end)) return fail(h, NRB200_ERR_INVALID, std::string(what) + ": null coordinate pointer");
    const int C = h->n_seq;
    bool sorted = true, stranded = false;
    int32_t wmax = 0, wmin = INT32_MAX;
    for (int64_t i = 0; i < n; ++i) {
        if (seqid[i] < 0 || seqid[i] >= C) return fail(h, NRB200_ERR_INVALID, std::string(what) + ": seqname not in seqlevels(y)");
        if (start[i] < 1) return fail(h, NRB200_ERR_UNSUPPORTED, std::string(what) + ": ranges must start at position >= 1");
        if (end[i] < start[i]) return fail(h, NRB200_ERR_UNSUPPORTED, std::string(what) + ": zero-width ranges are not supported");
        if (end[i] >= (1 << 30)) return fail(h, NRB200_ERR_UNSUPPORTED, std::string(what) + ": coordinates must be below 2^30");
        if (strand) {
            if (strand[i] < 0 || strand[i] > 2) return fail(h, NRB200_ERR_INVALID, std::string(what) + ": strand codes are 0 '*', 1 '+', 2 '-'");
            stranded |= strand[i] != 0;
        }
        wmax = std::max(wmax, end[i] - start[i] + 1);
        wmin = std::min(wmin, end[i] - start[i] + 1);
        if (i > 0) {
            const int64_t j = i - 1;
            if (seqid[i] < seqid[j] || (seqid[i] == seqid[j] && (start[i] < start[j] || (start[i] == start[j] && end[i] < end[j]))))
                sorted = false;
        }
    }
    rs.n = n;
    rs.stranded = stranded;
    rs.wmax = wmax;
    rs.wmin = n ? wmin : 0;
    rs.h_src.resize(n);
    std::iota(rs.h_src.begin(), rs.h_src.end(), 0);
    if (!sorted)
        std::sort(rs.h_src.begin(), rs.h_src.end(), [&](int32_t a, int32_t b) {
            if (seqid[a] != seqid[b]) return seqid[a] < seqid[b];
            if (start[a] != start[b]) return start[a] < start[b];
            if (end[a] != end[b]) return end[a] < end[b];
            return a < b;
        });
    rs.h_start.resize(n);
    rs.h_end.resize(n);
    rs.h_pmax.resize(n);
    rs.h_strand.clear();
    if (stranded) rs.h_strand.resize(n);
    rs.h_seq_off.assign(C + 1, 0);
    std::vector<int32_t> h_seqid(keep_seqid_dev ? n : 0);
    for (int64_t k = 0; k < n; ++k) {
        const int32_t i = rs.h_src[k];
        rs.h_start[k] = start[i] - widen;  // widen > 0: query with maxgap (see nrb200_set_query_ex)
        rs.h_end[k] = end[i] + widen;
        if (stranded) rs.h_strand[k] = strand[i];
        if (keep_seqid_dev) h_seqid[k] = seqid[i];
        rs.h_seq_off[seqid[i] + 1]++;
    }
    for (int c = 0; c < C; ++c) rs.h_seq_off[c + 1] += rs.h_seq_off[c];
    for (int c = 0; c < C; ++c) {
        int32_t m = INT32_MIN;
        for (int32_t k = rs.h_seq_off[c]; k < rs.h_seq_off[c + 1]; ++k) {
            m = std::max(m, rs.h_end[k]);
            rs.h_pmax[k] = m;
        }
    }
    rs.sorted_input = sorted;
    rs.max_end.assign(C, INT32_MIN);
    for (int c = 0; c < C; ++c)
        if (rs.h_seq_off[c + 1] > rs.h_seq_off[c]) rs.max_end[c] = rs.h_pmax[rs.h_seq_off[c + 1] - 1];
    rs.on_device = false;
    if (defer_device) return NRB200_OK;
    int rc;
    if ((rc = range_set_to_device(h, rs))) return rc;
    if (keep_seqid_dev && (rc = upload(h, rs.seqid, h_seqid.data(), n))) return rc;
    return NRB200_OK;  // uploads were staged or made with the synchronising copy: the host vectors may go
}

// y on the fast path: the caller's arrays go to the device as they are and ONE kernel does the
// argument checks and finds out whether they are already in canonical order (the usual case:
// the reference insists on sort(y) for the across-chromosome samplers, R/bootUnsegmented.R:31).
// Unsorted input is sorted on the device (two stable radix sorts); only INVALID input falls back to the host
// path above, which words the error.  Two halves (nrb200_set_ranges_begin / _end): the first only queues the
// copies, the inspection and the copy of its verdict into pinned memory, so the caller's host-side work on the
// query, the exclude set and the plan runs while y is on the wire.
int load_y_begin(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start, const int32_t *end,
                 const int8_t *strand) {
    RangeSet &y = h->y;
    const int C = h->n_seq;
    h->pend.n = n;
    h->pend.seqid = seqid;
    h->pend.start = start;
    h->pend.end = end;
    h->pend.strand = strand;
    h->pend.on_device = false;
    if (n <= 0 || n > (int64_t)INT32_MAX - 64 || !seqid || !start || !end) return NRB200_OK;  // host path words the errors
    int rc;
    // scratch: InspectOut (5 words, padded to 8) | seq_count[C] | seq_max_end[C]
    const size_t kHead = 8;
    const size_t words = kHead + 2 * (size_t)C;
    NRB_CUDA(h, h->inspect_dev.reserve(words * 4));
    if (h->inspect_host_words < words) {
        if (h->inspect_host) cudaFreeHost(h->inspect_host);
        h->inspect_host = nullptr;
        h->inspect_host_words = 0;
        NRB_CUDA(h, cudaMallocHost((void **)&h->inspect_host, 2 * words * 4));  // verdict | initial values
        h->inspect_host_words = words;
    }
    h->timeline.mark("begin", h->stream);
    int32_t *init = h->inspect_host + words;
    std::fill(init, init + words, 0);
    init[4] = INT32_MAX;  // wmin
    for (int c = 0; c < C; ++c) init[kHead + C + c] = INT32_MIN;
    NRB_CUDA(h, cudaMemcpyAsync(h->inspect_dev.p, init, words * 4, cudaMemcpyHostToDevice, h->stream));
    int32_t *base = h->inspect_dev.as<int32_t>();
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
    const HostColumn cols[4] = {{&y.start, start, (size_t)n * 4}, {&y.end, end, (size_t)n * 4}, {&y.seqid, seqid, (size_t)n * 4},
                                {&y.strand, strand, (size_t)n}};
    if (h->y_preloaded) {  // member of a device group: the columns are already here (group_scatter_y; this stream waits for them)
        for (int c = 0; c < (strand ? 4 : 3); ++c) NRB_CUDA(h, cols[c].dst->reserve(cols[c].bytes));
    } else if ((rc = upload_columns(h, cols, strand ? 4 : 3))) {
        return rc;
    }
    inspect_ranges_kernel<<<grid, 256, 0, h->stream>>>(y.seqid.as<int32_t>(), y.start.as<int32_t>(), y.end.as<int32_t>(),
                                                       strand ? y.strand.as<int8_t>() : nullptr, n, C,
                                                       reinterpret_cast<InspectOut *>(base), reinterpret_cast<unsigned *>(base + kHead),
                                                       base + kHead + C);
    NRB_CUDA(h, cudaGetLastError());
    NRB_CUDA(h, cudaMemcpyAsync(h->inspect_host, base, words * 4, cudaMemcpyDeviceToHost, h->stream));
    h->timeline.mark("verdict copied", h->stream);
    h->pend.on_device = true;
    return NRB200_OK;
}

// Second half: waits for the verdict.  Returns 1 for "fall back to the host path".
int load_y_end(nrb200_handle *h) {
    if (!h->pend.on_device) return 1;
    RangeSet &y = h->y;
    const int C = h->n_seq;
    const int64_t n = h->pend.n;
    const int8_t *strand = h->pend.strand;
    const size_t kHead = 8;
    int rc;
    Trace tr("load_y_end");
    if ((rc = sync_stream(h))) return rc;  // the caller's arrays are no longer read after this point
    tr.lap("sync");
    const int32_t *res = h->inspect_host;
    if (res[0] != 0) return 1;  // invalid: the host path names the problem
    y.n = n;
    y.sorted_input = res[1] == 0;
    y.stranded = res[2] != 0;
    if (!y.sorted_input) {
        // Canonical order on the device: stable LSD sort, first on end, then on (seqid << 32 | start); the
        // sorted payload is src, the canonical -> caller's-order map; the columns are gathered through it.
        const unsigned grid = (unsigned)((n + 255) / 256);
        NRB_CUDA(h, h->sort_keys_a.reserve((size_t)n * 8));
        NRB_CUDA(h, h->sort_keys_b.reserve((size_t)n * 8));
        for (DevMem *m : {&h->r_idx, &h->r_order, &h->r_keys_a, &h->r_keys_b, &y.src}) NRB_CUDA(h, m->reserve((size_t)n * 4));
        canon_keys_end_kernel<<<grid, 256, 0, h->stream>>>(y.end.as<int32_t>(), n, h->r_keys_a.as<unsigned>(), h->r_idx.as<int32_t>());
        size_t tmp = 0;
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->r_keys_a.as<unsigned>(), h->r_keys_b.as<unsigned>(),
                                                    h->r_idx.as<int32_t>(), h->r_order.as<int32_t>(), n, 0, 31, h->stream));
        NRB_CUDA(h, h->sort_tmp.reserve(tmp));
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp, h->r_keys_a.as<unsigned>(), h->r_keys_b.as<unsigned>(),
                                                    h->r_idx.as<int32_t>(), h->r_order.as<int32_t>(), n, 0, 31, h->stream));
        canon_keys_seq_start_kernel<<<grid, 256, 0, h->stream>>>(h->r_order.as<int32_t>(), y.seqid.as<int32_t>(), y.start.as<int32_t>(), n,
                                                                 h->sort_keys_a.as<unsigned long long>());
        int seq_bits = 1;
        while ((1 << seq_bits) < C) ++seq_bits;
        tmp = 0;
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->sort_keys_a.as<unsigned long long>(), h->sort_keys_b.as<unsigned long long>(),
                                                    h->r_order.as<int32_t>(), y.src.as<int32_t>(), n, 0, 32 + seq_bits, h->stream));
        NRB_CUDA(h, h->sort_tmp.reserve(tmp));
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp, h->sort_keys_a.as<unsigned long long>(), h->sort_keys_b.as<unsigned long long>(),
                                                    h->r_order.as<int32_t>(), y.src.as<int32_t>(), n, 0, 32 + seq_bits, h->stream));
        // gather the columns: the caller-order copies move aside, the canonical ones take their place
        y.seqid.swap(h->raw_seqid);
        y.start.swap(h->raw_start);
        y.end.swap(h->raw_end);
        for (DevMem *m : {&y.seqid, &y.start, &y.end}) NRB_CUDA(h, m->reserve((size_t)n * 4));
        gather_ranges_kernel<<<grid, 256, 0, h->stream>>>(y.src.as<int32_t>(), h->raw_seqid.as<int32_t>(), h->raw_start.as<int32_t>(),
                                                          h->raw_end.as<int32_t>(), n, y.seqid.as<int32_t>(), y.start.as<int32_t>(),
                                                          y.end.as<int32_t>());
        if (strand) {
            y.strand.swap(h->raw_strand);
            NRB_CUDA(h, y.strand.reserve((size_t)n));
            gather_i8_kernel<<<grid, 256, 0, h->stream>>>(y.src.as<int32_t>(), h->raw_strand.as<int8_t>(), n, y.strand.as<int8_t>());
        }
        NRB_CUDA(h, cudaGetLastError());
    }
    y.wmax = res[3];
    y.wmin = res[4];
    y.h_seq_off.assign(C + 1, 0);
    y.max_end.assign(C, INT32_MIN);
    for (int c = 0; c < C; ++c) {
        y.h_seq_off[c + 1] = y.h_seq_off[c] + res[kHead + c];
        y.max_end[c] = res[kHead + C + c];
    }
    if ((rc = upload(h, y.seq_off, y.h_seq_off.data(), (size_t)C + 1))) return rc;
    NRB_CUDA(h, y.pmax.reserve(n * 4));  // src stays unset: identity (kernels take nullptr for it)
    // running max of end within each sequence
    size_t tmp_bytes = 0;
    NRB_CUDA(h, cub::DeviceScan::InclusiveScanByKey(nullptr, tmp_bytes, y.seqid.as<int32_t>(), y.end.as<int32_t>(),
                                                    y.pmax.as<int32_t>(), MaxOp(), n, cub::Equality(), h->stream));
    NRB_CUDA(h, h->sort_tmp.reserve(tmp_bytes));
    NRB_CUDA(h, cub::DeviceScan::InclusiveScanByKey(h->sort_tmp.p, tmp_bytes, y.seqid.as<int32_t>(), y.end.as<int32_t>(),
                                                    y.pmax.as<int32_t>(), MaxOp(), n, cub::Equality(), h->stream));
    return NRB200_OK;  // the index build that follows stays queued: the host moves on to the plan and the lists
}

int build_rank_tables(nrb200_handle *h) {
    const int C = h->n_seq;
    RangeSet &y = h->y;
    // 2-4 buckets per range (about 1 above 4M ranges, to keep the tables inside the L2): a rank
    // query is one table probe plus at most three values read together; clustered ranges seldom
    // put more than three values into a bucket.
    int64_t genome = 0;
    std::vector<int32_t> max_pos(C);
    for (int c = 0; c < C; ++c) {
        const int64_t m = std::max<int64_t>(h->seqlen[c], y.max_end[c]);
        max_pos[c] = (int32_t)std::min<int64_t>(m, INT32_MAX - 1);
        genome += max_pos[c];
    }
    int shift = 0;
    const int64_t want = std::max<int64_t>(std::min<int64_t>((int64_t)(h->tune_table_mult * y.n), std::max<int64_t>(y.n, h->tune_table_cap)), 1024);
    while (shift < 30 && (genome >> shift) > want) ++shift;
    h->rank_shift = shift;
    h->rank_off.assign(C + 1, 0);
    for (int c = 0; c < C; ++c) h->rank_off[c + 1] = h->rank_off[c] + ((int64_t)max_pos[c] >> shift) + 3;  // +1: probes read entry u + 1
    const int64_t entries = h->rank_off[C];
    // classes of y the rank path counts separately: strand class x (narrower than the query's minoverlap?)
    h->n_width = h->query_minoverlap > 0 ? 2 : 1;
    h->n_cls = (y.stranded ? 3 : 1) * h->n_width;
    h->index_minoverlap = h->query_minoverlap;
    h->weights_dirty = true;  // the class-major order of the weights follows the index
    if (entries * h->n_cls > (int64_t)INT32_MAX) return fail(h, NRB200_ERR_UNSUPPORTED, "rank tables exceed 2^31 entries");
    for (int c = 0; c < C; ++c)  // packed table entries hold a per-sequence index in 29 bits
        if (y.h_seq_off[c + 1] - y.h_seq_off[c] >= (1 << 29)) return fail(h, NRB200_ERR_UNSUPPORTED, "more than 2^29 ranges on one sequence");
    int rc;
    if ((rc = upload(h, h->rank_off_dev, h->rank_off.data(), (size_t)C + 1))) return rc;
    if ((rc = upload(h, h->max_pos_dev, max_pos.data(), (size_t)C))) return rc;
    std::vector<int32_t> seq_info(4 * (size_t)C);
    for (int c = 0; c < C; ++c) {
        seq_info[4 * c + 0] = y.h_seq_off[c];
        seq_info[4 * c + 1] = y.h_seq_off[c + 1];
        seq_info[4 * c + 2] = max_pos[c];
        seq_info[4 * c + 3] = (int32_t)h->rank_off[c];
    }
    if ((rc = upload(h, h->seq_info_dev, seq_info.data(), seq_info.size()))) return rc;
    NRB_CUDA(h, h->rank_start.reserve(entries * sizeof(RankEntry)));
    NRB_CUDA(h, h->rank_pmax.reserve(entries * sizeof(RankEntry)));
    NRB_CUDA(h, h->end_sorted.reserve(std::max<int64_t>(y.n, 1) * sizeof(int32_t)));
    const int threads = 256;
    const unsigned n_grid = (unsigned)((std::max<int64_t>(y.n, 1) + threads - 1) / threads);
    const unsigned t_grid = rank_table_grid(entries);
    // canonical start table: locate_block() of the streaming kernels, and the rank path of an unstranded y
    build_rank_table_kernel<<<t_grid, threads, 0, h->stream>>>(y.start.as<int32_t>(), y.seq_off.as<int32_t>(),
                                                               h->rank_off_dev.as<int64_t>(), C, shift, h->rank_start.as<RankEntry>());
    h->have_pmax_table = false;  // only the streaming kernels need it: built on first use
    h->have_se = false;          // (the same for the interleaved copy of y)
    NRB_CUDA(h, h->sort_keys_a.reserve(std::max<int64_t>(y.n, 1) * 8));
    NRB_CUDA(h, h->sort_keys_b.reserve(std::max<int64_t>(y.n, 1) * 8));
    auto sort_u64 = [&](int bits) -> int {  // sort_keys_a -> sort_keys_b
        size_t tmp_bytes = 0;
        NRB_CUDA(h, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, h->sort_keys_a.as<unsigned long long>(),
                                                   h->sort_keys_b.as<unsigned long long>(), y.n, 0, bits, h->stream));
        NRB_CUDA(h, h->sort_tmp.reserve(tmp_bytes));
        NRB_CUDA(h, cub::DeviceRadixSort::SortKeys(h->sort_tmp.p, tmp_bytes, h->sort_keys_a.as<unsigned long long>(),
                                                   h->sort_keys_b.as<unsigned long long>(), y.n, 0, bits, h->stream));
        return NRB200_OK;
    };
    // Equally wide ranges (points, fixed-width windows): the end order is the start order, so the start
    // array and its table answer #{end < b} as #{start < b - (w - 1)}: no sorted-ends copy, half the index.
    h->end_shift = (y.n > 0 && y.wmin == y.wmax) ? y.wmax - 1 : 0;
    h->ends_alias_starts = y.n > 0 && y.wmin == y.wmax;
    if (h->n_cls == 1) {
        if (h->ends_alias_starts) return NRB200_OK;
        // ends in ascending order within each sequence: radix sort of (seqid << 32 | end)
        NRB_CUDA(h, h->rank_end.reserve(entries * sizeof(RankEntry)));
        // ranges that are short against their spacing (expected ranges per widest-range window <= 4): the ends are
        // nearly in order already and every range finds its rank with a short scan (end_rank_scatter_kernel);
        // otherwise a radix sort of (seqid << 32 | end)
        bool by_rank = y.n > 0 && h->tune_end_rank;
        for (int c = 0; c < C && by_rank; ++c)  // judged per sequence: a crowded one would make its ranges scan long windows
            by_rank = (double)y.wmax * (double)(y.h_seq_off[c + 1] - y.h_seq_off[c]) <= 4.0 * (double)std::max<int64_t>(max_pos[c], 1);
        if (by_rank) {
            end_rank_scatter_kernel<<<n_grid, threads, 0, h->stream>>>(y.seqid.as<int32_t>(), y.start.as<int32_t>(), y.end.as<int32_t>(),
                                                                       h->seq_info_dev.as<int4>(), h->rank_start.as<RankEntry>(), shift, y.n,
                                                                       y.wmax, h->end_sorted.as<int32_t>());
        } else if (y.n > 0) {
            int seq_bits = 1;
            while ((1 << seq_bits) < C) ++seq_bits;
            pack_end_keys_kernel<<<n_grid, threads, 0, h->stream>>>(y.seqid.as<int32_t>(), y.end.as<int32_t>(), y.n,
                                                                    h->sort_keys_a.as<unsigned long long>());
            if ((rc = sort_u64(32 + seq_bits))) return rc;
            unpack_end_keys_kernel<<<n_grid, threads, 0, h->stream>>>(h->sort_keys_b.as<unsigned long long>(), y.n, h->end_sorted.as<int32_t>());
        }
        build_rank_table_kernel<<<t_grid, threads, 0, h->stream>>>(h->end_sorted.as<int32_t>(), y.seq_off.as<int32_t>(),
                                                                   h->rank_off_dev.as<int64_t>(), C, shift, h->rank_end.as<RankEntry>());
        NRB_CUDA(h, cudaGetLastError());
        return NRB200_OK;
    }
    // ---- several classes: a second copy grouped by class inside each sequence (virtual sequences)
    const int V = h->n_cls * C;
    std::vector<int64_t> roff(V + 1, 0);
    for (int v = 0; v < V; ++v) roff[v + 1] = roff[v] + ((int64_t)max_pos[v / h->n_cls] >> shift) + 3;
    const int64_t r_entries = roff[V];
    if ((rc = upload(h, h->rrank_off_dev, roff.data(), (size_t)V + 1))) return rc;
    for (DevMem *m : {&h->r_start, &h->r_end, &h->r_pmax, &h->r_order, &h->r_idx, &h->r_keys_a, &h->r_keys_b})
        NRB_CUDA(h, m->reserve(std::max<int64_t>(y.n, 1) * 4));
    NRB_CUDA(h, h->r_voff_dev.reserve(((size_t)V + 1) * 4));
    NRB_CUDA(h, h->rseq_info_dev.reserve((size_t)V * sizeof(int4)));
    NRB_CUDA(h, h->rrank_start.reserve(r_entries * sizeof(RankEntry)));
    NRB_CUDA(h, h->rank_end.reserve(r_entries * sizeof(RankEntry)));
    int v_bits = 1;
    while ((1 << v_bits) < V) ++v_bits;
    if (y.n > 0) {
        vseq_keys_kernel<<<n_grid, threads, 0, h->stream>>>(y.seqid.as<int32_t>(), y.stranded ? y.strand.as<int8_t>() : nullptr,
                                                            y.start.as<int32_t>(), y.end.as<int32_t>(), y.n, h->n_cls, h->n_width,
                                                            h->query_minoverlap, h->r_keys_a.as<unsigned>(), h->r_idx.as<int32_t>());
        size_t tmp_bytes = 0;  // stable: inside a virtual sequence the canonical (start, end) order survives
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->r_keys_a.as<unsigned>(), h->r_keys_b.as<unsigned>(),
                                                    h->r_idx.as<int32_t>(), h->r_order.as<int32_t>(), y.n, 0, v_bits, h->stream));
        NRB_CUDA(h, h->sort_tmp.reserve(tmp_bytes));
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp_bytes, h->r_keys_a.as<unsigned>(), h->r_keys_b.as<unsigned>(),
                                                    h->r_idx.as<int32_t>(), h->r_order.as<int32_t>(), y.n, 0, v_bits, h->stream));
        gather_by_order_kernel<<<n_grid, threads, 0, h->stream>>>(h->r_order.as<int32_t>(), y.start.as<int32_t>(), y.end.as<int32_t>(), y.n,
                                                                  h->r_start.as<int32_t>(), h->r_end.as<int32_t>());
        tmp_bytes = 0;
        NRB_CUDA(h, cub::DeviceScan::InclusiveScanByKey(nullptr, tmp_bytes, h->r_keys_b.as<unsigned>(), h->r_end.as<int32_t>(),
                                                        h->r_pmax.as<int32_t>(), MaxOp(), y.n, cub::Equality(), h->stream));
        NRB_CUDA(h, h->sort_tmp.reserve(tmp_bytes));
        NRB_CUDA(h, cub::DeviceScan::InclusiveScanByKey(h->sort_tmp.p, tmp_bytes, h->r_keys_b.as<unsigned>(), h->r_end.as<int32_t>(),
                                                        h->r_pmax.as<int32_t>(), MaxOp(), y.n, cub::Equality(), h->stream));
        if (!h->ends_alias_starts) {
            pack_vend_keys_kernel<<<n_grid, threads, 0, h->stream>>>(h->r_keys_b.as<unsigned>(), h->r_end.as<int32_t>(), y.n,
                                                                     h->sort_keys_a.as<unsigned long long>());
            if ((rc = sort_u64(32 + v_bits))) return rc;
            unpack_end_keys_kernel<<<n_grid, threads, 0, h->stream>>>(h->sort_keys_b.as<unsigned long long>(), y.n, h->end_sorted.as<int32_t>());
        }
    }
    vseq_info_kernel<<<(unsigned)((V + 1 + threads - 1) / threads), threads, 0, h->stream>>>(
        h->r_keys_b.as<unsigned>(), y.n, V, h->n_cls, h->max_pos_dev.as<int32_t>(), h->rrank_off_dev.as<int64_t>(), h->r_voff_dev.as<int32_t>(),
        h->rseq_info_dev.as<int4>());
    const unsigned rt_grid = rank_table_grid(r_entries);
    build_rank_table_kernel<<<rt_grid, threads, 0, h->stream>>>(h->r_start.as<int32_t>(), h->r_voff_dev.as<int32_t>(),
                                                                h->rrank_off_dev.as<int64_t>(), V, shift, h->rrank_start.as<RankEntry>());
    if (!h->ends_alias_starts)
        build_rank_table_kernel<<<rt_grid, threads, 0, h->stream>>>(h->end_sorted.as<int32_t>(), h->r_voff_dev.as<int32_t>(),
                                                                    h->rrank_off_dev.as<int64_t>(), V, shift, h->rank_end.as<RankEntry>());
    NRB_CUDA(h, cudaGetLastError());
    return NRB200_OK;
}


// y as interleaved {start, end} pairs: what boot_stream_kernel streams
int ensure_interleaved(nrb200_handle *h) {
    if (h->have_se) return NRB200_OK;
    const int64_t n = h->y.n;
    NRB_CUDA(h, h->y_se.reserve((size_t)(n + kSePad) * sizeof(int2)));
    interleave_ranges_kernel<<<(unsigned)((n + kSePad + 255) / 256), 256, 0, h->stream>>>(h->y.start.as<int32_t>(), h->y.end.as<int32_t>(), n,
                                                                                     h->y_se.as<int2>());
    NRB_CUDA(h, cudaGetLastError());
    h->have_se = true;
    return NRB200_OK;
}

// rank table over the running max of end: locate_block() of the streaming kernels
int ensure_pmax_table(nrb200_handle *h) {
    if (h->have_pmax_table) return NRB200_OK;
    const int64_t entries = h->rank_off[h->n_seq];
    build_rank_table_kernel<<<rank_table_grid(entries), 256, 0, h->stream>>>(
        h->y.pmax.as<int32_t>(), h->y.seq_off.as<int32_t>(), h->rank_off_dev.as<int64_t>(), h->n_seq, h->rank_shift,
        h->rank_pmax.as<RankEntry>());
    NRB_CUDA(h, cudaGetLastError());
    h->have_pmax_table = true;
    return NRB200_OK;
}

}  // namespace

// Weights in the rank path's orders and their running sums (see WeightView, kernels.cuh).
int build_weight_index(nrb200_handle *h) {
    if (!h->weights_dirty) return NRB200_OK;
    RangeSet &y = h->y;
    const int64_t n = y.n;
    if (n == 0) {
        h->weights_dirty = false;
        return NRB200_OK;
    }
    const unsigned grid = (unsigned)((n + 255) / 256);
    const bool classes = h->n_cls > 1;
    for (DevMem *m : {&h->w_r, &h->w_e, &h->w_ps, &h->w_pe}) NRB_CUDA(h, m->reserve((size_t)n * 8));
    const double *w_r = h->w_c.as<double>();
    if (classes) {
        gather_f64_kernel<<<grid, 256, 0, h->stream>>>(h->w_c.as<double>(), h->r_order.as<int32_t>(), n, h->w_r.as<double>());
        w_r = h->w_r.as<double>();
    }
    auto scan_by_key = [&](const unsigned *keys, const double *in, double *out) -> int {
        size_t tmp = 0;
        NRB_CUDA(h, cub::DeviceScan::InclusiveScanByKey(nullptr, tmp, keys, in, out, AddF64(), n, cub::Equality(), h->stream));
        NRB_CUDA(h, h->sort_tmp.reserve(tmp));
        NRB_CUDA(h, cub::DeviceScan::InclusiveScanByKey(h->sort_tmp.p, tmp, keys, in, out, AddF64(), n, cub::Equality(), h->stream));
        return NRB200_OK;
    };
    // virtual-sequence key of every position of the (class-major) start order
    const unsigned *vkeys = classes ? h->r_keys_b.as<unsigned>() : reinterpret_cast<const unsigned *>(y.seqid.as<int32_t>());
    int rc;
    if ((rc = scan_by_key(vkeys, w_r, h->w_ps.as<double>()))) return rc;
    if (!h->ends_alias_starts) {
        NRB_CUDA(h, h->sort_keys_a.reserve((size_t)n * 8));
        NRB_CUDA(h, h->sort_keys_b.reserve((size_t)n * 8));
        NRB_CUDA(h, h->w_perm_a.reserve((size_t)n * 4));
        NRB_CUDA(h, h->w_perm_b.reserve((size_t)n * 4));
        NRB_CUDA(h, h->w_keys_e.reserve((size_t)n * 4));
        const int32_t *rend = classes ? h->r_end.as<int32_t>() : y.end.as<int32_t>();
        pack_vend_pairs_kernel<<<grid, 256, 0, h->stream>>>(classes ? h->r_keys_b.as<unsigned>() : nullptr, y.seqid.as<int32_t>(), rend, n,
                                                            h->sort_keys_a.as<unsigned long long>(), h->w_perm_a.as<int32_t>());
        size_t tmp = 0;
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->sort_keys_a.as<unsigned long long>(), h->sort_keys_b.as<unsigned long long>(),
                                                    h->w_perm_a.as<int32_t>(), h->w_perm_b.as<int32_t>(), n, 0, 64, h->stream));
        NRB_CUDA(h, h->sort_tmp.reserve(tmp));
        NRB_CUDA(h, cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp, h->sort_keys_a.as<unsigned long long>(), h->sort_keys_b.as<unsigned long long>(),
                                                    h->w_perm_a.as<int32_t>(), h->w_perm_b.as<int32_t>(), n, 0, 64, h->stream));
        gather_f64_kernel<<<grid, 256, 0, h->stream>>>(w_r, h->w_perm_b.as<int32_t>(), n, h->w_e.as<double>());
        high32_kernel<<<grid, 256, 0, h->stream>>>(h->sort_keys_b.as<unsigned long long>(), n, h->w_keys_e.as<unsigned>());
        if ((rc = scan_by_key(h->w_keys_e.as<unsigned>(), h->w_e.as<double>(), h->w_pe.as<double>()))) return rc;
    }
    NRB_CUDA(h, cudaGetLastError());
    h->weights_dirty = false;
    return NRB200_OK;
}
