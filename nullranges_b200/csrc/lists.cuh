// libnrb200: host-built tables of a run -- per-tile slices (streaming path), gaps(exclude), and the
// signed window lists of the rank path.  (textually included by engine.cu)
namespace {

// Per-tile candidate window of a query/exclude set: everything that can overlap a range
// landing on the tile.  A shifted range overlaps [tile_start, tile_start + width - 1] before
// trimming, so it lies within (wmax - 1) of it on either side.
int build_slices(nrb200_handle *h, const RangeSet &rs, DevMem &lo_dev, DevMem &hi_dev) {
    const Plan &p = h->plan;
    const int64_t B = p.n_blocks();
    std::vector<int32_t> lo(B), hi(B);
    const int64_t reach = std::max<int32_t>(h->y.wmax, 1) - 1;
    for (int64_t t = 0; t < B; ++t) {
        const int c = p.tile_seq[t];
        const int64_t r_lo = (int64_t)p.tile_start[t] - reach;
        const int64_t r_hi = (int64_t)p.tile_start[t] + p.bait_width[t] - 1 + reach;
        const int32_t *sb = rs.h_start.data() + rs.h_seq_off[c], *se = rs.h_start.data() + rs.h_seq_off[c + 1];
        const int32_t *pb = rs.h_pmax.data() + rs.h_seq_off[c], *pe = rs.h_pmax.data() + rs.h_seq_off[c + 1];
        int32_t a = (int32_t)(std::lower_bound(pb, pe, r_lo, [](int32_t v, int64_t key) { return (int64_t)v < key; }) - rs.h_pmax.data());
        int32_t b = (int32_t)(std::upper_bound(sb, se, r_hi, [](int64_t key, int32_t v) { return key < (int64_t)v; }) - rs.h_start.data());
        if (a > b) a = b;
        lo[t] = a;
        hi[t] = b;
    }
    int rc;
    if ((rc = upload(h, lo_dev, lo.data(), (size_t)B))) return rc;
    if ((rc = upload(h, hi_dev, hi.data(), (size_t)B))) return rc;
    if (&rs == &h->query) {  // the tile groups of the streaming kernel are cut from these
        h->q_lo_host = std::move(lo);
        h->q_hi_host = std::move(hi);
    }
    return NRB200_OK;
}

// Tile groups of boot_stream_kernel: the tiles of every sequence in ascending start order, cut into runs of at most
// kGroupTiles whose query slices (the union of the tiles' candidate windows: contiguous, because both orders follow the
// genome) fit the shared-memory arrays.  A single tile with a larger slice makes the plan fall back to boot_count_kernel.
int build_tile_groups(nrb200_handle *h) {
    h->groups_ok = false;
    if (h->plan.permute()) return NRB200_OK;  // block b lands on tile dst_tile[b]: block-centric kernel
    const bool has_query = h->query.n > 0;
    std::vector<int32_t> sorted;
    std::vector<int32_t> groups;  // 4 words per group
    int cap = 16;
    for (int c = 0; c < h->n_seq; ++c) {
        const auto &ids = h->tiles_by_seq.id[c];
        size_t at = 0;
        while (at < ids.size()) {
            size_t n = std::min<size_t>(kGroupTiles, ids.size() - at);
            int32_t lo = 0, hi = 0;
            for (;;) {
                lo = INT32_MAX, hi = 0;
                for (size_t k = 0; has_query && k < n; ++k) {
                    lo = std::min(lo, h->q_lo_host[ids[at + k]]);
                    hi = std::max(hi, h->q_hi_host[ids[at + k]]);
                }
                if (!has_query || hi <= lo) { lo = hi = 0; break; }
                lo &= ~15;
                hi = (hi + 15) & ~15;
                if (hi - lo <= kMaxSlice) break;
                if (n == 1) return NRB200_OK;  // does not fit: the block-centric kernel takes this plan
                n = (n + 1) / 2;
            }
            groups.insert(groups.end(), {(int32_t)sorted.size(), (int32_t)n, lo, hi});
            cap = std::max(cap, hi - lo);
            for (size_t k = 0; k < n; ++k) sorted.push_back(ids[at + k]);
            at += n;
        }
    }
    int rc;
    if ((rc = upload(h, h->tiles_sorted_dev, sorted.data(), sorted.size()))) return rc;
    if ((rc = upload(h, h->groups_dev, groups.data(), groups.size()))) return rc;
    h->n_groups = (int64_t)groups.size() / 4;
    h->group_slice_cap = cap;
    h->groups_ok = h->n_groups > 0 && h->n_groups <= INT32_MAX;
    return NRB200_OK;
}

// ---- rank path tables --------------------------------------------------------------------
// Per feature: the signed windows to count on each tile within reach (two int4 per window, rank_kernels.cuh).
// Tile t is within reach of a window when the window comes within (wmax - 1) of the tile, i.e.
// when a range landing on the tile can overlap it.  Windows that start beyond their sequence end
// are dropped (trim() leaves nothing there).  With an exclude set, the merged exclude regions
// near both the tile and the feature add their pseudo-features with sign -1 (+1 for the overlap
// of two consecutive regions), and each tile gets the same for the row counts.
struct Interval {
    int64_t s, e;
};

inline bool strands_compatible(int a, int b) { return a == 0 || b == 0 || a == b; }

// reduce(): merged, sorted, disjoint (adjacent regions are joined: the union is what matters) -- of the
// exclude ranges a y range of strand class `cls` can overlap ('*' y: all of them).
static std::vector<std::vector<Interval>> merged_exclude(const nrb200_handle *h, int cls = 0) {
    const RangeSet &x = h->excl;
    std::vector<std::vector<Interval>> out(h->n_seq);
    for (int c = 0; c < h->n_seq; ++c) {
        for (int32_t k = x.h_seq_off[c]; k < x.h_seq_off[c + 1]; ++k) {  // canonical order: by start
            if (x.stranded && !strands_compatible(cls, x.h_strand[k])) continue;
            const int64_t s = x.h_start[k], e = x.h_end[k];
            if (s > h->seqlen[c]) continue;  // nothing survives trim() out there
            if (!out[c].empty() && s <= out[c].back().e + 1) out[c].back().e = std::max(out[c].back().e, e);
            else out[c].push_back({s, e});
        }
    }
    return out;
}

// gaps(exclude, end = seqlengths(y)) on strand "*" (R/bootRanges.R:112-113): complement of the merged
// exclude regions inside [1, L] on EVERY sequence, as a range set of its own.
int build_gaps(nrb200_handle *h) {
    const auto X = merged_exclude(h);
    std::vector<int32_t> sid, st, en;
    for (int c = 0; c < h->n_seq; ++c) {
        const int64_t L = h->seqlen[c];
        int64_t pos = 1;
        for (const Interval &x : X[c]) {
            if (x.s > pos) {
                sid.push_back(c);
                st.push_back((int32_t)pos);
                en.push_back((int32_t)std::min<int64_t>(x.s - 1, L));
            }
            pos = std::max(pos, x.e + 1);
        }
        if (pos <= L) {
            sid.push_back(c);
            st.push_back((int32_t)pos);
            en.push_back((int32_t)L);
        }
    }
    h->gaps.n = 0;
    if (sid.empty()) return NRB200_OK;
    return load_range_set(h, h->gaps, "gaps", (int64_t)sid.size(), sid.data(), st.data(), en.data(), nullptr, false);
}

int build_pairs(nrb200_handle *h) {
    Trace tr("build_pairs");
    const Plan &p = h->plan;
    const RangeSet &q = h->query;
    const int C = h->n_seq;
    const int64_t B = p.n_blocks(), G = q.n;
    if (h->lists_from) {  // member of a device group: same inputs as the leader, so the leader's lists are this handle's lists
        const HostLists &L = h->lists_from->lists;
        if (!L.valid) return fail(h, NRB200_ERR_STATE, "device group: the leader has no window lists");
        int rc;
        if (h->excl.n > 0) {
            if ((rc = upload(h, h->tile_x_off, L.x_off.data(), L.x_off.size()))) return rc;
            if ((rc = upload(h, h->tile_x_rec, L.x_rec.data(), L.x_rec.size()))) return rc;
        }
        if (G == 0) return NRB200_OK;
        h->n_pairs = L.n_pairs;
        if ((rc = upload(h, h->f_src, L.f_src.data(), L.f_src.size()))) return rc;
        if ((rc = upload(h, h->pair_off, L.f_off.data(), L.f_off.size()))) return rc;
        return upload(h, h->pair_rec, L.recs.data(), L.recs.size());
    }
    h->lists.valid = false;
    const int64_t reach = std::max<int32_t>(h->y.wmax, 1) - 1;
    const bool has_excl = h->excl.n > 0, trim = trim_mode(h);
    const int n_cls = h->n_cls;
    // Per strand class of y -- "drop": the merged exclude regions that class can overlap; "trim": the gaps
    // (already disjoint and sorted, strand-free).
    std::vector<std::vector<std::vector<Interval>>> XC(n_cls, std::vector<std::vector<Interval>>(C));
    const int n_width = h->n_width;
    const int64_t kk = h->query_minoverlap;  // 0: any overlap counts
    for (int cls = 0; cls < n_cls; ++cls) {
        if (has_excl && !trim) XC[cls] = (cls % n_width) ? XC[cls - 1] : merged_exclude(h, cls / n_width);
        if (trim)
            for (int c = 0; c < C; ++c)
                for (int32_t k = h->gaps.h_seq_off[c]; k < h->gaps.h_seq_off[c + 1]; ++k)
                    XC[cls][c].push_back({h->gaps.h_start[k], h->gaps.h_end[k]});
    }

    // exclude regions that intersect [lo, hi]: a contiguous run [first, last) of the merged list
    auto x_run = [&](const std::vector<Interval> &v, int64_t lo, int64_t hi, size_t &first, size_t &last) {
        first = std::lower_bound(v.begin(), v.end(), lo, [](const Interval &a, int64_t key) { return a.e < key; }) - v.begin();
        last = std::upper_bound(v.begin(), v.end(), hi, [](int64_t key, const Interval &a) { return key < a.s; }) - v.begin();
        if (last < first) last = first;
    };

    // ---- per (tile, strand class): exclusion windows for the row counts
    std::vector<int32_t> x_off((size_t)B * n_cls + 1, 0), x_rec;
    if (has_excl) {
        for (int64_t t = 0; t < B; ++t) {
            const int c = p.tile_seq[t];
            const int64_t ts = p.tile_start[t], te = ts + p.bait_width[t] - 1;
            for (int cls = 0; cls < n_cls; ++cls) {  // rows count every class, narrow ranges included
                const auto &X = XC[cls][c];
                size_t f, l;
                x_run(X, ts - reach, te + reach, f, l);
                for (size_t k = f; k < l; ++k) {
                    if (trim) {  // rows = one per (range, gap) overlap
                        x_rec.insert(x_rec.end(), {(int32_t)X[k].s, (int32_t)X[k].e, +1, 0});
                        continue;
                    }
                    x_rec.insert(x_rec.end(), {(int32_t)X[k].s, (int32_t)std::min<int64_t>(X[k].e, INT32_MAX / 2), -1, 0});
                    if (k + 1 < l && X[k + 1].s - X[k].e <= reach) x_rec.insert(x_rec.end(), {(int32_t)X[k + 1].s, (int32_t)X[k].e, +1, 0});
                }
                x_off[t * n_cls + cls + 1] = (int32_t)(x_rec.size() / 4);
            }
        }
    }
    int rc;
    if (has_excl) {
        if ((rc = upload(h, h->tile_x_off, x_off.data(), x_off.size()))) return rc;
        if ((rc = upload(h, h->tile_x_rec, x_rec.data(), x_rec.size()))) return rc;
    }
    if (G == 0) {
        if (h->keep_lists) {
            h->lists.x_off = std::move(x_off);
            h->lists.x_rec = std::move(x_rec);
            h->lists.n_pairs = 0;
            h->lists.valid = true;
        }
        return NRB200_OK;
    }

    // ---- per feature.  Tiles of each sequence in ascending start order with the running max of
    // their ends (tiles_by_seq, made with the plan); features come in canonical (start-sorted) order, so the
    // first tile that can still reach a feature only moves forward.  Two passes over the same enumeration: list
    // lengths, then the records, written straight into their final place.
    const auto &by_seq = h->tiles_by_seq.id;
    const auto &ts = h->tiles_by_seq.ts, &te = h->tiles_by_seq.te, &te_max = h->tiles_by_seq.te_max;
    // calls emit(tile, ws, we, sign) for every signed window of every feature, features in canonical order
    // (one sequence at a time: the sequences are independent, so the two passes run them in parallel)
    auto enumerate_seq = [&](int c, auto &&begin_feature, auto &&emit) {
        {
            const auto &v = by_seq[c];
            const int64_t L = h->seqlen[c], nt = (int64_t)v.size();
            int64_t lo = 0;
            for (int32_t k = q.h_seq_off[c]; k < q.h_seq_off[c + 1]; ++k) {
                begin_feature(k);
                const int64_t gs = q.h_start[k], ge = q.h_end[k];
                if (gs > L) continue;
                while (lo < nt && te_max[c][lo] + reach < gs) ++lo;
                const int g_strand = q.stranded ? q.h_strand[k] : 0;
                // minoverlap = kk: start' <= we - kk + 1, end' >= ws + kk - 1, and the trimmed range itself at
                // least kk wide (start' <= L - kk + 1, end' >= kk; the ranges narrower than kk sit in classes of
                // their own that no window names).  All of it folds into the window's end points.
                auto window = [&](int64_t ws, int64_t we, int64_t &lo_w, int64_t &hi_w) {
                    if (kk <= 0) { lo_w = ws; hi_w = we; return true; }
                    if (we - ws + 1 < kk) return false;
                    lo_w = std::max(ws + kk - 1, kk);
                    hi_w = std::min(we - kk + 1, L - kk + 1);
                    return true;
                };
                int64_t g_lo, g_hi;
                if (!window(gs, ge, g_lo, g_hi)) continue;
                for (int64_t j = lo; j < nt && ts[c][j] <= ge + reach; ++j) {
                    if (te[c][j] + reach < gs) continue;
                    for (int cls = 0; cls < n_cls; ++cls) {  // classes of y this feature counts
                        if ((cls % n_width) != 0 || !strands_compatible(cls / n_width, g_strand)) continue;
                        const auto &X = XC[cls][c];
                        if (trim) {  // the pieces that can land on the feature: one window per gap it intersects
                            size_t f, l;
                            x_run(X, std::max(ts[c][j] - reach, gs), std::min(te[c][j] + reach, ge), f, l);
                            for (size_t i = f; i < l; ++i) {
                                int64_t p_lo, p_hi;
                                if (window(std::max(X[i].s, gs), std::min(X[i].e, ge), p_lo, p_hi)) emit(k, v[j], p_lo, p_hi, +1, cls);
                            }
                            continue;
                        }
                        emit(k, v[j], g_lo, g_hi, +1, cls);
                        if (has_excl) {  // regions a range landing on the tile can overlap together with the feature
                            size_t f, l;
                            x_run(X, std::max(ts[c][j] - reach, gs - reach), std::min(te[c][j] + reach, ge + reach), f, l);
                            for (size_t i = f; i < l; ++i) {
                                emit(k, v[j], std::max(X[i].s, g_lo), std::min(X[i].e, g_hi), -1, cls);
                                if (i + 1 < l && X[i + 1].s - X[i].e <= reach) emit(k, v[j], std::max(X[i + 1].s, g_lo), std::min(X[i].e, g_hi), +1, cls);
                            }
                        }
                    }
                }
            }
        }
    };
    // Visit order of the count kernel: features grouped by list length (bins 0..62 hold exactly that length, the
    // last bin everything longer), inside a bin in canonical order, so neighbours in a group are neighbours on the
    // genome and the lanes of a warp loop equally often.  Sequences are enumerated in parallel: pass 1 also counts
    // each sequence's features per bin, which fixes where every (bin, sequence) group starts -- both among the
    // features and among the records (length x position inside a bin; a running sum in the last one).
    constexpr int kBins = 64;
    std::vector<int32_t> len(G, 0);
    std::vector<int64_t> hist((size_t)C * kBins, 0), long_sum(C, 0);
    auto bin = [&](int64_t k) { return std::min<int>(len[k], kBins - 1); };
    // body(c) for every sequence, in parallel: this handle's host team, or -- the leader of a device group -- all members' teams
    auto for_each_sequence = [&](const std::function<void(int)> &body) {
        if (h->seq_runner && G >= 4096) {
            h->seq_runner(C, body);
            return;
        }
#pragma omp parallel for schedule(dynamic, 1) num_threads(host_threads(h)) if (G >= 4096)
        for (int c = 0; c < C; ++c) body(c);
    };
    for_each_sequence([&](int c) {
        enumerate_seq(c, [](int32_t) {}, [&](int32_t k, int32_t, int64_t, int64_t, int, int) { ++len[k]; });
        for (int32_t k = q.h_seq_off[c]; k < q.h_seq_off[c + 1]; ++k) {
            hist[(size_t)c * kBins + bin(k)]++;
            if (len[k] >= kBins - 1) long_sum[c] += len[k];
        }
    });
    tr.lap("pass1");
    std::vector<int64_t> pos_start((size_t)C * kBins);  // first position of the (sequence, bin) group
    int64_t bin_first[kBins], rec_first[kBins];
    std::vector<int64_t> long_rec(C);  // first record of the sequence's features in the last bin
    int64_t pos = 0, total = 0;
    for (int b = 0; b < kBins; ++b) {
        bin_first[b] = pos;
        rec_first[b] = total;
        for (int c = 0; c < C; ++c) {
            pos_start[(size_t)c * kBins + b] = pos;
            pos += hist[(size_t)c * kBins + b];
            if (b == kBins - 1) {
                long_rec[c] = total;
                total += long_sum[c];
            }
        }
        if (b < kBins - 1) total += (pos - bin_first[b]) * b;
    }
    if (total > (int64_t)INT32_MAX / 2) return fail(h, NRB200_ERR_UNSUPPORTED, "too many (feature, tile) pairs");
    h->n_pairs = total;
    std::vector<int32_t> f_src(G), f_off(G + 1);
    f_off[G] = (int32_t)total;
    // records are written straight into pinned staging memory when they fit there
    struct Rec4 {
        int32_t x, y, z, w;
    };
    std::vector<int32_t> f_recs_vec;  // 16-byte window records (rank_kernels.cuh)
    Rec4 *f_recs = h->keep_lists ? nullptr : static_cast<Rec4 *>(arena_alloc(h, (size_t)std::max<int64_t>(total, 1) * sizeof(Rec4)));
    const bool staged = f_recs != nullptr;
    if (!staged) {  // (the leader of a device group keeps the host copy for the other members)
        f_recs_vec.resize((size_t)(4 * total));
        f_recs = reinterpret_cast<Rec4 *>(f_recs_vec.data());
    }
    for_each_sequence([&](int c) {
        int64_t cursor = 0, long_cursor = long_rec[c];
        int64_t *next_pos = &pos_start[(size_t)c * kBins];
        enumerate_seq(c,
                      [&](int32_t k) {
                          const int b = bin(k);
                          const int64_t at = next_pos[b]++;
                          if (b < kBins - 1) {
                              cursor = rec_first[b] + (at - bin_first[b]) * b;
                          } else {
                              cursor = long_cursor;
                              long_cursor += len[k];
                          }
                          f_src[at] = q.h_src[k];
                          f_off[at] = (int32_t)cursor;
                      },
                      [&](int32_t, int32_t t, int64_t ws, int64_t we, int sign, int cls) {
                          f_recs[cursor] = Rec4{(t << kRecTileShift) | (cls << 1) | (sign < 0 ? 1 : 0), (int32_t)ws,
                                                (int32_t)std::min<int64_t>(we, INT32_MAX / 2), p.tile_start[t]};
                          ++cursor;
                      });
    });
    tr.lap("lists");
    if ((rc = upload(h, h->f_src, f_src.data(), (size_t)G))) return rc;
    if ((rc = upload(h, h->pair_off, f_off.data(), (size_t)G + 1))) return rc;
    if ((rc = staged ? upload_staged(h, h->pair_rec, f_recs, (size_t)total) : upload(h, h->pair_rec, f_recs, (size_t)total))) return rc;
    tr.lap("upload");
    if (h->keep_lists) {
        HostLists &L = h->lists;
        L.n_pairs = total;
        L.f_src = std::move(f_src);
        L.f_off = std::move(f_off);
        L.x_off = std::move(x_off);
        L.x_rec = std::move(x_rec);
        L.recs = std::move(f_recs_vec);
        L.valid = true;
    }
    return NRB200_OK;
}

}  // namespace
