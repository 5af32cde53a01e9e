// libnrb200: host side of the C ABI (include/nrb200.h) and the kernel launches.
//
// Memory layout in HBM (all int32 unless noted, canonical = sorted by (seqname, start, end)):
//   y        : start[N] end[N] pmax[N] src[N] (+ strand[N] int8 when stranded), seq_off[C+1],
//              end_sorted[N] (ends ascending within each sequence), three direct-address rank
//              tables (start, running-max end, sorted end), ~N entries each
//   query    : start[G] end[G] pmax[G] src[G] (+strand) + per-tile window lo/hi [B]
//              + feature -> tiles-within-reach lists (pair_off[G+1], pair_tile[])
//   exclude  : same shape as query
//   plan     : tile_seq[B] tile_start[B] bait_width[B], sampler tables (int64, a few KB)
//   draws    : validation mode only, bait_seq/bait_start(/dst_tile) [R x B]
//   results  : iter_nranges[R] iter_totals[R] (uint64), counts[R x G]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/nrb200.h"
#include "kernels.cuh"
#include "plan.h"

namespace nrb {

// NRB200_TRACE=1: host-side stage timings on stderr (development aid)
struct Trace {
    static bool on() {
        static const bool v = getenv("NRB200_TRACE") != nullptr;
        return v;
    }
    const char *name;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(const char *n) : name(n), t0(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on()) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[nrb200] %s/%s %.3f ms\n", name, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static thread_local std::string g_create_error;

// Growth-only device buffer.
struct DevMem {
    void *p = nullptr;
    size_t cap = 0;
    ~DevMem() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// One set of ranges in canonical order, host mirror + device copy.
struct RangeSet {
    int64_t n = 0;
    bool stranded = false;
    int32_t wmax = 0, wmin = 0;
    std::vector<int32_t> h_start, h_end, h_pmax, h_src, h_seq_off;  // host mirror (kept for query/exclude)
    std::vector<int8_t> h_strand;
    std::vector<int32_t> max_end;  // [n_seq] largest end on the sequence (INT32_MIN when empty)
    bool on_device = false;        // query: device copies are made on first use (the rank path needs none)
    bool sorted_input = false;     // the caller's order was already canonical
    DevMem start, end, pmax, src, strand, seq_off, seqid;
};

}  // namespace nrb

using namespace nrb;

struct nrb200_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    // Pinned staging for small host->device uploads (block tables, window lists, ...): a copy from
    // pageable memory would first wait for everything queued on the stream; through this arena the
    // uploads queue behind the GPU work and the host carries on.  Reset whenever the stream is known
    // to be idle; when full, uploads fall back to the direct (synchronising) copy.
    char *arena = nullptr;
    size_t arena_cap = 0, arena_used = 0;
    cudaEvent_t split_ev = nullptr;       // set while a stats run wants the rows / count kernel split
    std::vector<cudaEvent_t> batch_ev;    // 3 per batch: start, rows done, count done
    cudaEvent_t chunk_ev[16] = {};
    std::string error;
    int sm_count = 148;

    int n_seq = 0;
    std::vector<int64_t> seqlen;
    DevMem seqlen_dev;

    int32_t query_widen = 0;  // maxgap + 1 of the stored query
    RangeSet y, query, excl, gaps;  // gaps = gaps(exclude) inside [1, seqlength], built for excludeOption "trim"
    int excl_option = NRB200_EXCLUDE_DROP;
    bool have_y = false, have_plan = false;
    // rank tables of y
    int rank_shift = 0;
    std::vector<int64_t> rank_off;
    DevMem inspect_dev, seq_info_dev, draw_tab;
    // stranded y only: class-major copy for the rank path (kernels.cuh, RangesView)
    int n_cls = 1, n_width = 1;
    int32_t query_minoverlap = 0, index_minoverlap = 0;  // the index groups y by "narrower than minoverlap": rebuilt when it changes
    int32_t end_shift = 0;
    bool ends_alias_starts = false;
    // weights of y (nrb200_set_weights): canonical order, class-major order, running sums along start / end order
    bool have_weights = false, weights_dirty = true;
    DevMem w_in, w_c, w_r, w_e, w_ps, w_pe, w_keys_e, w_perm_a, w_perm_b, ws_sums, ws_iter;
    DevMem r_start, r_end, r_pmax, r_order, r_idx, r_keys_a, r_keys_b, r_voff_dev, rseq_info_dev, rrank_off_dev, rrank_start;
    DevMem rank_off_dev, max_pos_dev, rank_start, rank_pmax, rank_end, end_sorted, sort_keys_a, sort_keys_b, sort_tmp;
    uint32_t flags = 0;
    // tuning knobs (environment, read at create): NRB200_RPT iterations per thread of the rank count
    // kernel, NRB200_MINB its min CTAs/SM, NRB200_TABLE_MULT rank-table entries per range
    int tune_rpt = 1, tune_minb = 1;
    double tune_table_mult = 4.0;
    int64_t tune_max_batch = 2048;     // NRB200_MAX_BATCH: iterations per launch group
    int64_t tune_table_cap = 8 << 20;  // NRB200_TABLE_CAP: the multiplier applies up to this many entries

    Plan plan;
    DevMem tile_seq, tile_start, bait_width, sampler_i64, sampler_i32;
    SamplerView sampler_view{};
    bool slices_dirty = true, have_pmax_table = false;
    DevMem q_tile_lo, q_tile_hi, x_tile_lo, x_tile_hi, pair_off, pair_rec, f_src, tile_x_off, tile_x_rec;
    int64_t n_pairs = 0;
    bool tiles_in_bounds = false;

    // run scratch / results
    DevMem d_bait_seq, d_bait_start, d_dst_tile, perm_vals_a, perm_vals_b, np_seg, np_start, np_state, np_list, np_cnt;
    DevMem d_iter_nranges, d_iter_totals, d_counts, d_n_hits;
    bool res_has_counts = false;
    // moments
    DevMem m_sum, m_sumsq, m_nge, m_observed;
    bool moments_ready = false;
    // rows
    DevMem item_rec, item_rows, within_off, iter_off_dev, row_seq, row_start, row_end, row_src, row_block;
    int64_t n_rows = 0;
};

namespace {

int fail(nrb200_handle *h, int code, const std::string &msg) {
    if (h) h->error = msg; else g_create_error = msg;
    return code;
}

// The set the kernels test bootstrapped ranges against: the exclude ranges ("drop") or their gaps ("trim").
inline bool trim_mode(const nrb200_handle *h) { return h->excl_option == NRB200_EXCLUDE_TRIM && h->excl.n > 0; }
inline const RangeSet &active_excl(const nrb200_handle *h) { return trim_mode(h) ? h->gaps : h->excl; }
inline int excl_kind(const nrb200_handle *h) { return h->excl.n == 0 ? kExclNone : (trim_mode(h) ? kExclTrim : kExclDrop); }

#define NRB_CUDA(h, call)                                                                          \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess)                                                                  \
            return fail((h), NRB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)); \
    } while (0)

constexpr size_t kArenaBytes = 16u << 20;

// Waits for the stream; the staging arena can be reused afterwards.
int sync_stream(nrb200_handle *h) {
    NRB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->arena_used = 0;
    return NRB200_OK;
}

// Host -> device copy of a library-owned (pageable) array, staged through the pinned arena when it fits.
template <class T>
int upload(nrb200_handle *h, DevMem &dst, const T *src, size_t n) {
    const size_t bytes = n * sizeof(T);
    NRB_CUDA(h, dst.reserve(bytes));
    if (!n) return NRB200_OK;
    if (!h->arena && cudaMallocHost((void **)&h->arena, kArenaBytes) == cudaSuccess) h->arena_cap = kArenaBytes;
    const size_t at = (h->arena_used + 255) & ~(size_t)255;
    if (h->arena && at + bytes <= h->arena_cap) {
        std::memcpy(h->arena + at, src, bytes);
        h->arena_used = at + bytes;
        NRB_CUDA(h, cudaMemcpyAsync(dst.p, h->arena + at, bytes, cudaMemcpyHostToDevice, h->stream));
    } else {
        NRB_CUDA(h, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    }
    return NRB200_OK;
}

// Room inside the pinned arena to BUILD an array in place (then upload_staged): nullptr when it does not fit.
void *arena_alloc(nrb200_handle *h, size_t bytes) {
    if (!h->arena && cudaMallocHost((void **)&h->arena, kArenaBytes) == cudaSuccess) h->arena_cap = kArenaBytes;
    const size_t at = (h->arena_used + 255) & ~(size_t)255;
    if (!h->arena || at + bytes > h->arena_cap) return nullptr;
    h->arena_used = at + bytes;
    return h->arena + at;
}
template <class T>
int upload_staged(nrb200_handle *h, DevMem &dst, const T *arena_src, size_t n) {
    NRB_CUDA(h, dst.reserve(n * sizeof(T)));
    if (n) NRB_CUDA(h, cudaMemcpyAsync(dst.p, arena_src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return NRB200_OK;
}

// Host -> device copy straight from a caller-owned array (pinned or pageable): never staged.
template <class T>
int upload_direct(nrb200_handle *h, DevMem &dst, const T *src, size_t n) {
    NRB_CUDA(h, dst.reserve(n * sizeof(T)));
    if (n) NRB_CUDA(h, cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return NRB200_OK;
}

// Canonicalise one set of ranges on the host and mirror it to the device.
int range_set_to_device(nrb200_handle *h, RangeSet &rs) {
    if (rs.on_device) return NRB200_OK;
    int rc;
    const size_t n = (size_t)rs.n;
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
// Unsorted input falls back to the host canonicalisation above.  Returns 1 for "fall back".
int load_y_device(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start, const int32_t *end,
                  const int8_t *strand) {
    RangeSet &y = h->y;
    const int C = h->n_seq;
    if (n <= 0 || n > (int64_t)INT32_MAX - 64 || !seqid || !start || !end) return 1;  // host path words the errors
    int rc;
    if ((rc = upload_direct(h, y.seqid, seqid, (size_t)n))) return rc;
    if ((rc = upload_direct(h, y.start, start, (size_t)n))) return rc;
    if ((rc = upload_direct(h, y.end, end, (size_t)n))) return rc;
    if (strand && (rc = upload_direct(h, y.strand, strand, (size_t)n))) return rc;
    Trace tr("load_y_device");
    tr.lap("enqueue_h2d");
    // scratch: InspectOut (5 words, padded to 8) | seq_count[C] | seq_max_end[C]
    const size_t kHead = 8;
    const size_t words = kHead + 2 * (size_t)C;
    NRB_CUDA(h, h->inspect_dev.reserve(words * 4));
    std::vector<int32_t> init(words, 0);
    init[4] = INT32_MAX;  // wmin
    for (int c = 0; c < C; ++c) init[kHead + C + c] = INT32_MIN;
    NRB_CUDA(h, cudaMemcpyAsync(h->inspect_dev.p, init.data(), words * 4, cudaMemcpyHostToDevice, h->stream));
    int32_t *base = h->inspect_dev.as<int32_t>();
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
    inspect_ranges_kernel<<<grid, 256, 0, h->stream>>>(y.seqid.as<int32_t>(), y.start.as<int32_t>(), y.end.as<int32_t>(),
                                                       strand ? y.strand.as<int8_t>() : nullptr, n, C,
                                                       reinterpret_cast<InspectOut *>(base), reinterpret_cast<unsigned *>(base + kHead),
                                                       base + kHead + C);
    NRB_CUDA(h, cudaGetLastError());
    std::vector<int32_t> res(words);
    NRB_CUDA(h, cudaMemcpyAsync(res.data(), base, words * 4, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = sync_stream(h))) return rc;  // the caller's arrays are no longer read after this point
    tr.lap("inspect+sync");
    if (res[0] != 0 || res[1] != 0) return 1;  // invalid (host path names the problem) or unsorted
    y.n = n;
    y.sorted_input = true;
    y.stranded = res[2] != 0;
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
    NRB_CUDA(h, h->rank_start.reserve(entries * sizeof(int32_t)));
    NRB_CUDA(h, h->rank_pmax.reserve(entries * sizeof(int32_t)));
    NRB_CUDA(h, h->end_sorted.reserve(std::max<int64_t>(y.n, 1) * sizeof(int32_t)));
    const int threads = 256;
    const unsigned n_grid = (unsigned)((std::max<int64_t>(y.n, 1) + threads - 1) / threads);
    const unsigned t_grid = (unsigned)((entries + threads - 1) / threads);
    // canonical start table: locate_block() of the streaming kernels, and the rank path of an unstranded y
    build_rank_table_kernel<<<t_grid, threads, 0, h->stream>>>(y.start.as<int32_t>(), y.seq_off.as<int32_t>(),
                                                               h->rank_off_dev.as<int64_t>(), C, shift, h->rank_start.as<uint32_t>());
    h->have_pmax_table = false;  // only the streaming kernels need it: built on first use
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
        NRB_CUDA(h, h->rank_end.reserve(entries * sizeof(int32_t)));
        if (y.n > 0) {
            int seq_bits = 1;
            while ((1 << seq_bits) < C) ++seq_bits;
            pack_end_keys_kernel<<<n_grid, threads, 0, h->stream>>>(y.seqid.as<int32_t>(), y.end.as<int32_t>(), y.n,
                                                                    h->sort_keys_a.as<unsigned long long>());
            if ((rc = sort_u64(32 + seq_bits))) return rc;
            unpack_end_keys_kernel<<<n_grid, threads, 0, h->stream>>>(h->sort_keys_b.as<unsigned long long>(), y.n, h->end_sorted.as<int32_t>());
        }
        build_rank_table_kernel<<<t_grid, threads, 0, h->stream>>>(h->end_sorted.as<int32_t>(), y.seq_off.as<int32_t>(),
                                                                   h->rank_off_dev.as<int64_t>(), C, shift, h->rank_end.as<uint32_t>());
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
    NRB_CUDA(h, h->rrank_start.reserve(r_entries * sizeof(int32_t)));
    NRB_CUDA(h, h->rank_end.reserve(r_entries * sizeof(int32_t)));
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
    const unsigned rt_grid = (unsigned)((r_entries + threads - 1) / threads);
    build_rank_table_kernel<<<rt_grid, threads, 0, h->stream>>>(h->r_start.as<int32_t>(), h->r_voff_dev.as<int32_t>(),
                                                                h->rrank_off_dev.as<int64_t>(), V, shift, h->rrank_start.as<uint32_t>());
    if (!h->ends_alias_starts)
        build_rank_table_kernel<<<rt_grid, threads, 0, h->stream>>>(h->end_sorted.as<int32_t>(), h->r_voff_dev.as<int32_t>(),
                                                                    h->rrank_off_dev.as<int64_t>(), V, shift, h->rank_end.as<uint32_t>());
    NRB_CUDA(h, cudaGetLastError());
    return NRB200_OK;
}

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
    return NRB200_OK;
}

// ---- rank path tables --------------------------------------------------------------------
// Per feature: the signed windows to count on each tile within reach (PairRec, kernels.cuh).
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
    if (G == 0) return NRB200_OK;

    // ---- per feature.  Tiles of each sequence in ascending start order with the running max of
    // their ends; features come in canonical (start-sorted) order, so the first tile that can still
    // reach a feature only moves forward.  Two passes over the same enumeration: list lengths, then
    // (after grouping features by length) the records, written straight into their final place.
    std::vector<std::vector<int32_t>> by_seq(C);
    for (int64_t t = 0; t < B; ++t) by_seq[p.tile_seq[t]].push_back((int32_t)t);
    std::vector<std::vector<int64_t>> ts(C), te(C), te_max(C);
    for (int c = 0; c < C; ++c) {
        auto &v = by_seq[c];
        bool sorted = true;
        for (size_t j = 1; j < v.size(); ++j) sorted &= p.tile_start[v[j - 1]] <= p.tile_start[v[j]];
        if (!sorted)
            std::sort(v.begin(), v.end(), [&](int32_t a, int32_t b) {
                return p.tile_start[a] != p.tile_start[b] ? p.tile_start[a] < p.tile_start[b] : a < b;
            });
        ts[c].resize(v.size());
        te[c].resize(v.size());
        te_max[c].resize(v.size());
        int64_t m = INT64_MIN;
        for (size_t j = 0; j < v.size(); ++j) {
            ts[c][j] = p.tile_start[v[j]];
            te[c][j] = ts[c][j] + p.bait_width[v[j]] - 1;
            m = std::max(m, te[c][j]);
            te_max[c][j] = m;
        }
    }
    tr.lap("tiles_by_seq");
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
    std::vector<int32_t> len(G, 0);
#pragma omp parallel for schedule(dynamic, 1) if (G >= 4096)
    for (int c = 0; c < C; ++c) enumerate_seq(c, [](int32_t) {}, [&](int32_t k, int32_t, int64_t, int64_t, int, int) { ++len[k]; });
    tr.lap("pass1");
    // Visit order of the count kernel: features grouped by list length (stable counting sort, so
    // neighbours in a group are neighbours on the genome): the lanes of a warp then loop equally often.
    constexpr int kBins = 64;
    int64_t bin_start[kBins + 1] = {0};
    auto bin = [&](int64_t k) { return std::min<int>(len[k], kBins - 1); };
    for (int64_t k = 0; k < G; ++k) bin_start[bin(k) + 1]++;
    for (int b = 0; b < kBins; ++b) bin_start[b + 1] += bin_start[b];
    std::vector<int32_t> pos_of(G), f_src(G), f_off(G + 1, 0);
    for (int64_t k = 0; k < G; ++k) pos_of[k] = (int32_t)bin_start[bin(k)]++;
    for (int64_t k = 0; k < G; ++k) {
        f_src[pos_of[k]] = q.h_src[k];
        f_off[pos_of[k] + 1] = len[k];
    }
    int64_t total = 0;
    for (int64_t i = 0; i < G; ++i) {
        total += f_off[i + 1];
        if (total > (int64_t)INT32_MAX / 2) return fail(h, NRB200_ERR_UNSUPPORTED, "too many (feature, tile) pairs");
        f_off[i + 1] = (int32_t)total;
    }
    h->n_pairs = total;
    tr.lap("binning");
    // records are written straight into pinned staging memory when they fit there
    std::vector<PairRec> f_recs_vec;
    PairRec *f_recs = static_cast<PairRec *>(arena_alloc(h, (size_t)std::max<int64_t>(total, 1) * sizeof(PairRec)));
    const bool staged = f_recs != nullptr;
    if (!staged) {
        f_recs_vec.resize((size_t)total);
        f_recs = f_recs_vec.data();
    }
#pragma omp parallel for schedule(dynamic, 1) if (G >= 4096)
    for (int c = 0; c < C; ++c) {
        int64_t cursor = 0;
        enumerate_seq(c, [&](int32_t k) { cursor = f_off[pos_of[k]]; },
                      [&](int32_t, int32_t t, int64_t ws, int64_t we, int sign, int cls) {
                          f_recs[cursor++] = PairRec{t, (int32_t)ws, (int32_t)std::min<int64_t>(we, INT32_MAX / 2), sign, p.tile_start[t],
                                                     p.bait_width[t], (int32_t)h->seqlen[p.tile_seq[t]], cls};
                      });
    }
    tr.lap("lists");
    if ((rc = upload(h, h->f_src, f_src.data(), (size_t)G))) return rc;
    if ((rc = upload(h, h->pair_off, f_off.data(), (size_t)G + 1))) return rc;
    if ((rc = staged ? upload_staged(h, h->pair_rec, f_recs, (size_t)total) : upload(h, h->pair_rec, f_recs, (size_t)total))) return rc;
    tr.lap("upload");
    return NRB200_OK;
}

// The rank path covers every sampler (bootstrap and permute, with or without exclusion, any strands);
// only a segmentation reaching beyond its sequence falls back to streaming the hits (boot_count_kernel).
bool rank_path_ok(const nrb200_handle *h, bool with_query) {
    (void)with_query;
    if (h->flags & NRB200_FLAG_FORCE_STREAM) return false;
    return h->tiles_in_bounds;
}

// rank table over the running max of end: locate_block() of the streaming kernels
int ensure_pmax_table(nrb200_handle *h) {
    if (h->have_pmax_table) return NRB200_OK;
    const int64_t entries = h->rank_off[h->n_seq];
    build_rank_table_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, h->stream>>>(
        h->y.pmax.as<int32_t>(), h->y.seq_off.as<int32_t>(), h->rank_off_dev.as<int64_t>(), h->n_seq, h->rank_shift,
        h->rank_pmax.as<uint32_t>());
    NRB_CUDA(h, cudaGetLastError());
    h->have_pmax_table = true;
    return NRB200_OK;
}

int prepare(nrb200_handle *h) {
    if (!h->have_y) return fail(h, NRB200_ERR_STATE, "nrb200_set_ranges() has not been called");
    if (!h->have_plan) return fail(h, NRB200_ERR_STATE, "nrb200_set_plan() has not been called");
    if (h->slices_dirty) {
        int rc;
        Trace tr("prepare");
        h->tiles_in_bounds = true;
        for (int64_t t = 0; t < h->plan.n_blocks(); ++t)
            if (h->plan.tile_start[t] > h->seqlen[h->plan.tile_seq[t]]) h->tiles_in_bounds = false;
        if (h->index_minoverlap != h->query_minoverlap && (rc = build_rank_tables(h))) return rc;  // the width classes changed
        if (trim_mode(h) && h->excl.stranded) return fail(h, NRB200_ERR_INVALID, "all(strand(exclude) == \"*\") is not TRUE");
        if (trim_mode(h) && (rc = build_gaps(h))) return rc;
        // rank_path_ok(h, false) holds whenever rank_path_ok(h, true) does
        const bool rank_q = h->query.n && rank_path_ok(h, true), rank_rows = rank_path_ok(h, false);
        if (rank_q || rank_rows) {
            const int64_t g_keep = h->query.n;
            if (!rank_q) h->query.n = 0;  // tables for the row counts only
            rc = build_pairs(h);
            h->query.n = g_keep;
            if (rc) return rc;
            tr.lap("build_pairs");
        }
        // the streaming kernels (count without the rank path, and the fill pass) test exclusion per range
        if (h->excl.n && (rc = build_slices(h, active_excl(h), h->x_tile_lo, h->x_tile_hi))) return rc;
        if (h->query.n && !rank_q && (rc = range_set_to_device(h, h->query))) return rc;
        if (h->query.n && !rank_q && (rc = build_slices(h, h->query, h->q_tile_lo, h->q_tile_hi))) return rc;
        if (((h->query.n && !rank_q) || !rank_rows) && (rc = ensure_pmax_table(h))) return rc;
        h->slices_dirty = false;
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
    int rc;
    if ((rc = upload(h, h->d_bait_seq, d->bait_seqid, (size_t)n))) return rc;
    if ((rc = upload(h, h->d_bait_start, d->bait_start, (size_t)n))) return rc;
    if (d->dst_tile && (rc = upload(h, h->d_dst_tile, d->dst_tile, (size_t)n))) return rc;
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
    Y.rank_start = h->rank_start.as<uint32_t>();
    Y.rank_pmax = h->rank_pmax.as<uint32_t>();
    Y.rank_shift = h->rank_shift;
    Y.n_cls = h->n_cls;
    Y.end_sorted = h->end_sorted.as<int32_t>();
    Y.rank_end = h->rank_end.as<uint32_t>();
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
        Y.rrank_start = h->rrank_start.as<uint32_t>();
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
void launch_count_t(nrb200_handle *h, const BootParams &P, unsigned grid, int excl) {
    if (excl == kExclNone) boot_count_kernel<ST, HQ, kExclNone><<<grid, 256, 0, h->stream>>>(P);
    else if (excl == kExclDrop) boot_count_kernel<ST, HQ, kExclDrop><<<grid, 256, 0, h->stream>>>(P);
    else boot_count_kernel<ST, HQ, kExclTrim><<<grid, 256, 0, h->stream>>>(P);
}

constexpr int64_t kMaxBatch = 2048;  // iterations per launch at most (gridDim.y and scratch stay small); NRB200_MAX_BATCH lowers it
constexpr int kChunkEvents = 16;
constexpr int kChunks = 8;  // pipeline depth of nrb200_run_counts (compute chunk k+1 while chunk k copies out)

// Returns the number of kernels launched.
int launch_rank(nrb200_handle *h, const BootParams &P, bool has_query) {
    if (P.n_iter == 0) return 0;
    int launches = 0;
    const int rpt = h->tune_rpt;
    const unsigned gy = (unsigned)((P.n_iter + rpt - 1) / rpt);
    if (P.n_blocks > 0) {
        const TileExcl X{h->excl.n > 0 ? h->tile_x_off.as<int32_t>() : nullptr, h->tile_x_rec.as<int4>()};  // drop: minus, trim: plus
        boot_block_rows_kernel<1><<<dim3((unsigned)((P.n_blocks + 255) / 256), (unsigned)P.n_iter), 256, 0, h->stream>>>(P, X);
        ++launches;
    }
    if (h->split_ev) cudaEventRecord(h->split_ev, h->stream);  // stats: end of the rows kernel
    if (has_query) {
        const FeatView F{h->f_src.as<int32_t>(), h->pair_off.as<int32_t>(), h->pair_rec.as<PairRec>()};
        const dim3 grid((unsigned)((P.n_query + 255) / 256), gy);
        switch (rpt * 10 + h->tune_minb) {
        case 11: boot_rank_count_kernel<1, 1><<<grid, 256, 0, h->stream>>>(P, F); break;
        case 21: boot_rank_count_kernel<2, 1><<<grid, 256, 0, h->stream>>>(P, F); break;
        case 24: boot_rank_count_kernel<2, 4><<<grid, 256, 0, h->stream>>>(P, F); break;
        case 41: boot_rank_count_kernel<4, 1><<<grid, 256, 0, h->stream>>>(P, F); break;
        case 43: boot_rank_count_kernel<4, 3><<<grid, 256, 0, h->stream>>>(P, F); break;
        case 44: boot_rank_count_kernel<4, 4><<<grid, 256, 0, h->stream>>>(P, F); break;
        case 81: boot_rank_count_kernel<8, 1><<<grid, 256, 0, h->stream>>>(P, F); break;
        default: boot_rank_count_kernel<1, 1><<<grid, 256, 0, h->stream>>>(P, F); break;
        }
        ++launches;
    }
    return launches;
}

int launch_count(nrb200_handle *h, const BootParams &P, bool has_query) {
    if (rank_path_ok(h, has_query)) return launch_rank(h, P, has_query);
    const int64_t items = P.n_iter * P.n_blocks;
    if (items == 0) return 0;
    const int64_t ctas = (items + 7) / 8;
    const unsigned grid = (unsigned)std::min<int64_t>(ctas, (int64_t)h->sm_count * 64);
    const bool st = h->y.stranded && ((has_query && h->query.stranded) || (h->excl.n && active_excl(h).stranded));
    const int excl = excl_kind(h);
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
    if (P.item_rows) Q.item_rows = P.item_rows + r0 * P.n_blocks;
    if (P.item_rec) Q.item_rec = P.item_rec + r0 * P.n_blocks;
    return Q;
}

// Core of every counting entry point: results stay on the device.  With host_counts set, the
// iterations run in chunks and each chunk of the [R x G] matrix starts its way to the host
// (copy stream) while the next chunk computes.
int run_counts_core(nrb200_handle *h, const nrb200_draws *d, bool want_counts, int32_t *host_counts,
                    nrb200_stats *stats) {
    int rc;
    if ((rc = prepare(h))) return rc;
    NRB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if ((rc = stage_draws(h, d))) return rc;
    const int64_t R = d->n_iter, G = h->query.n;
    const bool has_query = G > 0;
    want_counts = want_counts && has_query;
    NRB_CUDA(h, h->d_iter_nranges.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->d_iter_totals.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->d_n_hits.reserve(8));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_nranges.p, 0, R * 8, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_totals.p, 0, R * 8, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->d_n_hits.p, 0, 8, h->stream));
    int32_t *counts = nullptr;
    if (want_counts) {
        NRB_CUDA(h, h->d_counts.reserve((size_t)R * G * sizeof(int32_t)));
        counts = h->d_counts.as<int32_t>();
        // the rank path writes every element; the streaming path accumulates with atomics
        if (!rank_path_ok(h, true)) NRB_CUDA(h, cudaMemsetAsync(counts, 0, (size_t)R * G * sizeof(int32_t), h->stream));
    }
    BootParams P = make_params(h, d);
    P.iter_nranges = h->d_iter_nranges.as<unsigned long long>();
    P.iter_totals = h->d_iter_totals.as<unsigned long long>();
    P.counts = counts;
    P.n_hits = h->d_n_hits.as<unsigned long long>();
    NRB_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    int launches = 0;
    const bool pipelined = host_counts && counts && R > 0;
    // Iterations go in batches: bounded draw-table scratch on the rank path and, when the matrix
    // is wanted on the host, chunk k copies out (copy stream) while chunk k + 1 computes.
    int64_t chunk = h->tune_max_batch;
    if (pipelined) {
        if (!h->copy_stream) NRB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        chunk = std::min<int64_t>(h->tune_max_batch, std::max<int64_t>(((R + kChunks - 1) / kChunks + 7) / 8 * 8, 64));
    }
    if (rank_path_ok(h, has_query) && has_query) {
        NRB_CUDA(h, h->draw_tab.reserve((size_t)std::min(chunk, std::max<int64_t>(R, 1)) * h->plan.n_blocks() * sizeof(int2)));
        P.draw_tab = h->draw_tab.as<int2>();
    }
    int k = 0;
    const bool split = stats && rank_path_ok(h, has_query);
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
        launches += launch_count(h, slice_params(P, r0, n), has_query);
        h->split_ev = nullptr;
        if (split) NRB_CUDA(h, cudaEventRecord(h->batch_ev[3 * k + 2], h->stream));
        NRB_CUDA(h, cudaGetLastError());
        if (pipelined) {
            cudaEvent_t &ev = h->chunk_ev[k % kChunkEvents];
            if (!ev) NRB_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            NRB_CUDA(h, cudaEventRecord(ev, h->stream));
            NRB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, ev, 0));
            NRB_CUDA(h, cudaMemcpyAsync(host_counts + r0 * G, counts + r0 * G, (size_t)n * G * sizeof(int32_t),
                                        cudaMemcpyDeviceToHost, h->copy_stream));
        }
    }
    NRB_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    h->res_has_counts = want_counts;
    if (stats) {
        NRB_CUDA(h, cudaEventSynchronize(h->ev[2]));
        float k_ms = 0, t_ms = 0;
        cudaEventElapsedTime(&k_ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&t_ms, h->ev[0], h->ev[2]);
        unsigned long long hits = 0;
        NRB_CUDA(h, cudaMemcpyAsync(&hits, h->d_n_hits.p, 8, cudaMemcpyDeviceToHost, h->stream));
        if (int rc_sync = sync_stream(h)) return rc_sync;
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

extern "C" {

const char *nrb200_version(void) { return "nrb200 0.1.0 (sm_100a)"; }

const char *nrb200_last_error(const nrb200_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int nrb200_create(const nrb200_config *cfg, nrb200_handle **out) {
    if (!out) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: out is NULL");
    *out = nullptr;
    const int dev = cfg ? cfg->device : 0;
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
    if (const char *e = getenv("NRB200_RPT")) h->tune_rpt = atoi(e);
    if (const char *e = getenv("NRB200_MINB")) h->tune_minb = atoi(e);
    if (const char *e = getenv("NRB200_TABLE_MULT")) h->tune_table_mult = atof(e);
    if (const char *e = getenv("NRB200_TABLE_CAP")) h->tune_table_cap = atoll(e);
    if (const char *e = getenv("NRB200_MAX_BATCH")) h->tune_max_batch = std::min<int64_t>(std::max<int64_t>(atoll(e), 8), 2048);
    if (h->tune_rpt != 1 && h->tune_rpt != 2 && h->tune_rpt != 4 && h->tune_rpt != 8) h->tune_rpt = 1;
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

void nrb200_destroy(nrb200_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto &ev : h->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->chunk_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->batch_ev) cudaEventDestroy(ev);
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamDestroy(h->copy_stream);
    }
    if (h->arena) cudaFreeHost(h->arena);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

int nrb200_set_ranges(nrb200_handle *h, int64_t n, int32_t n_seq, const int64_t *seqlengths, const int32_t *seqid,
                      const int32_t *start, const int32_t *end, const int8_t *strand) {
    if (!h) return NRB200_ERR_INVALID;
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (n_seq <= 0 || !seqlengths) return fail(h, NRB200_ERR_INVALID, "seqlengths(y) must be known");
    for (int c = 0; c < n_seq; ++c) {
        // stopifnot(all(!is.na(chr_lens)))                          R/bootRanges.R:81
        if (seqlengths[c] <= 0) return fail(h, NRB200_ERR_INVALID, "all(!is.na(seqlengths(y))) is not TRUE");
        if (seqlengths[c] >= (1 << 30)) return fail(h, NRB200_ERR_UNSUPPORTED, "seqlengths must be below 2^30");
    }
    h->have_y = h->have_plan = false;
    h->query.n = h->excl.n = 0;  // seqlevels may have changed
    h->query_minoverlap = 0;
    h->n_seq = n_seq;
    h->seqlen.assign(seqlengths, seqlengths + n_seq);
    int rc;
    if ((rc = upload(h, h->seqlen_dev, seqlengths, (size_t)n_seq))) return rc;
    Trace tr("set_ranges");
    rc = load_y_device(h, n, seqid, start, end, strand);
    if (rc < 0) return rc;
    tr.lap("load_y_device");
    if (rc == 1 && (rc = load_range_set(h, h->y, "y", n, seqid, start, end, strand, true))) return rc;
    if ((rc = build_rank_tables(h))) return rc;
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

int nrb200_ranges_info(const nrb200_handle *h, int32_t *is_sorted, int32_t *is_stranded, int32_t *max_width) {
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
    if (!h->have_y) return fail(h, NRB200_ERR_STATE, "call nrb200_set_ranges() first (it defines the seqlevels)");
    h->slices_dirty = true;
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
    const int rc = set_side(h, h->query, "query", n, seqid, start, end, strand, maxgap + 1);
    if (rc == NRB200_OK) {
        h->query_widen = maxgap + 1;
        h->query_minoverlap = n > 0 ? minoverlap : 0;
    }
    return rc;
}

int nrb200_set_exclude(nrb200_handle *h, int64_t n, const int32_t *seqid, const int32_t *start, const int32_t *end,
                       const int8_t *strand) {
    return set_side(h, h->excl, "exclude", n, seqid, start, end, strand);
}

int nrb200_set_exclude_option(nrb200_handle *h, int32_t option) {
    if (!h) return NRB200_ERR_INVALID;
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
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (!h->have_y) return fail(h, NRB200_ERR_STATE, "call nrb200_set_ranges() first (the plan needs seqlengths)");
    Plan p;
    const std::string msg = build_plan(*spec, h->seqlen, p);
    if (!msg.empty()) return fail(h, NRB200_ERR_INVALID, msg);
    if (p.n_blocks() > (int64_t)INT32_MAX / 2) return fail(h, NRB200_ERR_UNSUPPORTED, "too many blocks per iteration");
    h->plan = std::move(p);
    const Plan &pl = h->plan;
    const size_t B = (size_t)pl.n_blocks();
    int rc;
    if ((rc = upload(h, h->tile_seq, pl.tile_seq.data(), B))) return rc;
    if ((rc = upload(h, h->tile_start, pl.tile_start.data(), B))) return rc;
    if ((rc = upload(h, h->bait_width, pl.bait_width.data(), B))) return rc;
    // sampler tables packed into one int64 and one int32 buffer
    std::vector<int64_t> i64;
    std::vector<int32_t> i32;
    auto put64 = [&](const std::vector<int64_t> &v) { size_t o = i64.size(); i64.insert(i64.end(), v.begin(), v.end()); return o; };
    auto put32 = [&](const std::vector<int32_t> &v) { size_t o = i32.size(); i32.insert(i32.end(), v.begin(), v.end()); return o; };
    const size_t o_cum = put64(pl.seq_cum), o_tps = put64(pl.tiles_per_seq), o_sbo = put64(pl.state_block_off),
                 o_sso = put64(pl.state_seg_off), o_sl = put64(pl.state_len), o_sbl = put64(pl.state_block_len),
                 o_sgc = put64(pl.sg_cum);
    const size_t o_sq = put32(pl.sg_seq), o_ss = put32(pl.sg_start), o_st = put32(pl.sg_tiles);
    if ((rc = upload(h, h->sampler_i64, i64.data(), i64.size()))) return rc;
    if ((rc = upload(h, h->sampler_i32, i32.data(), i32.size()))) return rc;
    const int64_t *b64 = h->sampler_i64.as<int64_t>();
    const int32_t *b32 = h->sampler_i32.as<int32_t>();
    h->sampler_view = SamplerView{pl.kind, pl.n_seq, pl.n_states, pl.block_length, pl.genome_length,
                                  b64 + o_cum, b64 + o_tps, b64 + o_sbo, b64 + o_sso, b64 + o_sl, b64 + o_sbl,
                                  b64 + o_sgc, b32 + o_sq, b32 + o_ss, b32 + o_st};
    h->have_plan = true;
    h->slices_dirty = true;
    if (n_blocks) *n_blocks = pl.n_blocks();
    return NRB200_OK;
}

int nrb200_get_tiles(const nrb200_handle *h, int32_t *tile_seqid, int32_t *tile_start, int32_t *bait_width) {
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
    NRB_CUDA(h, cudaSetDevice(h->device));
    return run_counts_core(h, draws, want_feature_counts != 0, nullptr, stats);
}

int nrb200_resident_pointers(const nrb200_handle *h, void **iter_nranges_dev, void **iter_totals_dev,
                             void **feature_counts_dev) {
    if (!h) return NRB200_ERR_INVALID;
    if (iter_nranges_dev) *iter_nranges_dev = h->d_iter_nranges.p;
    if (iter_totals_dev) *iter_totals_dev = h->d_iter_totals.p;
    if (feature_counts_dev) *feature_counts_dev = h->res_has_counts ? h->d_counts.p : nullptr;
    return NRB200_OK;
}

int nrb200_run_counts(nrb200_handle *h, const nrb200_draws *draws, int64_t *iter_nranges, int64_t *iter_totals,
                      int32_t *feature_counts, nrb200_stats *stats) {
    if (!h) return NRB200_ERR_INVALID;
    NRB_CUDA(h, cudaSetDevice(h->device));
    int rc = run_counts_core(h, draws, feature_counts != nullptr, feature_counts, stats);
    if (rc) return rc;
    const int64_t R = draws->n_iter, G = h->query.n;
    if ((rc = fetch_u64_as_i64(h, h->d_iter_nranges, R, iter_nranges))) return rc;
    if ((rc = fetch_u64_as_i64(h, h->d_iter_totals, R, iter_totals))) return rc;
    if (int rc_sync = sync_stream(h)) return rc_sync;
    if (h->copy_stream) NRB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    if (stats) stats->d2h_bytes = (iter_nranges ? R * 8 : 0) + (iter_totals ? R * 8 : 0) + (feature_counts ? R * G * 4 : 0);
    return NRB200_OK;
}

int nrb200_moments_reset(nrb200_handle *h) {
    if (!h) return NRB200_ERR_INVALID;
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
    return NRB200_OK;
}

int nrb200_run_moments(nrb200_handle *h, const nrb200_draws *draws, int64_t batch_iter, const int32_t *observed,
                       int64_t *iter_nranges, int64_t *iter_totals, nrb200_stats *stats) {
    if (!h || !draws) return NRB200_ERR_INVALID;
    NRB_CUDA(h, cudaSetDevice(h->device));
    const int64_t G = h->query.n;
    if (G == 0) return fail(h, NRB200_ERR_STATE, "moments need a query (nrb200_set_query)");
    int rc;
    if (!h->moments_ready && (rc = nrb200_moments_reset(h))) return rc;
    if (observed && (rc = upload(h, h->m_observed, observed, (size_t)G))) return rc;
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
    if (!h->moments_ready) return fail(h, NRB200_ERR_STATE, "no moments accumulated");
    NRB_CUDA(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->query.n * 8;
    if (feat_sum) NRB_CUDA(h, cudaMemcpyAsync(feat_sum, h->m_sum.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (feat_sumsq) NRB_CUDA(h, cudaMemcpyAsync(feat_sumsq, h->m_sumsq.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (feat_n_ge_observed) NRB_CUDA(h, cudaMemcpyAsync(feat_n_ge_observed, h->m_nge.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (int rc_sync = sync_stream(h)) return rc_sync;
    return NRB200_OK;
}

int nrb200_set_weights(nrb200_handle *h, const double *weights) {
    if (!h) return NRB200_ERR_INVALID;
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

int nrb200_run_weighted(nrb200_handle *h, const nrb200_draws *draws, double *iter_sums, double *feature_sums, nrb200_stats *stats) {
    if (!h || !draws) return NRB200_ERR_INVALID;
    NRB_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = prepare(h))) return rc;
    if (!h->have_weights) return fail(h, NRB200_ERR_STATE, "nrb200_set_weights() has not been called");
    if (h->query.n == 0) return fail(h, NRB200_ERR_STATE, "nrb200_set_query() has not been called");
    if (!rank_path_ok(h, true))
        return fail(h, NRB200_ERR_UNSUPPORTED, "weighted sums run on the rank path only (segmentation within seqlengths, no NRB200_FLAG_FORCE_STREAM)");
    if ((rc = build_weight_index(h))) return rc;
    NRB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if ((rc = stage_draws(h, draws))) return rc;
    const int64_t R = draws->n_iter, G = h->query.n, B = h->plan.n_blocks();
    const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(h->tune_max_batch, ((int64_t)1 << 27) / std::max<int64_t>(G, 1)));
    NRB_CUDA(h, h->d_iter_nranges.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->d_n_hits.reserve(16));
    NRB_CUDA(h, h->ws_iter.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->ws_sums.reserve((size_t)std::min(batch, std::max<int64_t>(R, 1)) * G * 8));
    NRB_CUDA(h, h->draw_tab.reserve((size_t)std::min(batch, std::max<int64_t>(R, 1)) * B * sizeof(int2)));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_nranges.p, 0, R * 8, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->d_n_hits.p, 0, 16, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->ws_iter.p, 0, R * 8, h->stream));
    BootParams P = make_params(h, draws);
    P.iter_nranges = h->d_iter_nranges.as<unsigned long long>();
    P.n_hits = h->d_n_hits.as<unsigned long long>();
    P.draw_tab = h->draw_tab.as<int2>();
    const FeatView F{h->f_src.as<int32_t>(), h->pair_off.as<int32_t>(), h->pair_rec.as<PairRec>()};
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
        if (feature_sums)
            NRB_CUDA(h, cudaMemcpyAsync(feature_sums + r0 * G, h->ws_sums.p, (size_t)n * G * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    NRB_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    if (iter_sums && R > 0) NRB_CUDA(h, cudaMemcpyAsync(iter_sums, h->ws_iter.p, R * 8, cudaMemcpyDeviceToHost, h->stream));
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

int nrb200_count_observed(nrb200_handle *h, int32_t *feature_counts, int64_t *total) {
    return nrb200_count_observed_ex(h, 0, feature_counts, total);
}

int nrb200_count_observed_ex(nrb200_handle *h, int32_t minoverlap, int32_t *feature_counts, int64_t *total) {
    if (!h) return NRB200_ERR_INVALID;
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
    DevMem counts, tot;
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
    NRB_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = prepare(h))) return rc;
    if ((rc = ensure_pmax_table(h))) return rc;  // the fill kernel locates the hit windows
    NRB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if ((rc = stage_draws(h, draws))) return rc;
    const int64_t R = draws->n_iter, B = h->plan.n_blocks(), items = R * B;
    NRB_CUDA(h, h->d_iter_nranges.reserve(std::max<int64_t>(R, 1) * 8));
    NRB_CUDA(h, h->d_n_hits.reserve(16));  // [0] hits, [1] fill-overflow flag
    NRB_CUDA(h, h->item_rows.reserve(std::max<int64_t>(items, 1) * 4));
    NRB_CUDA(h, h->within_off.reserve(std::max<int64_t>(items, 1) * 4));
    NRB_CUDA(h, h->iter_off_dev.reserve((R + 1) * 8));
    NRB_CUDA(h, cudaMemsetAsync(h->d_iter_nranges.p, 0, R * 8, h->stream));
    NRB_CUDA(h, cudaMemsetAsync(h->d_n_hits.p, 0, 16, h->stream));
    BootParams P = make_params(h, draws);
    P.n_query = 0;
    P.iter_nranges = h->d_iter_nranges.as<unsigned long long>();
    P.iter_totals = nullptr;
    P.item_rows = h->item_rows.as<int32_t>();
    P.n_hits = h->d_n_hits.as<unsigned long long>();
    NRB_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    int count_launches = 0;
    const bool rank_rows = rank_path_ok(h, false);
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
    for (int64_t r = 0; r < R; ++r) offs[r + 1] = offs[r] + per_iter[r];
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
                            h->item_rows.as<int32_t>() + r0 * B, reinterpret_cast<int *>(h->d_n_hits.as<char>() + 8)};
            const dim3 grid((unsigned)std::min<int64_t>((B + 7) / 8, 65535), (unsigned)Q.n_iter);
            if (excl == kExclTrim) boot_fill_trim_kernel<<<grid, 256, 0, h->stream>>>(Q, O);
            else if (st) boot_fill_kernel<true, kExclDrop><<<grid, 256, 0, h->stream>>>(Q, O);
            else if (excl == kExclDrop) boot_fill_kernel<false, kExclDrop><<<grid, 256, 0, h->stream>>>(Q, O);
            else boot_fill_kernel<false, kExclNone><<<grid, 256, 0, h->stream>>>(Q, O);
            NRB_CUDA(h, cudaGetLastError());
        }
    }
    NRB_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    unsigned long long hits_and_flag[2] = {0, 0};
    NRB_CUDA(h, cudaMemcpyAsync(hits_and_flag, h->d_n_hits.p, 16, cudaMemcpyDeviceToHost, h->stream));
    if (int rc_sync = sync_stream(h)) return rc_sync;
    if (hits_and_flag[1] != 0) {
        h->n_rows = 0;
        return fail(h, NRB200_ERR_STATE, "internal error: the count and fill passes of the table disagree (no row was written out of place)");
    }
    h->n_rows = rows;
    if (iter_offsets) std::memcpy(iter_offsets, offs.data(), (size_t)(R + 1) * 8);
    if (stats) {
        float k_ms = 0, t_ms = 0;
        cudaEventElapsedTime(&k_ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&t_ms, h->ev[0], h->ev[2]);
        const unsigned long long hits = hits_and_flag[0];
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
    NRB_CUDA(h, cudaSetDevice(h->device));
    if (row0 < 0 || n < 0 || row0 + n > h->n_rows) return fail(h, NRB200_ERR_INVALID, "fetch_rows: range outside the materialised table");
    struct { int32_t *dst; DevMem *src; } cols[] = {{seqid, &h->row_seq}, {start, &h->row_start}, {end, &h->row_end},
                                                     {src_idx, &h->row_src}, {block, &h->row_block}};
    for (auto &c : cols)
        if (c.dst && n)
            NRB_CUDA(h, cudaMemcpyAsync(c.dst, c.src->as<int32_t>() + row0, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (int rc_sync = sync_stream(h)) return rc_sync;
    return NRB200_OK;
}

}  // extern "C"
