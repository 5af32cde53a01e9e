// libnrb200: the handle (device buffers, host mirrors, tuning knobs) and small helpers shared by the
// parts of the engine.  (textually included by engine.cu)
#pragma once
#include <algorithm>
#include <chrono>
#include <sys/mman.h>
#include <sched.h>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <omp.h>
#include <nvtx3/nvToolsExt.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/nrb200.h"
#include "kernels.cuh"
#include "plan.h"

namespace nrb {

void widen_u16_to_i32(const uint16_t *src, int32_t *dst, size_t n);  // host_simd.cpp
const char *widen_variant();
void copy_streaming(const void *src, void *dst, size_t bytes);  // host_simd.cpp: memcpy with non-temporal stores (staging for DMA)

// NVTX ranges around the stages of a call (upload, index build, window lists, chunks, drain): visible in Nsight
// Systems / Compute timelines, no-ops when no tool is attached (NVTX3 is header-only).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct GroupSpin {
    static void pause() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
};

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

// NRB200_TRACE_DEV=1: device-side timeline (development aid).  mark() records a timing event on a stream together
// with the host clock; dump() prints, for every mark, when the host queued it and when the device reached it, both
// relative to the first mark.
struct DevTimeline {
    static bool on() {
        static const bool v = getenv("NRB200_TRACE_DEV") != nullptr;
        return v;
    }
    struct Mark {
        const char *what;
        cudaEvent_t ev;
        std::chrono::steady_clock::time_point host;
    };
    std::vector<Mark> marks;
    std::vector<cudaEvent_t> pool;
    void mark(const char *what, cudaStream_t s) {
        if (!on()) return;
        cudaEvent_t e = nullptr;
        if (!pool.empty()) {
            e = pool.back();
            pool.pop_back();
        } else if (cudaEventCreate(&e) != cudaSuccess) {
            return;
        }
        cudaEventRecord(e, s);
        marks.push_back({what, e, std::chrono::steady_clock::now()});
    }
    void dump() {
        if (!on() || marks.empty()) return;
        for (auto &m : marks) cudaEventSynchronize(m.ev);
        for (auto &m : marks) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[0].ev, m.ev);
            fprintf(stderr, "[nrb200-dev] %-28s queued %7.3f ms   reached %7.3f ms\n", m.what,
                    std::chrono::duration<double, std::milli>(m.host - marks[0].host).count(), ms);
            pool.push_back(m.ev);
        }
        marks.clear();
    }
};

static thread_local std::string g_create_error;

// Growth-only device buffer.
// NRB200_GUARD=1 (development aid; compute-sanitizer is not available on the test pool): every buffer gets a
// 256-byte tail filled with a pattern, checked whenever the buffer is released; an out-of-bounds WRITE behind
// any device buffer aborts with a message.
inline bool guard_on() {
    static const bool v = getenv("NRB200_GUARD") != nullptr;
    return v;
}
constexpr size_t kGuardBytes = 256;

struct DevMem {
    void *p = nullptr;
    size_t cap = 0;
    ~DevMem() { release(); }
    void check_guard() const {
        if (!p || !guard_on()) return;
        unsigned char tail[kGuardBytes];
        if (cudaMemcpy(tail, static_cast<char *>(p) + cap, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return;
        for (size_t i = 0; i < kGuardBytes; ++i)
            if (tail[i] != 0xA5) {
                fprintf(stderr, "[nrb200] GUARD: write past the end of a device buffer of %zu bytes (offset +%zu)\n", cap, i);
                abort();
            }
    }
    void release() {
        check_guard();
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        const size_t want = bytes ? bytes : 1;
        cudaError_t e = cudaMalloc(&p, want + (guard_on() ? kGuardBytes : 0));
        if (e == cudaSuccess) {
            cap = want;
            if (guard_on()) cudaMemset(static_cast<char *>(p) + cap, 0xA5, kGuardBytes);
        }
        return e;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    void swap(DevMem &o) {
        std::swap(p, o.p);
        std::swap(cap, o.cap);
    }
    DevMem() = default;
    DevMem(const DevMem &) = delete;  // owns its allocation
    DevMem &operator=(const DevMem &) = delete;
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

// Window lists of the rank path as the host built them (lists.cuh): kept by the leader of a device group so that the
// other members upload the same lists instead of building them again.
struct HostLists {
    bool valid = false;
    int64_t n_pairs = 0;
    std::vector<int32_t> f_src, f_off, x_off, x_rec;
    std::vector<int32_t> recs;  // 4 words per record
};

struct GroupPool;

}  // namespace nrb

using namespace nrb;

struct nrb200_handle {
    // ---- device group (nrb200_config.n_devices > 1): this handle owns no GPU state, it shards the iterations over
    // its members (one nrb200_handle per GPU, each driven by its own host thread) -- see group.cuh
    std::vector<nrb200_handle *> members;
    GroupPool *pool = nullptr;
    void *g_stage = nullptr;  // pinned copy of a pageable y, so that it crosses the host memory once for all members
    size_t g_stage_cap = 0;
    // ---- member of a group
    nrb200_handle *lists_from = nullptr;  // take the window lists this handle (the group's leader) built
    bool keep_lists = false;              // leader: keep the host copy of the lists for the others
    HostLists lists;
    int omp_threads = 0;                  // host threads of this handle's parallel loops (0 = all)

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
    cudaEvent_t chunk_ev[64] = {};
    cudaEvent_t copied_ev[64] = {};   // pageable destinations: chunk k has reached the pinned staging buffer
    void *stage_buf = nullptr;        // that buffer (pinned, grown on demand)
    size_t stage_cap = 0;
    void *stage_in = nullptr;         // pinned staging for large pageable INPUT columns (4 regions)
    size_t stage_in_cap = 0;
    // device groups, leader only, set around prepare(): runs body(0..n-1) on the driver threads of ALL members, each with its own
    // host team (the window lists are built once, by every core of the box, without a team larger than a member's share)
    std::function<void(int, const std::function<void(int)> &)> seq_runner;
    bool y_preloaded = false;         // device groups: the group has already put y's columns on this GPU (group_scatter_y)
    cudaEvent_t y_ev = nullptr;       // device groups: this member's slice of y has reached every member
    void *m_stage = nullptr;          // device groups: pinned landing place of this member's three per-feature moment vectors
    size_t m_stage_cap = 0;
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
    DevTimeline timeline;
    // nrb200_set_ranges_begin() .. _end(): the upload and the inspection are queued, their verdict not yet read
    bool y_pending = false;
    struct {
        int64_t n = 0;
        const int32_t *seqid = nullptr, *start = nullptr, *end = nullptr;
        const int8_t *strand = nullptr;
        bool on_device = false;  // false: degenerate input, the host path takes it at _end()
    } pend;
    int32_t *inspect_host = nullptr;  // pinned landing place of the inspection verdict (+ the initial values)
    size_t inspect_host_words = 0;
    DevMem raw_seqid, raw_start, raw_end, raw_strand;  // caller-order copies of an unsorted y while it is being sorted
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
    // tuning knobs (environment, read at create)
    bool tune_end_rank = true;         // NRB200_END_RANK=0: always radix-sort the ends (index build)
    int tune_ipt = 4;                  // NRB200_IPT: iterations a thread of the count kernel stays on its feature (<= 64)
    int tune_loop_minb = 12;           // NRB200_LOOP_MINB: its min CTAs per SM (16 = 32 registers, 12 = 40)
    bool tune_stream_old = false;      // NRB200_STREAM_OLD=1: streaming path through boot_count_kernel only (A/B)
    int tune_rows_ipc = 4;             // NRB200_ROWS_IPC: iterations a CTA of the rows kernel loops over
    double tune_table_mult = 4.0;
    int tune_wire16 = 2;               // NRB200_WIRE16: uint16 wire format of a host-bound matrix (0 never, 1 always, 2 = where it pays: run.cuh)
    int group_w = 1;                   // members of the device group this handle belongs to (1: a handle of its own)
    int tune_chunks = 8;               // NRB200_CHUNKS: pipeline depth of nrb200_run_counts
    int64_t tune_max_batch = 2048;     // NRB200_MAX_BATCH: iterations per launch group
    int64_t tune_table_cap = 8 << 20;  // NRB200_TABLE_CAP: the multiplier applies up to this many entries

    Plan plan;
    DevMem tile_seq, tile_start, bait_width, tile_info, sampler_i64, sampler_i32;
    SamplerView sampler_view{};
    bool slices_dirty = true, have_pmax_table = false;
    DevMem q_tile_lo, q_tile_hi, x_tile_lo, x_tile_hi, pair_off, pair_rec, f_src, tile_x_off, tile_x_rec;
    // streaming path, tile-group form (boot_stream_kernel): y interleaved, tiles in genome order, groups of tiles
    DevMem y_se, tiles_sorted_dev, groups_dev;
    bool have_se = false, groups_ok = false;
    int64_t n_groups = 0;
    int group_slice_cap = 0;
    std::vector<int32_t> q_lo_host, q_hi_host;
    int64_t n_pairs = 0;
    // tiles of each sequence in ascending start order, with the running max of their ends (made by nrb200_set_plan
    // for build_pairs)
    struct TilesBySeq {
        std::vector<std::vector<int32_t>> id;
        std::vector<std::vector<int64_t>> ts, te, te_max;
    } tiles_by_seq;
    bool tiles_in_bounds = false;

    // run scratch / results
    DevMem d_bait_seq, d_bait_start, d_dst_tile, perm_vals_a, perm_vals_b, np_seg, np_start, np_state, np_list, np_cnt;
    DevMem d_iter_nranges, d_iter_totals, d_counts, d_counts16, redo_scratch, d_n_hits;
    int *flags_host = nullptr;         // pinned: per-chunk overflow flags of the uint16 wire format
    uint16_t *out16 = nullptr;         // set by nrb200_run_counts_u16 around its run: the caller's uint16 matrix
    int out16_overflow = 0;            // ... some count did not fit (it was stored as 65535)
    bool last_wire16 = false;          // the last host-bound matrix travelled as uint16
    bool res_has_counts = false;
    // moments
    DevMem m_sum, m_sumsq, m_nge, m_observed, obs_counts, obs_total;
    DevMem wm_sum, wm_sumsq, wm_nge, wm_observed;  // moments of the weighted sums (double; n_ge int64)
    bool moments_ready = false;
    int64_t moments_G = 0;  // features the moment buffers were sized (and zeroed) for
    // rows
    DevMem item_rec, item_rows, within_off, iter_off_dev, row_seq, row_start, row_end, row_src, row_block;
    // device group, gathered result (group.cuh): every member holds the full [R x G] matrix and [R] vectors
    DevMem gat_counts, gat_nranges, gat_totals;
    int64_t gat_R = 0, gat_G = 0;
    struct PeerOut {
        int n = 0;
        int32_t *counts[kMaxPeers] = {};
    } peer_out;                       // set by the group around a gathered run of this member
    bool peers_enabled = false;       // group handle: every pair of members can store into each other's memory
    int peers_state = 0;              // group handle: 0 = not probed, 1 = peer access is on, -1 = not available
    int64_t n_rows = 0;
    int64_t dst_tile_items = 0;  // [iterations x blocks] the last run left in d_dst_tile (permute kinds)
};

namespace {

inline int host_threads(const nrb200_handle *h) { return h && h->omp_threads > 0 ? h->omp_threads : omp_get_max_threads(); }

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
        if (bytes >= ((size_t)64 << 10)) copy_streaming(src, h->arena + at, bytes);  // (a DMA engine reads it next: keep it out of the caches)
        else std::memcpy(h->arena + at, src, bytes);
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

// A large pageable destination is usually freshly allocated: ask for huge pages before the first touch, so that the
// staging copies below fault in 2 MB at a time instead of 4 KB (no effect when the system does not allow it).
inline void advise_huge_pages(void *dst, size_t bytes) {
    static const bool off = getenv("NRB200_HUGEPAGES") && atoi(getenv("NRB200_HUGEPAGES")) == 0;  // NRB200_HUGEPAGES=0: leave the pages as they are (A/B)
    if (off) return;
    constexpr uintptr_t kHuge = (uintptr_t)2 << 20;
    const uintptr_t lo = ((uintptr_t)dst + kHuge - 1) & ~(kHuge - 1), hi = ((uintptr_t)dst + bytes) & ~(kHuge - 1);
    if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
}

// Fault in the pages of [p, p + bytes) for writing with ONE call (MADV_POPULATE_WRITE, Linux 5.14+): the kernel maps and
// zeroes them without a trap per page.  Returns false where the call is not available; the caller then touches the pages.
inline bool populate_pages(void *p, size_t bytes) {
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
    if (!bytes) return true;
    const uintptr_t lo = (uintptr_t)p & ~(uintptr_t)4095, hi = ((uintptr_t)p + bytes + 4095) & ~(uintptr_t)4095;
    return madvise((void *)lo, hi - lo, MADV_POPULATE_WRITE) == 0;
}

// Device -> host copy into a caller-owned buffer, complete on return.  Pinned destinations take one DMA; large
// pageable ones (R's vectors, plain numpy arrays) go through a double-buffered pinned staging area: chunk i
// travels while chunk i - 1 is moved on with a parallel memcpy.
inline int download(nrb200_handle *h, void *dst, const void *src_dev, size_t bytes) {
    if (!bytes) return NRB200_OK;
    cudaPointerAttributes attr{};
    const bool pageable = cudaPointerGetAttributes(&attr, dst) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    constexpr size_t kChunk = (size_t)32 << 20;
    if (pageable && bytes >= ((size_t)1 << 20)) {
        advise_huge_pages(dst, bytes);
        if (h->stage_cap < 2 * kChunk) {
            if (h->stage_buf) cudaFreeHost(h->stage_buf);
            h->stage_buf = nullptr;
            h->stage_cap = 0;
            if (cudaMallocHost((void **)&h->stage_buf, 2 * kChunk) == cudaSuccess) h->stage_cap = 2 * kChunk;
            else cudaGetLastError();
        }
    }
    if (!pageable || bytes < ((size_t)1 << 20) || h->stage_cap < 2 * kChunk) {
        NRB_CUDA(h, cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, h->stream));
        NRB_CUDA(h, cudaStreamSynchronize(h->stream));
        return NRB200_OK;
    }
    for (int k = 0; k < 2; ++k)
        if (!h->copied_ev[k]) NRB_CUDA(h, cudaEventCreateWithFlags(&h->copied_ev[k], cudaEventDisableTiming));
    auto move_out = [&](size_t chunk) {
        const size_t off = chunk * kChunk, len = std::min(kChunk, bytes - off);
        const char *from = static_cast<const char *>(h->stage_buf) + (chunk & 1) * kChunk;
        char *to = static_cast<char *>(dst) + off;
        const size_t piece = (size_t)1 << 20;
        const int64_t pieces = (int64_t)((len + piece - 1) / piece);
#pragma omp parallel for schedule(static) num_threads(host_threads(h))
        for (int64_t i = 0; i < pieces; ++i) std::memcpy(to + i * piece, from + i * piece, std::min(piece, len - (size_t)i * piece));
    };
    const size_t chunks = (bytes + kChunk - 1) / kChunk;
    for (size_t c = 0; c < chunks; ++c) {
        const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
        NRB_CUDA(h, cudaMemcpyAsync(static_cast<char *>(h->stage_buf) + (c & 1) * kChunk, static_cast<const char *>(src_dev) + off, len,
                                    cudaMemcpyDeviceToHost, h->stream));
        NRB_CUDA(h, cudaEventRecord(h->copied_ev[c & 1], h->stream));
        if (c > 0) {  // chunk c - 1 has its own slot: move it out while chunk c is on the wire
            NRB_CUDA(h, cudaEventSynchronize(h->copied_ev[(c - 1) & 1]));
            move_out(c - 1);
        }
    }
    NRB_CUDA(h, cudaEventSynchronize(h->copied_ev[(chunks - 1) & 1]));
    move_out(chunks - 1);
    return NRB200_OK;
}

// Host -> device copy of a (large) caller-owned array.  Pinned memory goes straight to the DMA engine.  Pageable
// memory (R's vectors, plain numpy arrays) would be staged by the driver at a few GB/s: it is copied into a pinned
// buffer of the handle with a parallel memcpy first.  `slot` (0..3) picks the region of that buffer, so that the
// columns of one call do not overwrite each other before the stream is synchronised.
template <class T>
int upload_direct(nrb200_handle *h, DevMem &dst, const T *src, size_t n, int slot = 0) {
    const size_t bytes = n * sizeof(T);
    NRB_CUDA(h, dst.reserve(bytes));
    if (!n) return NRB200_OK;
    const void *from = src;
    if (bytes >= ((size_t)256 << 10) && bytes <= ((size_t)256 << 20)) {
        cudaPointerAttributes attr{};
        const bool pageable = cudaPointerGetAttributes(&attr, src) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        if (pageable) {
            // regions are sized for 8-byte elements whatever T is: the columns of one call (same n) never overlap
            const size_t region = (n * 8 + 4095) & ~(size_t)4095, need = 4 * region;
            if (h->stage_in_cap < need) {
                if (h->stage_in) cudaFreeHost(h->stage_in);
                h->stage_in = nullptr;
                h->stage_in_cap = 0;
                if (cudaMallocHost((void **)&h->stage_in, need) == cudaSuccess) h->stage_in_cap = need;
                else cudaGetLastError();
            }
            if (h->stage_in_cap >= need) {
                char *to = static_cast<char *>(h->stage_in) + (size_t)slot * region;
                const size_t piece = (size_t)1 << 20;
                const int64_t pieces = (int64_t)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static) num_threads(host_threads(h))
                for (int64_t i = 0; i < pieces; ++i)
                    copy_streaming(reinterpret_cast<const char *>(src) + i * piece, to + i * piece, std::min(piece, bytes - (size_t)i * piece));
                from = to;
            }
        }
    }
    NRB_CUDA(h, cudaMemcpyAsync(dst.p, from, bytes, cudaMemcpyHostToDevice, h->stream));
    return NRB200_OK;
}

// Several caller-owned columns of one call (the columns of y).  Pinned memory goes straight to the DMA engine.  Pageable
// memory is staged through the handle's pinned buffer by ONE team of host threads for all columns: every column is
// copied in parallel and handed to the DMA engine by the team's master as soon as it is complete, while the others
// already copy the next column.
struct HostColumn {
    DevMem *dst;
    const void *src;
    size_t bytes;
};
inline int upload_columns(nrb200_handle *h, const HostColumn *cols, int n_cols) {
    size_t total = 0, largest = 0;
    for (int c = 0; c < n_cols; ++c) {
        NRB_CUDA(h, cols[c].dst->reserve(cols[c].bytes));
        total += (cols[c].bytes + 4095) & ~(size_t)4095;
        largest = std::max(largest, cols[c].bytes);
    }
    bool staged = false;
    if (largest >= ((size_t)256 << 10) && total <= ((size_t)1 << 30)) {
        cudaPointerAttributes attr{};
        const bool pageable = cudaPointerGetAttributes(&attr, cols[0].src) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        if (pageable) {
            if (h->stage_in_cap < total) {
                if (h->stage_in) cudaFreeHost(h->stage_in);
                h->stage_in = nullptr;
                h->stage_in_cap = 0;
                if (cudaMallocHost((void **)&h->stage_in, total) == cudaSuccess) h->stage_in_cap = total;
                else cudaGetLastError();
            }
            staged = h->stage_in_cap >= total;
        }
    }
    if (!staged) {
        for (int c = 0; c < n_cols; ++c)
            if (cols[c].bytes) NRB_CUDA(h, cudaMemcpyAsync(cols[c].dst->p, cols[c].src, cols[c].bytes, cudaMemcpyHostToDevice, h->stream));
        return NRB200_OK;
    }
    cudaError_t err = cudaSuccess;
    const size_t piece = (size_t)256 << 10;
#pragma omp parallel num_threads(host_threads(h))
    {
        size_t at = 0;
        for (int c = 0; c < n_cols; ++c) {
            char *to = static_cast<char *>(h->stage_in) + at;
            const char *from = static_cast<const char *>(cols[c].src);
            const size_t bytes = cols[c].bytes;
            const int64_t pieces = (int64_t)((bytes + piece - 1) / piece);
#pragma omp for schedule(static)
            for (int64_t i = 0; i < pieces; ++i) copy_streaming(from + i * piece, to + i * piece, std::min(piece, bytes - (size_t)i * piece));
#pragma omp master
            {
                if (bytes) {
                    const cudaError_t e = cudaMemcpyAsync(cols[c].dst->p, to, bytes, cudaMemcpyHostToDevice, h->stream);
                    if (e != cudaSuccess) err = e;
                    h->timeline.mark("column uploaded", h->stream);
                }
            }
            at += (bytes + 4095) & ~(size_t)4095;
        }
    }
    if (err != cudaSuccess) return fail(h, NRB200_ERR_CUDA, std::string("cudaMemcpyAsync (staged upload): ") + cudaGetErrorString(err));
    return NRB200_OK;
}

// The H statistic: kHitSlots partial sums followed by one flag word (the table's fill-overflow flag).
constexpr size_t kHitBytes = ((size_t)64 + 1) * 8;
inline int hits_reset(nrb200_handle *h) {
    NRB_CUDA(h, h->d_n_hits.reserve(kHitBytes));
    NRB_CUDA(h, cudaMemsetAsync(h->d_n_hits.p, 0, kHitBytes, h->stream));
    return NRB200_OK;
}
// queues the copy; after the next synchronisation hits_total() sums what landed in `buf` ([65] words)
inline int hits_fetch(nrb200_handle *h, unsigned long long *buf) {
    NRB_CUDA(h, cudaMemcpyAsync(buf, h->d_n_hits.p, kHitBytes, cudaMemcpyDeviceToHost, h->stream));
    return NRB200_OK;
}
inline unsigned long long hits_total(const unsigned long long *buf) {
    unsigned long long s = 0;
    for (int i = 0; i < 64; ++i) s += buf[i];
    return s;
}

}  // namespace
