// libnrb200: device groups -- one handle that shards the bootstrap iterations over several GPUs of the box.
// (textually included by engine.cu, before api.cuh)
//
// The reference's iterations are independent (replicate(R, ...), R/bootRanges.R:90) and its result is ONE table /
// matrix for one caller (R/bootRanges.R:120-122).  A handle created with nrb200_config.n_devices > 1 keeps one
// ordinary single-GPU handle per device ("members") and serves the same entry points:
//   * inputs (y, query, exclude, plan, weights) go to every member from the caller's ONE host copy, in parallel
//     (one host thread and one PCIe link per GPU); a pageable y is first copied into pinned memory once;
//   * the window lists of the rank path are built once, by the leader (member 0), and uploaded by the others;
//   * member k runs iterations [ceil(k R / W), ceil((k + 1) R / W)) -- Philox is keyed by the GLOBAL iteration, so the
//     result does not depend on W -- and copies its rows of the count matrix straight into the caller's matrix
//     (W device-to-host copies in parallel, no collective: the result lives on the host);
//   * per-feature moments are summed over the members on the host (3 G words per member).
// The caller's thread drives member 0 itself; members 1.. have a worker thread each.  R's API is single-threaded:
// workers never touch anything but the plain C arrays the caller handed in.
namespace nrb {

struct GroupPool {
    int n_workers = 0;                      // worker threads
    int first = 1;                          // member driven by worker 0's thread (1: the caller drives member 0 itself; 0: workers drive all)
    std::vector<std::vector<int>> cpus;     // per member: CPUs to bind its driver thread (and so its host team) to; empty = unbound
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> generation{0};
    std::atomic<int> pending{0};
    std::atomic<bool> stop{false};
    std::function<int(int)> task;           // task(member index) -> status
    std::vector<int> rc;

    static void spin_pause() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }

    void worker(int k) {
        if (k < (int)cpus.size() && !cpus[k].empty()) {  // before the thread's first parallel region: its host team inherits the mask
            cpu_set_t set;
            CPU_ZERO(&set);
            for (int c : cpus[k]) CPU_SET(c, &set);
            sched_setaffinity(0, sizeof set, &set);
        }
        uint64_t seen = 0;
        for (;;) {
            // calls of one front-end step follow each other within microseconds: spin briefly before sleeping
            bool got = false;
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0;; ++i) {
                if (generation.load(std::memory_order_acquire) != seen || stop.load(std::memory_order_acquire)) {
                    got = true;
                    break;
                }
                spin_pause();
                if ((i & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) break;
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return generation.load(std::memory_order_acquire) != seen || stop.load(std::memory_order_acquire); });
            }
            if (stop.load(std::memory_order_acquire)) return;
            seen = generation.load(std::memory_order_acquire);
            rc[k] = task(k);
            pending.fetch_sub(1, std::memory_order_acq_rel);
        }
    }

    // members: group size; caller_drives_first: member 0 runs on the calling thread (else every member has a worker)
    void start(int members, bool caller_drives_first) {
        first = caller_drives_first ? 1 : 0;
        n_workers = members - first;
        rc.assign((size_t)members, 0);
        for (int k = first; k < members; ++k) threads.emplace_back([this, k] { worker(k); });
    }

    // Runs fn(k) for every member k; returns when all are done.
    void run(const std::function<int(int)> &fn) {
        task = fn;
        pending.store(n_workers, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(mu);
            generation.fetch_add(1, std::memory_order_acq_rel);
        }
        cv.notify_all();
        if (first == 1) rc[0] = fn(0);
        for (int i = 0; pending.load(std::memory_order_acquire) > 0; ++i) {
            spin_pause();
            if ((i & 1023) == 1023) std::this_thread::yield();
        }
    }

    ~GroupPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop.store(true, std::memory_order_release);
        }
        cv.notify_all();
        for (auto &t : threads) t.join();
    }
};

}  // namespace nrb

namespace {

inline bool is_group(const nrb200_handle *h) { return h && !h->members.empty(); }
inline int group_size(const nrb200_handle *h) { return (int)h->members.size(); }

// [lo, hi) of member k: iterations {r : floor(r W / R) == k}  (the same split as nullranges_b200/sharding.py)
inline void shard_iterations(int64_t R, int W, int k, int64_t &lo, int64_t &hi) {
    lo = ((int64_t)k * R + W - 1) / W;
    hi = ((int64_t)(k + 1) * R + W - 1) / W;
}

// fn(k, member) on every member in parallel; the first failure's status and message become the group's.
int group_all(nrb200_handle *g, const std::function<int(int, nrb200_handle *)> &fn) {
    g->pool->run([&](int k) { return fn(k, g->members[k]); });
    for (int k = 0; k < group_size(g); ++k)
        if (g->pool->rc[k] != NRB200_OK) {
            g->error = "GPU " + std::to_string(g->members[k]->device) + ": " + g->members[k]->error;
            return g->pool->rc[k];
        }
    return NRB200_OK;
}

// The draws of member k's shard.
nrb200_draws shard_draws(const nrb200_handle *g, const nrb200_draws *d, int k, int64_t &lo, int64_t &n) {
    int64_t hi;
    shard_iterations(d->n_iter, group_size(g), k, lo, hi);
    n = hi - lo;
    nrb200_draws s = *d;
    s.n_iter = n;
    s.iter0 = d->iter0 + lo;
    const int64_t B = g->members[0]->plan.n_blocks();
    if (d->rng_mode == NRB200_RNG_HOST) {
        if (d->bait_seqid) s.bait_seqid = d->bait_seqid + lo * B;
        if (d->bait_start) s.bait_start = d->bait_start + lo * B;
        if (d->dst_tile) s.dst_tile = d->dst_tile + lo * B;
    }
    return s;
}

void merge_stats(nrb200_stats *out, const std::vector<nrb200_stats> &st) {
    if (!out) return;
    std::memset(out, 0, sizeof *out);
    for (const nrb200_stats &s : st) {  // times: the slowest member; volumes: all members
        out->kernel_ms = std::max(out->kernel_ms, s.kernel_ms);
        out->total_ms = std::max(out->total_ms, s.total_ms);
        out->rows_kernel_ms = std::max(out->rows_kernel_ms, s.rows_kernel_ms);
        out->count_kernel_ms = std::max(out->count_kernel_ms, s.count_kernel_ms);
        out->n_launches += s.n_launches;
        out->n_hits += s.n_hits;
        out->h2d_bytes += s.h2d_bytes;
        out->d2h_bytes += s.d2h_bytes;
        out->n_windows = std::max(out->n_windows, s.n_windows);
    }
}

// fn(leader) on the calling thread: entry points a group serves with its first device only.
int group_lead(nrb200_handle *g, const std::function<int(nrb200_handle *)> &fn) {
    nrb200_handle *lead = g->members[0];
    const int rc = fn(lead);
    if (rc) g->error = "GPU " + std::to_string(lead->device) + ": " + lead->error;
    return rc;
}


// Before a run: the leader prepares first (it builds the window lists), the others then take its lists.  Nothing happens when no input changed since the last run.
int group_prepare(nrb200_handle *g) {
    nrb200_handle *lead = g->members[0];
    if (lead->have_y && lead->have_plan && !lead->slices_dirty) return NRB200_OK;  // nothing changed since the last run: no dispatch
    // The leader prepares on the calling thread; the per-sequence passes of its list building are handed to the driver
    // threads of all members, each with its own share of the host threads: every core builds, and no team is larger than a
    // member's share (a larger team's idle threads would spin on the cores the members' driver threads need right after).
    NRB_CUDA(g, cudaSetDevice(lead->device));
    std::atomic<int> next{0};
    lead->seq_runner = [&](int n, const std::function<void(int)> &body) {
        next.store(0);
        group_all(g, [&](int, nrb200_handle *m) {
#pragma omp parallel num_threads(host_threads(m))
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n) break;
                body(i);
            }
            return NRB200_OK;
        });
    };
    const int rc = prepare(lead);
    lead->seq_runner = nullptr;
    if (rc) g->error = "GPU " + std::to_string(lead->device) + ": " + lead->error;
    return rc;
}

int group_create(const nrb200_config *cfg, nrb200_handle **out) {
    const int W = cfg->n_devices;
    if (!cfg->devices) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: n_devices > 1 needs the device list");
    if (cfg->stream) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: a device group runs on private streams (stream must be NULL)");
    // NRB200_ALLOW_DUP_DEVICES=1 (tests on a one-GPU box): members may share a GPU -- same code path, no speed-up
    const bool dup_ok = getenv("NRB200_ALLOW_DUP_DEVICES") != nullptr;
    for (int a = 0; a < W && !dup_ok; ++a)
        for (int b = a + 1; b < W; ++b)
            if (cfg->devices[a] == cfg->devices[b]) return fail(nullptr, NRB200_ERR_INVALID, "nrb200_create: a device is listed twice");
    nrb200_handle *g = new nrb200_handle();
    g->device = cfg->devices[0];
    g->flags = cfg->flags;
    const int cores = omp_get_max_threads();
    for (int k = 0; k < W; ++k) {
        nrb200_config one{};
        one.device = cfg->devices[k];
        one.flags = cfg->flags;
        nrb200_handle *m = nullptr;
        const int rc = nrb200_create(&one, &m);
        if (rc) {
            for (nrb200_handle *x : g->members) nrb200_destroy(x);
            delete g;
            return rc;  // the message is already in place
        }
        m->omp_threads = std::max(1, cores / W);
        m->group_w = W;
        m->keep_lists = k == 0;
        m->lists_from = k == 0 ? nullptr : g->members[0];
        g->members.push_back(m);
    }
    g->pool = new GroupPool();
    // NUMA placement (NRB200_NUMA=0 turns it off): when the GPUs hang off different NUMA nodes, every member gets its own
    // driver thread bound to the CPUs of its GPU's node, so that its host team, its pinned staging buffers and the pages
    // it touches first sit next to the PCIe root its GPU uses.
    const char *numa_env = getenv("NRB200_NUMA");
    if (!numa_env || atoi(numa_env) != 0) {
        std::vector<int> node((size_t)W, -1);
        std::vector<std::vector<int>> cpus((size_t)W);
        cpu_set_t allowed;
        CPU_ZERO(&allowed);
        sched_getaffinity(0, sizeof allowed, &allowed);
        bool known = true;
        for (int k = 0; k < W && known; ++k) {
            char bus[32] = {0};
            if (cudaDeviceGetPCIBusId(bus, sizeof bus, cfg->devices[k]) != cudaSuccess) { known = false; break; }
            for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
            FILE *f = fopen((std::string("/sys/bus/pci/devices/") + bus + "/numa_node").c_str(), "r");
            if (!f || fscanf(f, "%d", &node[k]) != 1 || node[k] < 0) known = false;
            if (f) fclose(f);
            if (!known) break;
            f = fopen(("/sys/devices/system/node/node" + std::to_string(node[k]) + "/cpulist").c_str(), "r");
            if (!f) { known = false; break; }
            int a = 0, b = 0;
            while (fscanf(f, "%d", &a) == 1) {  // "0-15,32-47"
                b = a;
                int ch = fgetc(f);
                if (ch == '-') {
                    if (fscanf(f, "%d", &b) != 1) break;
                    ch = fgetc(f);
                }
                for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
                    if (CPU_ISSET(c, &allowed)) cpus[k].push_back(c);
                if (ch != ',') break;
            }
            fclose(f);
            if (cpus[k].empty()) known = false;
        }
        cudaGetLastError();
        bool several = false;
        for (int k = 1; k < W && known; ++k) several |= node[k] != node[0];
        if (known && several) {
            for (int k = 0; k < W; ++k) {
                int sharing = 0;
                for (int j = 0; j < W; ++j) sharing += node[j] == node[k];
                // the members of one node split its CPUs among them
                int rank_on_node = 0;
                for (int j = 0; j < k; ++j) rank_on_node += node[j] == node[k];
                const size_t per = std::max<size_t>(1, cpus[k].size() / (size_t)sharing);
                std::vector<int> mine(cpus[k].begin() + std::min(cpus[k].size(), per * rank_on_node),
                                      cpus[k].begin() + std::min(cpus[k].size(), per * (rank_on_node + 1)));
                if (mine.empty()) mine = cpus[k];
                g->members[k]->omp_threads = (int)mine.size();
                g->pool->cpus.push_back(std::move(mine));
            }
        }
    }
    g->pool->start(W, g->pool->cpus.empty());
    *out = g;
    return NRB200_OK;
}

void group_destroy(nrb200_handle *g) {
    delete g->pool;  // joins the workers
    for (nrb200_handle *m : g->members) nrb200_destroy(m);
    if (g->g_stage) {
        cudaSetDevice(g->device);
        cudaFreeHost(g->g_stage);
    }
    delete g;
}

int group_enable_peers(nrb200_handle *g);  // (below, with the gathered result)

// Peer access between all members, probed once and without raising an error (the callers have a path without it).
bool group_peers_quiet(nrb200_handle *g) {
    if (g->peers_state == 0) {
        const std::string keep = g->error;
        g->peers_state = group_enable_peers(g) == NRB200_OK ? 1 : -1;
        g->error = keep;
        cudaGetLastError();
    }
    return g->peers_state > 0;
}

// A large y crosses PCIe ONCE for the whole group: member k uploads the k-th slice of every column to its own GPU and
// passes it on to the other W - 1 members with peer copies queued behind it on the same stream (NVLink through the
// NVSwitch: W x faster than W full uploads through the host); a member's stream takes up the inspection of y when all
// W slices have arrived.  (start, end, seqid[, strand]) are pinned host arrays of n elements.
int group_scatter_y(nrb200_handle *g, int64_t n, const int32_t *seqid, const int32_t *start, const int32_t *end, const int8_t *strand) {
    const int W = group_size(g);
    int rc = group_all(g, [&](int, nrb200_handle *m) {
        NRB_CUDA(m, cudaSetDevice(m->device));
        if (int rs = sync_stream(m)) return rs;  // nothing of an earlier call reads the columns any more
        NRB_CUDA(m, m->y.start.reserve((size_t)n * 4));
        NRB_CUDA(m, m->y.end.reserve((size_t)n * 4));
        NRB_CUDA(m, m->y.seqid.reserve((size_t)n * 4));
        if (strand) NRB_CUDA(m, m->y.strand.reserve((size_t)n));
        if (!m->y_ev) NRB_CUDA(m, cudaEventCreateWithFlags(&m->y_ev, cudaEventDisableTiming));
        return NRB200_OK;
    });
    if (rc) return rc;
    auto cut = [&](int k) { return k >= W ? n : (n * k / W) & ~(int64_t)63; };  // slice bounds: multiples of 64 elements
    rc = group_all(g, [&](int k, nrb200_handle *m) {
        NRB_CUDA(m, cudaSetDevice(m->device));
        const int64_t lo = cut(k), cnt = cut(k + 1) - lo;
        struct Col {
            DevMem RangeSet::*dst;
            const void *src;
            size_t elem;
        };
        const Col cols[4] = {{&RangeSet::start, start, 4}, {&RangeSet::end, end, 4}, {&RangeSet::seqid, seqid, 4}, {&RangeSet::strand, strand, 1}};
        for (const Col &c : cols) {
            if (!c.src || cnt <= 0) continue;
            const size_t off = (size_t)lo * c.elem, bytes = (size_t)cnt * c.elem;
            char *own = static_cast<char *>((m->y.*(c.dst)).p) + off;
            NRB_CUDA(m, cudaMemcpyAsync(own, static_cast<const char *>(c.src) + off, bytes, cudaMemcpyHostToDevice, m->stream));
            for (int j = 0; j < W; ++j) {
                nrb200_handle *peer = g->members[j];
                if (j == k) continue;
                NRB_CUDA(m, cudaMemcpyPeerAsync(static_cast<char *>((peer->y.*(c.dst)).p) + off, peer->device, own, m->device, bytes, m->stream));
            }
        }
        NRB_CUDA(m, cudaEventRecord(m->y_ev, m->stream));
        return NRB200_OK;
    });
    if (rc) return rc;
    return group_all(g, [&](int k, nrb200_handle *m) {
        NRB_CUDA(m, cudaSetDevice(m->device));
        for (int j = 0; j < W; ++j)
            if (j != k) NRB_CUDA(m, cudaStreamWaitEvent(m->stream, g->members[j]->y_ev, 0));
        return NRB200_OK;
    });
}

// y for every member.  A large pageable y is copied into pinned memory once (all host threads); from pinned memory
// (that copy, or the caller's own) a large y is scattered over the members and exchanged between the GPUs
// (group_scatter_y), a small one or one without peer access is uploaded by every member from the same pinned copy.
int group_set_ranges_begin(nrb200_handle *g, int64_t n, int32_t n_seq, const int64_t *seqlengths, const int32_t *seqid,
                           const int32_t *start, const int32_t *end, const int8_t *strand) {
    NRB_CUDA(g, cudaSetDevice(g->device));
    const size_t col = ((size_t)std::max<int64_t>(n, 0) * 4 + 4095) & ~(size_t)4095;
    if (n > 0 && seqid && start && end && (size_t)n * 4 >= ((size_t)256 << 10) && (size_t)n * 4 <= ((size_t)1 << 30)) {
        cudaPointerAttributes attr{};
        const bool pageable = cudaPointerGetAttributes(&attr, start) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        if (pageable) {
            const size_t need = 4 * col;
            if (g->g_stage_cap < need) {
                if (g->g_stage) cudaFreeHost(g->g_stage);
                g->g_stage = nullptr;
                g->g_stage_cap = 0;
                if (cudaMallocHost(&g->g_stage, need) == cudaSuccess) g->g_stage_cap = need;
                else cudaGetLastError();
            }
            if (g->g_stage_cap >= need) {
                // every member copies its 1/W-th of each column with its own host threads (all cores busy, no team larger
                // than a member's share), then all of them upload from the one pinned copy
                char *base = static_cast<char *>(g->g_stage);
                const void *src[4] = {seqid, start, end, strand};
                const size_t bytes[4] = {(size_t)n * 4, (size_t)n * 4, (size_t)n * 4, strand ? (size_t)n : 0};
                const int W = group_size(g);
                group_all(g, [&](int k, nrb200_handle *m) {
                    const size_t piece = (size_t)256 << 10;
                    for (int c = 0; c < 4; ++c) {
                        const int64_t pieces = (int64_t)((bytes[c] + piece - 1) / piece);
                        const int64_t p0 = pieces * k / W, p1 = pieces * (k + 1) / W;
#pragma omp parallel for schedule(static) num_threads(host_threads(m))
                        for (int64_t i = p0; i < p1; ++i)
                            copy_streaming(static_cast<const char *>(src[c]) + i * piece, base + c * col + i * piece,
                                           std::min(piece, bytes[c] - (size_t)i * piece));
                    }
                    return NRB200_OK;
                });
                seqid = reinterpret_cast<const int32_t *>(base);
                start = reinterpret_cast<const int32_t *>(base + col);
                end = reinterpret_cast<const int32_t *>(base + 2 * col);
                if (strand) strand = reinterpret_cast<const int8_t *>(base + 3 * col);
            }
        }
    }
    // NRB200_SCATTER_MIN (ranges; default 2^20, 0 = never): from this size on y is scattered and exchanged between the GPUs
    const char *scatter_env = getenv("NRB200_SCATTER_MIN");
    const int64_t scatter_min = scatter_env ? atoll(scatter_env) : ((int64_t)1 << 20);
    bool scattered = false;
    if (scatter_min > 0 && n >= scatter_min && n <= (int64_t)INT32_MAX - 64 && seqid && start && end) {
        cudaPointerAttributes attr{};
        const bool pinned = cudaPointerGetAttributes(&attr, start) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (pinned && group_peers_quiet(g)) {
            if (int rc = group_scatter_y(g, n, seqid, start, end, strand)) return rc;
            scattered = true;
        }
    }
    return group_all(g, [&](int, nrb200_handle *m) {
        m->y_preloaded = scattered;
        const int rc = nrb200_set_ranges_begin(m, n, n_seq, seqlengths, seqid, start, end, strand);
        m->y_preloaded = false;
        return rc;
    });
}

// ---- runs: member k takes its shard and writes straight into the caller's arrays ------------------------------

int group_run_counts(nrb200_handle *g, const nrb200_draws *d, int64_t *iter_nranges, int64_t *iter_totals, int32_t *feature_counts,
                     nrb200_stats *stats) {
    if (!d || d->n_iter < 0) return fail(g, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
    int rc;
    if ((rc = group_prepare(g))) return rc;
    const int64_t G = g->members[0]->query.n;
    std::vector<nrb200_stats> st((size_t)group_size(g));
    rc = group_all(g, [&](int k, nrb200_handle *m) {
        int64_t lo, n;
        const nrb200_draws s = shard_draws(g, d, k, lo, n);
        return nrb200_run_counts(m, &s, iter_nranges ? iter_nranges + lo : nullptr, iter_totals ? iter_totals + lo : nullptr,
                                 feature_counts ? feature_counts + lo * G : nullptr, stats ? &st[k] : nullptr);
    });
    if (rc) return rc;
    merge_stats(stats, st);
    return NRB200_OK;
}

int group_run_counts_u16(nrb200_handle *g, const nrb200_draws *d, int64_t *iter_nranges, int64_t *iter_totals, uint16_t *feature_counts,
                         int32_t *overflow, nrb200_stats *stats) {
    if (!d || d->n_iter < 0) return fail(g, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
    int rc;
    if ((rc = group_prepare(g))) return rc;
    const int64_t G = g->members[0]->query.n;
    std::vector<nrb200_stats> st((size_t)group_size(g));
    std::vector<int32_t> ov((size_t)group_size(g), 0);
    rc = group_all(g, [&](int k, nrb200_handle *m) {
        int64_t lo, n;
        const nrb200_draws s = shard_draws(g, d, k, lo, n);
        return nrb200_run_counts_u16(m, &s, iter_nranges ? iter_nranges + lo : nullptr, iter_totals ? iter_totals + lo : nullptr,
                                     feature_counts + lo * G, &ov[k], stats ? &st[k] : nullptr);
    });
    if (rc) return rc;
    if (overflow) {
        *overflow = 0;
        for (int32_t v : ov) *overflow |= v;
    }
    merge_stats(stats, st);
    return NRB200_OK;
}

int group_run_moments(nrb200_handle *g, const nrb200_draws *d, int64_t batch_iter, const int32_t *observed, int64_t *iter_nranges,
                      int64_t *iter_totals, nrb200_stats *stats) {
    if (!d || d->n_iter < 0) return fail(g, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
    int rc;
    if ((rc = group_prepare(g))) return rc;
    std::vector<nrb200_stats> st((size_t)group_size(g));
    rc = group_all(g, [&](int k, nrb200_handle *m) {
        int64_t lo, n;
        const nrb200_draws s = shard_draws(g, d, k, lo, n);
        return nrb200_run_moments(m, &s, batch_iter, observed, iter_nranges ? iter_nranges + lo : nullptr,
                                  iter_totals ? iter_totals + lo : nullptr, stats ? &st[k] : nullptr);
    });
    if (rc) return rc;
    merge_stats(stats, st);
    return NRB200_OK;
}

// Pinned landing place of a member's three per-feature vectors (3 x G x 8 bytes), kept between calls.
int member_moment_stage(nrb200_handle *m, size_t G) {
    const size_t need = std::max<size_t>(G, 1) * 24;
    if (m->m_stage_cap >= need) return NRB200_OK;
    NRB_CUDA(m, cudaSetDevice(m->device));
    if (m->m_stage) cudaFreeHost(m->m_stage);
    m->m_stage = nullptr;
    m->m_stage_cap = 0;
    NRB_CUDA(m, cudaMallocHost(&m->m_stage, need));
    m->m_stage_cap = need;
    return NRB200_OK;
}

// out[j][i] = sum over the members, in member order, of vector j of their staging buffers; every member's driver thread adds
// its share of the features with its own host team (all cores at work, no oversized team left spinning).
template <class T0, class T1, class T2>
int group_add_staged(nrb200_handle *g, size_t G, T0 *out0, T1 *out1, T2 *out2) {
    const int W = group_size(g);
    return group_all(g, [&](int k, nrb200_handle *m) {
        const int64_t lo = (int64_t)(G * (size_t)k / (size_t)W), hi = (int64_t)(G * (size_t)(k + 1) / (size_t)W);
#pragma omp parallel for schedule(static) num_threads(host_threads(m))
        for (int64_t i = lo; i < hi; ++i) {
            T0 a = 0;
            T1 b = 0;
            T2 c = 0;
            for (int w = 0; w < W; ++w) {
                const char *base = static_cast<const char *>(g->members[w]->m_stage);
                a += reinterpret_cast<const T0 *>(base)[i];
                b += reinterpret_cast<const T1 *>(base + G * 8)[i];
                c += reinterpret_cast<const T2 *>(base + G * 16)[i];
            }
            if (out0) out0[i] = a;
            if (out1) out1[i] = b;
            if (out2) out2[i] = c;
        }
        return NRB200_OK;
    });
}

// Per-feature moment vectors of all members, summed on the host in member order: the W x 3 device-to-host copies land in
// pinned staging and are in flight together, then the members' threads add them.
int group_moments_fetch(nrb200_handle *g, int64_t *feat_sum, int64_t *feat_sumsq, int64_t *feat_n_ge) {
    const size_t G = (size_t)g->members[0]->query.n;
    const int rc = group_all(g, [&](int, nrb200_handle *m) {
        if (int rs = member_moment_stage(m, G)) return rs;
        int64_t *st = static_cast<int64_t *>(m->m_stage);
        return nrb200_moments_fetch(m, st, st + G, st + 2 * G);
    });
    if (rc) return rc;
    return group_add_staged<int64_t, int64_t, int64_t>(g, G, feat_sum, feat_sumsq, feat_n_ge);
}

int group_run_weighted(nrb200_handle *g, const nrb200_draws *d, double *iter_sums, double *feature_sums, nrb200_stats *stats) {
    if (!d || d->n_iter < 0) return fail(g, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
    int rc;
    if ((rc = group_prepare(g))) return rc;
    const int64_t G = g->members[0]->query.n;
    std::vector<nrb200_stats> st((size_t)group_size(g));
    rc = group_all(g, [&](int k, nrb200_handle *m) {
        int64_t lo, n;
        const nrb200_draws s = shard_draws(g, d, k, lo, n);
        return nrb200_run_weighted(m, &s, iter_sums ? iter_sums + lo : nullptr, feature_sums ? feature_sums + lo * G : nullptr,
                                   stats ? &st[k] : nullptr);
    });
    if (rc) return rc;
    merge_stats(stats, st);
    return NRB200_OK;
}

// Moments of the weighted sums: every member reduces its shard, the host adds the W partial vectors in member order
// (double sums: deterministic for a given W; across different W they agree to rounding, as a weighted sum does anyway).
int group_run_weighted_moments(nrb200_handle *g, const nrb200_draws *d, const double *observed, double *iter_sums, double *feat_sum,
                               double *feat_sumsq, int64_t *feat_n_ge, nrb200_stats *stats) {
    if (!d || d->n_iter < 0) return fail(g, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
    int rc;
    if ((rc = group_prepare(g))) return rc;
    const int W = group_size(g);
    const size_t G = (size_t)g->members[0]->query.n;
    std::vector<nrb200_stats> st((size_t)W);
    rc = group_all(g, [&](int k, nrb200_handle *m) {
        int64_t lo, n;
        const nrb200_draws s = shard_draws(g, d, k, lo, n);
        if (int rs = member_moment_stage(m, G)) return rs;
        char *base = static_cast<char *>(m->m_stage);
        return nrb200_run_weighted_moments(m, &s, observed, iter_sums ? iter_sums + lo : nullptr, reinterpret_cast<double *>(base),
                                           reinterpret_cast<double *>(base + G * 8), reinterpret_cast<int64_t *>(base + G * 16),
                                           stats ? &st[k] : nullptr);
    });
    if (rc) return rc;
    if ((rc = group_add_staged<double, double, int64_t>(g, G, feat_sum, feat_sumsq, feat_n_ge))) return rc;
    merge_stats(stats, st);
    return NRB200_OK;
}

// ---- gathered result: every member ends up with the whole [R x G] matrix in ITS memory ---------------------------
// SURVEY 8e's "ncclAllGather of the [R/W x G] tiles" without NCCL and without a second process: the count kernel of
// GPU k writes the rows of its iterations into its own copy of the matrix, and a push kernel on the same stream
// stores them into the copies of the other W - 1 members through peer-mapped pointers (NVLink 5 through the NVSwitch,
// 16-byte coalesced stores); a last small kernel spreads the per-iteration vectors.  Every GPU pushes while the others
// compute and push.  The call returns when every member's stream has drained: all matrices are complete.
int group_enable_peers(nrb200_handle *g) {
    if (g->peers_enabled) return NRB200_OK;
    const int W = group_size(g);
    if (W > kMaxPeers) return fail(g, NRB200_ERR_UNSUPPORTED, "a gathered result supports at most 16 GPUs");
    for (int a = 0; a < W; ++a) {
        NRB_CUDA(g, cudaSetDevice(g->members[a]->device));
        for (int b = 0; b < W; ++b) {
            if (a == b || g->members[a]->device == g->members[b]->device) continue;
            int can = 0;
            NRB_CUDA(g, cudaDeviceCanAccessPeer(&can, g->members[a]->device, g->members[b]->device));
            if (!can) return fail(g, NRB200_ERR_UNSUPPORTED, "the GPUs of this group cannot address each other's memory (no peer access)");
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->members[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) NRB_CUDA(g, e);
            cudaGetLastError();
        }
    }
    g->peers_enabled = true;
    return NRB200_OK;
}

int group_run_counts_gathered(nrb200_handle *g, const nrb200_draws *d, nrb200_stats *stats) {
    if (!d || d->n_iter < 0) return fail(g, NRB200_ERR_INVALID, "draws: n_iter must be >= 0");
    int rc;
    if ((rc = group_enable_peers(g))) return rc;
    if ((rc = group_prepare(g))) return rc;
    const int W = group_size(g);
    const int64_t R = d->n_iter, G = g->members[0]->query.n;
    if (G == 0) return fail(g, NRB200_ERR_STATE, "a gathered result needs a query (nrb200_set_query)");
    bool sized = true;
    for (int k = 0; k < W; ++k) sized = sized && g->members[k]->gat_R == R && g->members[k]->gat_G == G;
    if (!sized) rc = group_all(g, [&](int, nrb200_handle *m) {  // every member makes room for the whole result
        NRB_CUDA(m, cudaSetDevice(m->device));
        NRB_CUDA(m, m->gat_counts.reserve((size_t)std::max<int64_t>(R * G, 1) * 4));
        NRB_CUDA(m, m->gat_nranges.reserve((size_t)std::max<int64_t>(R, 1) * 8));
        NRB_CUDA(m, m->gat_totals.reserve((size_t)std::max<int64_t>(R, 1) * 8));
        m->gat_R = R;
        m->gat_G = G;
        return NRB200_OK;
    });
    if (rc) return rc;
    std::vector<nrb200_stats> st((size_t)W);
    rc = group_all(g, [&](int k, nrb200_handle *m) {
        int64_t lo, n;
        const nrb200_draws s = shard_draws(g, d, k, lo, n);
        // the count kernel writes this GPU's rows into its OWN copy of the matrix ...
        m->peer_out.n = 1;
        m->peer_out.counts[0] = m->gat_counts.as<int32_t>() + lo * G;
        PeerVectors V{};
        PeerTiles T{};
        for (int w = 0; w < W; ++w) {
            V.nranges[V.n] = g->members[w]->gat_nranges.as<unsigned long long>() + lo;
            V.totals[V.n++] = g->members[w]->gat_totals.as<unsigned long long>() + lo;
            if (w != k) T.dst[T.n++] = g->members[w]->gat_counts.as<int32_t>() + lo * G;
        }
        NRB_CUDA(m, cudaSetDevice(m->device));
        int rc2 = run_counts_core(m, &s, true, nullptr, stats ? &st[k] : nullptr);
        m->peer_out.n = 0;
        if (rc2) return rc2;
        if (n > 0) {
            // ... and two small kernels spread the rows and the per-iteration vectors to the other members
            static const bool push_by_copy_engine = getenv("NRB200_PUSH_MEMCPY") != nullptr;  // A/B: cudaMemcpyPeerAsync per peer
            if (T.n > 0 && push_by_copy_engine) {
                for (int w = 0; w < W; ++w)
                    if (w != k)
                        NRB_CUDA(m, cudaMemcpyPeerAsync(g->members[w]->gat_counts.as<int32_t>() + lo * G, g->members[w]->device,
                                                        m->gat_counts.as<int32_t>() + lo * G, m->device, (size_t)n * G * 4, m->stream));
            } else if (T.n > 0) {
                push_tile_kernel<<<(unsigned)m->sm_count * 4, 256, 0, m->stream>>>(m->gat_counts.as<int32_t>() + lo * G, n * G, T);
            }
            broadcast_iteration_vectors_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 64), 256, 0, m->stream>>>(
                m->d_iter_nranges.as<unsigned long long>(), m->d_iter_totals.as<unsigned long long>(), n, V);
            NRB_CUDA(m, cudaGetLastError());
        }
        if (stats) NRB_CUDA(m, cudaEventRecord(m->ev[3], m->stream));
        if (int rc3 = sync_stream(m)) return rc3;
        if (stats) {  // device time of this member's whole step: draws, counts, push to the peers
            float ms = 0;
            cudaEventElapsedTime(&ms, m->ev[0], m->ev[3]);
            st[k].total_ms = ms;
        }
        return NRB200_OK;
    });
    if (rc) return rc;
    merge_stats(stats, st);
    return NRB200_OK;
}

}  // namespace
