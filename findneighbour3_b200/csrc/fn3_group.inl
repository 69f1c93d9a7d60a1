// fn3_group.inl — an engine GROUP: one process drives the GPUs of one box (include/fn3.h, "engine groups").
// Included at the end of fn3_engine.cu (it needs the engine's internals).
//
// What it gives the reference: findNeighbour3-server.py:340-345 constructs ONE seqComparer and calls sc.mcompare from that
// one process (:530).  A group lets that unmodified single process use every GPU: the candidate set is sharded by
// insertion order — global row g lives on engine g % G as local row g / G — and every group call fans
// out to one HOST THREAD PER DEVICE that enqueues on its own device's stream and waits for its own completion flag in C;
// nothing on the data plane runs in Python and no collective is launched: the new sample goes host -> every device, the
// thresholded neighbour records come back through each device's mapped pinned block.
//
// Rows are read ACROSS devices where a call needs it (a stored query against the other shards, all-vs-all of a sharded
// store, consensus / MSA over rows of several shards): with peer access enabled the kernels dereference the owner's HBM
// directly over NVLink (the row pointer in a QuerySlot or a virtual head tile is a peer address) — no staging copy.

#include <condition_variable>
#include <functional>
#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define FN3_CPU_RELAX() _mm_pause()
#else
#define FN3_CPU_RELAX() std::this_thread::yield()
#endif

struct fn3_group {
    int G = 0;
    std::vector<fn3_engine*> eng;
    std::vector<int> dev;
    long long L = 0;
    std::mutex mu;                       // one group call at a time (they share the per-device result buffers)
    std::mutex err_mu; std::string err;
    long long n_global = 0;              // rows ever appended through the group
    bool peer_all = true;                // every engine can dereference every other engine's rows
    // workers: device d > 0 is served by its own host thread, device 0 by the calling thread
    struct Worker {
        std::thread th;
        std::atomic<uint64_t> done{0};
        int rc = 0;
        std::vector<fn3_hit> hits; int64_t n_out = 0;
        std::vector<fn3_edge> edges;
        std::vector<int64_t> rows;       // per-device candidate list of a call
    };
    std::vector<Worker> w;
    std::atomic<uint64_t> job_seq{0};
    const std::function<int(int)>* job = nullptr;
    std::atomic<bool> stop{false};
    std::mutex cv_mu; std::condition_variable cv; std::atomic<int> sleepers{0};
};

static int gset_err(fn3_group* g, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (g) { std::lock_guard<std::mutex> l(g->err_mu); g->err = buf; }
    else g_create_error = buf;
    return code;
}

// error of engine d becomes the group's error
static int gfrom(fn3_group* g, int d, int rc) {
    if (rc != FN3_OK) gset_err(g, rc, "device %d: %s", g->dev[(size_t)d], fn3_last_error(g->eng[(size_t)d]));
    return rc;
}

// a job on device d; a C++ exception (the per-device buffers are std::vectors) becomes an error code, not std::terminate
static int group_call(fn3_group* g, const std::function<int(int)>& f, int d) {
    try { return f(d); }
    catch (const std::bad_alloc&) { return set_err(g->eng[(size_t)d], FN3_ENOMEM, "out of host memory"); }
    catch (const std::exception& ex) { return set_err(g->eng[(size_t)d], FN3_ECUDA, "unexpected exception: %s", ex.what()); }
}

static void group_worker(fn3_group* g, int d) {
    cudaSetDevice(g->dev[(size_t)d]);
    uint64_t seen = 0;
    for (;;) {
        uint64_t s;
        unsigned spins = 0;
        while ((s = g->job_seq.load(std::memory_order_acquire)) == seen) {
            if (g->stop.load(std::memory_order_acquire)) return;
            if (++spins < 20000) { FN3_CPU_RELAX(); continue; }
            // idle for a while: sleep until the next job is posted (bounded, so a missed wake-up costs a millisecond at most)
            g->sleepers.fetch_add(1);
            {
                std::unique_lock<std::mutex> lk(g->cv_mu);
                g->cv.wait_for(lk, std::chrono::milliseconds(1),
                               [&] { return g->job_seq.load() != seen || g->stop.load(); });
            }
            g->sleepers.fetch_sub(1);
            spins = 0;
        }
        seen = s;
        g->w[(size_t)d].rc = group_call(g, *g->job, d);
        g->w[(size_t)d].done.store(s, std::memory_order_release);
    }
}

// run f(d) for every device: d = 0 on the calling thread, the others on their workers; returns the first failure
static int group_run(fn3_group* g, const std::function<int(int)>& f) {
    uint64_t seq = 0;
    if (g->G > 1) {
        g->job = &f;
        seq = g->job_seq.load(std::memory_order_relaxed) + 1;
        g->job_seq.store(seq);
        if (g->sleepers.load() > 0) { std::lock_guard<std::mutex> lk(g->cv_mu); g->cv.notify_all(); }
    }
    g->w[0].rc = group_call(g, f, 0);
    for (int d = 1; d < g->G; ++d)
        while (g->w[(size_t)d].done.load(std::memory_order_acquire) != seq) FN3_CPU_RELAX();
    for (int d = 0; d < g->G; ++d)
        if (g->w[(size_t)d].rc != FN3_OK) return gfrom(g, d, g->w[(size_t)d].rc);
    return FN3_OK;
}

extern "C" fn3_group* fn3_group_create(int64_t genome_length, int32_t n_devices, const int32_t* device_ids) {
    if (n_devices < 1 || n_devices > 64 || !device_ids) {
        set_err(nullptr, FN3_EARG, "fn3_group_create: 1..64 devices");
        return nullptr;
    }
    fn3_group* g = new fn3_group();
    g->G = n_devices; g->L = genome_length;
    g->w = std::vector<fn3_group::Worker>((size_t)n_devices);
    for (int d = 0; d < n_devices; ++d) {
        fn3_engine* e = fn3_create(genome_length, device_ids[d]);
        if (!e) { fn3_group_destroy(g); return nullptr; }            // fn3_last_error(NULL) holds the reason
        g->eng.push_back(e); g->dev.push_back(device_ids[d]);
    }
    // peer access between every pair of distinct devices (NVLink / NVSwitch on an HGX board)
    DeviceGuard dg(g->dev[0]);
    for (int a = 0; a < n_devices; ++a)
        for (int b = 0; b < n_devices; ++b) {
            if (g->dev[(size_t)a] == g->dev[(size_t)b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->dev[(size_t)a], g->dev[(size_t)b]);
            if (can) {
                cudaSetDevice(g->dev[(size_t)a]);
                const cudaError_t s = cudaDeviceEnablePeerAccess(g->dev[(size_t)b], 0);
                if (s != cudaSuccess && s != cudaErrorPeerAccessAlreadyEnabled) can = 0;
                (void)cudaGetLastError();
            }
            if (!can) g->peer_all = false;
        }
    for (int d = 1; d < n_devices; ++d) g->w[(size_t)d].th = std::thread(group_worker, g, d);
    return g;
}

extern "C" int fn3_group_destroy(fn3_group* g) {
    if (!g) return FN3_OK;
    g->stop.store(true);
    { std::lock_guard<std::mutex> lk(g->cv_mu); g->cv.notify_all(); }
    for (auto& w : g->w) if (w.th.joinable()) w.th.join();
    for (auto e : g->eng) fn3_destroy(e);
    delete g;
    return FN3_OK;
}

extern "C" const char* fn3_group_last_error(fn3_group* g) {
    if (!g) return g_create_error.c_str();
    std::lock_guard<std::mutex> l(g->err_mu);
    static thread_local std::string copy;
    copy = g->err;
    return copy.c_str();
}

extern "C" int32_t fn3_group_n_devices(fn3_group* g) { return g ? g->G : 0; }
extern "C" int32_t fn3_group_peer_access(fn3_group* g) { return g && g->peer_all ? 1 : 0; }
extern "C" fn3_engine* fn3_group_engine(fn3_group* g, int32_t i) { return (g && i >= 0 && i < g->G) ? g->eng[(size_t)i] : nullptr; }

extern "C" int64_t fn3_group_launch_count(fn3_group* g) {
    long long n = 0;
    if (g) for (auto e : g->eng) n += e->launches.load();
    return n;
}

// ---- store ---------------------------------------------------------------------------------------------

extern "C" int fn3_group_append(fn3_group* g, const int32_t* pos, const int32_t off[7], int32_t invalid, int64_t* out_row) {
    if (!g) return FN3_EARG;
    std::lock_guard<std::mutex> l(g->mu);
    const int d = (int)(g->n_global % g->G);
    int64_t local = -1;
    const int rc = fn3_append(g->eng[(size_t)d], pos, off, invalid, &local);
    if (rc) return gfrom(g, d, rc);
    if (local != g->n_global / g->G) return gset_err(g, FN3_EARG, "fn3_group_append: engine %d was appended to behind the group's back", d);
    if (out_row) *out_row = g->n_global;
    g->n_global++;
    return FN3_OK;
}

// sample i of the call becomes global row first + i: dealt round-robin, every device packs and copies its share on its
// own host thread
extern "C" int fn3_group_append_bulk(fn3_group* g, int64_t n, const int32_t* pos, const int64_t* off, const uint8_t* invalid,
                                     int64_t* out_first_row) {
    if (!g || n < 0 || !off) return gset_err(g, FN3_EARG, "fn3_group_append_bulk: bad arguments");
    std::lock_guard<std::mutex> l(g->mu);
    const long long first = g->n_global;
    if (out_first_row) *out_first_row = first;
    if (n == 0) return FN3_OK;
    for (int64_t i = 0; i < 6 * n; ++i)
        if (off[i + 1] < off[i]) return gset_err(g, FN3_EARG, "list offsets must be non-decreasing");
    const int G = g->G;
    // validate EVERYTHING before anything is stored: a sample that only one shard would reject must not leave the other
    // shards appended and the group out of step (the per-engine call validates its own share again; that is the cheap part)
    {
        fn3_engine* e0 = g->eng[0];
        std::atomic<int> bad{0};
        const int host_threads = (int)std::max<long long>(1, std::min<long long>(env_ll("FN3_HOST_THREADS", 16),
                                                                                 (long long)std::thread::hardware_concurrency()));
        parallel_rows(n, host_threads, [&](long long lo, long long hi) {
            for (long long i = lo; i < hi && !bad.load(std::memory_order_relaxed); ++i) {
                if (invalid && invalid[i]) continue;
                long long o[7];
                for (int b = 0; b < 7; ++b) o[b] = off[6 * i + b];
                if (row_words_needed(e0, pos, o, 0) < 0) { bad.store(1); return; }
                if (!raw_disjoint(pos, o)) { set_err(e0, FN3_EARG, "the six lists of a sample must be mutually disjoint"); bad.store(1); return; }
            }
        });
        if (bad.load()) return gfrom(g, 0, FN3_EARG);
    }
    std::vector<int64_t> got((size_t)G, -1);
    const int rc = group_run(g, [&](int d) -> int {
        // samples i with (first + i) % G == d, in order
        long long i0 = ((d - first) % G + G) % G;
        const long long cnt = i0 < n ? (n - i0 + G - 1) / G : 0;
        if (cnt == 0) return FN3_OK;
        std::vector<int64_t> o((size_t)(6 * cnt + 1));
        long long total = 0;
        for (long long k = 0; k < cnt; ++k) { const long long i = i0 + k * G; total += off[6 * i + 6] - off[6 * i]; }
        std::vector<int32_t> p((size_t)total);
        std::vector<uint8_t> inv((size_t)cnt, 0);
        long long at = 0;
        o[0] = 0;
        for (long long k = 0; k < cnt; ++k) {
            const long long i = i0 + k * G;
            const long long base = off[6 * i], len = off[6 * i + 6] - base;
            if (len) memcpy(p.data() + at, pos + base, (size_t)len * sizeof(int32_t));
            for (int b = 1; b <= 6; ++b) o[(size_t)(6 * k + b)] = at + (off[6 * i + b] - base);
            at += len;
            if (invalid) inv[(size_t)k] = invalid[i];
        }
        return fn3_append_bulk(g->eng[(size_t)d], cnt, p.data(), o.data(), invalid ? inv.data() : nullptr, &got[(size_t)d]);
    });
    if (rc) return rc;                     // (only an allocation / CUDA failure can get here: the samples were validated above)
    for (int d = 0; d < G; ++d) {
        const long long i0 = ((d - first) % G + G) % G;
        if (i0 < n && got[(size_t)d] != (first + i0) / G)
            return gset_err(g, FN3_EARG, "fn3_group_append_bulk: engine %d is out of step with the group", d);
    }
    g->n_global = first + n;
    return FN3_OK;
}

extern "C" int fn3_group_remove(fn3_group* g, int64_t row) {
    if (!g) return FN3_EARG;
    std::lock_guard<std::mutex> l(g->mu);
    if (row < 0 || row >= g->n_global) return gset_err(g, FN3_ENOTFOUND, "fn3_group_remove: row %lld does not exist", (long long)row);
    const int d = (int)(row % g->G);
    return gfrom(g, d, fn3_remove(g->eng[(size_t)d], row / g->G));
}

extern "C" int fn3_group_row_get(fn3_group* g, int64_t row, int32_t* pos, int64_t pos_cap, int32_t off[7], int32_t* invalid,
                                 int64_t* n_pos) {
    if (!g) return FN3_EARG;
    std::lock_guard<std::mutex> l(g->mu);
    if (row < 0 || row >= g->n_global) return gset_err(g, FN3_ENOTFOUND, "fn3_group_row_get: row %lld does not exist", (long long)row);
    const int d = (int)(row % g->G);
    return gfrom(g, d, fn3_row_get(g->eng[(size_t)d], row / g->G, pos, pos_cap, off, invalid, n_pos));
}

extern "C" int fn3_group_stats_get(fn3_group* g, fn3_stats* out) {
    if (!g || !out) return FN3_EARG;
    std::lock_guard<std::mutex> l(g->mu);
    memset(out, 0, sizeof *out);
    for (int d = 0; d < g->G; ++d) {
        fn3_stats s;
        const int rc = fn3_stats_get(g->eng[(size_t)d], &s);
        if (rc) return gfrom(g, d, rc);
        out->n_live += s.n_live; out->n_invalid += s.n_invalid; out->store_words += s.store_words;
        out->store_bytes += s.store_bytes; out->positions += s.positions; out->h2d_bytes += s.h2d_bytes; out->d2h_bytes += s.d2h_bytes;
        out->t_pack_ns += s.t_pack_ns; out->t_enqueue_ns += s.t_enqueue_ns; out->t_wait_ns += s.t_wait_ns; out->n_timed += s.n_timed;
        out->head_refreshes += s.head_refreshes; out->scan_mode += s.scan_mode;
    }
    out->n_rows = g->n_global;
    return FN3_OK;
}

extern "C" int fn3_group_set_reference(fn3_group* g, const char* reference, const int32_t* excluded, int64_t n_excluded) {
    if (!g) return FN3_EARG;
    std::lock_guard<std::mutex> l(g->mu);
    return group_run(g, [&](int d) { return fn3_set_reference(g->eng[(size_t)d], reference, excluded, n_excluded); });
}

// ---- the hot path ----------------------------------------------------------------------------------------

// header + touched-block bound of a live global row (from its owner's host mirror)
static int group_hdr(fn3_group* g, long long row, RowHdr* h, uint32_t* nblk, uint8_t* flags, const char* who) {
    if (row < 0 || row >= g->n_global) return gset_err(g, FN3_ENOTFOUND, "%s: row %lld does not exist", who, row);
    fn3_engine* e = g->eng[(size_t)(row % g->G)];
    const long long local = row / g->G;
    std::lock_guard<std::mutex> s(e->store_mu);
    if (local >= e->n_rows) return gset_err(g, FN3_ENOTFOUND, "%s: row %lld does not exist", who, row);
    const uint8_t fl = e->h_flags[(size_t)local];
    if (fl & 2) return gset_err(g, FN3_ENOTFOUND, "%s: row %lld was removed", who, row);
    *h = e->h_hdr[(size_t)local];
    if (nblk) *nblk = e->h_nblk[(size_t)local];
    if (flags) *flags = fl;
    return FN3_OK;
}

// A query that is a row of ANOTHER engine (header qh; data on device src_dev) against this engine's store.  With peer access
// the kernel reads the row where it lives; without, the row is first copied into this engine's scratch.
static int foreign_vs_all(fn3_engine* e, const RowHdr& qh, int src_dev, bool peer, uint32_t nblk, long long tag,
                          int32_t ceiling, const int64_t* rows, int64_t n_rows, fn3_hit* out, int64_t cap, int64_t* n_out) {
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    const int rcr = maybe_refresh_heads(e);
    if (rcr) return rcr;
    RowHdr h = qh;
    if (h.data && !peer && src_dev != e->device) {
        const size_t nv = h.end[3];
        const size_t need = row_alloc_words(h.end[5], nv);
        if (e->qstage_busy) { CU(e, cudaStreamSynchronize(e->q_stream)); e->qstage_busy = false; }
        if (need * 2 > e->scratch_words) {
            const size_t want = std::max(need * 2, std::max<size_t>(e->scratch_words * 2, 1u << 16));
            CU(e, cudaStreamSynchronize(e->q_stream));
            if (e->d_scratch) cudaFree(e->d_scratch);
            e->d_scratch = nullptr; e->scratch_words = 0;
            CU(e, cudaMalloc(&e->d_scratch, want * 4));
            e->scratch_words = want;
        }
        CU(e, cudaMemcpyPeerAsync(e->d_scratch, e->device, (const void*)(uintptr_t)h.data, src_dev, need * 4, e->q_stream));
        h.data = (unsigned long long)(uintptr_t)e->d_scratch;
    }
    QueryJob job; job.ceiling = ceiling;
    snapshot_store(e, job);
    set_slot0(e, h, tag, -1);
    e->slot_blocks[0] = nblk;
    return run_one_vs_all(e, job, rows, n_rows, out, cap, n_out);
}

// per-device result buffers sized for the worst case (every local row a neighbour); local rows -> global rows on the way out
static void group_size_hits(fn3_group* g, int d, long long extra = 1) {
    fn3_group::Worker& w = g->w[(size_t)d];
    long long n;
    { std::lock_guard<std::mutex> s(g->eng[(size_t)d]->store_mu); n = g->eng[(size_t)d]->n_rows + extra; }
    if ((long long)w.hits.size() < n) w.hits.resize((size_t)std::max<long long>(n, 2 * (long long)w.hits.size()));
}

static int group_merge_hits(fn3_group* g, fn3_hit* out, int64_t cap, int64_t* n_out) {
    long long total = 0;
    for (int d = 0; d < g->G; ++d) total += g->w[(size_t)d].n_out;
    *n_out = total;
    if (total > cap) return gset_err(g, FN3_ECAPACITY, "output buffer holds %lld hits, %lld needed", (long long)cap, total);
    long long at = 0;
    for (int d = 0; d < g->G; ++d) {
        const fn3_group::Worker& w = g->w[(size_t)d];
        for (int64_t i = 0; i < w.n_out; ++i) {
            out[at] = w.hits[(size_t)i];
            out[at].row = w.hits[(size_t)i].row * g->G + d;
            ++at;
        }
    }
    return FN3_OK;
}

// explicit candidate list (global rows) -> one local list per device; the query itself is left out (mcompare, :166)
static int group_split_rows(fn3_group* g, const int64_t* rows, int64_t n_rows, long long skip) {
    for (int d = 0; d < g->G; ++d) g->w[(size_t)d].rows.clear();
    for (int64_t i = 0; i < n_rows; ++i) {
        const long long r = rows[i];
        if (r < 0 || r >= g->n_global) return gset_err(g, FN3_ENOTFOUND, "candidate row %lld does not exist", r);
        if (r == skip) continue;
        g->w[(size_t)(r % g->G)].rows.push_back(r / g->G);
    }
    return FN3_OK;
}

extern "C" int fn3_group_insert_query(fn3_group* g, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                                      int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!g || !n_out || !out_row || (cap > 0 && !out) || (!invalid && !off))
        return gset_err(g, FN3_EARG, "fn3_group_insert_query: bad arguments");
    std::lock_guard<std::mutex> l(g->mu);
    const long long gid = g->n_global;
    const int owner = (int)(gid % g->G);
    *n_out = 0; *out_row = -1;
    int64_t local = -1;
    const int rc = group_run(g, [&](int d) -> int {
        fn3_group::Worker& w = g->w[(size_t)d];
        group_size_hits(g, d);
        w.n_out = 0;
        fn3_engine* e = g->eng[(size_t)d];
        if (d == owner) return fn3_insert_query(e, pos, off, invalid, ceiling, &local, w.hits.data(), (int64_t)w.hits.size(), &w.n_out);
        return fn3_lists_vs_all(e, pos, off, invalid, ceiling, nullptr, 0, w.hits.data(), (int64_t)w.hits.size(), &w.n_out);
    });
    if (local >= 0) {                       // the owner stored the row (even if another device then failed)
        if (local != gid / g->G) return gset_err(g, FN3_EARG, "fn3_group_insert_query: engine %d is out of step with the group", owner);
        g->n_global = gid + 1;
        *out_row = gid;
    }
    if (rc) return rc;
    return group_merge_hits(g, out, cap, n_out);
}

extern "C" int fn3_group_lists_vs_all(fn3_group* g, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                                      const int64_t* rows, int64_t n_rows, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!g || !n_out || (cap > 0 && !out) || (!invalid && !off) || (rows && n_rows < 0))
        return gset_err(g, FN3_EARG, "fn3_group_lists_vs_all: bad arguments");
    std::lock_guard<std::mutex> l(g->mu);
    *n_out = 0;
    int rc;
    if (rows && (rc = group_split_rows(g, rows, n_rows, -1))) return rc;
    rc = group_run(g, [&](int d) -> int {
        fn3_group::Worker& w = g->w[(size_t)d];
        group_size_hits(g, d);
        w.n_out = 0;
        return fn3_lists_vs_all(g->eng[(size_t)d], pos, off, invalid, ceiling, rows ? w.rows.data() : nullptr,
                                rows ? (int64_t)w.rows.size() : 0, w.hits.data(), (int64_t)w.hits.size(), &w.n_out);
    });
    if (rc) return rc;
    return group_merge_hits(g, out, cap, n_out);
}

extern "C" int fn3_group_one_vs_all(fn3_group* g, int64_t query_row, int32_t ceiling, const int64_t* rows, int64_t n_rows,
                                    fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!g || !n_out || (cap > 0 && !out) || (rows && n_rows < 0)) return gset_err(g, FN3_EARG, "fn3_group_one_vs_all: bad arguments");
    std::lock_guard<std::mutex> l(g->mu);
    *n_out = 0;
    RowHdr qh; uint32_t nblk = 0;
    int rc;
    if ((rc = group_hdr(g, query_row, &qh, &nblk, nullptr, "fn3_group_one_vs_all"))) return rc;
    if (rows && (rc = group_split_rows(g, rows, n_rows, query_row))) return rc;
    const int owner = (int)(query_row % g->G);
    rc = group_run(g, [&](int d) -> int {
        fn3_group::Worker& w = g->w[(size_t)d];
        group_size_hits(g, d);
        w.n_out = 0;
        fn3_engine* e = g->eng[(size_t)d];
        const int64_t* r = rows ? w.rows.data() : nullptr;
        const int64_t nr = rows ? (int64_t)w.rows.size() : 0;
        if (rows && nr == 0) return FN3_OK;
        if (d == owner) return fn3_one_vs_all(e, query_row / g->G, ceiling, r, nr, w.hits.data(), (int64_t)w.hits.size(), &w.n_out);
        return foreign_vs_all(e, qh, g->dev[(size_t)owner], g->peer_all, nblk, query_row, ceiling, r, nr, w.hits.data(),
                              (int64_t)w.hits.size(), &w.n_out);
    });
    if (rc) return rc;
    return group_merge_hits(g, out, cap, n_out);
}

extern "C" int fn3_group_pair(fn3_group* g, int64_t row1, int64_t row2, int32_t ceiling, fn3_hit* out, int32_t* has_dist) {
    if (!g || !out || !has_dist) return gset_err(g, FN3_EARG, "fn3_group_pair: bad arguments");
    std::lock_guard<std::mutex> l(g->mu);
    *has_dist = 0;
    RowHdr qh, ch; uint32_t nblk = 0;
    int rc;
    if ((rc = group_hdr(g, row1, &qh, &nblk, nullptr, "fn3_group_pair"))) return rc;
    if ((rc = group_hdr(g, row2, &ch, nullptr, nullptr, "fn3_group_pair"))) return rc;
    const int d1 = (int)(row1 % g->G), d2 = (int)(row2 % g->G);
    fn3_engine* e = g->eng[(size_t)d2];                 // the pair is computed where the candidate (seq2) lives
    if (d1 == d2) {
        rc = gfrom(g, d2, fn3_pair(e, row1 / g->G, row2 / g->G, ceiling, out, has_dist));
    } else {
        int64_t n = 0; const int64_t r2 = row2 / g->G;
        rc = gfrom(g, d2, foreign_vs_all(e, qh, g->dev[(size_t)d1], g->peer_all, nblk, row1, ceiling, &r2, 1, out, 1, &n));
        *has_dist = (rc == FN3_OK && n == 1) ? 1 : 0;
    }
    if (rc == FN3_OK && *has_dist) out->row = row2;
    return rc;
}

extern "C" int fn3_group_ingest_query(fn3_group* g, const char* sequence, int64_t max_ns, int32_t ceiling, int64_t* hist,
                                      int32_t* pos, int64_t pos_cap, int32_t off[7], char* m_chars, int64_t m_cap, int64_t* n_pos,
                                      int32_t* invalid, int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!g || !sequence || !off || !n_pos || !invalid || !out_row || !n_out || (cap > 0 && !out))
        return gset_err(g, FN3_EARG, "fn3_group_ingest_query: bad arguments");
    *n_out = 0; *out_row = -1; *invalid = 0;
    int owner;
    { std::lock_guard<std::mutex> l(g->mu); owner = (int)(g->n_global % g->G); }
    {
        // compress on the device that will own the row (its copy of the sequence, its ingest stream)
        fn3_engine* e = g->eng[(size_t)owner];
        std::lock_guard<std::mutex> il(e->ingest_mu);
        const int rc = compress_locked(e, sequence, pos, pos_cap, off, m_chars, m_cap, n_pos, hist, "fn3_group_ingest_query");
        if (rc) return gfrom(g, owner, rc);
    }
    *invalid = (int64_t)(off[6] - off[4]) > max_ns ? 1 : 0;
    return fn3_group_insert_query(g, pos, off, *invalid, ceiling, out_row, out, cap, n_out);
}

// distmat over a SHARDED store: every device takes every live row of the group as a query — the rows of the other shards
// are read over NVLink where they live — against its own shard.  upper_only keeps global row1 < global row2: candidate
// c of device d is global row c*G + d, so the query g_q skips local rows <= floor((g_q - d) / G).
extern "C" int fn3_group_all_vs_all(fn3_group* g, int32_t ceiling, int32_t upper_only, fn3_edge* out, int64_t cap, int64_t* n_out) {
    if (!g || !n_out || (cap > 0 && !out)) return gset_err(g, FN3_EARG, "fn3_group_all_vs_all: bad arguments");
    std::lock_guard<std::mutex> l(g->mu);
    *n_out = 0;
    const int G = g->G;
    if (G > 1 && !g->peer_all)
        return gset_err(g, FN3_ECUDA, "fn3_group_all_vs_all: the devices of this group cannot read each other's memory (no peer access)");
    struct GQ { RowHdr hdr; uint32_t nblk; long long row; };
    std::vector<GQ> all;
    for (int d = 0; d < G; ++d) {
        fn3_engine* e = g->eng[(size_t)d];
        std::lock_guard<std::mutex> s(e->store_mu);
        for (long long r = 0; r < e->n_rows; ++r)
            if (e->h_flags[(size_t)r] == 0) { GQ q; q.hdr = e->h_hdr[(size_t)r]; q.nblk = e->h_nblk[(size_t)r]; q.row = r * G + d; all.push_back(q); }
    }
    std::sort(all.begin(), all.end(), [](const GQ& a, const GQ& b) { return a.row < b.row; });
    std::vector<long long> counts((size_t)G, 0);
    int rc = group_run(g, [&](int d) -> int {
        fn3_engine* e = g->eng[(size_t)d];
        std::lock_guard<std::mutex> q(e->query_mu);
        FN3_ON_DEVICE(e);
        const int rcr = maybe_refresh_heads(e);
        if (rcr) return rcr;
        QueryJob job; job.ceiling = ceiling;
        snapshot_store(e, job);
        std::vector<AvaQuery> qs;
        qs.reserve(all.size());
        for (const GQ& s : all) {
            AvaQuery a; a.hdr = s.hdr; a.tag = s.row; a.nblk = s.nblk;
            a.exclude = (s.row % G == d) ? s.row / G : -1;
            a.skip_le = !upper_only ? -1 : (s.row >= d ? (s.row - d) / G : -1);
            if (a.skip_le >= job.n_rows - 1 && upper_only) continue;          // nothing above it in this shard
            qs.push_back(a);
        }
        long long total = 0;
        int r2 = all_vs_all_core(e, job, qs, &total);
        if (r2) return r2;
        fn3_group::Worker& w = g->w[(size_t)d];
        if ((long long)w.edges.size() < total) w.edges.resize((size_t)total);
        counts[(size_t)d] = total;
        return total > 0 ? fetch_edges(e, w.edges.data(), total) : FN3_OK;
    });
    if (rc) return rc;
    long long total = 0;
    for (int d = 0; d < G; ++d) total += counts[(size_t)d];
    *n_out = total;
    // local candidate rows -> global rows (kept in the workers' buffers so that a too-small `out` can be served by a second call)
    if (total > cap) return gset_err(g, FN3_ECAPACITY, "output buffer holds %lld edges, %lld needed", (long long)cap, total);
    long long at = 0;
    for (int d = 0; d < G; ++d) {
        const fn3_group::Worker& w = g->w[(size_t)d];
        for (long long i = 0; i < counts[(size_t)d]; ++i) { out[at] = w.edges[(size_t)i]; out[at].row2 = w.edges[(size_t)i].row2 * G + d; ++at; }
    }
    return FN3_OK;
}

// ---- cluster statistics over rows of several shards: virtual head tiles on device 0, rows read over NVLink --------------

static int group_virtual(fn3_group* g, const int64_t* rows, int64_t n, bool reject_invalid, const char* who, std::vector<RowHdr>& hdrs) {
    if (g->G > 1 && !g->peer_all)
        return gset_err(g, FN3_ECUDA, "%s: the devices of this group cannot read each other's memory (no peer access)", who);
    hdrs.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        uint8_t fl = 0;
        const int rc = group_hdr(g, rows[i], &hdrs[(size_t)i], nullptr, &fl, who);
        if (rc) return rc;
        if (reject_invalid && (fl & 1)) return gset_err(g, FN3_EARG, "%s: row %lld is an invalid sample", who, (long long)rows[i]);
    }
    return FN3_OK;
}

extern "C" int fn3_group_consensus(fn3_group* g, const int64_t* rows, int64_t n_rows, int64_t min_votes, int32_t* pos,
                                   int64_t pos_cap, int32_t off[6], int64_t* n_pos) {
    if (!g || !off || !n_pos || n_rows < 0 || (n_rows > 0 && !rows) || min_votes < 1)
        return gset_err(g, FN3_EARG, "fn3_group_consensus: bad arguments (min_votes must be >= 1)");
    std::lock_guard<std::mutex> l(g->mu);
    std::vector<RowHdr> hdrs;
    int rc;
    if ((rc = group_virtual(g, rows, n_rows, false, "fn3_group_consensus", hdrs))) return rc;
    fn3_engine* e = g->eng[0];
    std::lock_guard<std::mutex> q(e->query_mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return gset_err(g, FN3_ECUDA, "cudaSetDevice failed");
    HdrTable ht;
    if ((rc = virtual_rows(e, hdrs.data(), n_rows, &ht))) return gfrom(g, 0, rc);
    return gfrom(g, 0, consensus_core(e, ht, n_rows, min_votes, pos, pos_cap, off, n_pos, "fn3_group_consensus"));
}

extern "C" int fn3_group_consensus_lists(fn3_group* g, int64_t n, const int32_t* pos_in, const int64_t* off_in, int64_t min_votes,
                                         int32_t* pos, int64_t pos_cap, int32_t off[6], int64_t* n_pos) {
    if (!g) return FN3_EARG;
    return gfrom(g, 0, fn3_consensus_lists(g->eng[0], n, pos_in, off_in, min_votes, pos, pos_cap, off, n_pos));
}

extern "C" int fn3_group_msa_sites_rule(fn3_group* g, const int64_t* rows, int64_t n_rows, int32_t rule, int32_t* sites,
                                        int64_t sites_cap, int64_t* n_sites) {
    if (!g || !n_sites || n_rows < 0 || (n_rows > 0 && !rows) || (rule != FN3_SITES_CALLS && rule != FN3_SITES_CALLS_OR_N))
        return gset_err(g, FN3_EARG, "fn3_group_msa_sites: bad arguments");
    *n_sites = 0;
    if (n_rows == 0) return FN3_OK;
    std::lock_guard<std::mutex> l(g->mu);
    std::vector<RowHdr> hdrs;
    int rc;
    if ((rc = group_virtual(g, rows, n_rows, true, "fn3_group_msa_sites", hdrs))) return rc;
    fn3_engine* e = g->eng[0];
    std::lock_guard<std::mutex> q(e->query_mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return gset_err(g, FN3_ECUDA, "cudaSetDevice failed");
    HdrTable ht;
    if ((rc = virtual_rows(e, hdrs.data(), n_rows, &ht))) return gfrom(g, 0, rc);
    return gfrom(g, 0, msa_sites_core(e, ht, n_rows, rule, sites, sites_cap, n_sites));
}

extern "C" int fn3_group_msa_align(fn3_group* g, const int64_t* rows, int64_t n_rows, const int32_t* sites, int64_t n_sites,
                                   uint8_t* codes, int32_t* counts) {
    if (!g || n_rows < 0 || n_sites < 0 || (n_rows > 0 && (!rows || !counts)) || (n_sites > 0 && !sites))
        return gset_err(g, FN3_EARG, "fn3_group_msa_align: bad arguments");
    int rc;
    if ((rc = msa_align_check_sites(g->eng[0], sites, n_sites))) return gfrom(g, 0, rc);
    if (n_rows == 0) return FN3_OK;
    std::lock_guard<std::mutex> l(g->mu);
    std::vector<RowHdr> hdrs;
    if ((rc = group_virtual(g, rows, n_rows, true, "fn3_group_msa_align", hdrs))) return rc;
    fn3_engine* e = g->eng[0];
    std::lock_guard<std::mutex> q(e->query_mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return gset_err(g, FN3_ECUDA, "cudaSetDevice failed");
    HdrTable ht;
    if ((rc = virtual_rows(e, hdrs.data(), n_rows, &ht))) return gfrom(g, 0, rc);
    return gfrom(g, 0, msa_align_core(e, ht, n_rows, sites, n_sites, codes, counts));
}
