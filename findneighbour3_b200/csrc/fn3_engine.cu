// fn3_engine.cu — B200 (sm_100a) SNV-distance engine: HBM-resident sample store + the C-ABI of include/fn3.h.
// Kernels live in fn3_kernels.cuh (query build, k_query) and fn3_compress.cuh (device compress).
//
// What it replaces in the reference (davidhwyllie/findNeighbour3, src/seqComparer.py):
//   store            seqProfile dict + persist/remove               :108, :119-134
//   compare          countDifferences (set algebra per base)        :424-458
//   pair statistics  countDifferences_byKey + _setStats             :382-421, :354-380
//   one-vs-all       mcompare                                       :149-172
//   all-vs-all       distmat                                        :174-197
//   ingest           compress                                       :274-307
//
// Layout in HBM: fn3_types.h / DESIGN.md §3.  No CPU path exists in this file: every distance is produced
// by k_query on the device.

#include "fn3.h"
#include "fn3_types.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <chrono>

static_assert(sizeof(RowHdr) == 32, "RowHdr must be one 32-byte sector");
static_assert(sizeof(RowHead) == 96, "RowHead must be 96 bytes");
static_assert(sizeof(QEntry) == 32, "QEntry must be one 32-byte sector");
static_assert(sizeof(EdgeRec) == sizeof(fn3_edge), "EdgeRec/fn3_edge mismatch");

#include "fn3_kernels.cuh"
#include "fn3_scan.cuh"
#include "fn3_compress.cuh"
#include "fn3_cluster.cuh"
#include "fn3_pickle.h"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

static thread_local std::string g_create_error;

struct fn3_engine {
    int device = 0;
    int sm_count = 148;
    long long L = 0;
    int pos_bits = 32;
    uint32_t max_run = 1;

    std::mutex store_mu;        // guards store metadata + append path
    std::mutex query_mu;        // guards query state (slots, survivor buffers, q_stream); taken before store_mu
    std::string err;
    std::mutex err_mu;

    cudaStream_t copy_stream = nullptr, q_stream = nullptr;   // appends / queries

    // store
    size_t chunk_words = 0;                 // words per data chunk
    std::vector<uint32_t*> data_chunks;
    std::vector<size_t> data_chunk_cap;     // words
    size_t cur_used = 0;                    // words used in the last chunk
    int hdr_shift = 20;
    int sb_shift = 13;                      // rows per superblock of head tiles (<= hdr_shift)
    std::vector<unsigned char*> hdr_chunks; // device: head tiles (fn3_types.h)
    std::vector<RowHdr> h_hdr;              // host mirror (published rows)
    std::vector<uint8_t> h_flags;           // bit0 invalid, bit1 removed
    std::vector<uint32_t> h_npos;           // canonical position count per row
    std::vector<uint32_t> h_nblk;           // number of 32-position blocks the row touches (sizes query staging)
    long long n_rows = 0;                   // published
    long long n_live = 0, n_invalid = 0;
    long long store_words = 0, store_bytes = 0, positions = 0;

    // staging
    uint32_t* stage = nullptr; size_t stage_words = 0;     // pinned (append path, store_mu)
    uint32_t* istage = nullptr; size_t istage_words = 0;   // pinned (fn3_insert_query, query_mu)
    RowHead* stage_hdr = nullptr; size_t stage_hdr_n = 0;  // pinned, row order
    RowHead* d_heads = nullptr; size_t d_heads_n = 0;      // device staging for k_scatter_heads

    // query state
    int n_slots = 128;                     // queries per all-vs-all tile (one k_query launch)
    QueryState qstate;                     // global-memory query structures, fallback path only (ensure_qstate)
    int qstate_slots = 0;
    int nw = 0, ncw = 0, n_blk = 0, fshift = 5;   // bitmap words, coarse map words, 32-position blocks, fingerprint shift
    uint16_t fp_pad = 0;                   // coarse id carried by padded fingerprints (its map byte is always 1)
    uint16_t fp_dead = 0;                  // coarse id carried by the heads of invalid / removed rows (its map byte is always 0)
    int max_smem = 0;                      // opt-in dynamic shared memory per CTA
    int stage_limit = 0;                   // = max_smem, or 0 to force the global-memory query variants
    int q_static_smem = 0;                 // static shared memory of k_query
    int debug_stop = 0;                    // env FN3_DEBUG_STOP (profiling aid, wrong results when set)
    int scan_ss = -1;                      // env FN3_SCAN_SS: forces k_scan's split of the 32-row tiles (profiling aid)
    uint32_t slot_blocks[1];               // touched-block bound of the single query of a one-vs-all call (sizes smem)
    QuerySlot* d_slots = nullptr; QuerySlot* h_slots = nullptr;   // [FN3_SLOT_RING tiles][FN3_MAX_SLOTS]; h_slots pinned
    // survivor buffer: one allocation = [32-byte counter block | records]; the pinned mirror is used by the
    // copy-based paths (all-vs-all, FN3_ZERO_COPY=0)
    unsigned char* d_outbuf = nullptr; unsigned char* h_outbuf = nullptr;
    EdgeRec* d_out = nullptr; long long out_cap = 0;
    unsigned long long* d_count = nullptr;   // [0] hit counter, [1] touched bytes
    unsigned long long* h_count = nullptr;
    EdgeRec* h_out = nullptr; long long h_out_cap = 0;
    std::atomic<long long> h2d_bytes{0}, d2h_bytes{0};
    // zero-copy completion of one-vs-all queries (mapped pinned memory)
    unsigned char* zc_buf = nullptr;         // [64-byte flag block | FN3_ZC_CAP records]
    unsigned int* d_done = nullptr;
    unsigned long long zc_seq = 0;
    bool count_dirty = false;                // the device counter was left non-zero by a path that does not re-arm it
    bool zero_copy = true;
    long long t_pack_ns = 0, t_enqueue_ns = 0, t_wait_ns = 0, n_timed = 0;   // fn3_insert_query host-time split
    long long edges_ready = -1;              // edges of the last fn3_all_vs_all still in d_out (-1: none)
    long long* d_rows = nullptr; long long d_rows_cap = 0;
    long long* h_rows = nullptr; long long h_rows_cap = 0;       // pinned
    // scratch rows for samples that are not stored (pair_lists / lists_vs_all)
    uint32_t* d_scratch = nullptr; size_t scratch_words = 0;
    uint32_t* qstage = nullptr; size_t qstage_words = 0;      // pinned staging of the scratch samples (two halves)
    bool qstage_busy = false;                                // a copy from qstage may still be in flight
    unsigned char* d_scratch_tile = nullptr; // one head tile for fn3_pair_lists' candidate
    uint4* d_flush = nullptr; long long flush_n = 0;
    // ingest (fn3_compress)
    std::mutex ingest_mu;
    cudaStream_t in_stream = nullptr;
    long long Lpad = 0; int cmp_ctas = 0;
    uint8_t* d_refmask = nullptr; uint8_t* d_seq = nullptr; uint8_t* d_mchars = nullptr;
    int32_t* d_cpos = nullptr; uint32_t* d_ccounts = nullptr; uint32_t* d_ctotals = nullptr;
    uint8_t* h_seq = nullptr; uint32_t* h_ctotals = nullptr;   // pinned
    int32_t* h_cpos = nullptr; size_t h_cpos_cap = 0; uint8_t* h_mchars = nullptr; size_t h_mchars_cap = 0;   // pinned
    unsigned long long* d_hist = nullptr; unsigned long long* h_hist = nullptr;   // byte histogram of the sequence (fn3_ingest_query)

    // cluster statistics (fn3_consensus*, fn3_msa_*): vote planes + selection scratch, allocated on first use
    uint32_t* d_votes = nullptr; long long vLpad = 0; int sel_ctas = 0;
    uint8_t* d_refcode = nullptr;          // reference base codes 0..3, unmasked (fn3_set_reference; needed by fn3_msa_sites)
    uint32_t* d_sel_counts = nullptr; uint32_t* d_sel_totals = nullptr; uint32_t* h_sel_totals = nullptr;   // pinned
    int32_t* d_sel_out = nullptr; size_t sel_out_cap = 0;
    int32_t* d_sites = nullptr; size_t sites_cap = 0;
    uint8_t* d_codes = nullptr; size_t codes_cap = 0;
    int32_t* d_acounts = nullptr; size_t acounts_cap = 0;
    int32_t* d_lpos = nullptr; size_t lpos_cap = 0;
    unsigned char* d_vtiles = nullptr; size_t vtiles_cap = 0;   // head tiles of a virtual row set (engine groups)
    long long* d_loff = nullptr; size_t loff_cap = 0;

    // background-aware heads (fn3_refresh_heads): positions that are non-reference in many rows do not go into a head
    std::vector<uint32_t> bg_bits;          // host copy of the bitmap (fill_head); empty = no background known
    uint32_t* d_bg = nullptr;
    bool bg_any = false;
    bool bg_auto = true;                    // refresh when the store has doubled since the last one (env FN3_BG_AUTO)
    long long bg_next_at = 4096;            // live rows at which the next automatic refresh is due
    long long bg_ppm = 10000;               // background = non-reference in >= 1 % of the live rows (env FN3_BG_PPM)
    long long bg_refreshes = 0;
    // FILTER <-> STREAM steering of full one-vs-all scans at low ceilings (an outbreak: every candidate passes the filter)
    double pass_ema = 0.0;
    bool prefer_scan = false;
    int scan_streak = 0;
    long long last_passed = 0;

    std::atomic<long long> launches{0};
};

static int set_err(fn3_engine* e, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (e) { std::lock_guard<std::mutex> g(e->err_mu); e->err = buf; }
    else g_create_error = buf;
    return code;
}

#define CU(e, call) do { cudaError_t _s = (call); if (_s != cudaSuccess) \
    return set_err((e), FN3_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_s), __FILE__, __LINE__); } while (0)

// Every entry point works on its engine's device and then gives the calling thread its own device back: a caller that also
// uses CUDA itself (PyTorch in bench.py; another engine on another GPU) must not find its current device changed.  The
// previous device is restored only if it has a live primary context (cuDevicePrimaryCtxGetState through dlsym: libcuda is
// not linked) — a thread that never chose a device reports 0, and "restoring" that would create a context on GPU 0.
static bool primary_ctx_active(int dev) {
    typedef int (*state_fn)(int, unsigned int*, int*);
    static state_fn fn = []() -> state_fn {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libcuda.so.1", RTLD_NOW);
        return h ? (state_fn)dlsym(h, "cuDevicePrimaryCtxGetState") : nullptr;
    }();
    if (!fn) return true;
    unsigned int flags = 0; int active = 0;
    return fn(dev, &flags, &active) != 0 || active != 0;
}

struct DeviceGuard {
    int prev = -1, dev;
    bool ok;
    explicit DeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; (void)cudaGetLastError(); }
        ok = prev == d || cudaSetDevice(d) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev < 0 || cudaGetDevice(&cur) != cudaSuccess || cur == prev) return;
        if (primary_ctx_active(prev)) (void)cudaSetDevice(prev);
    }
};
#define FN3_ON_DEVICE(e) DeviceGuard _dg((e)->device); \
    if (!_dg.ok) return set_err((e), FN3_ECUDA, "cudaSetDevice(%d) failed (%s:%d)", (e)->device, __FILE__, __LINE__)

static long long env_ll(const char* name, long long dflt) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return atoll(s);
}

extern "C" int fn3_abi_version(void) { return FN3_ABI_VERSION; }

extern "C" const char* fn3_last_error(fn3_engine* e) {
    if (!e) return g_create_error.c_str();
    std::lock_guard<std::mutex> g(e->err_mu);
    static thread_local std::string copy;
    copy = e->err;
    return copy.c_str();
}

extern "C" int64_t fn3_launch_count(fn3_engine* e) { return e ? e->launches.load() : 0; }

#define FN3_SLOT_RING 32        // all-vs-all: query tiles queued between two host synchronisations
#define FN3_ZC_CAP 4096         // survivor records the mapped host buffer holds (longer lists: one extra copy)
#define FN3_OUT_HDR 32
#define FN3_FIRST_HITS 256      // records fetched together with the counter

static int ensure_out(fn3_engine* e, long long cap) {
    if (cap <= e->out_cap) return FN3_OK;
    long long want = std::max<long long>(cap, e->out_cap * 2);
    want = std::max<long long>(want, 4096);
    if (e->d_outbuf) { cudaStreamSynchronize(e->q_stream); cudaFree(e->d_outbuf); }
    e->d_outbuf = nullptr; e->out_cap = 0;
    CU(e, cudaMalloc(&e->d_outbuf, FN3_OUT_HDR + (size_t)want * sizeof(EdgeRec)));
    CU(e, cudaMemsetAsync(e->d_outbuf, 0, FN3_OUT_HDR, e->q_stream));
    e->d_count = reinterpret_cast<unsigned long long*>(e->d_outbuf);
    e->d_out = reinterpret_cast<EdgeRec*>(e->d_outbuf + FN3_OUT_HDR);
    e->out_cap = want;
    return FN3_OK;
}

static int ensure_hout(fn3_engine* e, long long cap) {
    if (cap <= e->h_out_cap) return FN3_OK;
    long long want = std::max<long long>(cap, e->h_out_cap * 2);
    want = std::max<long long>(want, 4096);
    if (e->h_outbuf) cudaFreeHost(e->h_outbuf);
    e->h_outbuf = nullptr; e->h_out_cap = 0;
    CU(e, cudaHostAlloc(&e->h_outbuf, FN3_OUT_HDR + (size_t)want * sizeof(EdgeRec), cudaHostAllocDefault));
    e->h_count = reinterpret_cast<unsigned long long*>(e->h_outbuf);
    e->h_out = reinterpret_cast<EdgeRec*>(e->h_outbuf + FN3_OUT_HDR);
    e->h_out_cap = want;
    return FN3_OK;
}

static int ensure_rows(fn3_engine* e, long long n) {
    if (n > e->d_rows_cap) {
        long long want = std::max<long long>(n, e->d_rows_cap * 2);
        if (e->d_rows) cudaFree(e->d_rows);
        e->d_rows = nullptr; e->d_rows_cap = 0;
        CU(e, cudaMalloc(&e->d_rows, (size_t)want * sizeof(long long)));
        e->d_rows_cap = want;
    }
    if (n > e->h_rows_cap) {
        long long want = std::max<long long>(n, e->h_rows_cap * 2);
        if (e->h_rows) cudaFreeHost(e->h_rows);
        e->h_rows = nullptr; e->h_rows_cap = 0;
        CU(e, cudaHostAlloc(&e->h_rows, (size_t)want * sizeof(long long), cudaHostAllocDefault));
        e->h_rows_cap = want;
    }
    return FN3_OK;
}

static int ensure_stage(fn3_engine* e, size_t words, size_t hdrs) {
    if (words > e->stage_words) {
        size_t want = std::max(words, e->stage_words * 2);
        want = std::max<size_t>(want, 1u << 16);
        if (e->stage) cudaFreeHost(e->stage);
        e->stage = nullptr; e->stage_words = 0;
        CU(e, cudaHostAlloc(&e->stage, want * 4, cudaHostAllocDefault));
        e->stage_words = want;
    }
    if (hdrs > e->stage_hdr_n) {
        size_t want = std::max(hdrs, e->stage_hdr_n * 2);
        want = std::max<size_t>(want, 1024);
        if (e->stage_hdr) cudaFreeHost(e->stage_hdr);
        e->stage_hdr = nullptr; e->stage_hdr_n = 0;
        CU(e, cudaHostAlloc(&e->stage_hdr, want * sizeof(RowHead), cudaHostAllocDefault));
        e->stage_hdr_n = want;
    }
    return FN3_OK;
}

extern "C" fn3_engine* fn3_create(int64_t genome_length, int32_t device) {
    if (genome_length <= 0 || genome_length >= (1ll << 31)) {
        set_err(nullptr, FN3_EARG, "fn3_create: genome_length %lld out of range (1 .. 2^31-1)", (long long)genome_length);
        return nullptr;
    }
    int ndev = 0;
    cudaError_t s = cudaGetDeviceCount(&ndev);
    if (s != cudaSuccess || ndev <= 0) {
        set_err(nullptr, FN3_ECUDA, "fn3_create: no CUDA device (%s); this engine has no CPU path",
                s == cudaSuccess ? "device count 0" : cudaGetErrorString(s));
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        set_err(nullptr, FN3_EARG, "fn3_create: device %d out of range (0..%d)", device, ndev - 1);
        return nullptr;
    }
    fn3_engine* e = new fn3_engine();
    e->device = device;
    e->L = genome_length;
    auto fail = [&](const char* what, cudaError_t st) -> fn3_engine* {
        set_err(nullptr, FN3_ECUDA, "fn3_create: %s failed: %s", what, cudaGetErrorString(st));
        fn3_destroy(e);
        return nullptr;
    };
    DeviceGuard dg(device);
    if (!dg.ok) return fail("cudaSetDevice", cudaGetLastError());
    cudaDeviceProp prop;
    if ((s = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", s);
    e->sm_count = prop.multiProcessorCount;
    int pb = 1; while ((1ll << pb) < genome_length) ++pb;
    e->pos_bits = pb;
    e->max_run = (pb >= 32) ? 1u : (uint32_t)std::min<long long>(1ll << (32 - pb), 1ll << 30);
    e->chunk_words = (size_t)env_ll("FN3_CHUNK_MB", 256) * (1u << 20) / 4;
    if (e->chunk_words < 1024) e->chunk_words = 1024;
    e->hdr_shift = (int)env_ll("FN3_HDR_SHIFT", 20);
    if (e->hdr_shift < 5) e->hdr_shift = 5;                     // a chunk holds whole tiles of 32 rows
    if (e->hdr_shift > 26) e->hdr_shift = 26;
    e->sb_shift = std::min(13, e->hdr_shift);                   // 8192-row superblocks: neighbours in insertion
                                                                // order sit 256 tiles apart
    e->debug_stop = (int)env_ll("FN3_DEBUG_STOP", 0);
    e->scan_ss = (int)env_ll("FN3_SCAN_SS", -1);
    e->bg_auto = env_ll("FN3_BG_AUTO", 1) != 0;
    e->bg_ppm = std::max<long long>(1, env_ll("FN3_BG_PPM", 10000));
    // queries per all-vs-all tile (one k_query launch): one per SM, so that with two CTAs per query every resident CTA
    // slot of the GPU (2 per SM) is filled by one wave — 148 on a B200 (128 left 40 of the 296 slots idle)
    e->n_slots = (int)env_ll("FN3_SLOTS", e->sm_count);
    if (e->n_slots < 1) e->n_slots = 1;
    if (e->n_slots > FN3_MAX_SLOTS) e->n_slots = FN3_MAX_SLOTS;

    if ((s = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", s);
    if ((s = cudaStreamCreateWithFlags(&e->q_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", s);

    // query structure sizes; the last 32-position block is the end sentinel for rank(L)
    {
        e->n_blk = (int)((genome_length + 31) / 32 + 1);
        e->nw = ((e->n_blk + 31) / 32 + 3) & ~3;                 // entries start 32-byte aligned behind bmpref
        e->fshift = 5;
        while (((genome_length - 1) >> e->fshift) + 2 > 16383) ++e->fshift;  // <= 16 K coarse blocks: a 16 KB byte map
        const long long n_coarse = ((genome_length - 1) >> e->fshift) + 2;   // + the block rank(L) may fall in
        e->ncw = (int)((((n_coarse + 2) + 3) / 4 + 7) & ~7ll);               // bytes / 4, + 2 for the pad and dead ids
        e->fp_pad = (uint16_t)(e->ncw * 4 - 1);                              // byte always 1: never a mismatch
        e->fp_dead = (uint16_t)(e->ncw * 4 - 2);                             // byte always 0: always a mismatch
        QueryState& q = e->qstate;
        memset(&q, 0, sizeof q);
        q.ent_stride = (long long)e->n_blk + 1;
        // (the global-memory copies behind `q` are only needed by queries too large for shared memory:
        //  ensure_qstate allocates them on first use)
        e->max_smem = (int)prop.sharedMemPerBlockOptin;
        e->stage_limit = env_ll("FN3_FORCE_GLOBAL_QUERY", 0) ? 0 : e->max_smem;   // tests: exercise the global-memory variants
        // opt in to the full shared memory for the variants that build the query in shared memory
        // (k_scan always needs dynamic shared memory: its row rings)
        const void* staged[6] = {(const void*)k_query<false, 1>, (const void*)k_query<true, 1>,
                                 (const void*)k_scan<false, 1>, (const void*)k_scan<true, 1>,
                                 (const void*)k_scan<false, 0>, (const void*)k_scan<true, 0>};
        e->q_static_smem = 0;
        for (int i = 0; i < 6; ++i) {
            cudaFuncAttributes fa;
            if ((s = cudaFuncGetAttributes(&fa, staged[i])) != cudaSuccess) return fail("cudaFuncGetAttributes", s);
            e->q_static_smem = std::max(e->q_static_smem, (int)fa.sharedSizeBytes);
            if ((s = cudaFuncSetAttribute(staged[i], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          e->max_smem - (int)fa.sharedSizeBytes)) != cudaSuccess)
                return fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", s);
        }
    }
    if (env_ll("FN3_CARVEOUT", 1)) {
        // every kernel of the query path asks for the same (maximal) shared-memory carve-out: an SM that must switch its
        // L1 / shared split between two kernels drains first, which showed up as microseconds of launch floor
        const void* ks[] = {(const void*)k_query<false, 0>, (const void*)k_query<true, 0>, (const void*)k_query<false, 1>,
                            (const void*)k_query<true, 1>, (const void*)k_scan<false, 0>, (const void*)k_scan<true, 0>,
                            (const void*)k_scan<false, 1>, (const void*)k_scan<true, 1>, (const void*)k_put_head,
                            (const void*)k_scatter_heads, (const void*)k_fill, (const void*)k_build_query};
        for (const void* k : ks) (void)cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        (void)cudaGetLastError();
    }
    if ((s = cudaMalloc(&e->d_slots, sizeof(QuerySlot) * FN3_MAX_SLOTS * FN3_SLOT_RING)) != cudaSuccess) return fail("cudaMalloc(slots)", s);
    if ((s = cudaHostAlloc(&e->h_slots, sizeof(QuerySlot) * FN3_MAX_SLOTS * FN3_SLOT_RING, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc(slots)", s);
    if (ensure_out(e, 4096) != FN3_OK || ensure_hout(e, 4096) != FN3_OK) {
        set_err(nullptr, FN3_ENOMEM, "fn3_create: survivor buffers: %s", e->err.c_str());
        fn3_destroy(e);
        return nullptr;
    }
    e->zero_copy = env_ll("FN3_ZERO_COPY", 1) != 0 && e->debug_stop == 0;
    if ((s = cudaHostAlloc(&e->zc_buf, 64 + (size_t)FN3_ZC_CAP * sizeof(EdgeRec), cudaHostAllocMapped)) != cudaSuccess) return fail("cudaHostAlloc(mapped)", s);
    memset(e->zc_buf, 0, 64);
    // the insert path's pinned staging row: allocated now, not by the first insert this engine owns (a few ms of cudaHostAlloc
    // showed up as the three slowest steps of an eight-GPU group: each device's first turn as the owner)
    e->istage_words = 1u << 16;
    if ((s = cudaHostAlloc(&e->istage, e->istage_words * 4, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc(istage)", s);
    if ((s = cudaMalloc(&e->d_done, sizeof(unsigned int))) != cudaSuccess) return fail("cudaMalloc(done)", s);
    if ((s = cudaMemset(e->d_done, 0, sizeof(unsigned int))) != cudaSuccess) return fail("cudaMemset(done)", s);
    if ((s = cudaMalloc(&e->d_scratch_tile, FN3_TILE_BYTES)) != cudaSuccess) return fail("cudaMalloc(scratch tile)", s);
    if ((s = cudaMemset(e->d_scratch_tile, 0, FN3_TILE_BYTES)) != cudaSuccess) return fail("cudaMemset(scratch tile)", s);
    if ((s = cudaDeviceSynchronize()) != cudaSuccess) return fail("cudaDeviceSynchronize", s);
    e->store_bytes = 0;
    return e;
}

extern "C" int fn3_destroy(fn3_engine* e) {
    if (!e) return FN3_OK;
    DeviceGuard dg(e->device);
    if (e->q_stream) cudaStreamSynchronize(e->q_stream);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    for (auto p : e->data_chunks) cudaFree(p);
    for (auto p : e->hdr_chunks) cudaFree(p);
    cudaFree(e->qstate.bm); cudaFree(e->qstate.coarse); cudaFree(e->qstate.ent); cudaFree(e->qstate.info);
    cudaFree(e->d_heads);
    cudaFree(e->d_slots);
    cudaFree(e->d_outbuf); cudaFree(e->d_rows); cudaFree(e->d_scratch);
    cudaFree(e->d_scratch_tile); cudaFree(e->d_flush); cudaFree(e->d_done); cudaFree(e->d_bg);
    if (e->zc_buf) cudaFreeHost(e->zc_buf);
    cudaFree(e->d_refmask); cudaFree(e->d_seq); cudaFree(e->d_mchars); cudaFree(e->d_cpos);
    cudaFree(e->d_ccounts); cudaFree(e->d_ctotals);
    if (e->h_seq) cudaFreeHost(e->h_seq);
    if (e->h_ctotals) cudaFreeHost(e->h_ctotals);
    if (e->h_cpos) cudaFreeHost(e->h_cpos);
    if (e->h_mchars) cudaFreeHost(e->h_mchars);
    cudaFree(e->d_hist);
    if (e->h_hist) cudaFreeHost(e->h_hist);
    cudaFree(e->d_refcode); cudaFree(e->d_votes); cudaFree(e->d_sel_counts); cudaFree(e->d_sel_totals); cudaFree(e->d_sel_out);
    cudaFree(e->d_sites); cudaFree(e->d_codes); cudaFree(e->d_acounts); cudaFree(e->d_lpos); cudaFree(e->d_loff); cudaFree(e->d_vtiles);
    if (e->h_sel_totals) cudaFreeHost(e->h_sel_totals);
    if (e->in_stream) cudaStreamDestroy(e->in_stream);
    if (e->h_slots) cudaFreeHost(e->h_slots);
    if (e->h_outbuf) cudaFreeHost(e->h_outbuf);
    if (e->h_rows) cudaFreeHost(e->h_rows);
    if (e->stage) cudaFreeHost(e->stage);
    if (e->istage) cudaFreeHost(e->istage);
    if (e->qstage) cudaFreeHost(e->qstage);
    if (e->stage_hdr) cudaFreeHost(e->stage_hdr);
    if (e->q_stream) cudaStreamDestroy(e->q_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    (void)cudaGetLastError();          // a failed release must not surface as the next engine's error
    delete e;
    return FN3_OK;
}

// ---- packing: six plain lists -> row words ------------------------------------------------

// validates one sample; returns number of row words (0 for invalid), or -1 with error set
// *nblk receives an upper bound of the number of distinct 32-position blocks the sample touches
static long long row_words_needed(fn3_engine* e, const int32_t* pos, const long long off[7], int invalid,
                                  uint32_t* nblk = nullptr) {
    if (nblk) *nblk = 0;
    if (invalid) return 0;
    long long blocks = 0;
    for (int b = 0; b < 6; ++b)
        if (off[b + 1] < off[b]) { set_err(e, FN3_EARG, "list offsets must be non-decreasing"); return -1; }
    long long words = off[4] - off[0];
    for (int b = 0; b < 6; ++b) {
        const int32_t* p = pos + off[b];
        const long long n = off[b + 1] - off[b];
        long long runs = 0; long long run_len = 0;
        for (long long i = 0; i < n; ++i) {
            if (p[i] < 0 || (long long)p[i] >= e->L) {
                set_err(e, FN3_EARG, "position %d out of range [0,%lld)", p[i], e->L); return -1;
            }
            if (i && p[i] <= p[i - 1]) {
                set_err(e, FN3_EARG, "positions must be strictly increasing within a list"); return -1;
            }
            if (!i || (p[i] >> 5) != (p[i - 1] >> 5)) ++blocks;
            if (b >= 4) {
                if (i && p[i] == p[i - 1] + 1 && run_len < (long long)e->max_run) ++run_len;
                else { ++runs; run_len = 1; }
            }
        }
        if (b >= 4) words += runs;
    }
    if (words >= (1ll << 31)) { set_err(e, FN3_EARG, "sample too large"); return -1; }
    if (nblk) *nblk = (uint32_t)blocks;
    return words;
}

// writes the row words; fills ends[6]
static void pack_row(const fn3_engine* e, const int32_t* pos, const long long off[7], uint32_t* dst, uint32_t ends[6]) {
    uint32_t w = 0;
    for (int b = 0; b < 4; ++b) {
        const long long n = off[b + 1] - off[b];
        memcpy(dst + w, pos + off[b], (size_t)n * 4);
        w += (uint32_t)n;
        ends[b] = w;
    }
    for (int b = 4; b < 6; ++b) {
        const int32_t* p = pos + off[b];
        const long long n = off[b + 1] - off[b];
        long long i = 0;
        while (i < n) {
            long long j = i + 1;
            while (j < n && p[j] == p[j - 1] + 1 && (j - i) < (long long)e->max_run) ++j;
            const uint32_t len = (uint32_t)(j - i);
            dst[w++] = (e->pos_bits >= 32) ? (uint32_t)p[i] : (((len - 1u) << e->pos_bits) | (uint32_t)p[i]);
            i = j;
        }
        ends[b] = w;
    }
}

// true iff the six lists of a packed row are mutually disjoint (include/fn3.h: a position carries ONE call; the reference's
// compress() cannot produce anything else, seqComparer.py:292-298, but a hand-made sample could, and the kernel's per-position
// code map would then answer differently from the reference's set algebra).  One merge over six sorted streams — the V lists
// word by word, the N and M lists run by run — every item an interval [start, start + len): ~6 compares per row word.
static bool packed_disjoint(const fn3_engine* e, const uint32_t* w, const uint32_t ends[6]) {
    uint32_t at[6], end[6];
    for (int b = 0; b < 6; ++b) { at[b] = b ? ends[b - 1] : 0u; end[b] = ends[b]; }
    const uint32_t pmask = (e->pos_bits >= 32) ? 0xffffffffu : ((1u << e->pos_bits) - 1u);
    unsigned long long last_end = 0;
    for (;;) {
        int best = -1; uint32_t bs = 0;
        for (int b = 0; b < 6; ++b)
            if (at[b] < end[b]) {
                const uint32_t st = b < 4 ? w[at[b]] : (w[at[b]] & pmask);
                if (best < 0 || st < bs) { best = b; bs = st; }
            }
        if (best < 0) return true;
        const uint32_t x = w[at[best]++];
        const unsigned long long len = (best < 4 || e->pos_bits >= 32) ? 1ull : (unsigned long long)(x >> e->pos_bits) + 1ull;
        if ((unsigned long long)bs < last_end) return false;
        last_end = (unsigned long long)bs + len;
    }
}

// the same test on the caller's raw lists (bulk path, pass 1: everything is validated before anything is stored); the N and M
// lists are walked run by run (maximal stretches of consecutive positions)
static bool raw_disjoint(const int32_t* pos, const long long off[7]) {
    long long at[6];
    for (int b = 0; b < 6; ++b) at[b] = off[b];
    long long last_end = 0;
    for (;;) {
        int best = -1; int32_t bs = 0;
        for (int b = 0; b < 6; ++b)
            if (at[b] < off[b + 1] && (best < 0 || pos[at[b]] < bs)) { best = b; bs = pos[at[b]]; }
        if (best < 0) return true;
        long long j = at[best] + 1;
        if (best >= 4) while (j < off[best + 1] && pos[j] == pos[j - 1] + 1) ++j;
        if ((long long)bs < last_end) return false;
        last_end = (long long)bs + (j - at[best]);
        at[best] = j;
    }
}

// single pass for the insert hot path: validates the six lists AND writes the row words (dst must hold one word per
// position, the worst case).  Returns the number of words, or -1 with the error set; *nblk = touched-block bound.
static long long pack_checked(fn3_engine* e, const int32_t* pos, const long long off[7], uint32_t* dst, uint32_t ends[6],
                              uint32_t* nblk) {
    for (int b = 0; b < 6; ++b)
        if (off[b + 1] < off[b]) { set_err(e, FN3_EARG, "list offsets must be non-decreasing"); return -1; }
    const long long L = e->L;
    long long blocks = 0;
    uint32_t w = 0;
    for (int b = 0; b < 4; ++b) {
        const int32_t* p = pos + off[b];
        const long long n = off[b + 1] - off[b];
        int32_t prev = -1;
        for (long long i = 0; i < n; ++i) {
            const int32_t v = p[i];
            if (v <= prev || (long long)v >= L) {
                set_err(e, FN3_EARG, v < 0 || (long long)v >= L ? "position out of range" : "positions must be strictly increasing within a list");
                return -1;
            }
            blocks += (i == 0) | ((v >> 5) != (prev >> 5));
            dst[w++] = (uint32_t)v;
            prev = v;
        }
        ends[b] = w;
    }
    for (int b = 4; b < 6; ++b) {
        const int32_t* p = pos + off[b];
        const long long n = off[b + 1] - off[b];
        long long i = 0;
        int32_t prev_last = -1;
        while (i < n) {
            const int32_t s0 = p[i];
            if (s0 <= prev_last || (long long)s0 >= L) {
                set_err(e, FN3_EARG, s0 < 0 || (long long)s0 >= L ? "position out of range" : "positions must be strictly increasing within a list");
                return -1;
            }
            long long j = i + 1;
            while (j < n && p[j] == p[j - 1] + 1 && (j - i) < (long long)e->max_run) ++j;
            const int32_t last = p[j - 1];
            if ((long long)last >= L) { set_err(e, FN3_EARG, "position out of range"); return -1; }
            blocks += (last >> 5) - (s0 >> 5) + 1 - ((prev_last >= 0 && (prev_last >> 5) == (s0 >> 5)) ? 1 : 0);
            const uint32_t len = (uint32_t)(j - i);
            dst[w++] = (e->pos_bits >= 32) ? (uint32_t)s0 : (((len - 1u) << e->pos_bits) | (uint32_t)s0);
            prev_last = last;
            i = j;
        }
        ends[b] = w;
    }
    if (!packed_disjoint(e, dst, ends)) { set_err(e, FN3_EARG, "the six lists of a sample must be mutually disjoint"); return -1; }
    *nblk = (uint32_t)std::min<long long>(blocks, 0x7fffffff);
    return (long long)w;
}

// words a row occupies in the arena: its list words padded to 4, then ONE 16-bit fingerprint per V word (the STREAM
// variant's certain-mismatch pre-pass reads these 2 bytes per word instead of the 4-byte positions), padded to 16 bytes
static size_t row_alloc_words(size_t words, size_t nv) {
    const size_t a = ((words + 3) & ~(size_t)3) + ((((nv + 1) / 2) + 3) & ~(size_t)3);
    return a ? a : 4;
}

// dst[0..words) holds the packed lists: write the zero padding and the fingerprint array behind them
static void finish_row(const fn3_engine* e, uint32_t* dst, size_t words, size_t nv) {
    const size_t w4 = (words + 3) & ~(size_t)3, total = row_alloc_words(words, nv);
    for (size_t z = words; z < w4; ++z) dst[z] = 0;
    uint16_t* fp = reinterpret_cast<uint16_t*>(dst + w4);
    const size_t nfp = (total - w4) * 2;
    for (size_t i = 0; i < nv; ++i) fp[i] = (uint16_t)(dst[i] >> e->fshift);
    for (size_t i = nv; i < nfp; ++i) fp[i] = e->fp_pad;
}

// header + 16-bit fingerprints (position >> fshift) of the row's first V words (fn3_types.h)
static void fill_head(const fn3_engine* e, RowHead& hd, const RowHdr& h, const uint32_t* words) {
    hd.hdr = h;
    if (!h.data) {                         // invalid (or removed) row: 32 certain mismatches, the filter never lets it pass
        for (uint32_t i = 0; i < FN3_HEAD_FP; ++i) hd.fp[i] = e->fp_dead;
        return;
    }
    uint32_t k = 0;
    if (!e->bg_any) {
        const uint32_t nv = std::min<uint32_t>(h.end[3], FN3_HEAD_FP);
        for (; k < nv; ++k) hd.fp[k] = (uint16_t)(words[k] >> e->fshift);
    } else {
        // the row's first 32 positions that are NOT background (private variants rule an unrelated query out; positions half
        // the collection shares do not), topped up with background ones — the same selection as k_rehead
        const uint32_t nv = h.end[3];
        const uint32_t* bg = e->bg_bits.data();
        for (uint32_t i = 0; i < nv && k < FN3_HEAD_FP; ++i)
            if (!((bg[words[i] >> 5] >> (words[i] & 31u)) & 1u)) hd.fp[k++] = (uint16_t)(words[i] >> e->fshift);
        for (uint32_t i = 0; i < nv && k < FN3_HEAD_FP; ++i)
            if ((bg[words[i] >> 5] >> (words[i] & 31u)) & 1u) hd.fp[k++] = (uint16_t)(words[i] >> e->fshift);
    }
    for (; k < FN3_HEAD_FP; ++k) hd.fp[k] = e->fp_pad;
}

// device address of a row's header inside its tile
static unsigned char* hdr_addr(const fn3_engine* e, long long row) {
    long long tile; int lane;
    fn3_slot_of(row & ((1ll << e->hdr_shift) - 1), e->sb_shift, tile, lane);
    return e->hdr_chunks[(size_t)(row >> e->hdr_shift)] + tile * FN3_TILE_BYTES + lane * 32;
}

// reserve `words` (rounded up to 4) in the arena; returns device pointer
static int arena_alloc(fn3_engine* e, size_t words, uint32_t** out) {
    size_t need = (words + 3) & ~(size_t)3;
    if (need == 0) need = 4;
    if (e->data_chunks.empty() || e->cur_used + need > e->data_chunk_cap.back()) {
        size_t cap = std::max(e->chunk_words, need);
        uint32_t* p = nullptr;
        cudaError_t s = cudaMalloc(&p, cap * 4);
        if (s != cudaSuccess) return set_err(e, FN3_ENOMEM, "cudaMalloc of a %zu-byte store chunk failed: %s", cap * 4, cudaGetErrorString(s));
        e->data_chunks.push_back(p);
        e->data_chunk_cap.push_back(cap);
        e->cur_used = 0;
        e->store_bytes += (long long)cap * 4;
    }
    *out = e->data_chunks.back() + e->cur_used;
    e->cur_used += need;
    e->store_words += (long long)need;
    return FN3_OK;
}

static int ensure_hdr_chunk(fn3_engine* e, long long row) {
    const long long c = row >> e->hdr_shift;
    if (c >= FN3_MAX_HDR_CHUNKS) return set_err(e, FN3_ECAPACITY, "store is full (%lld rows)", row);
    while ((long long)e->hdr_chunks.size() <= c) {
        unsigned char* p = nullptr;
        size_t bytes = ((size_t)FN3_TILE_BYTES / 32) << e->hdr_shift;      // 96 bytes per row, in tiles of 32
        cudaError_t s = cudaMalloc(&p, bytes);
        if (s != cudaSuccess) return set_err(e, FN3_ENOMEM, "cudaMalloc of a header chunk failed: %s", cudaGetErrorString(s));
        // same stream as the header copies that follow (the legacy default stream does not order against it)
        s = cudaMemsetAsync(p, 0, bytes, e->copy_stream);
        if (s != cudaSuccess) return set_err(e, FN3_ECUDA, "cudaMemsetAsync of a header chunk failed: %s", cudaGetErrorString(s));
        e->hdr_chunks.push_back(p);
        e->store_bytes += (long long)bytes;
    }
    return FN3_OK;
}

static void fill_hdr_table(fn3_engine* e, HdrTable& t);

// run fn(lo, hi) over [0, n) on up to `max_threads` host threads (bulk append: validation and packing are per-row work)
template <class F>
static void parallel_rows(long long n, int max_threads, F fn, long long min_rows = 2048) {
    int nt = (int)std::min<long long>(max_threads, (n + min_rows - 1) / min_rows);
    if (nt <= 1) { fn(0ll, n); return; }
    std::vector<std::thread> th;
    const long long per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const long long lo = t * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=]() { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

static int append_impl(fn3_engine* e, long long n, const int32_t* pos, const long long* off6, const uint8_t* invalid,
                       long long* out_first_row) {
    std::lock_guard<std::mutex> g(e->store_mu);
    FN3_ON_DEVICE(e);
    const int host_threads = (int)std::max<long long>(1, std::min<long long>(env_ll("FN3_HOST_THREADS", 16),
                                                                             (long long)std::thread::hardware_concurrency()));
    // pass 1: validate + size (rows are independent: host threads)
    std::vector<long long> words((size_t)n);
    std::vector<uint32_t> nblks((size_t)n);
    std::atomic<int> bad{0};
    parallel_rows(n, host_threads, [&](long long lo, long long hi) {
        for (long long i = lo; i < hi && !bad.load(std::memory_order_relaxed); ++i) {
            long long o[7];
            for (int b = 0; b < 7; ++b) o[b] = off6[6 * i + b];
            long long w = row_words_needed(e, pos, o, invalid ? invalid[i] : 0, &nblks[(size_t)i]);
            if (w >= 0 && !(invalid && invalid[i]) && !raw_disjoint(pos, o)) {
                set_err(e, FN3_EARG, "the six lists of a sample must be mutually disjoint");
                w = -1;
            }
            if (w < 0) { bad.store(1); return; }
            words[(size_t)i] = w;
        }
    });
    if (bad.load()) return FN3_EARG;
    const long long first = e->n_rows;
    // pass 2: pack into pinned staging in batches (each batch is one contiguous arena allocation), copy
    const size_t batch_limit = std::min<size_t>((size_t)16 << 20, e->chunk_words);   // words
    std::vector<size_t> at;                                                          // staging offset of each row of a batch
    long long i = 0;
    while (i < n) {
        size_t batch_words = 0; long long j = i;
        at.clear();
        while (j < n) {
            const bool invj = invalid && invalid[j];
            const size_t need = row_alloc_words((size_t)words[(size_t)j], invj ? 0 : (size_t)(off6[6 * j + 4] - off6[6 * j]));
            if (j > i && batch_words + need > batch_limit) break;
            at.push_back(batch_words);
            batch_words += need; ++j;
        }
        int rc = ensure_stage(e, batch_words, (size_t)(j - i));
        if (rc) return rc;
        uint32_t* dbase = nullptr;
        rc = arena_alloc(e, batch_words, &dbase);
        if (rc) return rc;
        parallel_rows(j - i, host_threads, [&](long long lo, long long hi) {
            for (long long r = i + lo; r < i + hi; ++r) {
                long long o[7];
                for (int b = 0; b < 7; ++b) o[b] = off6[6 * r + b];
                RowHdr h;
                const bool inv = invalid && invalid[r];
                const size_t nv = inv ? 0 : (size_t)(o[4] - o[0]);
                const size_t need = row_alloc_words((size_t)words[(size_t)r], nv);
                const size_t w = at[(size_t)(r - i)];
                if (inv) {
                    h.data = 0; for (int b = 0; b < 6; ++b) h.end[b] = 0;
                    memset(e->stage + w, 0, need * 4);
                } else {
                    pack_row(e, pos, o, e->stage + w, h.end);
                    finish_row(e, e->stage + w, (size_t)words[(size_t)r], nv);
                    h.data = (unsigned long long)(uintptr_t)(dbase + w);
                }
                fill_head(e, e->stage_hdr[r - i], h, e->stage + w);
            }
        });
        CU(e, cudaMemcpyAsync(dbase, e->stage, batch_words * 4, cudaMemcpyHostToDevice, e->copy_stream));
        e->h2d_bytes += (long long)batch_words * 4 + (j - i) * (long long)sizeof(RowHead);
        // heads: one copy in row order, then k_scatter_heads places every head in its tile
        for (long long r = first + i; r < first + j; r += (1ll << e->hdr_shift)) {
            if ((rc = ensure_hdr_chunk(e, r))) return rc;
        }
        if ((rc = ensure_hdr_chunk(e, first + j - 1))) return rc;
        if ((size_t)(j - i) > e->d_heads_n) {
            size_t want = std::max<size_t>((size_t)(j - i), std::max<size_t>(e->d_heads_n * 2, 1024));
            CU(e, cudaStreamSynchronize(e->copy_stream));
            if (e->d_heads) cudaFree(e->d_heads);
            e->d_heads = nullptr; e->d_heads_n = 0;
            CU(e, cudaMalloc(&e->d_heads, want * sizeof(RowHead)));
            e->d_heads_n = want;
        }
        CU(e, cudaMemcpyAsync(e->d_heads, e->stage_hdr, (size_t)(j - i) * sizeof(RowHead), cudaMemcpyHostToDevice, e->copy_stream));
        {
            HdrTable ht; fill_hdr_table(e, ht);
            const long long cnt = j - i;
            k_scatter_heads<<<(unsigned)((cnt + 255) / 256), 256, 0, e->copy_stream>>>(e->d_heads, first + i, cnt, ht);
            e->launches += 1;
            CU(e, cudaGetLastError());
        }
        CU(e, cudaStreamSynchronize(e->copy_stream));
        // publish host mirror
        for (long long r2 = i; r2 < j; ++r2) {
            const bool inv = invalid && invalid[r2];
            e->h_hdr.push_back(e->stage_hdr[r2 - i].hdr);
            e->h_flags.push_back(inv ? 1 : 0);
            const uint32_t np = inv ? 0u : (uint32_t)(off6[6 * r2 + 6] - off6[6 * r2]);
            e->h_npos.push_back(np);
            e->h_nblk.push_back(nblks[(size_t)r2]);
            e->positions += np;
            e->n_live++; if (inv) e->n_invalid++;
        }
        e->n_rows = first + j;
        i = j;
    }
    if (out_first_row) *out_first_row = first;
    return FN3_OK;
}

extern "C" int fn3_append(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int64_t* out_row) {
    if (!e) return FN3_EARG;
    if (!invalid && (!off || (!pos && off[6] > off[0]))) return set_err(e, FN3_EARG, "fn3_append: null lists");
    long long o[7] = {0, 0, 0, 0, 0, 0, 0};
    if (!invalid) for (int b = 0; b < 7; ++b) o[b] = off[b];
    uint8_t inv = invalid ? 1 : 0;
    long long first = -1;
    int rc = append_impl(e, 1, pos, o, &inv, &first);
    if (rc == FN3_OK && out_row) *out_row = first;
    return rc;
}

extern "C" int fn3_append_bulk(fn3_engine* e, int64_t n, const int32_t* pos, const int64_t* off, const uint8_t* invalid,
                               int64_t* out_first_row) {
    if (!e || n < 0 || !off) return set_err(e, FN3_EARG, "fn3_append_bulk: bad arguments");
    if (n == 0) { if (out_first_row) { std::lock_guard<std::mutex> g(e->store_mu); *out_first_row = e->n_rows; } return FN3_OK; }
    long long first = -1;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    int rc = append_impl(e, n, pos, reinterpret_cast<const long long*>(off), invalid, &first);
    if (rc == FN3_OK && out_first_row) *out_first_row = first;
    return rc;
}

extern "C" int fn3_remove(fn3_engine* e, int64_t row) {
    if (!e) return FN3_EARG;
    std::lock_guard<std::mutex> g(e->store_mu);
    if (row < 0 || row >= e->n_rows) return set_err(e, FN3_ENOTFOUND, "fn3_remove: row %lld does not exist", (long long)row);
    if (e->h_flags[(size_t)row] & 2) return FN3_OK;          // removing twice is silent (seqComparer.py:131-134)
    FN3_ON_DEVICE(e);
    RowHdr dead; memset(&dead, 0, sizeof dead);
    RowHead dead_head;
    fill_head(e, dead_head, dead, nullptr);           // zero header + fingerprints that always mismatch
    // queries may be in flight on q_stream: they see the old head, or the dead fingerprints (filtered out), or old
    // fingerprints with the zero header (skipped when the header is read) — the reference's own semantics for
    // remove-during-mcompare.
    HdrTable ht; fill_hdr_table(e, ht);
    k_put_head<<<1, 1, 0, e->copy_stream>>>(dead_head, row, ht);
    e->launches += 1;
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(e->copy_stream));
    if (e->h_flags[(size_t)row] & 1) e->n_invalid--;
    e->h_flags[(size_t)row] |= 2;
    e->n_live--;
    e->positions -= e->h_npos[(size_t)row];
    return FN3_OK;
}

extern "C" int fn3_stats_get(fn3_engine* e, fn3_stats* out) {
    if (!e || !out) return FN3_EARG;
    std::lock_guard<std::mutex> g(e->store_mu);
    out->n_rows = e->n_rows; out->n_live = e->n_live; out->n_invalid = e->n_invalid;
    out->store_words = e->store_words; out->store_bytes = e->store_bytes; out->positions = e->positions;
    out->h2d_bytes = e->h2d_bytes.load(); out->d2h_bytes = e->d2h_bytes.load();
    out->t_pack_ns = e->t_pack_ns; out->t_enqueue_ns = e->t_enqueue_ns; out->t_wait_ns = e->t_wait_ns; out->n_timed = e->n_timed;
    out->head_refreshes = e->bg_refreshes; out->scan_mode = e->prefer_scan ? 1 : 0;
    return FN3_OK;
}

extern "C" int fn3_row_get(fn3_engine* e, int64_t row, int32_t* pos, int64_t pos_cap, int32_t off[7], int32_t* invalid,
                           int64_t* n_pos) {
    if (!e || !off || !invalid || !n_pos) return FN3_EARG;
    RowHdr h; uint8_t fl; uint32_t np;
    {
        std::lock_guard<std::mutex> g(e->store_mu);
        if (row < 0 || row >= e->n_rows) return set_err(e, FN3_ENOTFOUND, "fn3_row_get: row %lld does not exist", (long long)row);
        h = e->h_hdr[(size_t)row]; fl = e->h_flags[(size_t)row]; np = e->h_npos[(size_t)row];
    }
    if (fl & 2) return set_err(e, FN3_ENOTFOUND, "fn3_row_get: row %lld was removed", (long long)row);
    for (int b = 0; b < 7; ++b) off[b] = 0;
    *invalid = (fl & 1) ? 1 : 0;
    *n_pos = np;
    if (fl & 1) return FN3_OK;
    if ((int64_t)np > pos_cap) return set_err(e, FN3_ECAPACITY, "fn3_row_get: need room for %u positions", np);
    std::vector<uint32_t> w(h.end[5]);
    FN3_ON_DEVICE(e);
    if (h.end[5]) CU(e, cudaMemcpy(w.data(), (const void*)(uintptr_t)h.data, (size_t)h.end[5] * 4, cudaMemcpyDeviceToHost));
    int64_t k = 0;
    const uint32_t pmask = (e->pos_bits >= 32) ? 0xffffffffu : ((1u << e->pos_bits) - 1u);
    uint32_t i = 0;
    for (int b = 0; b < 6; ++b) {
        for (; i < h.end[b]; ++i) {
            if (b < 4) pos[k++] = (int32_t)w[i];
            else {
                uint32_t s = w[i] & pmask, len = (e->pos_bits >= 32) ? 1u : ((w[i] >> e->pos_bits) + 1u);
                for (uint32_t j = 0; j < len; ++j) pos[k++] = (int32_t)(s + j);
            }
        }
        off[b + 1] = (int32_t)k;
    }
    return FN3_OK;
}

// ---- query machinery ----------------------------------------------------------------------

static void fill_hdr_table(fn3_engine* e, HdrTable& t) {   // store_mu held
    memset(&t, 0, sizeof t);
    for (size_t i = 0; i < e->hdr_chunks.size(); ++i) t.chunk[i] = e->hdr_chunks[i];
    t.shift = e->hdr_shift;
    t.sb_shift = e->sb_shift;
}

// global-memory query structures for `ns` slots (fallback path), allocated on first use
static int ensure_qstate(fn3_engine* e, int ns) {
    if (ns <= e->qstate_slots) return FN3_OK;
    QueryState& q = e->qstate;
    CU(e, cudaStreamSynchronize(e->q_stream));
    cudaFree(q.bm); cudaFree(q.coarse); cudaFree(q.ent); cudaFree(q.info);
    q.bm = nullptr; q.coarse = nullptr; q.ent = nullptr; q.info = nullptr; e->qstate_slots = 0;
    CU(e, cudaMalloc(&q.bm, (size_t)e->nw * 2 * sizeof(uint32_t) * ns));
    CU(e, cudaMalloc(&q.coarse, (size_t)e->ncw * sizeof(uint32_t) * ns));
    CU(e, cudaMalloc(&q.ent, (size_t)q.ent_stride * sizeof(QEntry) * ns));
    CU(e, cudaMalloc(&q.info, sizeof(uint4) * ns));
    e->qstate_slots = ns;
    return FN3_OK;
}

// the FILTER variant needs the 32 fingerprints of a head to be able to exceed the ceiling with some slack
static bool use_filter_for(int ceiling) { return ceiling <= FN3_HEAD_FP - 8; }

template <bool COUNT, int QMODE>
static void launch_query(bool filter, dim3 g, size_t smem, cudaStream_t st, const CompareParams& p) {
    if (filter) k_query<COUNT, QMODE><<<g, FN3_Q_THREADS, smem, st>>>(p);
    else k_scan<COUNT, QMODE><<<g, FN3_SCAN_THREADS, smem, st>>>(p);
}

struct QueryJob {                    // what enqueue_query needs beyond the engine state
    HdrTable ht;
    const long long* d_rows = nullptr;   // explicit candidate list (device) or nullptr = all rows < n_rows
    long long n_list = 0;
    long long n_rows = 0;
    int ns = 1;                          // query slots: 1 = passed by value; >1 = uploaded to e->d_slots here
    int slot_base = 0;                   // first slot of this tile inside the h_slots/d_slots ring
    uint32_t blocks = 0;                 // max touched-block bound over the tile's queries (sizes shared memory)
    int ceiling = 0;
    bool count_bytes = false;
    const RowHead* commit_head = nullptr;   // fn3_insert_query: head of the new row, placed by the kernel
    long long commit_row = -1;
    cudaEvent_t split = nullptr;         // measurement hook: recorded between the optional build kernel and k_query
    bool zero_copy = false;              // results + completion through mapped host memory (one-vs-all paths)
    bool adaptive = false;               // a full one-vs-all scan: may be steered from FILTER to STREAM (prefer_scan)
    bool used_filter = false;            // out: which variant enqueue_query chose
    bool split_recorded = false;         // out: a build kernel ran and `split` was recorded behind it
};

// enqueue one query tile on q_stream.  Normally ONE kernel (k_query builds the query structure in shared memory
// itself); queries too large for shared memory take k_build_query + the global-memory variant.
// k_scan: into how many sub-units (1 << ss, ss = 0..3) a 32-row head tile is split.  A warp takes whole sub-units, so the
// launch ends when the warp with the most rows ends: per split, (sub-units a warp must take) x (rows of a sub-unit + the cost
// of drawing one and fetching its headers, about half a row) is estimated and the cheapest split taken.  Sub-units that hold
// no row — the unused lanes of the last, partly filled superblock — are not counted.  Measured on the 50 000-row bench store
// (one-vs-all at ceiling 250 / none, scripts/variant_prof.py with FN3_SCAN_SS): ss 0: 45.0 / 84.8 us, 1: 34.9 / 51.3,
// 2: 35.8 / 52.1, 3: 36.8 / 53.3, 4: 36.9 / 53.8, 5: 39.8 / 57.3 — this rule picks 1 there, 3 for the 125 000 rows per GPU of
// configs[4] (measured: 0.57 of the HBM peak at ceiling 250) and whole tiles again once every warp gets several of them.
static int pick_scan_split(long long n_rows, int sb_shift, bool list, long long n_list, long long warps) {
    int best = 0;
    double best_cost = 1e300;
    for (int ss = 0; ss <= 3; ++ss) {
        const long long r = 32 >> ss;
        long long subs, rows_per;
        if (list) { subs = ((n_list + 31) / 32) << ss; rows_per = r; }
        else {
            const long long tiles = 1ll << (sb_shift - 5), full = n_rows >> sb_shift, rem = n_rows & ((1ll << sb_shift) - 1);
            const long long lanes = (rem + tiles - 1) / tiles;          // lanes of the last superblock's tiles that hold a row
            subs = full * (tiles << ss) + tiles * std::min<long long>(1ll << ss, (lanes + r - 1) / r);
            rows_per = full ? r : std::min<long long>(r, lanes);
        }
        const long long per_warp = (subs + warps - 1) / std::max<long long>(warps, 1);
        const double cost = (double)per_warp * ((double)rows_per + 0.5);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = ss; }
    }
    return best;
}

// host-only view of the rule above for tests/test_cabi_symbols.py (no device is touched)
extern "C" int fn3_debug_scan_split(int64_t n_rows, int32_t sb_shift, int64_t n_list, int64_t warps) {
    return pick_scan_split(n_rows, sb_shift, n_list >= 0, n_list < 0 ? 0 : n_list, warps);
}

static int enqueue_query(fn3_engine* e, QueryJob& job) {
    cudaStream_t st = e->q_stream;
    const int ns = job.ns;
    QuerySlot* h_sl = e->h_slots + job.slot_base;
    QuerySlot* d_sl = e->d_slots + job.slot_base;
    if (ns > 1) {
        CU(e, cudaMemcpyAsync(d_sl, h_sl, sizeof(QuerySlot) * ns, cudaMemcpyHostToDevice, st));
        e->h2d_bytes += (long long)sizeof(QuerySlot) * ns;
    }
    uint32_t blocks = job.blocks;
    if (ns == 1 && blocks == 0) blocks = e->slot_blocks[0];
    blocks = std::min<uint32_t>(blocks, (uint32_t)e->n_blk);
    const bool filter = use_filter_for(job.ceiling) && !(job.adaptive && e->prefer_scan);
    job.used_filter = filter;
    // dynamic shared memory: the query structure (when it fits) and, for k_scan, the row rings behind it
    const size_t q_smem = ((size_t)e->nw * sizeof(uint2) + (size_t)e->ncw * 4 + ((size_t)blocks + 1) * sizeof(QEntry) + 127) & ~(size_t)127;
    const size_t ring = filter ? 0 : FN3_SCAN_RING_BYTES;
    const bool in_smem = q_smem + ring + (size_t)e->q_static_smem + 256 <= (size_t)e->stage_limit;
    const size_t smem = (in_smem ? q_smem : 0) + ring;
    if (!in_smem) {
        int rcq = ensure_qstate(e, ns);
        if (rcq) return rcq;
        if (ns == 1) {                    // the build kernel reads its slots from device memory
            CU(e, cudaMemcpyAsync(d_sl, h_sl, sizeof(QuerySlot), cudaMemcpyHostToDevice, st));
            e->h2d_bytes += (long long)sizeof(QuerySlot);
        }
        k_build_query<<<ns, FN3_Q_THREADS, 0, st>>>(d_sl, e->qstate, e->nw, e->ncw, e->fshift, e->pos_bits);
        e->launches += 1;
        CU(e, cudaGetLastError());
        // (the split event only where there IS a build kernel: an event record between the step's first event and the compare
        //  kernel costs 2.7 us of stream time by itself — it used to be counted into every step)
        if (job.split) { CU(e, cudaEventRecord(job.split, st)); job.split_recorded = true; }
    }
    CompareParams p;
    memset(&p, 0, sizeof p);
    p.hdr = job.ht; p.qs = e->qstate; p.rows = job.d_rows; p.n_list = job.n_list; p.n_rows = job.n_rows;
    p.slots = d_sl; p.slot0 = h_sl[0]; p.slot_inline = ns == 1 ? 1 : 0;
    p.commit_row = job.commit_row;
    if (job.commit_head) p.commit_head = *job.commit_head;
    p.out = e->d_out; p.out_count = e->d_count; p.touched = e->d_count + 1; p.out_cap = e->out_cap;
    p.ceiling = job.ceiling; p.pos_bits = e->pos_bits; p.use_filter = filter ? 1 : 0;
    p.fshift = e->fshift; p.nw = e->nw; p.ncw = e->ncw; p.debug_stop = e->debug_stop;
    if (job.zero_copy) {
        p.host_out = reinterpret_cast<EdgeRec*>(e->zc_buf + 64);
        p.host_flag = reinterpret_cast<volatile unsigned long long*>(e->zc_buf);
        p.done_ctas = e->d_done; p.host_cap = FN3_ZC_CAP; p.seq = e->zc_seq;
    }
    const long long warps_per_cta = (filter ? FN3_Q_THREADS : FN3_SCAN_THREADS) / 32;
    const long long sb = 1ll << e->sb_shift;
    const long long slots_scanned = ((job.n_rows + sb - 1) / sb) * sb;          // whole superblocks of head tiles
    const long long cand = job.d_rows ? job.n_list : slots_scanned;
    const long long units = (cand + 31) / 32;                                   // warp-sized work units (32 candidates)
    long long per_sm = filter ? FN3_Q_MINB : 1;                     // __launch_bounds__(512, FN3_Q_MINB) / k_scan: (512, 1)
    if (in_smem && per_sm * (long long)(smem + (size_t)e->q_static_smem + 1024) > (long long)228 * 1024) per_sm = 1;
    // every CTA builds its query first: with several slots per launch give each slot its share of the resident
    // CTAs and let every CTA scan more candidates, rather than several waves that each rebuild the query
    const long long cta_cap = std::max<long long>(1, (long long)e->sm_count * per_sm / ns);
    long long ctas;
    if (filter) {
        ctas = std::min<long long>(std::max<long long>(1, (units + warps_per_cta - 1) / warps_per_cta), cta_cap);
    } else {
        // k_scan: a warp streams its rows one after the other; the 32-row tiles are split into sub-units so that the rows
        // spread over all warps of the launch (pick_scan_split)
        int ss = pick_scan_split(job.n_rows, e->sb_shift, job.d_rows != nullptr, job.n_list, cta_cap * warps_per_cta);
        if (e->scan_ss >= 0 && e->scan_ss <= 5) ss = e->scan_ss;
        p.split_shift = ss;
        p.ring_off = (unsigned int)(in_smem ? q_smem : 0);
        ctas = std::min<long long>(std::max<long long>(1, ((units << ss) + warps_per_cta - 1) / warps_per_cta), cta_cap);
    }
    size_t smem_launch = smem;
    if (e->debug_stop == 3) {                        // launch-floor experiments (the kernel returns at entry): grid / smem overrides
        ctas = env_ll("FN3_DBG_CTAS", ctas);
        smem_launch = (size_t)env_ll("FN3_DBG_SMEM", (long long)smem);
    }
    dim3 g((unsigned)ctas, (unsigned)ns);
    if (in_smem) {
        if (job.count_bytes) launch_query<true, 1>(filter, g, smem_launch, st, p);
        else launch_query<false, 1>(filter, g, smem_launch, st, p);
    } else {
        if (job.count_bytes) launch_query<true, 0>(filter, g, smem_launch, st, p);
        else launch_query<false, 0>(filter, g, smem_launch, st, p);
    }
    e->launches += 1;
    CU(e, cudaGetLastError());
    return FN3_OK;
}

// upload a non-stored sample into scratch slot `which` (0/1); returns its header (and optionally its head)
static int upload_scratch(fn3_engine* e, int which, const int32_t* pos, const int32_t off[7], int invalid, RowHdr* out,
                          uint32_t* nblk = nullptr, RowHead* head_out = nullptr) {
    RowHdr h; memset(&h, 0, sizeof h);
    if (nblk) *nblk = 0;
    if (invalid) { *out = h; if (head_out) fill_head(e, *head_out, h, nullptr); return FN3_OK; }
    long long o[7];
    for (int b = 0; b < 7; ++b) o[b] = off[b];
    const size_t npos_in = (size_t)(o[6] - o[0]), nv = (size_t)(o[4] - o[0]);
    const size_t worst = row_alloc_words(npos_in, nv) + 8;          // one word per position is the worst case
    // the query paths own a pinned staging buffer of two halves (guarded by query_mu, which every caller holds);
    // a copy still in flight from a call that returned without waiting is drained first
    if (e->qstage_busy) { CU(e, cudaStreamSynchronize(e->q_stream)); e->qstage_busy = false; }
    if (worst * 2 > e->qstage_words) {
        size_t want = std::max(worst * 2, std::max<size_t>(e->qstage_words * 2, 1u << 16));
        if (e->qstage) cudaFreeHost(e->qstage);
        e->qstage = nullptr; e->qstage_words = 0;
        CU(e, cudaHostAlloc(&e->qstage, want * 4, cudaHostAllocDefault));
        e->qstage_words = want;
    }
    if (worst * 2 > e->scratch_words) {
        size_t want = std::max(worst * 2, std::max<size_t>(e->scratch_words * 2, 1u << 16));
        CU(e, cudaStreamSynchronize(e->q_stream));
        if (e->d_scratch) cudaFree(e->d_scratch);
        e->d_scratch = nullptr; e->scratch_words = 0;
        CU(e, cudaMalloc(&e->d_scratch, want * 4));
        e->scratch_words = want;
    }
    uint32_t* src = e->qstage + (which ? e->qstage_words / 2 : 0);
    uint32_t nb = 0;
    const long long w = pack_checked(e, pos, o, src, h.end, &nb);   // single pass: validate + pack
    if (w < 0) return FN3_EARG;
    if (nblk) *nblk = nb;
    const size_t need = row_alloc_words((size_t)w, nv);
    finish_row(e, src, (size_t)w, nv);
    uint32_t* dst = e->d_scratch + (which ? e->scratch_words / 2 : 0);
    h.data = (unsigned long long)(uintptr_t)dst;
    if (head_out) fill_head(e, *head_out, h, src);
    CU(e, cudaMemcpyAsync(dst, src, need * 4, cudaMemcpyHostToDevice, e->q_stream));   // consumed by the query that follows
    e->qstage_busy = true;
    e->h2d_bytes += (long long)need * 4;
    *out = h;
    return FN3_OK;
}

static int ensure_votes(fn3_engine* e);
static int ensure_rows(fn3_engine* e, long long n);

// Recompute the background bitmap from the live rows and rewrite every head from it (query_mu held by the caller; store_mu is
// taken here for the whole pass: heads are rewritten in place).  ppm: a position is background when it is non-reference in
// at least ppm / 1e6 of the live valid rows (and in at least 8 of them); ppm <= 0 clears the background.
static int refresh_heads_locked(fn3_engine* e, long long ppm) {
    std::lock_guard<std::mutex> g(e->store_mu);
    FN3_ON_DEVICE(e);
    cudaStream_t st = e->q_stream;
    const long long n = e->n_rows;
    const long long n_words = (e->L + 31) / 32;
    int rc;
    if (!e->d_bg) {
        CU(e, cudaMalloc(&e->d_bg, (size_t)n_words * 4));
        CU(e, cudaMemsetAsync(e->d_bg, 0, (size_t)n_words * 4, st));
    }
    CU(e, cudaStreamSynchronize(e->copy_stream));          // heads of earlier appends are in place
    long long n_valid = 0;
    if ((rc = ensure_rows(e, std::max<long long>(n, 1)))) return rc;
    for (long long r = 0; r < n; ++r)
        if (e->h_flags[(size_t)r] == 0) e->h_rows[n_valid++] = r;
    HdrTable ht; fill_hdr_table(e, ht);
    if (ppm > 0 && n_valid > 0) {
        if ((rc = ensure_votes(e))) return rc;
        CU(e, cudaMemcpyAsync(e->d_rows, e->h_rows, (size_t)n_valid * sizeof(long long), cudaMemcpyHostToDevice, st));
        CU(e, cudaMemsetAsync(e->d_votes, 0, (size_t)e->vLpad * 4 * sizeof(uint32_t), st));
        k_votes_rows<<<(unsigned)std::min<long long>(n_valid, (long long)e->sm_count * 8), FN3_VOTE_THREADS, 0, st>>>(e->d_rows, n_valid, ht, e->d_votes, e->vLpad, 0, e->pos_bits);
        const long long thr = std::max<long long>(8, (ppm * n_valid + 999999) / 1000000);
        k_bg_bits<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(e->d_votes, e->vLpad, e->L, (uint32_t)std::min<long long>(thr, 0xffffffffll), e->d_bg, n_words);
        e->launches += 2;
        CU(e, cudaGetLastError());
    } else {
        CU(e, cudaMemsetAsync(e->d_bg, 0, (size_t)n_words * 4, st));
    }
    e->bg_bits.assign((size_t)n_words + 1, 0u);
    CU(e, cudaMemcpyAsync(e->bg_bits.data(), e->d_bg, (size_t)n_words * 4, cudaMemcpyDeviceToHost, st));
    if (n > 0) {
        k_rehead<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(ht, n, e->d_bg, e->fshift, (uint32_t)e->fp_pad);
        e->launches += 1;
        CU(e, cudaGetLastError());
    }
    CU(e, cudaStreamSynchronize(st));
    e->d2h_bytes += n_words * 4;
    bool any = false;
    for (long long w = 0; w < n_words && !any; ++w) any = e->bg_bits[(size_t)w] != 0u;
    e->bg_any = any;
    e->bg_refreshes++;
    e->bg_next_at = std::max<long long>(4096, 2 * e->n_live);
    return FN3_OK;
}

// remove(guid) leaves a tombstone: the row's words stay in HBM and its slot in the head tiles is scanned (and skipped) by every
// query.  fn3_compact rebuilds the store ON THE DEVICE: the live rows are copied into fresh chunks in row order (one warp per
// row), renumbered densely, their heads rebuilt from the row data (k_place_hdrs + k_rehead) — no row travels through the
// host.  new_row[r] receives the new number of old row r, or -1 if it was removed.
extern "C" int fn3_compact(fn3_engine* e, int64_t* new_row, int64_t cap, int64_t* n_before, int64_t* n_after) {
    if (!e || !n_before || !n_after) return set_err(e, FN3_EARG, "fn3_compact: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    std::lock_guard<std::mutex> g(e->store_mu);
    FN3_ON_DEVICE(e);
    const long long n = e->n_rows;
    *n_before = n; *n_after = e->n_live;
    if (n > cap || (n > 0 && !new_row)) return set_err(e, FN3_ECAPACITY, "fn3_compact: the row map needs %lld entries", n);
    if (e->n_live == n) { for (long long r = 0; r < n; ++r) new_row[r] = r; return FN3_OK; }
    cudaStream_t st = e->q_stream;
    CU(e, cudaStreamSynchronize(e->copy_stream));
    CU(e, cudaStreamSynchronize(st));
    // 1. lay the live rows out in fresh chunks
    std::vector<uint32_t*> chunks; std::vector<size_t> caps; size_t used = 0;
    std::vector<RowCopy> copies; std::vector<RowHdr> hdrs;
    std::vector<uint8_t> flags; std::vector<uint32_t> npos, nblk;
    long long words_total = 0, k = 0;
    auto fail = [&](int rc) { for (auto p : chunks) cudaFree(p); return rc; };
    for (long long r = 0; r < n; ++r) {
        if (e->h_flags[(size_t)r] & 2) { new_row[r] = -1; continue; }
        new_row[r] = k++;
        RowHdr h = e->h_hdr[(size_t)r];
        if (h.data) {
            const size_t need = row_alloc_words(h.end[5], h.end[3]);
            if (chunks.empty() || used + need > caps.back()) {
                const size_t c = std::max(e->chunk_words, need);
                uint32_t* p = nullptr;
                if (cudaMalloc(&p, c * 4) != cudaSuccess) { (void)cudaGetLastError(); return fail(set_err(e, FN3_ENOMEM, "fn3_compact: no room for a second copy of the store")); }
                chunks.push_back(p); caps.push_back(c); used = 0;
            }
            RowCopy c; c.src = h.data; c.dst = (unsigned long long)(uintptr_t)(chunks.back() + used); c.words = (uint32_t)need; c.pad = 0;
            copies.push_back(c);
            h.data = c.dst;
            used += need; words_total += (long long)need;
        }
        hdrs.push_back(h); flags.push_back(e->h_flags[(size_t)r]); npos.push_back(e->h_npos[(size_t)r]); nblk.push_back(e->h_nblk[(size_t)r]);
    }
    const long long m = k;
    // 2. copy the rows, place the headers in fresh tiles, rebuild the fingerprints
    RowCopy* d_copies = nullptr; RowHdr* d_hdrs = nullptr;
    std::vector<unsigned char*> tiles;
    auto fail2 = [&](int rc) { cudaFree(d_copies); cudaFree(d_hdrs); for (auto p : tiles) cudaFree(p); return fail(rc); };
    if (!copies.empty()) {
        if (cudaMalloc(&d_copies, copies.size() * sizeof(RowCopy)) != cudaSuccess) return fail2(set_err(e, FN3_ENOMEM, "fn3_compact: cudaMalloc failed"));
        cudaMemcpy(d_copies, copies.data(), copies.size() * sizeof(RowCopy), cudaMemcpyHostToDevice);
        k_copy_rows<<<(unsigned)std::min<long long>(((long long)copies.size() + 7) / 8, (long long)e->sm_count * 16), 256, 0, st>>>(d_copies, (long long)copies.size());
        e->launches += 1;
    }
    HdrTable ht; memset(&ht, 0, sizeof ht);
    ht.shift = e->hdr_shift; ht.sb_shift = e->sb_shift;
    const size_t tile_bytes = ((size_t)FN3_TILE_BYTES / 32) << e->hdr_shift;
    for (long long c = 0; c <= (std::max<long long>(m, 1) - 1) >> e->hdr_shift; ++c) {
        unsigned char* p = nullptr;
        if (cudaMalloc(&p, tile_bytes) != cudaSuccess) { (void)cudaGetLastError(); return fail2(set_err(e, FN3_ENOMEM, "fn3_compact: cudaMalloc of a header chunk failed")); }
        cudaMemsetAsync(p, 0, tile_bytes, st);
        tiles.push_back(p); ht.chunk[c] = p;
    }
    if (m > 0) {
        if (cudaMalloc(&d_hdrs, (size_t)m * sizeof(RowHdr)) != cudaSuccess) return fail2(set_err(e, FN3_ENOMEM, "fn3_compact: cudaMalloc failed"));
        cudaMemcpyAsync(d_hdrs, hdrs.data(), (size_t)m * sizeof(RowHdr), cudaMemcpyHostToDevice, st);
        k_place_hdrs<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_hdrs, m, ht, (uint32_t)e->fp_dead, (uint32_t)e->fp_pad);
        if (!e->d_bg) {
            const long long n_words = (e->L + 31) / 32;
            if (cudaMalloc(&e->d_bg, (size_t)n_words * 4) != cudaSuccess) return fail2(set_err(e, FN3_ENOMEM, "fn3_compact: cudaMalloc failed"));
            cudaMemsetAsync(e->d_bg, 0, (size_t)n_words * 4, st);
        }
        k_rehead<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(ht, m, e->d_bg, e->fshift, (uint32_t)e->fp_pad);
        e->launches += 2;
    }
    const cudaError_t err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) return fail2(set_err(e, FN3_ECUDA, "fn3_compact: %s", cudaGetErrorString(err)));
    cudaFree(d_copies); cudaFree(d_hdrs);
    // 3. switch over
    for (auto p : e->data_chunks) cudaFree(p);
    for (auto p : e->hdr_chunks) cudaFree(p);
    e->data_chunks = chunks; e->data_chunk_cap = caps; e->cur_used = used;
    e->hdr_chunks = tiles;
    e->h_hdr = hdrs; e->h_flags = flags; e->h_npos = npos; e->h_nblk = nblk;
    e->n_rows = m;
    e->store_words = words_total;
    e->store_bytes = 0;
    for (size_t c : caps) e->store_bytes += (long long)c * 4;
    e->store_bytes += (long long)tiles.size() * (long long)tile_bytes;
    e->edges_ready = -1;
    return FN3_OK;
}

// automatic refresh: when the store has doubled since the last one (query_mu held)
static int maybe_refresh_heads(fn3_engine* e) {
    if (!e->bg_auto) return FN3_OK;
    bool due;
    { std::lock_guard<std::mutex> g(e->store_mu); due = e->n_live >= e->bg_next_at; }
    return due ? refresh_heads_locked(e, e->bg_ppm) : FN3_OK;
}

extern "C" int fn3_refresh_heads(fn3_engine* e, int64_t min_ppm) {
    if (!e) return FN3_EARG;
    std::lock_guard<std::mutex> q(e->query_mu);
    return refresh_heads_locked(e, min_ppm);
}

// shared tail of one-vs-all style calls: the query is in h_slots[0] / slot_blocks[0]; job.ht / job.n_rows are set
static int run_one_vs_all(fn3_engine* e, QueryJob& job, const int64_t* rows, int64_t n_list, fn3_hit* out, int64_t cap,
                          int64_t* n_out) {
    const long long n = rows ? n_list : job.n_rows;
    *n_out = 0;
    e->edges_ready = -1;
    const bool nothing = e->h_slots[0].hdr.data == 0ull || n <= 0;   // invalid query: nothing is a neighbour
    if (nothing && job.commit_row < 0) return FN3_OK;
    int rc;
    if (rows && !nothing) {
        if ((rc = ensure_rows(e, n))) return rc;
        for (long long i = 0; i < n; ++i) {
            if (rows[i] < 0 || rows[i] >= job.n_rows)
                return set_err(e, FN3_ENOTFOUND, "candidate row %lld does not exist", (long long)rows[i]);
            e->h_rows[i] = rows[i];
        }
        CU(e, cudaMemcpyAsync(e->d_rows, e->h_rows, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, e->q_stream));
        e->h2d_bytes += n * (long long)sizeof(long long);
        job.d_rows = e->d_rows; job.n_list = n;
    }
    job.adaptive = !rows && n >= 1024;
    if ((rc = ensure_out(e, std::max<long long>(n, 1)))) return rc;
    const EdgeRec* recs;
    long long cnt;
    if (e->zero_copy) {
        // Results and completion travel through mapped pinned memory: no memset, no device-to-host copy, no
        // stream synchronisation — the host spins on the sequence number the last CTA writes.
        if (e->count_dirty) { CU(e, cudaMemsetAsync(e->d_count, 0, FN3_OUT_HDR, e->q_stream)); e->count_dirty = false; }
        job.zero_copy = true;
        ++e->zc_seq;
        const auto t_enq = std::chrono::steady_clock::now();
        if ((rc = enqueue_query(e, job))) return rc;
        const auto t_wait = std::chrono::steady_clock::now();
        e->t_enqueue_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(t_wait - t_enq).count();
        volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(e->zc_buf);
        for (unsigned long long spins = 1; flag[0] != e->zc_seq; ++spins) {
            if ((spins & 0x3fff) == 0) {
                const cudaError_t st = cudaStreamQuery(e->q_stream);
                if (st != cudaSuccess && st != cudaErrorNotReady)
                    return set_err(e, FN3_ECUDA, "k_query failed: %s", cudaGetErrorString(st));
                if (st == cudaSuccess && flag[0] != e->zc_seq)
                    return set_err(e, FN3_ECUDA, "k_query finished without publishing its result");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);      // records are read only after the flag
        e->qstage_busy = false;                                   // the stream has consumed everything queued before
        e->t_wait_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_wait).count();
        cnt = (long long)flag[1];
        e->last_passed = (long long)flag[2];
        if (job.adaptive && use_filter_for(job.ceiling)) {
            // FILTER <-> STREAM steering: when the head filter lets most candidates through (an outbreak: they ARE neighbours)
            // the next full scans go straight to k_scan; while there, the hit rate tells when to try the filter again
            const double cand = (double)std::max<long long>(n, 1);
            if (job.used_filter) {
                e->pass_ema = 0.5 * e->pass_ema + 0.5 * ((double)e->last_passed / cand);
                if (e->pass_ema > 0.25) { e->prefer_scan = true; e->scan_streak = 0; }
            } else if (e->prefer_scan) {
                if ((double)cnt / cand < 0.05 && ++e->scan_streak >= 8) { e->prefer_scan = false; e->pass_ema = 0.0; }
                else if ((double)cnt / cand >= 0.05) e->scan_streak = 0;
            }
        }
        e->d2h_bytes += 24 + std::min<long long>(cnt, FN3_ZC_CAP) * (long long)sizeof(EdgeRec);
        recs = reinterpret_cast<const EdgeRec*>(e->zc_buf + 64);
        *n_out = cnt;
        if (cnt > cap) return set_err(e, FN3_ECAPACITY, "output buffer holds %lld hits, %lld needed", (long long)cap, cnt);
        if (cnt > FN3_ZC_CAP) {                       // a very long neighbour list: fetch it from the device buffer
            if ((rc = ensure_hout(e, cnt))) return rc;
            CU(e, cudaMemcpyAsync(e->h_out, e->d_out, (size_t)cnt * sizeof(EdgeRec), cudaMemcpyDeviceToHost, e->q_stream));
            CU(e, cudaStreamSynchronize(e->q_stream));
            e->d2h_bytes += cnt * (long long)sizeof(EdgeRec);
            recs = e->h_out;
        }
    } else {
        CU(e, cudaMemsetAsync(e->d_count, 0, FN3_OUT_HDR, e->q_stream));
        if ((rc = enqueue_query(e, job))) return rc;
        // ONE device-to-host copy brings the counter and the first records; a second one only for long lists
        const long long first = std::min<long long>(std::max<long long>(n, 0), FN3_FIRST_HITS);
        CU(e, cudaMemcpyAsync(e->h_outbuf, e->d_outbuf, FN3_OUT_HDR + (size_t)first * sizeof(EdgeRec), cudaMemcpyDeviceToHost,
                              e->q_stream));
        CU(e, cudaStreamSynchronize(e->q_stream));
        e->d2h_bytes += FN3_OUT_HDR + first * (long long)sizeof(EdgeRec);
        cnt = (long long)e->h_count[0];
        e->count_dirty = true;
        e->qstage_busy = false;
        *n_out = cnt;
        if (cnt > cap) return set_err(e, FN3_ECAPACITY, "output buffer holds %lld hits, %lld needed", (long long)cap, cnt);
        if (cnt > first) {
            if ((rc = ensure_hout(e, cnt))) return rc;
            CU(e, cudaMemcpyAsync(e->h_out, e->d_out, (size_t)cnt * sizeof(EdgeRec), cudaMemcpyDeviceToHost, e->q_stream));
            CU(e, cudaStreamSynchronize(e->q_stream));
            e->d2h_bytes += cnt * (long long)sizeof(EdgeRec);
        }
        recs = e->h_out;
    }
    for (long long i = 0; i < cnt; ++i) {
        const EdgeRec& r = recs[i];
        out[i].row = r.row2; out[i].dist = r.dist; out[i].n1 = r.n1; out[i].n2 = r.n2; out[i].nboth = r.nboth;
    }
    return FN3_OK;
}

static void snapshot_store(fn3_engine* e, QueryJob& job) {
    std::lock_guard<std::mutex> g(e->store_mu);
    fill_hdr_table(e, job.ht);
    job.n_rows = e->n_rows;
}

static int stored_hdr(fn3_engine* e, long long row, RowHdr* h, const char* who, uint32_t* nblk = nullptr) {
    std::lock_guard<std::mutex> g(e->store_mu);
    if (row < 0 || row >= e->n_rows) return set_err(e, FN3_ENOTFOUND, "%s: row %lld does not exist", who, row);
    if (e->h_flags[(size_t)row] & 2) return set_err(e, FN3_ENOTFOUND, "%s: row %lld was removed", who, row);
    *h = e->h_hdr[(size_t)row];
    if (nblk) *nblk = e->h_nblk[(size_t)row];
    return FN3_OK;
}

static void set_slot0(fn3_engine* e, const RowHdr& h, long long row, long long exclude) {
    QuerySlot& s = e->h_slots[0];
    s.hdr = h; s.row = row; s.skip_le = -1; s.exclude = exclude; s.pad = 0;
}

extern "C" int fn3_one_vs_all(fn3_engine* e, int64_t query_row, int32_t ceiling, const int64_t* rows, int64_t n_rows,
                              fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || (rows && n_rows < 0)) return set_err(e, FN3_EARG, "fn3_one_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    RowHdr qh; int rc;
    if ((rc = maybe_refresh_heads(e))) return rc;
    if ((rc = stored_hdr(e, query_row, &qh, "fn3_one_vs_all", &e->slot_blocks[0]))) return rc;
    QueryJob job; job.ceiling = ceiling;
    snapshot_store(e, job);
    set_slot0(e, qh, query_row, query_row);
    return run_one_vs_all(e, job, rows, n_rows, out, cap, n_out);
}

// Body of fn3_insert_query (query_mu held, device set).  The sample is packed into the insert path's OWN pinned staging
// buffer (istage, guarded by query_mu), so store_mu is held only while the row is reserved and, after the kernel has
// answered, while it is published: fn3_append / fn3_remove / fn3_stats_get do not wait behind the query.  The row is
// reserved as "pending" (its host flags say removed) and becomes live only once k_query has placed its head — a call
// that fails before the launch leaves a tombstone, never a live row without a head.
static int insert_query_locked(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                               int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out) {
    RowHdr h; memset(&h, 0, sizeof h);
    const auto t_begin = std::chrono::steady_clock::now();
    uint32_t nblk = 0;
    long long o[7] = {0, 0, 0, 0, 0, 0, 0};
    if (!invalid) for (int b = 0; b < 7; ++b) o[b] = off[b];
    // one pass: validate + pack into pinned staging (one word per position is the worst case)
    const size_t npos_in = invalid ? 0 : (size_t)(o[6] - o[0]);
    const size_t nv_in = invalid ? 0 : (size_t)(o[4] - o[0]);
    const size_t worst = row_alloc_words(npos_in, nv_in) + 8;
    int rc;
    if (worst > e->istage_words) {
        const size_t want = std::max(worst, std::max<size_t>(e->istage_words * 2, 1u << 16));
        CU(e, cudaStreamSynchronize(e->q_stream));       // a copy out of the old buffer may still be in flight
        if (e->istage) cudaFreeHost(e->istage);
        e->istage = nullptr; e->istage_words = 0;
        CU(e, cudaHostAlloc(&e->istage, want * 4, cudaHostAllocDefault));
        e->istage_words = want;
    }
    uint32_t* const stage = e->istage;
    long long w = 0;
    if (!invalid) {
        w = pack_checked(e, pos, o, stage, h.end, &nblk);
        if (w < 0) return FN3_EARG;
    }
    const size_t need = row_alloc_words((size_t)w, nv_in);
    if (invalid) memset(stage, 0, need * 4);
    else finish_row(e, stage, (size_t)w, nv_in);
    uint32_t* dbase = nullptr;
    long long row;
    QueryJob job; job.ceiling = ceiling;
    const uint32_t np = invalid ? 0u : (uint32_t)(o[6] - o[0]);
    {
        std::lock_guard<std::mutex> g(e->store_mu);
        if ((rc = arena_alloc(e, need, &dbase))) return rc;
        row = e->n_rows;
        const size_t chunks_before = e->hdr_chunks.size();
        if ((rc = ensure_hdr_chunk(e, row))) return rc;
        if (e->hdr_chunks.size() != chunks_before)       // a new chunk is zeroed on copy_stream: order q_stream behind it
            CU(e, cudaStreamSynchronize(e->copy_stream));
        if (!invalid) h.data = (unsigned long long)(uintptr_t)dbase;
        e->h_hdr.push_back(h);
        e->h_flags.push_back((uint8_t)((invalid ? 1 : 0) | 2));      // pending: not live until the kernel has run
        e->h_npos.push_back(np);
        e->h_nblk.push_back(nblk);
        e->n_rows = row + 1;
        fill_hdr_table(e, job.ht);
    }
    RowHead head;
    fill_head(e, head, h, stage);
    e->t_pack_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
    e->n_timed++;
    // (passing the row's words by value in the kernel parameters and queueing this copy behind the kernel was measured:
    //  wait -1.0 us, enqueue +0.7 us — the DMA overlaps the launch latency anyway)
    CU(e, cudaMemcpyAsync(dbase, stage, need * 4, cudaMemcpyHostToDevice, e->q_stream));
    e->h2d_bytes += (long long)need * 4;
    *out_row = row;
    job.n_rows = row;                                // candidates: every earlier row
    job.commit_head = &head; job.commit_row = row;   // an invalid row's head too: zero header, fingerprints that always mismatch
    set_slot0(e, h, row, row);
    e->slot_blocks[0] = nblk;
    rc = run_one_vs_all(e, job, nullptr, 0, out, cap, n_out);
    if (rc == FN3_OK || (rc == FN3_ECAPACITY && *n_out > cap)) {   // the kernel ran: the row is in the store
        std::lock_guard<std::mutex> g(e->store_mu);
        e->h_flags[(size_t)row] = invalid ? 1 : 0;
        e->positions += np;
        e->n_live++; if (invalid) e->n_invalid++;
    } else {
        *out_row = -1;                               // the reserved row stays a tombstone
    }
    return rc;
}

// insert() of the reference's server (findNeighbour3-server.py:516 persist + :530 mcompare) as ONE call on ONE
// stream with ONE host synchronisation: the row data goes host->device, k_query places the row's head (passed by
// value) in its tile and compares the new sample with every earlier row, the neighbour records come back.
extern "C" int fn3_insert_query(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                                int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || !out_row || (cap > 0 && !out) || (!invalid && !off))
        return set_err(e, FN3_EARG, "fn3_insert_query: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    const int rcr = maybe_refresh_heads(e);
    if (rcr) return rcr;
    return insert_query_locked(e, pos, off, invalid, ceiling, out_row, out, cap, n_out);
}

extern "C" int fn3_lists_vs_all(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                                const int64_t* rows, int64_t n_rows, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || (!invalid && !off)) return set_err(e, FN3_EARG, "fn3_lists_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    RowHdr qh; int rc;
    if ((rc = maybe_refresh_heads(e))) return rc;
    if ((rc = upload_scratch(e, 0, pos, off, invalid, &qh, &e->slot_blocks[0]))) return rc;
    QueryJob job; job.ceiling = ceiling;
    snapshot_store(e, job);
    set_slot0(e, qh, -1, -1);
    return run_one_vs_all(e, job, rows, n_rows, out, cap, n_out);
}

extern "C" int fn3_pair(fn3_engine* e, int64_t row1, int64_t row2, int32_t ceiling, fn3_hit* out, int32_t* has_dist) {
    if (!e || !out || !has_dist) return set_err(e, FN3_EARG, "fn3_pair: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    RowHdr qh, ch; int rc;
    if ((rc = stored_hdr(e, row1, &qh, "fn3_pair", &e->slot_blocks[0]))) return rc;
    if ((rc = stored_hdr(e, row2, &ch, "fn3_pair"))) return rc;
    QueryJob job; job.ceiling = ceiling;
    snapshot_store(e, job);
    set_slot0(e, qh, row1, -1);                      // the self pair IS computed here (:382-421)
    int64_t n = 0; int64_t r2 = row2;
    rc = run_one_vs_all(e, job, &r2, 1, out, 1, &n);
    *has_dist = (rc == FN3_OK && n == 1) ? 1 : 0;
    return rc;
}

extern "C" int fn3_pair_lists(fn3_engine* e, const int32_t* pos1, const int32_t off1[7], int32_t invalid1,
                              const int32_t* pos2, const int32_t off2[7], int32_t invalid2, int32_t ceiling,
                              fn3_hit* out, int32_t* has_dist) {
    if (!e || !out || !has_dist) return set_err(e, FN3_EARG, "fn3_pair_lists: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    *has_dist = 0;
    RowHdr qh, ch; int rc;
    if ((rc = upload_scratch(e, 0, pos1, off1, invalid1, &qh, &e->slot_blocks[0]))) return rc;
    RowHead chead;
    if ((rc = upload_scratch(e, 1, pos2, off2, invalid2, &ch, nullptr, &chead))) return rc;
    if (qh.data == 0ull || ch.data == 0ull) return FN3_OK;   // either invalid: None (:438-442)
    // a one-row store: the candidate's head goes to lane 0 of the scratch tile
    QueryJob job; job.ceiling = ceiling;
    memset(&job.ht, 0, sizeof job.ht);
    job.ht.chunk[0] = e->d_scratch_tile; job.ht.shift = 5; job.ht.sb_shift = 5;
    job.n_rows = 1;
    // the candidate's head travels by value in the kernel parameters (k_put_head on the query stream): nothing is shared
    // with the append path, which owns d_heads / stage_hdr under store_mu
    k_put_head<<<1, 1, 0, e->q_stream>>>(chead, 0, job.ht);
    e->launches += 1;
    CU(e, cudaGetLastError());
    set_slot0(e, qh, -1, -1);
    int64_t n = 0;
    rc = run_one_vs_all(e, job, nullptr, 0, out, 1, &n);
    if (rc == FN3_OK && n == 1) { *has_dist = 1; out->row = -1; }
    return rc;
}

// grow the survivor buffer keeping its first `keep` records and its counter block
static int grow_out_keep(fn3_engine* e, long long cap, long long keep) {
    if (cap <= e->out_cap) return FN3_OK;
    unsigned char* nb = nullptr;
    CU(e, cudaStreamSynchronize(e->q_stream));
    CU(e, cudaMalloc(&nb, FN3_OUT_HDR + (size_t)cap * sizeof(EdgeRec)));
    CU(e, cudaMemcpy(nb, e->d_outbuf, FN3_OUT_HDR + (size_t)keep * sizeof(EdgeRec), cudaMemcpyDeviceToDevice));
    cudaFree(e->d_outbuf);
    e->d_outbuf = nb;
    e->d_count = reinterpret_cast<unsigned long long*>(nb);
    e->d_out = reinterpret_cast<EdgeRec*>(nb + FN3_OUT_HDR);
    e->out_cap = cap;
    return FN3_OK;
}

// One query of an all-vs-all pass: the row's header (its data may live on a peer device of an engine group), the row id the
// edges carry as row1, and which local candidates it is compared with.
struct AvaQuery { RowHdr hdr; long long tag; long long skip_le; long long exclude; uint32_t nblk; };

// Core of fn3_all_vs_all (query_mu held, device set, job.ht / job.n_rows snapshotted): every query of `qs` against the local
// store.  Tiles of T queries are queued FN3_SLOT_RING at a time with no host synchronisation in between; edges accumulate
// behind one device counter.  After each batch the counter is read back; if the batch overflowed the buffer (the kernels
// never write past it) the buffer grows and the batch is repeated.  The edges stay in d_out (e->edges_ready).
static int all_vs_all_core(fn3_engine* e, QueryJob& job, const std::vector<AvaQuery>& qs, long long* n_total) {
    int rc;
    const int T = e->n_slots;
    if ((rc = ensure_out(e, std::max<long long>(1 << 20, 4 * (long long)qs.size())))) return rc;
    CU(e, cudaMemsetAsync(e->d_count, 0, FN3_OUT_HDR, e->q_stream));
    e->count_dirty = true;
    long long total = 0;
    const size_t n_tiles = (qs.size() + (size_t)T - 1) / (size_t)T;
    for (size_t b0 = 0; b0 < n_tiles; b0 += FN3_SLOT_RING) {
        const size_t b1 = std::min(n_tiles, b0 + (size_t)FN3_SLOT_RING);
        for (;;) {
            for (size_t t = b0; t < b1; ++t) {
                const size_t q0 = t * (size_t)T;
                const int ns = (int)std::min<size_t>((size_t)T, qs.size() - q0);
                job.ns = ns;
                job.slot_base = (int)(t - b0) * FN3_MAX_SLOTS;
                job.blocks = 1;
                for (int s = 0; s < ns; ++s) {
                    const AvaQuery& q = qs[q0 + s];
                    QuerySlot& sl = e->h_slots[job.slot_base + s];
                    sl.hdr = q.hdr; sl.row = q.tag; sl.exclude = q.exclude; sl.skip_le = q.skip_le; sl.pad = 0;
                    job.blocks = std::max(job.blocks, q.nblk);
                }
                if ((rc = enqueue_query(e, job))) return rc;
            }
            CU(e, cudaMemcpyAsync(e->h_count, e->d_count, FN3_OUT_HDR, cudaMemcpyDeviceToHost, e->q_stream));
            CU(e, cudaStreamSynchronize(e->q_stream));
            e->d2h_bytes += FN3_OUT_HDR;
            const long long cnt = (long long)e->h_count[0];
            if (cnt <= e->out_cap) { total = cnt; break; }
            // overflow: records beyond the capacity were dropped; keep what earlier batches wrote, grow, redo the batch
            if ((rc = grow_out_keep(e, std::max<long long>(cnt * 2, e->out_cap * 4), total))) return rc;
            e->h_count[0] = (unsigned long long)total;
            CU(e, cudaMemcpyAsync(e->d_count, e->h_count, sizeof(unsigned long long), cudaMemcpyHostToDevice, e->q_stream));
            CU(e, cudaStreamSynchronize(e->q_stream));
        }
    }
    *n_total = total;
    e->edges_ready = total;                           // fetchable until the next query on this engine
    return FN3_OK;
}

// copy the pending edges to the caller (through pinned staging in pieces: a pageable cudaMemcpy of a large edge list
// runs at a fraction of the link rate)
static int fetch_edges(fn3_engine* e, fn3_edge* out, long long total) {
    const long long piece = 1 << 18;                  // 8 MB of records
    int rc;
    if ((rc = ensure_hout(e, std::min<long long>(total, 2 * piece)))) return rc;
    const long long half = e->h_out_cap / 2 >= piece ? piece : std::max<long long>(1, e->h_out_cap / 2);
    cudaStream_t st = e->q_stream;
    long long done = 0; int buf = 0;
    long long prev_n = 0, prev_at = 0; int prev_buf = -1;
    while (done < total || prev_buf >= 0) {
        long long n = 0; int this_buf = -1;
        if (done < total) {
            n = std::min(half, total - done);
            this_buf = buf;
            CU(e, cudaMemcpyAsync(e->h_out + (size_t)this_buf * half, e->d_out + done, (size_t)n * sizeof(EdgeRec), cudaMemcpyDeviceToHost, st));
        }
        if (prev_buf >= 0) memcpy(out + prev_at, e->h_out + (size_t)prev_buf * half, (size_t)prev_n * sizeof(EdgeRec));
        if (this_buf >= 0) CU(e, cudaStreamSynchronize(st));
        prev_buf = this_buf; prev_n = n; prev_at = done;
        done += n; buf ^= 1;
    }
    e->d2h_bytes += total * (long long)sizeof(EdgeRec);
    return FN3_OK;
}

extern "C" int fn3_all_vs_all(fn3_engine* e, int32_t ceiling, int32_t upper_only, int64_t row_begin, int64_t row_end,
                              int64_t stride, int64_t phase, fn3_edge* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || stride < 1 || phase < 0 || phase >= stride)
        return set_err(e, FN3_EARG, "fn3_all_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    const int rcr = maybe_refresh_heads(e);
    if (rcr) return rcr;
    QueryJob job; job.ceiling = ceiling;
    snapshot_store(e, job);
    const long long n_store = job.n_rows;
    if (row_begin < 0) row_begin = 0;
    if (row_end > n_store) row_end = n_store;
    *n_out = 0;
    std::vector<AvaQuery> qs;
    {
        std::lock_guard<std::mutex> g(e->store_mu);
        for (long long r = row_begin; r < row_end; ++r)
            if ((r % stride) == phase && e->h_flags[(size_t)r] == 0) {
                AvaQuery q; q.hdr = e->h_hdr[(size_t)r]; q.tag = r; q.exclude = r; q.skip_le = upper_only ? r : -1;
                q.nblk = e->h_nblk[(size_t)r];
                qs.push_back(q);
            }
    }
    long long total = 0;
    int rc = all_vs_all_core(e, job, qs, &total);
    if (rc) return rc;
    *n_out = total;
    if (total > cap) return set_err(e, FN3_ECAPACITY, "output buffer holds %lld edges, %lld needed", (long long)cap, total);
    if (total > 0) return fetch_edges(e, out, total);
    return FN3_OK;
}

extern "C" int fn3_all_vs_all_fetch(fn3_engine* e, fn3_edge* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out)) return set_err(e, FN3_EARG, "fn3_all_vs_all_fetch: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    if (e->edges_ready < 0) return set_err(e, FN3_ENOTFOUND, "fn3_all_vs_all_fetch: no all-vs-all result is pending");
    *n_out = e->edges_ready;
    if (e->edges_ready > cap) return set_err(e, FN3_ECAPACITY, "output buffer holds %lld edges, %lld needed", (long long)cap, e->edges_ready);
    if (e->edges_ready > 0) return fetch_edges(e, out, e->edges_ready);
    return FN3_OK;
}

// ---- ingest ----------------------------------------------------------------------------------

extern "C" int fn3_set_reference(fn3_engine* e, const char* reference, const int32_t* excluded, int64_t n_excluded) {
    if (!e || !reference || n_excluded < 0 || (n_excluded > 0 && !excluded))
        return set_err(e, FN3_EARG, "fn3_set_reference: bad arguments");
    std::lock_guard<std::mutex> g(e->ingest_mu);
    FN3_ON_DEVICE(e);
    const long long per_cta = (long long)FN3_CMP_THREADS * FN3_CMP_PER_THREAD;
    e->cmp_ctas = (int)((e->L + per_cta - 1) / per_cta);
    e->Lpad = (long long)e->cmp_ctas * per_cta;
    std::vector<uint8_t> rm((size_t)e->Lpad, 0), rcode((size_t)e->Lpad, 0);
    for (long long i = 0; i < e->L; ++i) {
        const char c = reference[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T')
            return set_err(e, FN3_EARG, "fn3_set_reference: reference holds a character other than ACGT at %lld", i);
        rm[(size_t)i] = (uint8_t)c;
        rcode[(size_t)i] = (uint8_t)(c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3);
    }
    for (int64_t i = 0; i < n_excluded; ++i) {
        if (excluded[i] < 0 || excluded[i] >= e->L) return set_err(e, FN3_EARG, "fn3_set_reference: excluded position out of range");
        rm[(size_t)excluded[i]] = 0;
    }
    if (!e->in_stream) CU(e, cudaStreamCreateWithFlags(&e->in_stream, cudaStreamNonBlocking));
    if (!e->d_refmask) {
        CU(e, cudaMalloc(&e->d_refmask, (size_t)e->Lpad));
        CU(e, cudaMalloc(&e->d_seq, (size_t)e->Lpad));
        CU(e, cudaMalloc(&e->d_mchars, (size_t)e->Lpad));
        CU(e, cudaMalloc(&e->d_cpos, (size_t)e->Lpad * 4));
        CU(e, cudaMalloc(&e->d_ccounts, (size_t)e->cmp_ctas * 6 * 4));
        CU(e, cudaMalloc(&e->d_ctotals, 8 * 4));
        CU(e, cudaHostAlloc(&e->h_seq, (size_t)e->Lpad, cudaHostAllocDefault));
        CU(e, cudaHostAlloc(&e->h_ctotals, 8 * 4, cudaHostAllocDefault));
        memset(e->h_seq, 0, (size_t)e->Lpad);
    }
    CU(e, cudaMemcpy(e->d_refmask, rm.data(), (size_t)e->Lpad, cudaMemcpyHostToDevice));
    {
        // the unmasked base codes travel under query_mu: fn3_msa_sites reads them on q_stream
        std::lock_guard<std::mutex> q(e->query_mu);
        if (!e->d_refcode) CU(e, cudaMalloc(&e->d_refcode, (size_t)e->Lpad));
        CU(e, cudaMemcpy(e->d_refcode, rcode.data(), (size_t)e->Lpad, cudaMemcpyHostToDevice));
    }
    CU(e, cudaMemsetAsync(e->d_seq, 0, (size_t)e->Lpad, e->in_stream));
    CU(e, cudaStreamSynchronize(e->in_stream));
    return FN3_OK;
}

#define FN3_STAGE_PIECE ((size_t)1 << 20)

// compress() proper (ingest_mu held): sequence -> device lists -> caller's buffers; optionally the byte histogram of the
// same device copy of the sequence (hist: 256 counts, or NULL)
static int compress_locked(fn3_engine* e, const char* sequence, int32_t* pos, int64_t pos_cap, int32_t off[7],
                           char* m_chars, int64_t m_cap, int64_t* n_pos, int64_t* hist, const char* who) {
    if (!e->d_refmask) return set_err(e, FN3_EARG, "%s: call fn3_set_reference first", who);
    FN3_ON_DEVICE(e);
    cudaStream_t st = e->in_stream;
    // staged through pinned memory in 1 MB pieces: the DMA of one piece runs under the host copy of the next
    for (size_t o = 0; o < (size_t)e->L; o += FN3_STAGE_PIECE) {
        const size_t n = std::min<size_t>(FN3_STAGE_PIECE, (size_t)e->L - o);
        memcpy(e->h_seq + o, sequence + o, n);
        CU(e, cudaMemcpyAsync(e->d_seq + o, e->h_seq + o, n, cudaMemcpyHostToDevice, st));
    }
    e->h2d_bytes += e->L;
    k_compress_count<<<e->cmp_ctas, FN3_CMP_THREADS, 0, st>>>(e->d_seq, e->d_refmask, e->d_ccounts);
    // (a one-warp serial scan over the 1 078 CTA counts took 38 us here — longer than the two passes over the sequence)
    k_sel_scan<<<1, 1024, 0, st>>>(e->d_ccounts, e->d_ctotals, e->cmp_ctas, 6);
    k_compress_write<<<e->cmp_ctas, FN3_CMP_THREADS, 0, st>>>(e->d_seq, e->d_refmask, e->d_ccounts, e->d_ctotals,
                                                               e->d_cpos, e->d_mchars);
    e->launches += 3;
    CU(e, cudaGetLastError());
    CU(e, cudaMemcpyAsync(e->h_ctotals, e->d_ctotals, 6 * 4, cudaMemcpyDeviceToHost, st));
    if (hist) {
        if (!e->d_hist) {
            CU(e, cudaMalloc(&e->d_hist, 256 * sizeof(unsigned long long)));
            CU(e, cudaHostAlloc(&e->h_hist, 256 * sizeof(unsigned long long), cudaHostAllocDefault));
        }
        CU(e, cudaMemsetAsync(e->d_hist, 0, 256 * sizeof(unsigned long long), st));
        const long long chunks = (e->L + 16ll * FN3_HIST_THREADS - 1) / (16ll * FN3_HIST_THREADS);
        const int grid = (int)std::max<long long>(1, std::min<long long>(chunks, (long long)e->sm_count * 8));
        k_byte_hist<<<grid, FN3_HIST_THREADS, 0, st>>>(e->d_seq, e->L, e->d_hist);
        e->launches += 1;
        CU(e, cudaGetLastError());
        CU(e, cudaMemcpyAsync(e->h_hist, e->d_hist, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }
    CU(e, cudaStreamSynchronize(st));
    if (hist) for (int b = 0; b < 256; ++b) hist[b] = (int64_t)e->h_hist[b];
    long long total = 0;
    off[0] = 0;
    for (int c = 0; c < 6; ++c) { total += e->h_ctotals[c]; off[c + 1] = (int32_t)total; }
    const long long n_m = e->h_ctotals[5];
    *n_pos = total;
    if (total > pos_cap || n_m > m_cap || (total > 0 && !pos) || (n_m > 0 && !m_chars))
        return set_err(e, FN3_ECAPACITY, "%s: need room for %lld positions and %lld mixed characters", who, total, n_m);
    if ((size_t)total > e->h_cpos_cap) {
        if (e->h_cpos) cudaFreeHost(e->h_cpos);
        e->h_cpos = nullptr; e->h_cpos_cap = 0;
        size_t want = std::max<size_t>((size_t)total * 2, 1u << 16);
        CU(e, cudaHostAlloc(&e->h_cpos, want * 4, cudaHostAllocDefault));
        e->h_cpos_cap = want;
    }
    if ((size_t)n_m > e->h_mchars_cap) {
        if (e->h_mchars) cudaFreeHost(e->h_mchars);
        e->h_mchars = nullptr; e->h_mchars_cap = 0;
        size_t want = std::max<size_t>((size_t)n_m * 2, 1u << 12);
        CU(e, cudaHostAlloc(&e->h_mchars, want, cudaHostAllocDefault));
        e->h_mchars_cap = want;
    }
    if (total) CU(e, cudaMemcpyAsync(e->h_cpos, e->d_cpos, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
    if (n_m) CU(e, cudaMemcpyAsync(e->h_mchars, e->d_mchars, (size_t)n_m, cudaMemcpyDeviceToHost, st));
    CU(e, cudaStreamSynchronize(st));
    e->d2h_bytes += total * 4 + n_m + 24 + (hist ? 2048 : 0);
    if (total) memcpy(pos, e->h_cpos, (size_t)total * 4);
    if (n_m) memcpy(m_chars, e->h_mchars, (size_t)n_m);
    return FN3_OK;
}

extern "C" int fn3_compress(fn3_engine* e, const char* sequence, int32_t* pos, int64_t pos_cap, int32_t off[7],
                            char* m_chars, int64_t m_cap, int64_t* n_pos) {
    if (!e || !sequence || !off || !n_pos) return set_err(e, FN3_EARG, "fn3_compress: bad arguments");
    std::lock_guard<std::mutex> g(e->ingest_mu);
    return compress_locked(e, sequence, pos, pos_cap, off, m_chars, m_cap, n_pos, nullptr, "fn3_compress");
}

extern "C" int fn3_ingest_query(fn3_engine* e, const char* sequence, int64_t max_ns, int32_t ceiling, int64_t* hist,
                                int32_t* pos, int64_t pos_cap, int32_t off[7], char* m_chars, int64_t m_cap, int64_t* n_pos,
                                int32_t* invalid, int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !sequence || !off || !n_pos || !invalid || !out_row || !n_out || (cap > 0 && !out))
        return set_err(e, FN3_EARG, "fn3_ingest_query: bad arguments");
    *n_out = 0; *out_row = -1; *invalid = 0;
    {
        std::lock_guard<std::mutex> g(e->ingest_mu);
        const int rc = compress_locked(e, sequence, pos, pos_cap, off, m_chars, m_cap, n_pos, hist, "fn3_ingest_query");
        if (rc) return rc;                                   // nothing was stored (FN3_ECAPACITY: *n_pos = required)
    }
    *invalid = (int64_t)(off[6] - off[4]) > max_ns ? 1 : 0;  // len(N) + len(M) > maxNs (:301): stored without its lists
    return fn3_insert_query(e, pos, off, *invalid, ceiling, out_row, out, cap, n_out);
}

// NucleicAcid.examine's character counts (src/NucleicAcid.py:50-58) for a string of any length, without an engine: one
// small context per device (stream, staging, 256 counters), created on first use.
struct HistCtx {
    std::mutex mu;
    cudaStream_t st = nullptr;
    uint8_t* h_buf = nullptr; uint8_t* d_buf = nullptr; size_t cap = 0;
    unsigned long long* d_hist = nullptr; unsigned long long* h_hist = nullptr;
    int sm_count = 0;
};
static HistCtx g_hist[64];

extern "C" int fn3_byte_histogram(int32_t device, const uint8_t* bytes, int64_t n, int64_t hist[256]) {
    if (!hist || n < 0 || (n > 0 && !bytes) || device < 0 || device >= 64) return FN3_EARG;
    for (int b = 0; b < 256; ++b) hist[b] = 0;
    if (n == 0) return FN3_OK;
    HistCtx& c = g_hist[device];
    std::lock_guard<std::mutex> g(c.mu);
    DeviceGuard dg(device);
    if (!dg.ok) { cudaGetLastError(); return FN3_ECUDA; }   // no device: no CPU path either
    if (!c.st) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FN3_ECUDA;
        c.sm_count = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking) != cudaSuccess) { c.st = nullptr; return FN3_ECUDA; }
        if (cudaMalloc(&c.d_hist, 256 * sizeof(unsigned long long)) != cudaSuccess) return FN3_ENOMEM;
        if (cudaHostAlloc(&c.h_hist, 256 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess) return FN3_ENOMEM;
    }
    if ((size_t)n > c.cap) {
        if (c.h_buf) cudaFreeHost(c.h_buf);
        if (c.d_buf) cudaFree(c.d_buf);
        c.h_buf = nullptr; c.d_buf = nullptr; c.cap = 0;
        const size_t want = std::max<size_t>((size_t)n + (size_t)n / 4, 1u << 20);
        if (cudaHostAlloc(&c.h_buf, want, cudaHostAllocDefault) != cudaSuccess) return FN3_ENOMEM;
        if (cudaMalloc(&c.d_buf, want) != cudaSuccess) return FN3_ENOMEM;
        c.cap = want;
    }
    for (size_t o = 0; o < (size_t)n; o += FN3_STAGE_PIECE) {
        const size_t m = std::min<size_t>(FN3_STAGE_PIECE, (size_t)n - o);
        memcpy(c.h_buf + o, bytes + o, m);
        if (cudaMemcpyAsync(c.d_buf + o, c.h_buf + o, m, cudaMemcpyHostToDevice, c.st) != cudaSuccess) return FN3_ECUDA;
    }
    if (cudaMemsetAsync(c.d_hist, 0, 256 * sizeof(unsigned long long), c.st) != cudaSuccess) return FN3_ECUDA;
    const long long chunks = (n + 16ll * FN3_HIST_THREADS - 1) / (16ll * FN3_HIST_THREADS);
    const int grid = (int)std::max<long long>(1, std::min<long long>(chunks, (long long)c.sm_count * 8));
    k_byte_hist<<<grid, FN3_HIST_THREADS, 0, c.st>>>(c.d_buf, n, c.d_hist);
    if (cudaGetLastError() != cudaSuccess) return FN3_ECUDA;
    if (cudaMemcpyAsync(c.h_hist, c.d_hist, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.st) != cudaSuccess) return FN3_ECUDA;
    if (cudaStreamSynchronize(c.st) != cudaSuccess) return FN3_ECUDA;
    for (int b = 0; b < 256; ++b) hist[b] = (int64_t)c.h_hist[b];
    return FN3_OK;
}

// ---- cluster statistics: consensus vote, MSA variant sites, alignment (fn3_cluster.cuh) --------------

template <class T>
static int ensure_dev(fn3_engine* e, T** p, size_t* cap, size_t need) {
    if (need <= *cap) return FN3_OK;
    const size_t want = std::max<size_t>(std::max(need, *cap * 2), 1024);
    if (*p) { cudaStreamSynchronize(e->q_stream); cudaFree(*p); }
    *p = nullptr; *cap = 0;
    CU(e, cudaMalloc(p, want * sizeof(T)));
    *cap = want;
    return FN3_OK;
}

static int ensure_votes(fn3_engine* e) {
    if (e->d_votes) return FN3_OK;
    const long long per_cta = (long long)FN3_SEL_THREADS * FN3_SEL_PER_THREAD;
    e->sel_ctas = (int)((e->L + per_cta - 1) / per_cta);
    e->vLpad = (long long)e->sel_ctas * per_cta;
    CU(e, cudaMalloc(&e->d_votes, (size_t)e->vLpad * 5 * sizeof(uint32_t)));
    CU(e, cudaMalloc(&e->d_sel_counts, (size_t)e->sel_ctas * 5 * sizeof(uint32_t)));
    CU(e, cudaMalloc(&e->d_sel_totals, 8 * sizeof(uint32_t)));
    CU(e, cudaHostAlloc(&e->h_sel_totals, 8 * sizeof(uint32_t), cudaHostAllocDefault));
    return FN3_OK;
}

// validate the rows against the host mirror, snapshot the head table, copy the list to d_rows (query_mu held)
static int cluster_rows(fn3_engine* e, const int64_t* rows, int64_t n, bool reject_invalid, const char* who, HdrTable* ht) {
    int rc;
    if ((rc = ensure_rows(e, std::max<long long>(n, 1)))) return rc;
    {
        std::lock_guard<std::mutex> g(e->store_mu);
        fill_hdr_table(e, *ht);
        for (int64_t i = 0; i < n; ++i) {
            const long long r = rows[i];
            if (r < 0 || r >= e->n_rows) return set_err(e, FN3_ENOTFOUND, "%s: row %lld does not exist", who, r);
            const uint8_t fl = e->h_flags[(size_t)r];
            if (fl & 2) return set_err(e, FN3_ENOTFOUND, "%s: row %lld was removed", who, r);
            if (reject_invalid && (fl & 1)) return set_err(e, FN3_EARG, "%s: row %lld is an invalid sample", who, r);
            e->h_rows[i] = r;
        }
    }
    if (n > 0) {
        CU(e, cudaMemcpyAsync(e->d_rows, e->h_rows, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, e->q_stream));
        e->h2d_bytes += n * (long long)sizeof(long long);
    }
    return FN3_OK;
}

static int grid_for(const fn3_engine* e, long long work_items, int threads) {
    const long long want = (work_items + threads - 1) / threads;
    return (int)std::max<long long>(1, std::min<long long>(want, (long long)e->sm_count * 8));
}

// count -> scan -> (host reads the list lengths) -> write; the selected positions stay in d_sel_out
template <int MODE>
static int run_select(fn3_engine* e, uint32_t thr, long long len[5], long long* total) {
    const int NL = SelLists<MODE>::N;
    cudaStream_t st = e->q_stream;
    k_sel_count<MODE><<<e->sel_ctas, FN3_SEL_THREADS, 0, st>>>(e->d_votes, e->d_refcode, e->vLpad, thr, e->d_sel_counts);
    k_sel_scan<<<1, 1024, 0, st>>>(e->d_sel_counts, e->d_sel_totals, e->sel_ctas, NL);
    e->launches += 2;
    CU(e, cudaGetLastError());
    CU(e, cudaMemcpyAsync(e->h_sel_totals, e->d_sel_totals, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(e, cudaStreamSynchronize(st));
    e->d2h_bytes += 8 * (long long)sizeof(uint32_t);
    *total = 0;
    for (int l = 0; l < 5; ++l) { len[l] = l < NL ? (long long)e->h_sel_totals[l] : 0; *total += len[l]; }
    int rc;
    if ((rc = ensure_dev(e, &e->d_sel_out, &e->sel_out_cap, (size_t)std::max<long long>(*total, 1)))) return rc;
    if (*total > 0) {
        k_sel_write<MODE><<<e->sel_ctas, FN3_SEL_THREADS, 0, st>>>(e->d_votes, e->d_refcode, e->vLpad, thr, e->d_sel_counts, e->d_sel_totals,
                                                                   e->d_sel_out);
        e->launches += 1;
        CU(e, cudaGetLastError());
    }
    return FN3_OK;
}

static int fetch_selected(fn3_engine* e, long long total, int32_t* out, int64_t cap, const char* who) {
    if (total > cap || (total > 0 && !out))
        return set_err(e, FN3_ECAPACITY, "%s: output buffer holds %lld positions, %lld needed", who, (long long)cap, total);
    if (total > 0) {
        CU(e, cudaMemcpyAsync(out, e->d_sel_out, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, e->q_stream));
        e->d2h_bytes += total * (long long)sizeof(int32_t);
    }
    CU(e, cudaStreamSynchronize(e->q_stream));
    return FN3_OK;
}

static int consensus_tail(fn3_engine* e, int64_t min_votes, int32_t* pos, int64_t pos_cap, int32_t off[6], int64_t* n_pos,
                          const char* who) {
    long long len[5], total = 0;
    const uint32_t thr = (uint32_t)std::min<int64_t>(min_votes, 0xffffffffll);
    int rc = run_select<1>(e, thr, len, &total);
    if (rc) return rc;
    off[0] = 0;
    for (int l = 0; l < 5; ++l) off[l + 1] = off[l] + (int32_t)len[l];
    *n_pos = total;
    return fetch_selected(e, total, pos, pos_cap, who);
}

// Rows of a "virtual store" for the cluster kernels: headers given by value — their row data may live on PEER devices of
// an engine group, read over NVLink — uploaded as head tiles of 32, with the identity row list in d_rows.  query_mu held.
static int virtual_rows(fn3_engine* e, const RowHdr* hdrs, int64_t n, HdrTable* ht) {
    int rc;
    if ((rc = ensure_rows(e, std::max<long long>(n, 1)))) return rc;
    const size_t n_tiles = (size_t)((n + 31) / 32);
    if ((rc = ensure_dev(e, &e->d_vtiles, &e->vtiles_cap, std::max<size_t>(n_tiles, 1) * FN3_TILE_BYTES))) return rc;
    std::vector<unsigned char> tiles(std::max<size_t>(n_tiles, 1) * FN3_TILE_BYTES, 0);
    for (int64_t i = 0; i < n; ++i) {
        memcpy(tiles.data() + (size_t)(i >> 5) * FN3_TILE_BYTES + (size_t)(i & 31) * 32, &hdrs[i], sizeof(RowHdr));
        e->h_rows[i] = i;
    }
    CU(e, cudaMemcpyAsync(e->d_vtiles, tiles.data(), tiles.size(), cudaMemcpyHostToDevice, e->q_stream));
    if (n > 0) CU(e, cudaMemcpyAsync(e->d_rows, e->h_rows, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, e->q_stream));
    CU(e, cudaStreamSynchronize(e->q_stream));             // `tiles` is pageable and leaves scope
    e->h2d_bytes += (long long)tiles.size() + n * (long long)sizeof(long long);
    memset(ht, 0, sizeof *ht);
    ht->chunk[0] = e->d_vtiles; ht->shift = 26; ht->sb_shift = 5;
    return FN3_OK;
}

// consensus over the rows in d_rows[0..n_rows) of the store described by ht (query_mu held, device set)
static int consensus_core(fn3_engine* e, const HdrTable& ht, int64_t n_rows, int64_t min_votes, int32_t* pos,
                          int64_t pos_cap, int32_t off[6], int64_t* n_pos, const char* who) {
    int rc;
    if ((rc = ensure_votes(e))) return rc;
    cudaStream_t st = e->q_stream;
    CU(e, cudaMemsetAsync(e->d_votes, 0, (size_t)e->vLpad * 5 * sizeof(uint32_t), st));
    if (n_rows > 0) {
        k_votes_rows<<<(unsigned)std::min<long long>(n_rows, (long long)e->sm_count * 8), FN3_VOTE_THREADS, 0, st>>>(e->d_rows, n_rows, ht, e->d_votes, e->vLpad, 1, e->pos_bits);
        e->launches += 1;
        CU(e, cudaGetLastError());
    }
    return consensus_tail(e, min_votes, pos, pos_cap, off, n_pos, who);
}

extern "C" int fn3_consensus(fn3_engine* e, const int64_t* rows, int64_t n_rows, int64_t min_votes, int32_t* pos,
                             int64_t pos_cap, int32_t off[6], int64_t* n_pos) {
    if (!e || !off || !n_pos || n_rows < 0 || (n_rows > 0 && !rows) || min_votes < 1)
        return set_err(e, FN3_EARG, "fn3_consensus: bad arguments (min_votes must be >= 1)");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    int rc;
    HdrTable ht;
    if ((rc = cluster_rows(e, rows, n_rows, false, "fn3_consensus", &ht))) return rc;
    return consensus_core(e, ht, n_rows, min_votes, pos, pos_cap, off, n_pos, "fn3_consensus");
}

extern "C" int fn3_consensus_lists(fn3_engine* e, int64_t n, const int32_t* pos_in, const int64_t* off_in, int64_t min_votes,
                                   int32_t* pos, int64_t pos_cap, int32_t off[6], int64_t* n_pos) {
    if (!e || !off || !n_pos || n < 0 || (n > 0 && !off_in) || min_votes < 1)
        return set_err(e, FN3_EARG, "fn3_consensus_lists: bad arguments (min_votes must be >= 1)");
    const long long total_in = n > 0 ? (long long)(off_in[6 * n] - off_in[0]) : 0;
    if (total_in > 0 && !pos_in) return set_err(e, FN3_EARG, "fn3_consensus_lists: null lists");
    for (int64_t l = 0; l < 6 * n; ++l) {
        if (off_in[l + 1] < off_in[l]) return set_err(e, FN3_EARG, "fn3_consensus_lists: list offsets must be non-decreasing");
        for (int64_t i = off_in[l]; i < off_in[l + 1]; ++i) {
            if (pos_in[i] < 0 || (long long)pos_in[i] >= e->L) return set_err(e, FN3_EARG, "fn3_consensus_lists: position out of range");
            if (i > off_in[l] && pos_in[i] <= pos_in[i - 1])
                return set_err(e, FN3_EARG, "fn3_consensus_lists: positions must be strictly increasing within a list");
        }
    }
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    int rc;
    if ((rc = ensure_votes(e))) return rc;
    cudaStream_t st = e->q_stream;
    CU(e, cudaMemsetAsync(e->d_votes, 0, (size_t)e->vLpad * 5 * sizeof(uint32_t), st));
    if (total_in > 0) {
        if ((rc = ensure_dev(e, &e->d_lpos, &e->lpos_cap, (size_t)total_in))) return rc;
        if ((rc = ensure_dev(e, &e->d_loff, &e->loff_cap, (size_t)(6 * n + 1)))) return rc;
        std::vector<long long> rel((size_t)(6 * n + 1));
        for (int64_t l = 0; l <= 6 * n; ++l) rel[(size_t)l] = (long long)(off_in[l] - off_in[0]);
        CU(e, cudaMemcpyAsync(e->d_lpos, pos_in + off_in[0], (size_t)total_in * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU(e, cudaMemcpyAsync(e->d_loff, rel.data(), rel.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
        CU(e, cudaStreamSynchronize(st));                       // `rel` is pageable and leaves scope
        e->h2d_bytes += total_in * (long long)sizeof(int32_t) + (long long)rel.size() * 8;
        k_votes_lists<<<grid_for(e, total_in, 256), 256, 0, st>>>(e->d_lpos, e->d_loff, 6 * n, total_in, e->d_votes, e->vLpad);
        e->launches += 1;
        CU(e, cudaGetLastError());
    }
    return consensus_tail(e, min_votes, pos, pos_cap, off, n_pos, "fn3_consensus_lists");
}

static int msa_sites_core(fn3_engine* e, const HdrTable& ht, int64_t n_rows, int32_t rule, int32_t* sites, int64_t sites_cap,
                          int64_t* n_sites) {
    if (!e->d_refcode) return set_err(e, FN3_EARG, "fn3_msa_sites: call fn3_set_reference first (the variant-site rule needs the reference bases)");
    int rc;
    if ((rc = ensure_votes(e))) return rc;
    cudaStream_t st = e->q_stream;
    const int planes = rule == FN3_SITES_CALLS_OR_N ? 5 : 4;          // the N plane only where the rule reads it
    CU(e, cudaMemsetAsync(e->d_votes, 0, (size_t)e->vLpad * planes * sizeof(uint32_t), st));
    k_votes_rows<<<(unsigned)std::min<long long>(n_rows, (long long)e->sm_count * 8), FN3_VOTE_THREADS, 0, st>>>(e->d_rows, n_rows, ht, e->d_votes, e->vLpad, planes == 5, e->pos_bits);
    e->launches += 1;
    CU(e, cudaGetLastError());
    long long len[5], total = 0;
    const uint32_t thr = (uint32_t)std::min<int64_t>(n_rows, 0xffffffffll);
    if ((rc = rule == FN3_SITES_CALLS_OR_N ? run_select<2>(e, thr, len, &total) : run_select<0>(e, thr, len, &total))) return rc;
    *n_sites = total;
    return fetch_selected(e, total, sites, sites_cap, "fn3_msa_sites");
}

extern "C" int fn3_msa_sites_rule(fn3_engine* e, const int64_t* rows, int64_t n_rows, int32_t rule, int32_t* sites,
                                  int64_t sites_cap, int64_t* n_sites) {
    if (!e || !n_sites || n_rows < 0 || (n_rows > 0 && !rows) || (rule != FN3_SITES_CALLS && rule != FN3_SITES_CALLS_OR_N))
        return set_err(e, FN3_EARG, "fn3_msa_sites: bad arguments");
    *n_sites = 0;
    if (n_rows == 0) return FN3_OK;
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    int rc;
    HdrTable ht;
    if ((rc = cluster_rows(e, rows, n_rows, true, "fn3_msa_sites", &ht))) return rc;
    return msa_sites_core(e, ht, n_rows, rule, sites, sites_cap, n_sites);
}

extern "C" int fn3_msa_sites(fn3_engine* e, const int64_t* rows, int64_t n_rows, int32_t* sites, int64_t sites_cap,
                             int64_t* n_sites) {
    return fn3_msa_sites_rule(e, rows, n_rows, FN3_SITES_CALLS, sites, sites_cap, n_sites);
}

static int msa_align_check_sites(fn3_engine* e, const int32_t* sites, int64_t n_sites) {
    for (int64_t k = 0; k < n_sites; ++k) {
        if (sites[k] < 0 || (long long)sites[k] >= e->L) return set_err(e, FN3_EARG, "fn3_msa_align: site out of range");
        if (k && sites[k] <= sites[k - 1]) return set_err(e, FN3_EARG, "fn3_msa_align: sites must be strictly increasing");
    }
    return FN3_OK;
}

static int msa_align_core(fn3_engine* e, const HdrTable& ht, int64_t n_rows, const int32_t* sites, int64_t n_sites,
                          uint8_t* codes, int32_t* counts) {
    int rc;
    cudaStream_t st = e->q_stream;
    if ((rc = ensure_dev(e, &e->d_sites, &e->sites_cap, (size_t)std::max<int64_t>(n_sites, 1)))) return rc;
    if ((rc = ensure_dev(e, &e->d_acounts, &e->acounts_cap, (size_t)n_rows * 4))) return rc;
    const size_t ncodes = (size_t)n_rows * (size_t)n_sites;
    const bool want_codes = codes != nullptr && ncodes > 0;
    if (want_codes) {
        if ((rc = ensure_dev(e, &e->d_codes, &e->codes_cap, ncodes))) return rc;
        CU(e, cudaMemsetAsync(e->d_codes, 0, ncodes, st));
    }
    if (n_sites > 0) {
        CU(e, cudaMemcpyAsync(e->d_sites, sites, (size_t)n_sites * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        e->h2d_bytes += n_sites * (long long)sizeof(int32_t);
    }
    k_align<<<grid_for(e, n_rows * 32, 256), 256, 0, st>>>(e->d_rows, n_rows, ht, e->d_sites, n_sites,
                                                          want_codes ? e->d_codes : nullptr, e->d_acounts, e->pos_bits);
    e->launches += 1;
    CU(e, cudaGetLastError());
    CU(e, cudaMemcpyAsync(counts, e->d_acounts, (size_t)n_rows * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (want_codes) CU(e, cudaMemcpyAsync(codes, e->d_codes, ncodes, cudaMemcpyDeviceToHost, st));
    CU(e, cudaStreamSynchronize(st));
    e->d2h_bytes += n_rows * 16 + (want_codes ? (long long)ncodes : 0);
    return FN3_OK;
}

extern "C" int fn3_msa_align(fn3_engine* e, const int64_t* rows, int64_t n_rows, const int32_t* sites, int64_t n_sites,
                             uint8_t* codes, int32_t* counts) {
    if (!e || n_rows < 0 || n_sites < 0 || (n_rows > 0 && (!rows || !counts)) || (n_sites > 0 && !sites))
        return set_err(e, FN3_EARG, "fn3_msa_align: bad arguments");
    int rc;
    if ((rc = msa_align_check_sites(e, sites, n_sites))) return rc;
    if (n_rows == 0) return FN3_OK;
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    HdrTable ht;
    if ((rc = cluster_rows(e, rows, n_rows, true, "fn3_msa_align", &ht))) return rc;
    return msa_align_core(e, ht, n_rows, sites, n_sites, codes, counts);
}

// ---- wire format: the reference's GridFS pickles -> packed lists (fn3_pickle.h) --------------------------

extern "C" int fn3_unpickle_samples(int64_t n, const uint8_t* blobs, const int64_t* blob_off, int32_t* pos, int64_t pos_cap,
                                    int64_t* off, uint8_t* status, int64_t* n_pos) {
    if (n < 0 || !blob_off || !off || !status || !n_pos || (n > 0 && !blobs)) return FN3_EARG;
    for (int64_t i = 0; i < n; ++i) if (blob_off[i + 1] < blob_off[i]) return FN3_EARG;
    const int host_threads = (int)std::max<long long>(1, std::min<long long>(env_ll("FN3_HOST_THREADS", 16),
                                                                             (long long)std::thread::hardware_concurrency()));
    // pass 1 (host threads): parse every pickle into its own lists; pass 2: lay them out one behind the other
    std::vector<fn3pickle::Parsed> parsed((size_t)n);
    parallel_rows(n, host_threads, [&](long long lo, long long hi) {
        for (long long i = lo; i < hi; ++i)
            status[i] = (uint8_t)fn3pickle::parse_sample(blobs + blob_off[i], blob_off[i + 1] - blob_off[i], parsed[(size_t)i]);
    }, 32);
    int64_t total = 0;
    off[0] = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int l = 0; l < 6; ++l) {
            if (status[i] == 0) total += (int64_t)parsed[(size_t)i].list[l].size();
            off[6 * i + l + 1] = total;
        }
    *n_pos = total;
    if (total > pos_cap || (total > 0 && !pos)) return FN3_ECAPACITY;
    parallel_rows(n, host_threads, [&](long long lo, long long hi) {
        for (long long i = lo; i < hi; ++i) {
            if (status[i] != 0) continue;
            for (int l = 0; l < 6; ++l) {
                const std::vector<int32_t>& v = parsed[(size_t)i].list[l];
                if (!v.empty()) memcpy(pos + off[6 * i + l], v.data(), v.size() * sizeof(int32_t));
            }
        }
    }, 256);
    return FN3_OK;
}

// ---- measurement hook ----------------------------------------------------------------------

extern "C" int fn3_profile_queries(fn3_engine* e, const int64_t* query_rows, int32_t n, int32_t ceiling, int32_t flush_l2,
                                   float* ms_build, float* ms_compare, int64_t* n_hits, int64_t* bytes_touched,
                                   int64_t* n_passed) {
    if (!e || !query_rows || n < 1) return set_err(e, FN3_EARG, "fn3_profile_queries: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    FN3_ON_DEVICE(e);
    int rc = maybe_refresh_heads(e);                 // as every query entry point does
    if (rc) return rc;
    QueryJob job; job.ceiling = ceiling;
    snapshot_store(e, job);
    job.adaptive = job.n_rows >= 1024;               // the variant a real full scan would take now (FILTER / STREAM steering)
    e->edges_ready = -1;
    e->count_dirty = true;
    if ((rc = ensure_out(e, std::max<long long>(job.n_rows, 1)))) return rc;
    if (flush_l2 && !e->d_flush) {
        e->flush_n = (256ll << 20) / 16;
        CU(e, cudaMalloc(&e->d_flush, (size_t)e->flush_n * 16));
        CU(e, cudaMemsetAsync(e->d_flush, 0, (size_t)e->flush_n * 16, e->q_stream));
    }
    cudaStream_t st = e->q_stream;
    std::vector<cudaEvent_t> evs((size_t)n * 3, nullptr);
    std::vector<uint8_t> has_split((size_t)n, 0);
    for (auto& ev : evs) CU(e, cudaEventCreate(&ev));
    unsigned long long* d_pc = nullptr;                       // per-step counters: [hits, touched, -, -]
    CU(e, cudaMalloc(&d_pc, (size_t)n * FN3_OUT_HDR));
    CU(e, cudaMemsetAsync(d_pc, 0, (size_t)n * FN3_OUT_HDR, st));
    unsigned long long* saved = e->d_count;
    std::vector<RowHdr> hdrs((size_t)n);
    std::vector<uint32_t> blks((size_t)n);
    for (int it = 0; it < n; ++it)
        if ((rc = stored_hdr(e, query_rows[it], &hdrs[(size_t)it], "fn3_profile_queries", &blks[(size_t)it]))) break;
    // the whole batch is queued without any host synchronisation (the query slot travels by value in the
    // kernel parameters); events bracket each step
    for (int pass = 0; pass < 2 && rc == FN3_OK; ++pass) {
        if (pass == 1) {
            if (!bytes_touched) break;
            // instrumented, untimed pass: bytes of heads + row data each k_query launch actually loads
            CU(e, cudaStreamSynchronize(st));
            std::vector<unsigned long long> pc((size_t)n * 4, 0);
            cudaMemcpy(pc.data(), d_pc, (size_t)n * FN3_OUT_HDR, cudaMemcpyDeviceToHost);
            if (n_hits) for (int it = 0; it < n; ++it) n_hits[it] = (int64_t)pc[(size_t)it * 4];
            cudaMemsetAsync(d_pc, 0, (size_t)n * FN3_OUT_HDR, st);
        }
        for (int it = 0; it < n && rc == FN3_OK; ++it) {
            // the query slot is staged in a ring of pinned slots (a query too large for shared memory is copied to the device
            // asynchronously): drain the stream before a ring position is reused
            job.slot_base = (it % FN3_SLOT_RING) * FN3_MAX_SLOTS;
            if (it && job.slot_base == 0) CU(e, cudaStreamSynchronize(st));
            QuerySlot& sl = e->h_slots[job.slot_base];
            sl.hdr = hdrs[(size_t)it]; sl.row = query_rows[it]; sl.skip_le = -1; sl.exclude = query_rows[it]; sl.pad = 0;
            e->slot_blocks[0] = blks[(size_t)it];
            if (pass == 0 && flush_l2) { k_fill<<<e->sm_count * 4, 512, 0, st>>>(e->d_flush, e->flush_n, (uint32_t)it); e->launches += 1; }
            e->d_count = d_pc + (size_t)it * (FN3_OUT_HDR / sizeof(unsigned long long));
            job.count_bytes = pass == 1;
            job.split = pass == 0 ? evs[(size_t)it * 3 + 1] : nullptr;
            if (pass == 0) cudaEventRecord(evs[(size_t)it * 3 + 0], st);
            job.split_recorded = false;
            rc = enqueue_query(e, job);
            if (pass == 0) { cudaEventRecord(evs[(size_t)it * 3 + 2], st); has_split[(size_t)it] = job.split_recorded ? 1 : 0; }
        }
        e->d_count = saved;
    }
    cudaError_t sync = cudaStreamSynchronize(st);
    if (rc == FN3_OK && sync == cudaSuccess) {
        for (int it = 0; it < n; ++it) {
            float a = 0, b = 0;
            if (has_split[(size_t)it]) {
                cudaEventElapsedTime(&a, evs[(size_t)it * 3 + 0], evs[(size_t)it * 3 + 1]);
                cudaEventElapsedTime(&b, evs[(size_t)it * 3 + 1], evs[(size_t)it * 3 + 2]);
            } else {
                cudaEventElapsedTime(&b, evs[(size_t)it * 3 + 0], evs[(size_t)it * 3 + 2]);      // no build kernel: one interval
            }
            if (ms_build) ms_build[it] = a;
            if (ms_compare) ms_compare[it] = b;
        }
        std::vector<unsigned long long> pc((size_t)n * 4, 0);
        cudaMemcpy(pc.data(), d_pc, (size_t)n * FN3_OUT_HDR, cudaMemcpyDeviceToHost);
        if (bytes_touched) for (int it = 0; it < n; ++it) bytes_touched[it] = (int64_t)pc[(size_t)it * 4 + 1];
        if (bytes_touched && n_passed) for (int it = 0; it < n; ++it) n_passed[it] = (int64_t)pc[(size_t)it * 4 + 2];
        else if (n_hits) for (int it = 0; it < n; ++it) n_hits[it] = (int64_t)pc[(size_t)it * 4];
    }
    for (auto ev : evs) cudaEventDestroy(ev);
    cudaFree(d_pc);
    if (rc != FN3_OK) return rc;
    if (sync != cudaSuccess) return set_err(e, FN3_ECUDA, "fn3_profile_queries: %s", cudaGetErrorString(sync));
    return FN3_OK;
}

#if FN3_TRACE
// trace build only (scripts/trace_prof.py): the SM clocks left by the last compare launch, then cleared
extern "C" int fn3_debug_trace(unsigned long long* out, int n_words) {
    const size_t bytes = std::min<size_t>((size_t)n_words * 8, sizeof(unsigned long long) * FN3_TRACE_CTAS * FN3_TRACE_PTS);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out, fn3_trace_buf, bytes) != cudaSuccess) return -1;
    std::vector<unsigned long long> z(FN3_TRACE_CTAS * FN3_TRACE_PTS, 0ull);
    return cudaMemcpyToSymbol(fn3_trace_buf, z.data(), z.size() * 8) == cudaSuccess ? 0 : -1;
}
#endif

#include "fn3_group.inl"
