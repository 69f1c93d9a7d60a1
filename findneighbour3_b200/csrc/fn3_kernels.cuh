// fn3_kernels.cuh — sm_100a kernels of the SNV-distance engine: query build, k_query (the compare kernel),
// head scatter.  Included by fn3_engine.cu only.
//
// Reference being replaced (davidhwyllie/findNeighbour3, src/seqComparer.py):
//   countDifferences :424-458, countDifferences_byKey + _setStats :382-421/:354-380, mcompare :149-172.
#pragma once
#include "fn3_types.h"

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

// -DFN3_TRACE=1 (scripts/trace_prof.py, never the shipped build): thread 0 of every CTA leaves its SM clock at numbered
// points of the compare kernels, read back through fn3_debug_trace()
#ifndef FN3_TRACE
#define FN3_TRACE 0
#endif
#if FN3_TRACE
#define FN3_TRACE_CTAS 512
#define FN3_TRACE_PTS 40
__device__ unsigned long long fn3_trace_buf[FN3_TRACE_CTAS * FN3_TRACE_PTS];
#define FN3_TR(i) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < FN3_TRACE_CTAS) \
        fn3_trace_buf[blockIdx.x * FN3_TRACE_PTS + (i)] = (unsigned long long)clock64(); } while (0)
__device__ __forceinline__ unsigned long long fn3_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FN3_TRG(i) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < FN3_TRACE_CTAS) \
        fn3_trace_buf[blockIdx.x * FN3_TRACE_PTS + (i)] = fn3_globaltimer(); } while (0)
#else
#define FN3_TR(i) do { } while (0)
#define FN3_TRG(i) do { } while (0)
#endif

__device__ __forceinline__ uint4 ld_stream_v4(const uint32_t* p) {
    // candidate row data is read once per comparison: do not let it displace anything in L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint4 ld_head_v4(const uint4* p) {
    // volatile: ptxas must not sink these loads below the early-exit tests of the fingerprint loop (it does so with
    // plain loads, turning one memory round trip into up to four dependent ones); they are served by L2
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t lowmask(uint32_t k) { return (1u << k) - 1u; }    // k in 0..31

// block-wide exclusive scan of one value per thread (blockDim.x multiple of 32, <= 1024);
// ws = 33 words of shared scratch; *total receives the block total.
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t* ws, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();                       // ws may still be read from a previous call
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t a = lane < nwarp ? ws[lane] : 0u, i2 = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, i2, o); if (lane >= o) i2 += t; }
        ws[lane] = i2 - a;
        if (lane == 31) ws[32] = i2;
    }
    __syncthreads();
    *total = ws[32];
    return ws[warp] + inc - v;
}

// three exclusive block scans sharing their barriers; ws = 99 words
__device__ __forceinline__ void block_exscan3(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t* ws,
                                              uint32_t& ta, uint32_t& tb, uint32_t& tc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t ia = a, ib = b, ic = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o), z = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) { ia += x; ib += y; ic += z; }
    }
    __syncthreads();
    if (lane == 31) { ws[warp] = ia; ws[33 + warp] = ib; ws[66 + warp] = ic; }
    __syncthreads();
    if (warp < 3) {
        uint32_t* w = ws + 33 * warp;
        uint32_t v = lane < nwarp ? w[lane] : 0u, i2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, i2, o); if (lane >= o) i2 += t; }
        w[lane] = i2 - v;
        if (lane == 31) w[32] = i2;
    }
    __syncthreads();
    ta = ws[32]; tb = ws[65]; tc = ws[98];
    a = ws[warp] + ia - a; b = ws[33 + warp] + ib - b; c = ws[66 + warp] + ic - c;
}

// One-barrier variants for the query build (called once each, on disjoint scratch: no guard barrier in front, and no third
// barrier behind — every warp sums the warp totals in front of it itself, two REDUX).  ws: 32 words per scanned value.
#define FN3_WS_WORDS 128
__device__ __forceinline__ uint32_t block_exscan_1b(uint32_t v, uint32_t* ws, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    const uint32_t w = lane < nwarp ? ws[lane] : 0u;
    *total = __reduce_add_sync(0xffffffffu, w);
    return __reduce_add_sync(0xffffffffu, lane < warp ? w : 0u) + inc - v;
}

__device__ __forceinline__ void block_exscan3_1b(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t* ws,
                                                 uint32_t& ta, uint32_t& tb, uint32_t& tc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t ia = a, ib = b, ic = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o), z = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) { ia += x; ib += y; ic += z; }
    }
    if (lane == 31) { ws[warp] = ia; ws[32 + warp] = ib; ws[64 + warp] = ic; }
    __syncthreads();
    const uint32_t wa = lane < nwarp ? ws[lane] : 0u, wb = lane < nwarp ? ws[32 + lane] : 0u, wc = lane < nwarp ? ws[64 + lane] : 0u;
    ta = __reduce_add_sync(0xffffffffu, wa); tb = __reduce_add_sync(0xffffffffu, wb); tc = __reduce_add_sync(0xffffffffu, wc);
    const bool before = lane < warp;
    a = __reduce_add_sync(0xffffffffu, before ? wa : 0u) + ia - a;
    b = __reduce_add_sync(0xffffffffu, before ? wb : 0u) + ib - b;
    c = __reduce_add_sync(0xffffffffu, before ? wc : 0u) + ic - c;
}

#ifndef FN3_PF
#define FN3_PF 1
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// one instruction asks `bytes` (multiple of 16) into L2 through the bulk-copy engine: nothing enters the load/store queue
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// address of the tile holding `row`, and the row's lane in it
__device__ __forceinline__ const unsigned char* tile_of(const HdrTable& ht, long long row, int& lane) {
    const long long local = row & ((1ll << ht.shift) - 1);
    long long tile;
    fn3_slot_of(local, ht.sb_shift, tile, lane);
    return ht.chunk[row >> ht.shift] + tile * FN3_TILE_BYTES;
}

__device__ __forceinline__ void put_head(const HdrTable& ht, long long row, const RowHead& h) {
    int lane;
    unsigned char* t = const_cast<unsigned char*>(tile_of(ht, row, lane));
    const uint4* src = reinterpret_cast<const uint4*>(&h);
    uint4* hdr = reinterpret_cast<uint4*>(t + lane * 32);
    // fingerprints first, header last: a concurrent reader sees either a zero (skipped) or a complete head
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(t + FN3_TILE_FP_OFF + j * 512 + lane * 16) = src[2 + j];
    __threadfence();
    hdr[1] = src[1];
    hdr[0] = src[0];
}

// one head by value (fn3_remove: zero header + fingerprints that always mismatch)
__global__ void k_put_head(const RowHead h, long long row, const HdrTable ht) { put_head(ht, row, h); }

// bulk append: heads arrive in row order, go to their tiles
__global__ void k_scatter_heads(const RowHead* __restrict__ src, long long first_row, long long n, const HdrTable ht) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) put_head(ht, first_row + i, src[i]);
}

// ------------------------------------------------------------------------------------------------
// query build: one CTA turns the query's own row into {bmpref, coarse, entries} — in shared memory
// (inside k_query) or in global memory (k_build_query, for queries that exceed shared memory)
// ------------------------------------------------------------------------------------------------

#define FN3_BUILD_KEEP 2

template <bool GLOBAL> __device__ __forceinline__ uint32_t ldq(const uint32_t* p) { return GLOBAL ? __ldcg(p) : *p; }

// All threads of the CTA must call.  bm: 2*nw words (bitmap, then prefix); coarse: ncw words; en: (touched+1)
// entries; ws: FN3_WS_WORDS words.
// tot[0..2] = |V_q|, |U_q|, |M_q|; returns the number of touched blocks.
template <bool GLOBAL, int NT>                               // NT = blockDim.x (a constant: the ranges below divide by it)
__device__ __forceinline__ uint32_t build_query(const QuerySlot& q, int nw, int ncw, int fshift, int pos_bits,
                                                uint32_t* bm, uint32_t* coarse, QEntry* en, uint32_t* ws, uint32_t tot[3]) {
    uint32_t* pref = bm + nw;
    const uint32_t* __restrict__ d = reinterpret_cast<const uint32_t*>(q.hdr.data);
    const uint32_t e0 = q.hdr.end[0], e1 = q.hdr.end[1], e2 = q.hdr.end[2], e3 = q.hdr.end[3],
                   e4 = q.hdr.end[4], e5 = d ? q.hdr.end[5] : 0u;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    const int tid = threadIdx.x;
    constexpr int nt = NT;
    // the row is read once: up to FN3_BUILD_KEEP words per thread stay in registers for the second pass
    uint32_t keep[FN3_BUILD_KEEP];
#pragma unroll
    for (int r = 0; r < FN3_BUILD_KEEP; ++r) { const uint32_t i = tid + r * nt; keep[r] = i < e5 ? __ldg(d + i) : 0u; }
    FN3_TR(1);
    {
        uint4* z = reinterpret_cast<uint4*>(bm);
        for (int i = tid; i < nw / 4; i += nt) z[i] = make_uint4(0u, 0u, 0u, 0u);
        uint4* zc = reinterpret_cast<uint4*>(coarse);
        for (int i = tid; i < ncw / 4; i += nt) zc[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (GLOBAL) __threadfence_block();
    __syncthreads();
    FN3_TR(2);
    const int cs = fshift - 5;                                   // coarse block = 1 << cs fine blocks
    uint8_t* cmap = reinterpret_cast<uint8_t*>(coarse);          // one BYTE per coarse block: 1 = touched
    if (tid == 0) cmap[ncw * 4 - 1] = 1;                         // the id padded fingerprints carry: never counts
    for (uint32_t i = tid, r = 0; i < e5; i += nt, ++r) {       // 1. touched blocks (fine and coarse)
        const uint32_t w = r < FN3_BUILD_KEEP ? keep[r < FN3_BUILD_KEEP ? r : 0] : __ldg(d + i);
        uint32_t b0, b1;
        if (i < e3) { b0 = b1 = w >> 5; }
        else {
            const uint32_t s = w & pmask, len = (pos_bits >= 32) ? 1u : ((w >> pos_bits) + 1u);
            b0 = s >> 5; b1 = (s + len - 1u) >> 5;
        }
        for (uint32_t blk = b0; blk <= b1; ++blk) atomicOr(&bm[blk >> 5], 1u << (blk & 31u));
        for (uint32_t cb = b0 >> cs; cb <= (b1 >> cs); ++cb) cmap[cb] = 1;      // plain byte stores: every writer writes 1
    }
    if (GLOBAL) __threadfence_block();
    __syncthreads();
    FN3_TR(3);
    // 2. prefix of touched-block counts: every thread owns `c4` consecutive groups of four words (nw is a multiple of 4;
    //    128-bit accesses a group apart per lane are conflict-free for c4 = 1, 3; the words stay in registers in between)
    constexpr int MAXC4 = 4;
    const int n4 = nw >> 2, c4 = (n4 + nt - 1) / nt;
    const int lo4 = min(n4, tid * c4), hi4 = min(n4, lo4 + c4);
    FN3_TR(32);
    uint32_t local = 0;
    uint4 own[MAXC4];
    if (!GLOBAL && c4 <= MAXC4) {
#pragma unroll
        for (int j = 0; j < MAXC4; ++j) {
            own[j] = (lo4 + j < hi4) ? reinterpret_cast<const uint4*>(bm)[lo4 + j] : make_uint4(0u, 0u, 0u, 0u);
            local += __popc(own[j].x) + __popc(own[j].y) + __popc(own[j].z) + __popc(own[j].w);
        }
    } else {
        for (int w = lo4 * 4; w < hi4 * 4; ++w) local += __popc(ldq<GLOBAL>(&bm[w]));
    }
    FN3_TR(24);
    uint32_t nE;
    uint32_t run = block_exscan_1b(local, ws, &nE);
    FN3_TR(25);
    if (!GLOBAL && c4 <= MAXC4) {
#pragma unroll
        for (int j = 0; j < MAXC4; ++j) {
            if (lo4 + j < hi4) {
                uint4 pr;
                pr.x = run; run += __popc(own[j].x);
                pr.y = run; run += __popc(own[j].y);
                pr.z = run; run += __popc(own[j].z);
                pr.w = run; run += __popc(own[j].w);
                reinterpret_cast<uint4*>(pref)[lo4 + j] = pr;
            }
        }
    } else {
        for (int w = lo4 * 4; w < hi4 * 4; ++w) { const uint32_t b = ldq<GLOBAL>(&bm[w]); pref[w] = run; run += __popc(b); }
    }
    FN3_TR(26);
    {
        uint4* z = reinterpret_cast<uint4*>(en);                 // 3. clear entries incl. the sentinel
        for (uint32_t i = tid; i < 2u * (nE + 1u); i += nt) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (GLOBAL) __threadfence_block();
    __syncthreads();
    FN3_TR(4);
    for (uint32_t i = tid, r = 0; i < e5; i += nt, ++r) {       // 4. per-position masks
        const uint32_t w = r < FN3_BUILD_KEEP ? keep[r < FN3_BUILD_KEEP ? r : 0] : __ldg(d + i);
        if (i < e3) {
            const uint32_t blk = w >> 5;
            const uint32_t bx = ldq<GLOBAL>(&bm[blk >> 5]), by = ldq<GLOBAL>(&pref[blk >> 5]);
            QEntry* t = en + (by + __popc(bx & lowmask(blk & 31u)));
            const uint32_t pb = 1u << (w & 31u);
            const uint32_t base = (i >= e0) + (i >= e1) + (i >= e2);
            atomicOr(&t->mDef, pb);
            if (base & 1u) atomicOr(&t->mB0, pb);
            if (base & 2u) atomicOr(&t->mB1, pb);
        } else {
            const uint32_t s = w & pmask, len = (pos_bits >= 32) ? 1u : ((w >> pos_bits) + 1u), e = s + len;
            for (uint32_t blk = s >> 5; blk <= ((e - 1u) >> 5); ++blk) {
                const uint32_t first = blk << 5;
                const uint32_t from = s > first ? s - first : 0u, to = min(e - first, 32u);
                const uint32_t m = (to >= 32u ? 0xffffffffu : lowmask(to)) & ~lowmask(from);
                const uint32_t bx = ldq<GLOBAL>(&bm[blk >> 5]), by = ldq<GLOBAL>(&pref[blk >> 5]);
                QEntry* t = en + (by + __popc(bx & lowmask(blk & 31u)));
                atomicOr(&t->mU, m);
                if (i >= e4) atomicOr(&t->mM, m);
            }
        }
    }
    if (GLOBAL) __threadfence_block();
    __syncthreads();
    FN3_TR(5);
    // 5. running counts.  Warp w owns a contiguous run of entries, its lanes take them INTERLEAVED (lane l: first + l,
    //    first + 32 + l, ...): neighbouring lanes read neighbouring 32-byte entries.  (One contiguous chunk per THREAD put
    //    the lanes of a warp a whole chunk apart — an 8- to 32-way bank conflict on every access; in-kernel clocks: this phase
    //    took 2.2 of the build's 4.3 us.)
    const uint32_t lane5 = tid & 31u, warp5 = tid >> 5;
    constexpr uint32_t nwarp5 = (uint32_t)NT >> 5;
    const uint32_t per_w = ((nE + nwarp5 * 32u - 1u) / (nwarp5 * 32u)) * 32u;
    const uint32_t wlo = min(nE, warp5 * per_w), whi = min(nE, wlo + per_w);
    // (in shared memory an entry is read as its two 128-bit halves: 32-bit reads of one field of 32 neighbouring entries
    //  fall into four banks)
    auto masks = [&](uint32_t i, uint32_t& d, uint32_t& u, uint32_t& m) {
        if (GLOBAL) { d = __ldcg(&en[i].mDef); u = __ldcg(&en[i].mU); m = __ldcg(&en[i].mM); }
        else {
            const uint4 a = reinterpret_cast<const uint4*>(en)[2u * i], b = reinterpret_cast<const uint4*>(en)[2u * i + 1u];
            d = a.x; u = a.y; m = b.x;
        }
    };
    uint32_t sD = 0, sU = 0, sM = 0;
    FN3_TR(30);
    for (uint32_t i = wlo + lane5; i < whi; i += 32u) {
        uint32_t d, u, m;
        masks(i, d, u, m);
        sD += __popc(d); sU += __popc(u); sM += __popc(m);
    }
    FN3_TR(31);
    sD = __reduce_add_sync(0xffffffffu, sD); sU = __reduce_add_sync(0xffffffffu, sU); sM = __reduce_add_sync(0xffffffffu, sM);
    if (lane5) { sD = 0u; sU = 0u; sM = 0u; }                    // one value per warp goes into the block scan
    FN3_TR(27);
    uint32_t tD, tU, tM;
    block_exscan3_1b(sD, sU, sM, ws + 32, tD, tU, tM);
    FN3_TR(28);
    uint32_t bD = __shfl_sync(0xffffffffu, sD, 0), bU = __shfl_sync(0xffffffffu, sU, 0), bM = __shfl_sync(0xffffffffu, sM, 0);
    for (uint32_t i0 = wlo; i0 < whi; i0 += 32u) {               // warp-uniform trip count
        const uint32_t i = i0 + lane5;
        const bool ok = i < whi;
        uint32_t d = 0u, u = 0u, m = 0u;
        if (ok) masks(i, d, u, m);
        const uint32_t a = __popc(d), b = __popc(u), c = __popc(m);
        uint32_t ia = a, ib = b, ic = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o),
                           z = __shfl_up_sync(0xffffffffu, ic, o);
            if ((int)lane5 >= o) { ia += x; ib += y; ic += z; }
        }
        if (ok) {
            if (GLOBAL) { en[i].cDef = bD + ia - a; en[i].cU = bU + ib - b; en[i].cM = bM + ic - c; }
            else reinterpret_cast<uint4*>(en)[2u * i + 1u] = make_uint4(m, bD + ia - a, bU + ib - b, bM + ic - c);
        }
        bD += __shfl_sync(0xffffffffu, ia, 31); bU += __shfl_sync(0xffffffffu, ib, 31); bM += __shfl_sync(0xffffffffu, ic, 31);
    }
    if (tid == 0) { en[nE].cDef = tD; en[nE].cU = tU; en[nE].cM = tM; }
    FN3_TR(29);
    __syncthreads();
    FN3_TR(6);
    tot[0] = tD; tot[1] = tU; tot[2] = tM;
    return nE;
}

#ifndef FN3_Q_THREADS
#define FN3_Q_THREADS 512
#endif

__global__ void __launch_bounds__(FN3_Q_THREADS) k_build_query(const QuerySlot* __restrict__ slots, const QueryState qs,
                                                               int nw, int ncw, int fshift, int pos_bits) {
    __shared__ uint32_t ws[FN3_WS_WORDS];
    const int slot = blockIdx.x;
    const QuerySlot q = slots[slot];
    uint32_t tot[3];
    const uint32_t nE = build_query<true, FN3_Q_THREADS>(q, nw, ncw, fshift, pos_bits, qs.bm + (long long)slot * 2 * nw,
                                          qs.coarse + (long long)slot * ncw, qs.ent + (long long)slot * qs.ent_stride, ws, tot);
    if (threadIdx.x == 0) qs.info[slot] = make_uint4(nE, tot[0], tot[1], tot[2]);
}

// ------------------------------------------------------------------------------------------------
// k_query — the compare kernel
// ------------------------------------------------------------------------------------------------
//
// Streaming identity (SURVEY.md §8 "Spec"), with V = definite non-reference positions,
// U = N u M positions, q = query (seq1), c = candidate (seq2):
//   S1    = #{p in V_c : code_q[p] is ref, or ACGT and != code_c[p]}         (monotone)
//   dist  = S1 + |V_q| - |V_c n V_q| - |U_c n V_q|
//   n1    = |U_q|,  n2 = |N_c| + |M_q| - |N_c n M_q|,  nboth = n1 + |N_c| - |N_c n U_q|
// S1 alone is a lower bound of dist, so a pair leaves as soon as S1 > ceiling.
//
// ONE launch per one-vs-all query.  Every CTA first builds the query's compact structure in its own SHARED
// memory straight from the query's ~2 KB row (QMODE 1) — every later lookup is an LDS, never an L2 gather.
// Queries whose structure exceeds shared memory use the copy k_build_query left in global memory (QMODE 0).
//   FILTER variant (ceiling <= 24): one THREAD per candidate reads the candidate's head from its tile — header
//     and 32 fingerprints, coalesced across the warp, one memory round trip — and counts the fingerprints that
//     fall in coarse blocks the query does not touch at all: each such position is a certain mismatch, so a
//     count above the ceiling proves "not a neighbour" (this is ~every unrelated pair).  Survivors are then
//     streamed one WARP per candidate.
//     (coalesced 128-bit loads, 256 words per warp step with the next 256 in flight, __reduce_add_sync tallies).
//   STREAM variant (larger ceilings): k_scan in fn3_scan.cuh — whole rows by bulk asynchronous copy into per-warp
//     shared-memory rings.
// Survivors are compacted into the neighbour buffer with one atomic each.

struct QView {                  // where the query structure lives (shared or global)
    const uint32_t* bm;         // bit per 32-position block: touched by the query
    const uint32_t* pref;       // touched blocks below block 32w
    const uint8_t* cmap;        // one byte per coarse (fingerprint) block: 1 = touched by the query
    const uint4* ent;           // two uint4 per entry
};

// is the 32-position block of p touched by the query?  (the common answer for an unrelated candidate is no)
__device__ __forceinline__ uint32_t touched_bit(const QView& qv, uint32_t p) {
    return (qv.bm[p >> 10] >> ((p >> 5) & 31u)) & 1u;
}

// V element at list index ii whose block IS touched: returns mismatch flag (bit0) | query-is-Def flag (bit1)
__device__ __forceinline__ uint32_t look_touched(const QView& qv, uint32_t p, uint32_t ii, uint32_t e0, uint32_t e1, uint32_t e2) {
    const uint32_t blk = p >> 5, wi = blk >> 5, bit = blk & 31u;
    const uint4 a = qv.ent[2u * (qv.pref[wi] + __popc(qv.bm[wi] & lowmask(bit)))];
    const uint32_t k = p & 31u;
    const uint32_t D = (a.x >> k) & 1u, U = (a.y >> k) & 1u;
    const uint32_t qb = ((a.z >> k) & 1u) | (((a.w >> k) & 1u) << 1);
    const uint32_t b = (ii >= e0) + (ii >= e1) + (ii >= e2);    // candidate's base code, needed only here
    const uint32_t mism = (U ^ 1u) & ((D ^ 1u) | (uint32_t)(qb != b));
    return mism | (D << 1);
}

// number of query positions of each kind strictly below x (Def, U, M)
__device__ __forceinline__ void rank3(const QView& qv, uint32_t x, uint32_t& rD, uint32_t& rU, uint32_t& rM, bool want_um) {
    const uint32_t blk = x >> 5, wi = blk >> 5, bit = blk & 31u;
    const uint32_t bits = qv.bm[wi];
    const uint32_t idx = qv.pref[wi] + __popc(bits & lowmask(bit));
    const bool touched = (bits >> bit) & 1u;
    const uint4 b = qv.ent[2u * idx + 1u];                      // {mM, cDef, cU, cM}
    rD = b.y; rU = b.z; rM = b.w;
    if (touched) {
        const uint4 a = qv.ent[2u * idx];                       // {mDef, mU, mB0, mB1}
        const uint32_t lm = lowmask(x & 31u);
        rD += __popc(a.x & lm);
        if (want_um) { rU += __popc(a.y & lm); rM += __popc(b.x & lm); }
    }
}

struct Tally { int S1, qdefV, qdefU, nNc, nNU, nNM; bool dead; };

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// One warp streams one candidate row against the query (all 32 lanes must call).
// The V words (A,C,G,T positions) and the run words (N, M) are handled by separate loops: V words 256 per warp step
// with two 128-bit loads per lane and the next step in flight; run words one per lane, the first 32 of them requested
// before the V loop starts.  (Handling both kinds in one loop made every lane of a mixed step walk through the V code
// AND the run code for each of its 8 words: ~970 warp instructions per 500-word row; profiles/README.md.)
template <bool COUNT>
__device__ __forceinline__ Tally stream_row(const QView& qv, const uint32_t* __restrict__ d, uint32_t e0, uint32_t e1,
                                            uint32_t e2, uint32_t e3, uint32_t e4, uint32_t e5, int k, int pos_bits,
                                            uint32_t pmask, int lane, unsigned long long& touched) {
    Tally t; t.S1 = 0; t.qdefV = 0; t.qdefU = 0; t.nNc = 0; t.nNU = 0; t.nNM = 0; t.dead = false;
    uint32_t r0 = 0u;
    if (e3 + (uint32_t)lane < e5) { r0 = ld_stream_u32(d + e3 + lane); if (COUNT) touched += 4; }
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
    if ((uint32_t)lane * 4u < e3) { c0 = ld_stream_v4(d + lane * 4); if (COUNT) touched += 16; }
    if ((uint32_t)lane * 4u + 128u < e3) { c1 = ld_stream_v4(d + lane * 4 + 128); if (COUNT) touched += 16; }
    for (uint32_t base = 0; base < e3; base += 256) {
        const uint32_t idx = base + lane * 4;
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
        if (idx + 256u < e3) { n0 = ld_stream_v4(d + idx + 256); if (COUNT) touched += 16; }
        if (idx + 384u < e3) { n1 = ld_stream_v4(d + idx + 384); if (COUNT) touched += 16; }
        const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        int mism = 0;
        if (base + 256u <= e3) {                                // warp-uniform: a full step
            // (a branch-free variant — all 8 bitmap tests first, touched words afterwards — measured slower on
            //  B200: more registers, spills at 64 regs/thread, 57 vs 42 us at ceiling 250; profiles/README.md)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (touched_bit(qv, w[j])) {
                    const uint32_t r = look_touched(qv, w[j], idx + (j & 3) + ((j >> 2) << 7), e0, e1, e2);
                    mism += r & 1u;
                    t.qdefV += r >> 1;
                } else mism += 1;
            }
        } else {                                                // the last, partial step
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t ii = idx + (j & 3) + ((j >> 2) << 7);
                if (ii < e3) {
                    if (touched_bit(qv, w[j])) {
                        const uint32_t r = look_touched(qv, w[j], ii, e0, e1, e2);
                        mism += r & 1u;
                        t.qdefV += r >> 1;
                    } else mism += 1;
                }
            }
        }
        t.S1 += __reduce_add_sync(0xffffffffu, mism);
        if (t.S1 > k) { t.dead = true; return t; }
        c0 = n0; c1 = n1;
    }
    // run words: [start, start+len) of N (below e4) or M; what the query holds inside each run comes from two rank lookups
    for (uint32_t rb = e3; rb < e5; rb += 32) {                 // warp-uniform
        const uint32_t ii = rb + lane;
        uint32_t x = r0;
        if (rb != e3 && ii < e5) { x = ld_stream_u32(d + ii); if (COUNT) touched += 4; }
        if (ii < e5) {
            const uint32_t s = x & pmask;
            const uint32_t len = (pos_bits >= 32) ? 1u : ((x >> pos_bits) + 1u);
            const bool isN = ii < e4;
            uint32_t aD, aU, aM, bD, bU, bM;
            rank3(qv, s, aD, aU, aM, isN);
            rank3(qv, s + len, bD, bU, bM, isN);
            t.qdefU += (int)(bD - aD);
            if (isN) { t.nNc += (int)len; t.nNU += (int)(bU - aU); t.nNM += (int)(bM - aM); }
        }
    }
    return t;
}
__device__ __forceinline__ void emit_if_close(const CompareParams& p, const Tally& t0, int k, uint32_t nVq, uint32_t nUq,
                                              uint32_t nMq, long long qrow, long long crow, int lane) {
    if (t0.dead) return;
    const int qdefV = __reduce_add_sync(0xffffffffu, t0.qdefV);
    const int qdefU = __reduce_add_sync(0xffffffffu, t0.qdefU);
    const int dist = t0.S1 + (int)nVq - qdefV - qdefU;
    if (dist > k) return;
    const int nNc = __reduce_add_sync(0xffffffffu, t0.nNc);
    const int nNU = __reduce_add_sync(0xffffffffu, t0.nNU);
    const int nNM = __reduce_add_sync(0xffffffffu, t0.nNM);
    if (lane == 0) {
        const unsigned long long at = atomicAdd(p.out_count, 1ull);
        if ((long long)at < p.out_cap) {
            EdgeRec r;
            r.row1 = qrow; r.row2 = crow; r.dist = dist;
            r.n1 = (int)nUq; r.n2 = nNc + (int)nMq - nNM; r.nboth = (int)nUq + nNc - nNU;
            p.out[at] = r;
            if (p.host_out && (long long)at < p.host_cap) p.host_out[at] = r;      // posted write over PCIe
        }
    }
}

// QMODE 0: query structure in global memory (built by k_build_query); 1: built here in shared memory
// zero-copy completion: every CTA makes its host-memory records visible, the last one publishes count + sequence
// number to the mapped host flag and re-arms the device counters (so the next query needs no memset).
__device__ __forceinline__ void signal_done(const CompareParams& p) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        const unsigned int t = atomicAdd(p.done_ctas, 1u);
        if (t == total - 1u) {
            __threadfence();
            const unsigned long long cnt = atomicExch(p.out_count, 0ull);
            const unsigned long long passed = atomicExch(p.out_count + 2, 0ull);
            *p.done_ctas = 0u;
            p.host_flag[1] = cnt;
            p.host_flag[2] = passed;
            __threadfence_system();
            p.host_flag[0] = p.seq;
        }
    }
}

#ifndef FN3_Q_MINB
#define FN3_Q_MINB 2
#endif
#ifndef FN3_SURV_CAP
#define FN3_SURV_CAP 96            // survivors a CTA queues for all its warps (beyond: streamed by the warp that found them)
#endif
template <bool COUNT, int QMODE>
__global__ void __launch_bounds__(FN3_Q_THREADS, FN3_Q_MINB) k_query(const CompareParams p) {
    extern __shared__ __align__(128) uint4 qsm[];
    __shared__ uint32_t ws[FN3_WS_WORDS];
    // survivors of the head filter, per CTA: phase A queues them, then ALL warps of the CTA draw from the queue — a warp
    // whose 32 candidates hold two relatives of the query no longer streams them one after the other while fifteen warps
    // idle (measured: the step went from 16.8 to 20.8 us once 400 appended relatives made such collisions common)
    __shared__ uint4 q_h0[FN3_SURV_CAP], q_h1[FN3_SURV_CAP];
    __shared__ int q_row[FN3_SURV_CAP];
    __shared__ unsigned int q_count, q_head;
    if (p.debug_stop == 3) return;
    FN3_TR(0); FN3_TRG(22);
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.y;
    const QuerySlot qs = p.slot_inline ? p.slot0 : p.slots[slot];
    if (threadIdx.x == 0) { q_count = 0u; q_head = 0u; }   // (ordered before their first use by the barriers of the query build /
                                                           //  the explicit one below)
    if (p.commit_row >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
        put_head(p.hdr, p.commit_row, p.commit_head);      // fn3_insert_query: the new row's head (never a candidate here)
    if (qs.hdr.data == 0ull) {                             // invalid query: no distances at all
        if (p.host_flag) signal_done(p);
        return;
    }
    if (p.debug_stop == 4) return;
    const int nw = p.nw, ncw = p.ncw;
    const int tshift = p.hdr.shift - 5;                    // tiles per chunk = 1 << tshift
    // FILTER: work unit = 32 candidates.  Consecutive units (tiles hold consecutively inserted rows, i.e. relatives)
    // go to different CTAs, so that the few surviving candidates of a family are streamed on different SMs.
    // (Requesting the first unit's heads before the query build was measured: no gain, +58 registers.)
    const uint32_t units = p.rows ? (uint32_t)((p.n_list + 31) >> 5)
                           : (uint32_t)(((p.n_rows + (1ll << p.hdr.sb_shift) - 1) >> p.hdr.sb_shift) << (p.hdr.sb_shift - 5));
    const uint32_t u0 = (uint32_t)(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
#if FN3_PF
    // the head tile this warp will test first is asked into L2 now (one bulk prefetch of its 3 072 bytes by one lane, no
    // register held, nothing in the load/store queue the build is about to fill): its way from HBM overlaps the query build
    // instead of following it (in-kernel clocks: head arrival 1.4 -> 0.7 us after the request; 24 per-line prefetch
    // instructions per warp did the same but slowed the build's shared-memory phases by as much)
    if (!p.rows) {
        const int sl0 = (int)qs.skip_le;
        const uint32_t uf = sl0 < 0 ? 0u : (((uint32_t)sl0 + 1u) >> p.hdr.sb_shift) << (p.hdr.sb_shift - 5);
        const uint32_t ust = gridDim.x * (FN3_Q_THREADS / 32);
        for (uint32_t uu = uf + u0, i = 0; i < (p.ceiling < 16 ? 2u : 1u) && uu < units; ++i, uu += ust)
            if (lane == 0)
                prefetch_l2_bulk(p.hdr.chunk[uu >> tshift] + (size_t)(uu & ((1u << tshift) - 1u)) * FN3_TILE_BYTES, FN3_TILE_BYTES);
    }
#endif
    uint32_t nVq, nUq, nMq;
    QView qv;
    if (QMODE == 1) {
        uint32_t* sm_bm = reinterpret_cast<uint32_t*>(qsm);                       // [bitmap nw | prefix nw]
        uint32_t* sm_co = reinterpret_cast<uint32_t*>(qsm + nw / 2);
        QEntry* sm_en = reinterpret_cast<QEntry*>(qsm + nw / 2 + ncw / 4);
        uint32_t tot[3];
        build_query<false, FN3_Q_THREADS>(qs, nw, ncw, p.fshift, p.pos_bits, sm_bm, sm_co, sm_en, ws, tot);
        nVq = tot[0]; nUq = tot[1]; nMq = tot[2];
        qv.bm = sm_bm; qv.pref = sm_bm + nw; qv.cmap = reinterpret_cast<const uint8_t*>(sm_co); qv.ent = reinterpret_cast<const uint4*>(sm_en);
    } else {
        const uint4 info = p.qs.info[slot];
        nVq = info.y; nUq = info.z; nMq = info.w;
        qv.bm = p.qs.bm + (long long)slot * 2 * nw;
        qv.pref = qv.bm + nw;
        qv.cmap = reinterpret_cast<const uint8_t*>(p.qs.coarse + (long long)slot * ncw);
        qv.ent = reinterpret_cast<const uint4*>(p.qs.ent + (long long)slot * p.qs.ent_stride);
    }
    if (p.debug_stop == 1) return;
    FN3_TR(7);
    if (QMODE != 1) __syncthreads();                       // the survivor queue's counters are initialised
    const int k = p.ceiling;
    const int pos_bits = p.pos_bits;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    unsigned long long touched = 0;

    {
        // upper triangle: superblocks that hold only rows <= skip_le cannot contain a candidate
        const int exclude = (int)qs.exclude, skip_le = (int)qs.skip_le;
        const uint32_t ufirst = (p.rows || skip_le < 0) ? 0u
                                : (((uint32_t)skip_le + 1u) >> p.hdr.sb_shift) << (p.hdr.sb_shift - 5);
        const uint32_t ustep = gridDim.x * (FN3_Q_THREADS / 32);
        const int sb_shift = p.hdr.sb_shift, gs = sb_shift - 5;
        const uint32_t gmask = (1u << gs) - 1u, tmask = (1u << tshift) - 1u, n_rows32 = (uint32_t)p.n_rows;
        // Unit = 32 candidates (a head tile of the full scan, or 32 strided entries of an explicit list).  All row / unit
        // arithmetic is 32-bit (a store holds at most FN3_MAX_HDR_CHUNKS << 20 = 2^25 rows): with 64-bit units and runtime
        // shifts the address computation cost as many instructions as the fingerprint tests it feeds.
        // -> this lane's candidate in unit uu: its tile, its lane in the tile, its row (-1: none, self, or below the triangle)
        auto locate = [&](uint32_t uu, const unsigned char*& tile, int& tl) -> int {
            int row = -1;
            tl = lane; tile = nullptr;
            if (p.rows) {
                const long long ci = (long long)uu + (long long)lane * units;
                if (ci < p.n_list) { row = (int)p.rows[ci]; tile = tile_of(p.hdr, row, tl); }
            } else {
                // uu = global tile index: chunk = uu >> tshift, tile inside the chunk = the low bits; row as in fn3_row_of
                const uint32_t r = ((uu >> gs) << sb_shift) + ((uint32_t)lane << gs) + (uu & gmask);
                tile = p.hdr.chunk[uu >> tshift] + (size_t)(uu & tmask) * FN3_TILE_BYTES;
                row = r < n_rows32 ? (int)r : -1;
            }
            return (row != exclude && row > skip_le) ? row : -1;
        };
        // certain mismatches among 8 fingerprints: a fingerprint whose coarse block the query does not touch at all.
        // Padded fingerprints (rows with fewer than 32 V words) carry an id whose byte is always 1; the heads of invalid
        // and removed rows carry an id whose byte is always 0, so they never pass and their header is never read.
        const uint8_t* cm = qv.cmap;
        auto count8 = [&](const uint4& g) -> int {
            const uint32_t w[4] = {g.x, g.y, g.z, g.w};
            int s = 8;
#pragma unroll
            for (int i = 0; i < 4; ++i) s -= (int)cm[w[i] & 0xffffu] + (int)cm[w[i] >> 16];
            return s;
        };
        // the groups beyond the eager ones, for the rare candidate that has not been ruled out yet
        auto rest = [&](int S, const uint4* fp, int from) -> bool {
            for (int j = from; j < 4 && S <= k; ++j) { S += count8(ld_head_v4(fp + j * 32)); if (COUNT) touched += 16; }
            return S <= k;
        };
        // ---- phase B: one warp per surviving candidate ---------------------------------------------------
        // Two candidate slots per thread, A and B (the single-candidate path passes B = nothing).  The survivor's header
        // either sits in its lane's registers already (single-candidate path: HAVE_HDR, nine shuffles) or is fetched now.
        // one survivor, its header known (h0.x | h0.y != 0) or to be fetched: stream its row, report it if it is close
        auto stream_one = [&](int crow, uint4 h0, uint4 h1, bool have_hdr) {
            if (!have_hdr) {
                int tl;
                const unsigned char* tile = tile_of(p.hdr, crow, tl);
                const uint4* hp = reinterpret_cast<const uint4*>(tile + tl * 32);
                h0 = __ldcg(hp); h1 = __ldcg(hp + 1);                      // all lanes, one address: a broadcast
                if (COUNT && lane == 0) touched += 32;
            }
            if ((h0.x | h0.y) == 0u) return;                                // removed while its fingerprints were still live
            const uint32_t* d = reinterpret_cast<const uint32_t*>(((unsigned long long)h0.y << 32) | h0.x);
            const Tally t = stream_row<COUNT>(qv, d, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, k, pos_bits, pmask, lane, touched);
            emit_if_close(p, t, k, nVq, nUq, nMq, qs.row, (long long)crow, lane);
        };
        auto survivors = [&](bool passA, int rowA, bool passB, int rowB, const bool HAVE_HDR, const uint4& ha0, const uint4& ha1) {
            uint32_t aliveA = __ballot_sync(0xffffffffu, passA), aliveB = __ballot_sync(0xffffffffu, passB);
            if ((aliveA | aliveB) == 0u) return;
            // candidates the head filter could not rule out: the host steers FILTER <-> STREAM by this rate (a few atomics per
            // launch on an ordinary collection; one per unit in an outbreak)
            if (lane == 0) atomicAdd(p.out_count + 2, (unsigned long long)(__popc(aliveA) + __popc(aliveB)));
            if (p.debug_stop == 2) return;
            // queue them for the whole CTA (phase B below); what does not fit is streamed here, as before
            {
                const uint32_t nA = __popc(aliveA), nB = __popc(aliveB);
                unsigned int base = 0u;
                if (lane == 0) base = atomicAdd(&q_count, nA + nB);
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t below = lowmask(lane);
                if (passA) {
                    const uint32_t at = base + __popc(aliveA & below);
                    if (at < FN3_SURV_CAP) {
                        q_row[at] = rowA;
                        q_h0[at] = HAVE_HDR ? ha0 : make_uint4(0u, 0u, 0u, 0u);
                        q_h1[at] = HAVE_HDR ? ha1 : make_uint4(0u, 0u, 0u, 1u);      // h1.w = 1 with a zero h0: "fetch the header"
                    }
                }
                if (passB) {
                    const uint32_t at = base + nA + __popc(aliveB & below);
                    if (at < FN3_SURV_CAP) { q_row[at] = rowB; q_h0[at] = make_uint4(0u, 0u, 0u, 0u); q_h1[at] = make_uint4(0u, 0u, 0u, 1u); }
                }
                // lanes whose slot lies beyond the queue stay alive for the inline path
                const uint32_t fitA = base >= FN3_SURV_CAP ? 0u : min(nA, FN3_SURV_CAP - base);
                const uint32_t usedA = base + nA;
                const uint32_t fitB = usedA >= FN3_SURV_CAP ? 0u : min(nB, FN3_SURV_CAP - usedA);
                for (uint32_t i = 0; i < fitA; ++i) aliveA &= aliveA - 1u;            // the first fitA set bits went to the queue
                for (uint32_t i = 0; i < fitB; ++i) aliveB &= aliveB - 1u;
            }
            while (aliveA | aliveB) {
                const bool fromA = aliveA != 0u;
                const uint32_t m = fromA ? aliveA : aliveB;
                const int src = __ffs(m) - 1;
                if (fromA) aliveA &= aliveA - 1u; else aliveB &= aliveB - 1u;
                const int crow = __shfl_sync(0xffffffffu, fromA ? rowA : rowB, src);
                uint4 h0 = make_uint4(0u, 0u, 0u, 0u), h1 = h0;
                if (HAVE_HDR) {
                    h0.x = __shfl_sync(0xffffffffu, ha0.x, src); h0.y = __shfl_sync(0xffffffffu, ha0.y, src);
                    h0.z = __shfl_sync(0xffffffffu, ha0.z, src); h0.w = __shfl_sync(0xffffffffu, ha0.w, src);
                    h1.x = __shfl_sync(0xffffffffu, ha1.x, src); h1.y = __shfl_sync(0xffffffffu, ha1.y, src);
                    h1.z = __shfl_sync(0xffffffffu, ha1.z, src); h1.w = __shfl_sync(0xffffffffu, ha1.w, src);
                    if ((h0.x | h0.y) == 0u) continue;                      // removed while its fingerprints were still live
                }
                stream_one(crow, h0, h1, HAVE_HDR);
            }
        };
        // ---- phase A: one thread per candidate, on the fingerprints of the candidate's head ---------------
        // k + 1 certain mismatches need ceil((k + 1) / 8) groups of 8 fingerprints.
        if (k < 16) {
            // low ceilings (all-vs-all at 12 SNV): the first two groups settle nearly every candidate, so they alone are
            // requested up front (one memory round trip; volatile loads, see ld_head_v4) and every thread takes TWO candidates
            // per iteration — twice the loads in flight per warp, 32 bytes per candidate instead of 96.  The other groups are
            // read only by a candidate that is still alive, the header only by a survivor.
            for (uint32_t u = ufirst + u0; u < units; u += 2u * ustep) {
                const unsigned char *tA, *tB = nullptr;
                int tlA, tlB = 0;
                const int rowA = locate(u, tA, tlA);
                const int rowB = (u + ustep < units) ? locate(u + ustep, tB, tlB) : -1;
                const uint4* fA = reinterpret_cast<const uint4*>(tA + FN3_TILE_FP_OFF + tlA * 16);
                const uint4* fB = reinterpret_cast<const uint4*>(tB + FN3_TILE_FP_OFF + tlB * 16);
                uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
                if (rowA >= 0) { a0 = ld_head_v4(fA); a1 = ld_head_v4(fA + 32); if (COUNT) touched += 32; }
                if (rowB >= 0) { b0 = ld_head_v4(fB); b1 = ld_head_v4(fB + 32); if (COUNT) touched += 32; }
                bool passA = false, passB = false;
                if (rowA >= 0) passA = rest(count8(a0) + count8(a1), fA, 2);
                if (rowB >= 0) passB = rest(count8(b0) + count8(b1), fB, 2);
                survivors(passA, rowA, passB, rowB, false, a0, a0);
            }
        } else {
            // one-vs-all at ceiling 20: one candidate per thread; its header and all 32 fingerprints travel together (one
            // round trip, and a surviving candidate is streamed without a second, dependent fetch)
            for (uint32_t u = ufirst + u0; u < units; u += ustep) {
                const unsigned char* tile;
                int tl;
                const int row = locate(u, tile, tl);
                FN3_TR(33);
                const uint4* hp = reinterpret_cast<const uint4*>(tile + tl * 32);
                const uint4* fp = reinterpret_cast<const uint4*>(tile + FN3_TILE_FP_OFF + tl * 16);
                uint4 h0 = make_uint4(0, 0, 0, 0), h1 = h0;
                bool pass = false;
                if (row >= 0) {
                    h0 = ld_head_v4(hp); h1 = ld_head_v4(hp + 1);
                    const uint4 f0 = ld_head_v4(fp), f1 = ld_head_v4(fp + 32), f2 = ld_head_v4(fp + 64), f3 = ld_head_v4(fp + 96);
                    if (COUNT) touched += 96;
                    int S = count8(f0);
                    FN3_TR(34);
                    S += count8(f1) + count8(f2);
                    if (S <= k) S += count8(f3);
                    pass = S <= k;
                }
                FN3_TR(8);
                survivors(pass, row, false, -1, true, h0, h1);
                FN3_TR(9);
            }
        }
        // ---- phase B: every warp of the CTA draws queued survivors ------------------------------------------
        FN3_TR(10);
        __syncthreads();
        FN3_TR(11);
        const unsigned int n_q = min(q_count, (unsigned int)FN3_SURV_CAP);
        for (;;) {
            unsigned int i = 0u;
            if (lane == 0) i = atomicAdd(&q_head, 1u);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= n_q) break;
            const uint4 h0 = q_h0[i], h1 = q_h1[i];
            stream_one(q_row[i], h0, h1, (h0.x | h0.y) != 0u);
        }
    }
    FN3_TR(12); FN3_TRG(23);
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, o);
        if (lane == 0) atomicAdd(p.touched, touched);
    }
    if (p.host_flag) signal_done(p);
}

// ------------------------------------------------------------------------------------------------
// background-aware heads (fn3_refresh_heads)
// ------------------------------------------------------------------------------------------------
// A head carries 32 of the row's V positions as fingerprints.  WHICH 32 does not matter for correctness (every V position of
// the candidate that falls in a block the query does not touch is a certain mismatch) but it decides the filter's power: two
// unrelated isolates of one lineage agree on most of their lineage-defining positions, so the 32 lowest coordinates of a real
// collection are mostly SHARED and rule nothing out.  The engine therefore keeps a bitmap of BACKGROUND positions — non-reference
// in at least a given fraction of the stored rows — and a head takes the row's first 32 positions that are NOT background
// (its private variants), topped up with background ones only when it has fewer.

// bit p of bg = the four base planes of the vote (fn3_cluster.cuh) add up to at least thr at position p
__global__ void k_bg_bits(const uint32_t* __restrict__ cnt, long long Lpad, long long L, uint32_t thr, uint32_t* __restrict__ bg,
                          long long n_words) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t bits = 0u;
    for (int k = 0; k < 32; ++k) {
        const long long pos = w * 32 + k;
        if (pos >= L) break;
        const uint32_t t = cnt[pos] + cnt[Lpad + pos] + cnt[2 * Lpad + pos] + cnt[3 * Lpad + pos];
        if (t >= thr) bits |= 1u << k;
    }
    bg[w] = bits;
}

// one warp per row: rewrite the 32 fingerprints of the row's head from its V words and the background bitmap
__global__ void __launch_bounds__(256) k_rehead(const HdrTable ht, long long n_rows, const uint32_t* __restrict__ bg, int fshift,
                                                uint32_t fp_pad) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    int tl;
    unsigned char* tile = const_cast<unsigned char*>(tile_of(ht, row, tl));
    const uint4* hp = reinterpret_cast<const uint4*>(tile + tl * 32);
    const uint4 h0 = __ldcg(hp), h1 = __ldcg(hp + 1);
    if ((h0.x | h0.y) == 0u) return;                        // invalid / removed rows keep their always-mismatching fingerprints
    const uint32_t* __restrict__ d = reinterpret_cast<const uint32_t*>(((unsigned long long)h0.y << 32) | h0.x);
    const uint32_t e3 = h1.y;
    unsigned char* fpb = tile + FN3_TILE_FP_OFF + tl * 16;  // fingerprint f lives at fpb + (f / 8) * 512 + (f % 8) * 2
    uint32_t k = 0;
    for (int pass = 0; pass < 2 && k < FN3_HEAD_FP; ++pass)
        for (uint32_t base = 0; base < e3 && k < FN3_HEAD_FP; base += 32) {
            const uint32_t i = base + lane;
            uint32_t w = 0u;
            bool want = false;
            if (i < e3) {
                w = __ldg(d + i);
                const bool isbg = (__ldg(bg + (w >> 5)) >> (w & 31u)) & 1u;
                want = pass == 0 ? !isbg : isbg;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, want);
            const uint32_t slot = k + __popc(m & lowmask(lane));
            if (want && slot < FN3_HEAD_FP)
                *reinterpret_cast<uint16_t*>(fpb + (slot >> 3) * 512 + (slot & 7u) * 2) = (uint16_t)(w >> fshift);
            k += __popc(m);
        }
    if (k < FN3_HEAD_FP && (uint32_t)lane >= k)
        *reinterpret_cast<uint16_t*>(fpb + ((uint32_t)lane >> 3) * 512 + ((uint32_t)lane & 7u) * 2) = (uint16_t)fp_pad;
}

// ------------------------------------------------------------------------------------------------
// device-side compaction (fn3_compact): live rows are copied into fresh chunks, heads rebuilt
// ------------------------------------------------------------------------------------------------
struct RowCopy { unsigned long long src, dst; uint32_t words; uint32_t pad; };     // words: multiple of 4

__global__ void __launch_bounds__(256) k_copy_rows(const RowCopy* __restrict__ list, long long n) {
    const int lane = threadIdx.x & 31;
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += (long long)gridDim.x * (blockDim.x >> 5)) {
        const RowCopy c = list[i];
        const uint4* s = reinterpret_cast<const uint4*>(c.src);
        uint4* d = reinterpret_cast<uint4*>(c.dst);
        for (uint32_t k = lane; k < c.words / 4; k += 32) d[k] = __ldcs(s + k);
    }
}

// headers of rows 0..n-1 into their tiles; every fingerprint starts as "pad" (valid rows: k_rehead fills them in) or as
// the always-mismatching id (invalid rows)
__global__ void k_place_hdrs(const RowHdr* __restrict__ hdrs, long long n, const HdrTable ht, uint32_t fp_dead, uint32_t fp_pad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RowHead h;
    h.hdr = hdrs[i];
    const uint16_t f = (uint16_t)(h.hdr.data ? fp_pad : fp_dead);
    for (int k = 0; k < FN3_HEAD_FP; ++k) h.fp[k] = f;
    put_head(ht, i, h);
}

// L2 flush helper for the measurement hook
__global__ void k_fill(uint4* p, long long n, uint32_t v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_uint4(v, v, v, v);
}
