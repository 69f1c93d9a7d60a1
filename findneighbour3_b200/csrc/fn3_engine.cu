// fn3_engine.cu — B200 (sm_100a) SNV-distance engine: HBM-resident sample store, query-map
// build, one-vs-N / tiled all-vs-all compare kernels, and the C-ABI of include/fn3.h.
//
// What it replaces in the reference (davidhwyllie/findNeighbour3, src/seqComparer.py):
//   store            seqProfile dict + persist/remove               :108, :119-134
//   compare          countDifferences (set algebra per base)        :424-458
//   pair statistics  countDifferences_byKey + _setStats             :382-421, :354-380
//   one-vs-all       mcompare                                       :149-172
//   all-vs-all       distmat                                        :174-197
//
// Layout in HBM:
//   row data   : 32-bit words, 16-byte aligned per row:  [A | C | G | T | Nruns | Mruns]
//                A..T = sorted int32 positions; runs = (len-1) << pos_bits | start
//   row header : 32 B (one sector): device pointer + 6 cumulative list ends
//   query      : compact, staged in SHARED memory by every compare CTA: bitmap of touched 32-position
//                blocks + prefix counts (uint2 per 1024 positions) and one 32-byte entry per touched block
//                {maskDef, maskU, maskB0, maskB1, maskM, cumDef, cumU, cumM}
//
// No CPU path exists in this file: every distance is produced by k_compare on the device.

#include "fn3.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------
// device-side types
// ------------------------------------------------------------------------------------------

struct __align__(32) RowHdr {
    unsigned long long data;   // device address of the row's first word; 0 = invalid or removed
    uint32_t end[6];           // cumulative ends (in words) of A,C,G,T,Nruns,Mruns
};
static_assert(sizeof(RowHdr) == 32, "RowHdr must be one 32-byte sector");

// What the header table holds per row: the header plus a copy of the row's first V words, so that the
// early-exit filter needs ONE contiguous 160-byte read per candidate and no dependent second access.
#define FN3_HEAD_WORDS 32
struct __align__(32) RowHead {
    RowHdr hdr;
    uint32_t first[FN3_HEAD_WORDS];   // row words 0 .. min(end[3], 32)-1 (A,C,G,T positions), zero padded
};
static_assert(sizeof(RowHead) == 160, "RowHead must be 160 bytes");

// The query, compacted.  Positions are grouped in blocks of 32.  `bmpref[w]` holds, for
// blocks 32w..32w+31, .x = bit per block "the query is non-reference / N / M somewhere in this block" and
// .y = number of such (touched) blocks below block 32w.  Entry i describes the i-th touched block:
struct __align__(32) QEntry {
    uint32_t mDef, mU, mB0, mB1;   // per position: ACGT-nonref / N-or-M / base code bit 0 / bit 1 (A0 C1 G2 T3)
    uint32_t mM, cDef, cU, cM;     // per position: M ; number of Def / U / M positions below this block
};
static_assert(sizeof(QEntry) == 32, "QEntry must be one 32-byte sector");

struct __align__(32) EdgeRec {   // device-side survivor record (== fn3_edge)
    long long row1, row2;
    int dist, n1, n2, nboth;
};
static_assert(sizeof(EdgeRec) == sizeof(fn3_edge), "EdgeRec/fn3_edge mismatch");

#define FN3_MAX_HDR_CHUNKS 128
#define FN3_MAX_SLOTS 16

struct HdrTable {
    const RowHead* chunk[FN3_MAX_HDR_CHUNKS];
    int shift;                 // rows per chunk = 1 << shift
};

struct QuerySlot {             // one query of a tile
    RowHdr hdr;                // where the query's own row lives (stored row or scratch)
    long long row;             // store row of the query, -1 if not stored
    long long skip_le;         // candidates with row <= skip_le are skipped (upper triangle), -1 = none
    long long exclude;         // candidate row to skip (self), -1 = none
    long long pad;
};

struct QueryState {            // device arrays holding the compact form of up to n_slots queries
    uint2* bmpref;             // slot s at bmpref + s*nw
    QEntry* ent;               // slot s at ent + s*ent_stride
    uint4* info;               // slot s: {nE, |V_q|, |U_q|, |M_q|}
    uint32_t* bm_scratch;      // slot s at bm_scratch + s*2*nw (build kernel, only when the bitmap exceeds smem)
    long long ent_stride;
    int nw;                    // 32-block words: ceil(n_blk / 32)
    int n_blk;                 // blocks incl. the end sentinel block
};

struct CompareParams {
    HdrTable hdr;
    QueryState qs;
    const long long* rows;     // optional explicit candidate list
    long long row_begin;       // else candidates row_begin .. row_begin+n-1
    long long n;
    const QuerySlot* slots;    // device array [n_slots]
    EdgeRec* out;
    unsigned long long* out_count;
    unsigned long long* touched;   // instrumented variant only
    long long out_cap;
    int ceiling;
    int pos_bits;
    int filter_len;            // V words examined per thread by the early-exit filter (0 = no filter)
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint4 ld_stream_v4(const uint32_t* p) {
    // candidate row data is read once per comparison: do not let it displace anything in L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t lowmask(uint32_t k) { return (1u << k) - 1u; }    // k in 0..31

// block-wide exclusive scan of one value per thread (blockDim.x <= 1024, multiple of 32); returns the
// exclusive prefix, *total receives the block total.  `ws` = 33 words of shared scratch.
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

// ------------------------------------------------------------------------------------------
// query build: ONE CTA turns the query's own row into {bmpref, entries} — no genome-sized pass
// ------------------------------------------------------------------------------------------

#define FN3_BUILD_THREADS 1024

template <bool BM_SMEM>
__global__ void __launch_bounds__(FN3_BUILD_THREADS) k_build_query(const QuerySlot* __restrict__ slots,
                                                                    const QueryState qs, int pos_bits) {
    extern __shared__ uint32_t dyn[];
    __shared__ uint32_t ws[33];
    const int slot = blockIdx.x;
    const QuerySlot& q = slots[slot];
    const int nw = qs.nw;
    uint2* __restrict__ bmpref = qs.bmpref + (long long)slot * nw;
    QEntry* __restrict__ ent = qs.ent + (long long)slot * qs.ent_stride;
    uint32_t* bm = BM_SMEM ? dyn : qs.bm_scratch + (long long)slot * 2 * nw;      // touched-block bitmap
    uint32_t* pf = bm + nw;                                                       // prefix per word
    const uint32_t* __restrict__ d = reinterpret_cast<const uint32_t*>(q.hdr.data);
    const uint32_t e0 = q.hdr.end[0], e1 = q.hdr.end[1], e2 = q.hdr.end[2], e3 = q.hdr.end[3],
                   e4 = q.hdr.end[4], e5 = d ? q.hdr.end[5] : 0u;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    const int tid = threadIdx.x;

    for (int w = tid; w < nw; w += FN3_BUILD_THREADS) bm[w] = 0u;
    __syncthreads();
    // 1. which blocks does the query touch
    for (uint32_t i = tid; i < e5; i += FN3_BUILD_THREADS) {
        const uint32_t w = d[i];
        if (i < e3) {
            const uint32_t blk = w >> 5;
            atomicOr(&bm[blk >> 5], 1u << (blk & 31u));
        } else {
            const uint32_t s = w & pmask, len = (pos_bits >= 32) ? 1u : ((w >> pos_bits) + 1u);
            for (uint32_t blk = s >> 5; blk <= ((s + len - 1u) >> 5); ++blk) atomicOr(&bm[blk >> 5], 1u << (blk & 31u));
        }
    }
    __threadfence_block();
    __syncthreads();
    // 2. prefix of touched-block counts per word -> bmpref
    const int chunk = (nw + FN3_BUILD_THREADS - 1) / FN3_BUILD_THREADS;
    const int lo = min(nw, tid * chunk), hi = min(nw, lo + chunk);
    uint32_t local = 0;
    for (int w = lo; w < hi; ++w) local += __popc(bm[w]);
    uint32_t nE;
    uint32_t run = block_exscan(local, ws, &nE);
    for (int w = lo; w < hi; ++w) { const uint32_t b = bm[w]; pf[w] = run; bmpref[w] = make_uint2(b, run); run += __popc(b); }
    // 3. clear the entries (incl. the sentinel)
    {
        uint4* z = reinterpret_cast<uint4*>(ent);
        for (uint32_t i = tid; i < 2u * (nE + 1u); i += FN3_BUILD_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __threadfence_block();
    __syncthreads();
    // 4. per-position masks
    for (uint32_t i = tid; i < e5; i += FN3_BUILD_THREADS) {
        const uint32_t w = d[i];
        if (i < e3) {
            const uint32_t blk = w >> 5, wi = blk >> 5, bit = blk & 31u;
            QEntry* en = ent + (pf[wi] + __popc(bm[wi] & lowmask(bit)));
            const uint32_t pb = 1u << (w & 31u);
            const uint32_t base = (i >= e0) + (i >= e1) + (i >= e2);
            atomicOr(&en->mDef, pb);
            if (base & 1u) atomicOr(&en->mB0, pb);
            if (base & 2u) atomicOr(&en->mB1, pb);
        } else {
            const uint32_t s = w & pmask, len = (pos_bits >= 32) ? 1u : ((w >> pos_bits) + 1u), e = s + len;
            for (uint32_t blk = s >> 5; blk <= ((e - 1u) >> 5); ++blk) {
                const uint32_t b0 = blk << 5;
                const uint32_t from = s > b0 ? s - b0 : 0u, to = min(e - b0, 32u);        // [from,to) inside the block
                const uint32_t m = (to >= 32u ? 0xffffffffu : lowmask(to)) & ~lowmask(from);
                const uint32_t wi = blk >> 5, bit = blk & 31u;
                QEntry* en = ent + (pf[wi] + __popc(bm[wi] & lowmask(bit)));
                atomicOr(&en->mU, m);
                if (i >= e4) atomicOr(&en->mM, m);
            }
        }
    }
    __threadfence_block();
    __syncthreads();
    // 5. running counts over the entries; the sentinel entry nE receives the totals
    const uint32_t echunk = (nE + FN3_BUILD_THREADS - 1) / FN3_BUILD_THREADS;
    const uint32_t elo = min(nE, tid * echunk), ehi = min(nE, elo + echunk);
    uint32_t sD = 0, sU = 0, sM = 0;
    for (uint32_t i = elo; i < ehi; ++i) {
        sD += __popc(__ldcg(&ent[i].mDef)); sU += __popc(__ldcg(&ent[i].mU)); sM += __popc(__ldcg(&ent[i].mM));
    }
    uint32_t tD, tU, tM;
    uint32_t rD = block_exscan(sD, ws, &tD), rU = block_exscan(sU, ws, &tU), rM = block_exscan(sM, ws, &tM);
    for (uint32_t i = elo; i < ehi; ++i) {
        const uint32_t a = __popc(__ldcg(&ent[i].mDef)), b = __popc(__ldcg(&ent[i].mU)), c = __popc(__ldcg(&ent[i].mM));
        ent[i].cDef = rD; ent[i].cU = rU; ent[i].cM = rM;
        rD += a; rU += b; rM += c;
    }
    if (tid == 0) {
        ent[nE].cDef = tD; ent[nE].cU = tU; ent[nE].cM = tM;
        qs.info[slot] = make_uint4(nE, tD, tU, tM);
    }
}

// ------------------------------------------------------------------------------------------
// the compare kernel (k_query)
// ------------------------------------------------------------------------------------------
//
// Streaming identity (SURVEY.md §8 "Spec"), with V = definite non-reference positions,
// U = N u M positions, q = query (seq1), c = candidate (seq2):
//   S1    = #{p in V_c : code_q[p] is ref, or ACGT and != code_c[p]}         (monotone)
//   dist  = S1 + |V_q| - |V_c n V_q| - |U_c n V_q|
//   n1    = |U_q|,  n2 = |N_c| + |M_q| - |N_c n M_q|,  nboth = n1 + |N_c| - |N_c n U_q|
// S1 alone is a lower bound of dist, so a pair leaves as soon as S1 > ceiling.
//
// ONE launch per one-vs-all query.  Every CTA first builds the query's compact structure in its own
// SHARED memory straight from the query's row (QMODE 1: ~1 us of block-local work, no global traffic
// beyond re-reading the ~4 KB row) — every later lookup is an LDS, never an L2 gather.  Queries whose
// structure exceeds shared memory use the copy k_build_query left in global memory (QMODE 0).
//   FILTER variant (small ceilings): one THREAD per candidate reads the candidate's 160-byte row head
//     (header + first 32 V words, contiguous, L2-resident across queries); a candidate whose first words
//     already give S1 > ceiling is done — that is ~every unrelated pair.  Survivors are then streamed
//     one WARP per candidate.  Lane l of group g takes candidate g + l*G so that consecutive rows
//     (relatives are inserted together) land in different warps.
//   STREAM variant (large ceilings): one WARP per candidate from the start, next header prefetched.
// Row streaming: coalesced 128-bit loads, next chunk in flight while the current one is tallied,
// __reduce_add_sync tallies, survivors compacted into the neighbour buffer with one atomic each.

struct QView {                  // where the query structure lives (shared or global)
    const uint2* bmpref;
    const uint4* ent;           // two uint4 per entry
};

// V element: returns mismatch flag (bit0) | query-is-Def flag (bit1)
__device__ __forceinline__ uint32_t look_v(const QView& qv, uint32_t p, uint32_t b) {
    const uint32_t blk = p >> 5;
    const uint2 bp = qv.bmpref[blk >> 5];
    const uint32_t bit = blk & 31u;
    if (!((bp.x >> bit) & 1u)) return 1u;                       // query is reference all over this block
    const uint4 a = qv.ent[2u * (bp.y + __popc(bp.x & lowmask(bit)))];
    const uint32_t k = p & 31u;
    const uint32_t D = (a.x >> k) & 1u, U = (a.y >> k) & 1u;
    const uint32_t qb = ((a.z >> k) & 1u) | (((a.w >> k) & 1u) << 1);
    const uint32_t mism = (U ^ 1u) & ((D ^ 1u) | (uint32_t)(qb != b));
    return mism | (D << 1);
}

// number of query positions of each kind strictly below x (Def, U, M)
__device__ __forceinline__ void rank3(const QView& qv, uint32_t x, uint32_t& rD, uint32_t& rU, uint32_t& rM, bool want_um) {
    const uint32_t blk = x >> 5;
    const uint2 bp = qv.bmpref[blk >> 5];
    const uint32_t bit = blk & 31u;
    const uint32_t idx = bp.y + __popc(bp.x & lowmask(bit));
    const bool touched = (bp.x >> bit) & 1u;
    const uint4 b = qv.ent[2u * idx + 1u];                      // {mM, cDef, cU, cM}
    rD = b.y; rU = b.z; rM = b.w;
    if (touched) {
        const uint4 a = qv.ent[2u * idx];                       // {mDef, mU, mB0, mB1}
        const uint32_t lm = lowmask(x & 31u);
        rD += __popc(a.x & lm);
        if (want_um) { rU += __popc(a.y & lm); rM += __popc(b.x & lm); }
    }
}

// three exclusive block scans sharing their barriers; ws = 3*33 words
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

#define FN3_BUILD_KEEP 2
// Build the query's compact structure in shared memory from its row.  sm_bp: nw uint2, sm_en: (cap+1) QEntry.
// Returns totals (|V_q|,|U_q|,|M_q|) in tot[3].  All threads of the CTA must call.
__device__ __forceinline__ void build_query_smem(const QuerySlot& q, int nw, int pos_bits, uint2* sm_bp, QEntry* sm_en,
                                                 uint32_t* ws, uint32_t tot[3]) {
    const uint32_t* __restrict__ d = reinterpret_cast<const uint32_t*>(q.hdr.data);
    const uint32_t e0 = q.hdr.end[0], e1 = q.hdr.end[1], e2 = q.hdr.end[2], e3 = q.hdr.end[3],
                   e4 = q.hdr.end[4], e5 = q.hdr.end[5];
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    const int tid = threadIdx.x, nt = blockDim.x;
    {
        uint4* z = reinterpret_cast<uint4*>(sm_bp);
        for (int i = tid; i < nw / 2; i += nt) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // the row is read once: up to FN3_BUILD_KEEP words per thread stay in registers for pass 4
    uint32_t keep[FN3_BUILD_KEEP];
#pragma unroll
    for (int r = 0; r < FN3_BUILD_KEEP; ++r) { const uint32_t i = tid + r * nt; keep[r] = i < e5 ? __ldg(d + i) : 0u; }
    __syncthreads();
    for (uint32_t i = tid, r = 0; i < e5; i += nt, ++r) {       // 1. touched blocks
        const uint32_t w = r < FN3_BUILD_KEEP ? keep[r < FN3_BUILD_KEEP ? r : 0] : __ldg(d + i);
        if (i < e3) {
            const uint32_t blk = w >> 5;
            atomicOr(&sm_bp[blk >> 5].x, 1u << (blk & 31u));
        } else {
            const uint32_t s = w & pmask, len = (pos_bits >= 32) ? 1u : ((w >> pos_bits) + 1u);
            for (uint32_t blk = s >> 5; blk <= ((s + len - 1u) >> 5); ++blk) atomicOr(&sm_bp[blk >> 5].x, 1u << (blk & 31u));
        }
    }
    __syncthreads();
    const int chunk = (nw + nt - 1) / nt;                        // 2. prefix of touched-block counts
    const int lo = min(nw, tid * chunk), hi = min(nw, lo + chunk);
    uint32_t local = 0;
    for (int w = lo; w < hi; ++w) local += __popc(sm_bp[w].x);
    uint32_t nE;
    uint32_t run = block_exscan(local, ws, &nE);
    for (int w = lo; w < hi; ++w) { const uint32_t b = sm_bp[w].x; sm_bp[w].y = run; run += __popc(b); }
    {
        uint4* z = reinterpret_cast<uint4*>(sm_en);              // 3. clear entries incl. the sentinel
        for (uint32_t i = tid; i < 2u * (nE + 1u); i += nt) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (uint32_t i = tid, r = 0; i < e5; i += nt, ++r) {       // 4. per-position masks
        const uint32_t w = r < FN3_BUILD_KEEP ? keep[r < FN3_BUILD_KEEP ? r : 0] : __ldg(d + i);
        if (i < e3) {
            const uint32_t blk = w >> 5;
            const uint2 bp = sm_bp[blk >> 5];
            QEntry* en = sm_en + (bp.y + __popc(bp.x & lowmask(blk & 31u)));
            const uint32_t pb = 1u << (w & 31u);
            const uint32_t base = (i >= e0) + (i >= e1) + (i >= e2);
            atomicOr(&en->mDef, pb);
            if (base & 1u) atomicOr(&en->mB0, pb);
            if (base & 2u) atomicOr(&en->mB1, pb);
        } else {
            const uint32_t s = w & pmask, len = (pos_bits >= 32) ? 1u : ((w >> pos_bits) + 1u), e = s + len;
            for (uint32_t blk = s >> 5; blk <= ((e - 1u) >> 5); ++blk) {
                const uint32_t b0 = blk << 5;
                const uint32_t from = s > b0 ? s - b0 : 0u, to = min(e - b0, 32u);
                const uint32_t m = (to >= 32u ? 0xffffffffu : lowmask(to)) & ~lowmask(from);
                const uint2 bp = sm_bp[blk >> 5];
                QEntry* en = sm_en + (bp.y + __popc(bp.x & lowmask(blk & 31u)));
                atomicOr(&en->mU, m);
                if (i >= e4) atomicOr(&en->mM, m);
            }
        }
    }
    __syncthreads();
    const uint32_t echunk = (nE + nt - 1) / nt;                  // 5. running counts
    const uint32_t elo = min(nE, tid * echunk), ehi = min(nE, elo + echunk);
    uint32_t sD = 0, sU = 0, sM = 0;
    for (uint32_t i = elo; i < ehi; ++i) { sD += __popc(sm_en[i].mDef); sU += __popc(sm_en[i].mU); sM += __popc(sm_en[i].mM); }
    uint32_t tD, tU, tM;
    block_exscan3(sD, sU, sM, ws, tD, tU, tM);
    for (uint32_t i = elo; i < ehi; ++i) {
        const uint32_t a = __popc(sm_en[i].mDef), b = __popc(sm_en[i].mU), c = __popc(sm_en[i].mM);
        sm_en[i].cDef = sD; sm_en[i].cU = sU; sm_en[i].cM = sM;
        sD += a; sU += b; sM += c;
    }
    if (tid == 0) { sm_en[nE].cDef = tD; sm_en[nE].cU = tU; sm_en[nE].cM = tM; }
    __syncthreads();
    tot[0] = tD; tot[1] = tU; tot[2] = tM;
}

struct Tally { int S1, qdefV, qdefU, nNc, nNU, nNM; bool dead; };

// One warp streams one candidate row against the query (all 32 lanes must call).
template <bool COUNT>
__device__ __forceinline__ Tally stream_row(const QView& qv, const uint32_t* __restrict__ d, uint32_t e0, uint32_t e1,
                                            uint32_t e2, uint32_t e3, uint32_t e4, uint32_t e5, int k, int pos_bits,
                                            uint32_t pmask, int lane, unsigned long long& touched) {
    Tally t; t.S1 = 0; t.qdefV = 0; t.qdefU = 0; t.nNc = 0; t.nNU = 0; t.nNM = 0; t.dead = false;
    // 256 words per warp step (two 128-bit loads per lane), the next 256 already in flight
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
    if ((uint32_t)lane * 4u < e5) { c0 = ld_stream_v4(d + lane * 4); if (COUNT) touched += 16; }
    if ((uint32_t)lane * 4u + 128u < e5) { c1 = ld_stream_v4(d + lane * 4 + 128); if (COUNT) touched += 16; }
    for (uint32_t base = 0; base < e5; base += 256) {
        const uint32_t idx = base + lane * 4;
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
        if (idx + 256u < e5) { n0 = ld_stream_v4(d + idx + 256); if (COUNT) touched += 16; }
        if (idx + 384u < e5) { n1 = ld_stream_v4(d + idx + 384); if (COUNT) touched += 16; }
        const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        int mism = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t ii = idx + (j & 3) + ((j >> 2) << 7);
            if (ii < e3) {
                const uint32_t r = look_v(qv, w[j], (ii >= e0) + (ii >= e1) + (ii >= e2));
                mism += r & 1u;
                t.qdefV += r >> 1;
            } else if (ii < e5) {
                const uint32_t s = w[j] & pmask;
                const uint32_t len = (pos_bits >= 32) ? 1u : ((w[j] >> pos_bits) + 1u);
                const bool isN = ii < e4;
                uint32_t aD, aU, aM, bD, bU, bM;
                rank3(qv, s, aD, aU, aM, isN);
                rank3(qv, s + len, bD, bU, bM, isN);
                t.qdefU += (int)(bD - aD);
                if (isN) { t.nNc += (int)len; t.nNU += (int)(bU - aU); t.nNM += (int)(bM - aM); }
            }
        }
        if (base < e3) {                                        // warp-uniform: this step touched V words
            t.S1 += __reduce_add_sync(0xffffffffu, mism);
            if (t.S1 > k) { t.dead = true; break; }
        }
        c0 = n0; c1 = n1;
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
        }
    }
}

#define FN3_Q_THREADS 512

// QMODE 0: query structure in global memory (built by k_build_query); 1: built here in shared memory
template <bool COUNT, int QMODE, bool FILTER>
__global__ void __launch_bounds__(FN3_Q_THREADS, 2) k_query(const CompareParams p) {
    extern __shared__ __align__(128) uint4 qsm[];
    __shared__ uint32_t ws[99];
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.y;
    const QuerySlot qs = p.slots[slot];
    if (qs.hdr.data == 0ull) return;                       // invalid query: no distances at all
    const int nw = p.qs.nw;
    uint32_t nVq, nUq, nMq;
    QView qv;
    if (QMODE == 1) {
        uint2* sm_bp = reinterpret_cast<uint2*>(qsm);
        QEntry* sm_en = reinterpret_cast<QEntry*>(qsm + nw / 2);
        uint32_t tot[3];
        build_query_smem(qs, nw, p.pos_bits, sm_bp, sm_en, ws, tot);
        nVq = tot[0]; nUq = tot[1]; nMq = tot[2];
        qv.bmpref = sm_bp;
        qv.ent = reinterpret_cast<const uint4*>(sm_en);
    } else {
        const uint4 info = p.qs.info[slot];
        nVq = info.y; nUq = info.z; nMq = info.w;
        qv.bmpref = p.qs.bmpref + (long long)slot * nw;
        qv.ent = reinterpret_cast<const uint4*>(p.qs.ent + (long long)slot * p.qs.ent_stride);
    }
    const int k = p.ceiling;
    const int pos_bits = p.pos_bits;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    const long long hmask = (1ll << p.hdr.shift) - 1;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    unsigned long long touched = 0;

    if (FILTER) {
        const long long G = (p.n + 31) >> 5;               // candidate groups of 32
        for (long long g = warp0; g < G; g += nwarps) {
            // ---- phase A: one thread per candidate, on the candidate's row head ---------------
            const long long ci = g + (long long)lane * G;
            long long row = -1;
            uint4 h0 = make_uint4(0, 0, 0, 0), h1 = make_uint4(0, 0, 0, 0);
            bool pass = false;
            if (ci < p.n) {
                row = p.rows ? p.rows[ci] : p.row_begin + ci;
                if (row != qs.exclude && row > qs.skip_le) {
                    const uint4* hp = reinterpret_cast<const uint4*>(p.hdr.chunk[row >> p.hdr.shift] + (row & hmask));
                    // header and head words are requested together: ONE memory round trip per candidate
                    h0 = __ldg(hp); h1 = __ldg(hp + 1);
                    uint4 f[FN3_HEAD_WORDS / 4];
#pragma unroll
                    for (int j = 0; j < FN3_HEAD_WORDS / 4; ++j) {
                        f[j] = make_uint4(0, 0, 0, 0);
                        if (4 * j < p.filter_len) { f[j] = __ldg(hp + 2 + j); if (COUNT) touched += 16; }
                    }
                    if (COUNT) touched += 32;
                    if ((h0.x | h0.y) != 0u) {             // data pointer 0: invalid or removed candidate
                        const uint32_t e0 = h0.z, e1 = h0.w, e2 = h1.x, e3 = h1.y;
                        const uint32_t lim = min(e3, (uint32_t)p.filter_len);
                        int S = 0;
#pragma unroll
                        for (int j = 0; j < FN3_HEAD_WORDS / 4; ++j) {
                            const uint32_t w[4] = {f[j].x, f[j].y, f[j].z, f[j].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t ii = 4 * j + i;
                                if (ii < lim && S <= k) S += look_v(qv, w[i], (ii >= e0) + (ii >= e1) + (ii >= e2)) & 1u;
                            }
                        }
                        pass = S <= k;
                    }
                }
            }
            // ---- phase B: one warp per surviving candidate -----------------------------------
            uint32_t alive = __ballot_sync(0xffffffffu, pass);
            while (alive) {
                const int src = __ffs(alive) - 1;
                alive &= alive - 1u;
                const uint32_t dlo = __shfl_sync(0xffffffffu, h0.x, src), dhi = __shfl_sync(0xffffffffu, h0.y, src);
                const uint32_t e0 = __shfl_sync(0xffffffffu, h0.z, src), e1 = __shfl_sync(0xffffffffu, h0.w, src);
                const uint32_t e2 = __shfl_sync(0xffffffffu, h1.x, src), e3 = __shfl_sync(0xffffffffu, h1.y, src);
                const uint32_t e4 = __shfl_sync(0xffffffffu, h1.z, src), e5 = __shfl_sync(0xffffffffu, h1.w, src);
                const long long crow = __shfl_sync(0xffffffffu, row, src);
                const uint32_t* d = reinterpret_cast<const uint32_t*>(((unsigned long long)dhi << 32) | dlo);
                const Tally t = stream_row<COUNT>(qv, d, e0, e1, e2, e3, e4, e5, k, pos_bits, pmask, lane, touched);
                emit_if_close(p, t, k, nVq, nUq, nMq, qs.row, crow, lane);
            }
        }
    } else {
        // one warp per candidate; the next candidate's header is requested before the current row is streamed
        long long i = warp0;
        long long row = -1;
        uint4 h0 = make_uint4(0, 0, 0, 0), h1 = make_uint4(0, 0, 0, 0);
        auto fetch = [&](long long ci, long long& r, uint4& a, uint4& b) {
            a = make_uint4(0, 0, 0, 0); b = a; r = -1;
            if (ci < p.n) {
                r = p.rows ? p.rows[ci] : p.row_begin + ci;
                if (r != qs.exclude && r > qs.skip_le) {
                    const uint4* hp = reinterpret_cast<const uint4*>(p.hdr.chunk[r >> p.hdr.shift] + (r & hmask));
                    a = __ldg(hp); b = __ldg(hp + 1);
                    if (COUNT && lane == 0) touched += 32;
                }
            }
        };
        fetch(i, row, h0, h1);
        while (i < p.n) {
            long long nrow; uint4 n0, n1;
            fetch(i + nwarps, nrow, n0, n1);
            if ((h0.x | h0.y) != 0u) {
                const uint32_t* d = reinterpret_cast<const uint32_t*>(((unsigned long long)h0.y << 32) | h0.x);
                const Tally t = stream_row<COUNT>(qv, d, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, k, pos_bits, pmask, lane, touched);
                emit_if_close(p, t, k, nVq, nUq, nMq, qs.row, row, lane);
            }
            i += nwarps; row = nrow; h0 = n0; h1 = n1;
        }
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, o);
        if (lane == 0) atomicAdd(p.touched, touched);
    }
}

// ------------------------------------------------------------------------------------------
// ingest: compress() on the device (seqComparer.py:274-307)
// ------------------------------------------------------------------------------------------
// refmask[p] = reference character, or 0 where p is excluded.  Class of a position:
// 0 same-as-reference or excluded, 1..4 A,C,G,T, 5 'N' or '-', 6 anything else (mixed base).

__device__ __forceinline__ uint32_t base_class(uint32_t s, uint32_t r) {
    if (r == 0u) return 0u;
    if (s == (uint32_t)'-') s = (uint32_t)'N';
    if (s == r) return 0u;
    if (s == (uint32_t)'A') return 1u;
    if (s == (uint32_t)'C') return 2u;
    if (s == (uint32_t)'G') return 3u;
    if (s == (uint32_t)'T') return 4u;
    if (s == (uint32_t)'N') return 5u;
    return 6u;
}

#define FN3_CMP_THREADS 256
#define FN3_CMP_PER_THREAD 16

// per-thread class counts packed as 3 x 16-bit fields in two 64-bit words (a CTA covers 4096 positions)
__device__ __forceinline__ void classify16(const uint8_t* __restrict__ seq, const uint8_t* __restrict__ refmask,
                                           long long p0, uint8_t cls[FN3_CMP_PER_THREAD],
                                           unsigned long long& lo, unsigned long long& hi) {
    const uint4 sv = __ldg(reinterpret_cast<const uint4*>(seq + p0));
    const uint4 rv = __ldg(reinterpret_cast<const uint4*>(refmask + p0));
    const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
    lo = 0ull; hi = 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = base_class((sw[k] >> (8 * j)) & 0xffu, (rw[k] >> (8 * j)) & 0xffu);
            cls[k * 4 + j] = (uint8_t)c;
            if (c >= 1u && c <= 3u) lo += 1ull << (16 * (c - 1u));
            else if (c >= 4u) hi += 1ull << (16 * (c - 4u));
        }
}

__global__ void __launch_bounds__(FN3_CMP_THREADS) k_compress_count(const uint8_t* __restrict__ seq,
                                                                     const uint8_t* __restrict__ refmask,
                                                                     uint32_t* __restrict__ counts) {
    const long long p0 = ((long long)blockIdx.x * FN3_CMP_THREADS + threadIdx.x) * FN3_CMP_PER_THREAD;
    uint8_t cls[FN3_CMP_PER_THREAD];
    unsigned long long lo, hi;
    classify16(seq, refmask, p0, cls, lo, hi);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, o); hi += __shfl_xor_sync(0xffffffffu, hi, o); }
    __shared__ unsigned long long slo[FN3_CMP_THREADS / 32], shi[FN3_CMP_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned long long t = 0;
        for (int w = 0; w < FN3_CMP_THREADS / 32; ++w) t += (threadIdx.x < 3) ? slo[w] : shi[w];
        counts[(long long)blockIdx.x * 6 + threadIdx.x] = (uint32_t)((t >> (16 * (threadIdx.x % 3))) & 0xffffull);
    }
}

// exclusive scan over CTAs, per class; totals[6] receives the list lengths
__global__ void k_compress_scan(uint32_t* __restrict__ counts, uint32_t* __restrict__ totals, int ncta) {
    const int c = threadIdx.x;
    if (c >= 6) return;
    uint32_t acc = 0;
    for (int i = 0; i < ncta; ++i) { const uint32_t v = counts[(long long)i * 6 + c]; counts[(long long)i * 6 + c] = acc; acc += v; }
    totals[c] = acc;
}

__global__ void __launch_bounds__(FN3_CMP_THREADS) k_compress_write(const uint8_t* __restrict__ seq,
                                                                     const uint8_t* __restrict__ refmask,
                                                                     const uint32_t* __restrict__ counts,
                                                                     const uint32_t* __restrict__ totals,
                                                                     int32_t* __restrict__ out_pos, uint8_t* __restrict__ out_m) {
    const long long p0 = ((long long)blockIdx.x * FN3_CMP_THREADS + threadIdx.x) * FN3_CMP_PER_THREAD;
    uint8_t cls[FN3_CMP_PER_THREAD];
    unsigned long long lo, hi;
    classify16(seq, refmask, p0, cls, lo, hi);
    // exclusive scan of (lo,hi) over the CTA's threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long ilo = lo, ihi = hi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, ilo, o), b = __shfl_up_sync(0xffffffffu, ihi, o);
        if (lane >= o) { ilo += a; ihi += b; }
    }
    __shared__ unsigned long long slo[FN3_CMP_THREADS / 32], shi[FN3_CMP_THREADS / 32];
    if (lane == 31) { slo[warp] = ilo; shi[warp] = ihi; }
    __syncthreads();
    unsigned long long wlo = 0, whi = 0;
    for (int w = 0; w < warp; ++w) { wlo += slo[w]; whi += shi[w]; }
    const unsigned long long elo = wlo + ilo - lo, ehi = whi + ihi - hi;
    uint32_t at[7];
    uint32_t start = 0;
#pragma unroll
    for (int c = 1; c <= 6; ++c) {
        const uint32_t within = (uint32_t)(((c <= 3 ? elo : ehi) >> (16 * ((c - 1) % 3))) & 0xffffull);
        at[c] = start + counts[(long long)blockIdx.x * 6 + (c - 1)] + within;
        start += totals[c - 1];
    }
    const uint32_t m_base = start - totals[5];
#pragma unroll
    for (int j = 0; j < FN3_CMP_PER_THREAD; ++j) {
        const uint32_t c = cls[j];
        if (c) {
            const uint32_t dst = at[c]++;
            out_pos[dst] = (int32_t)(p0 + j);
            if (c == 6u) out_m[dst - m_base] = seq[p0 + j];
        }
    }
}

// L2 flush helper for the measurement hook
__global__ void k_fill(uint4* p, long long n, uint32_t v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_uint4(v, v, v, v);
}

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
    std::mutex query_mu;        // guards query state (maps, slots, out buffers, q_stream)
    std::string err;
    std::mutex err_mu;

    cudaStream_t copy_stream = nullptr, q_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    // store
    size_t chunk_words = 0;                 // words per data chunk
    std::vector<uint32_t*> data_chunks;
    std::vector<size_t> data_chunk_cap;     // words
    size_t cur_used = 0;                    // words used in the last chunk
    int hdr_shift = 20;
    std::vector<RowHead*> hdr_chunks;       // device: header + first V words per row
    std::vector<RowHdr> h_hdr;              // host mirror (published rows)
    std::vector<uint8_t> h_flags;           // bit0 invalid, bit1 removed
    std::vector<uint32_t> h_npos;           // canonical position count per row
    std::vector<uint32_t> h_nblk;           // number of 32-position blocks the row touches (sizes query staging)
    long long n_rows = 0;                   // published
    long long n_live = 0, n_invalid = 0;
    long long store_words = 0, store_bytes = 0, positions = 0;

    // staging
    uint32_t* stage = nullptr; size_t stage_words = 0;     // pinned
    RowHead* stage_hdr = nullptr; size_t stage_hdr_n = 0;  // pinned

    // query state
    int n_slots = 8;
    QueryState qstate;                     // device arrays of the compact queries (n_slots of them)
    int max_smem = 0;                      // opt-in dynamic shared memory per CTA
    int stage_limit = 0;                   // = max_smem, or 0 to force the global-memory query variants
    int q_static_smem = 0;                 // static shared memory of k_query
    uint32_t slot_blocks[FN3_MAX_SLOTS];   // touched-block count of the query in each slot (sizes the staging)
    QuerySlot* d_slots = nullptr; QuerySlot* h_slots = nullptr;   // h_slots pinned
    // survivor buffer: one allocation = [32-byte counter block | records], mirrored in pinned host memory so
    // that the counter and the first records come back in ONE device-to-host copy
    unsigned char* d_outbuf = nullptr; unsigned char* h_outbuf = nullptr;
    EdgeRec* d_out = nullptr; long long out_cap = 0;
    unsigned long long* d_count = nullptr;   // [0] hit counter, [1] touched bytes
    unsigned long long* h_count = nullptr;
    EdgeRec* h_out = nullptr; long long h_out_cap = 0;
    std::atomic<long long> h2d_bytes{0}, d2h_bytes{0};
    long long* d_rows = nullptr; long long d_rows_cap = 0;
    long long* h_rows = nullptr; long long h_rows_cap = 0;       // pinned
    // scratch rows for samples that are not stored (pair_lists / lists_vs_all)
    uint32_t* d_scratch = nullptr; size_t scratch_words = 0;
    RowHead* d_scratch_hdr = nullptr;        // [2]
    uint4* d_flush = nullptr; long long flush_n = 0;
    // ingest (fn3_compress)
    std::mutex ingest_mu;
    cudaStream_t in_stream = nullptr;
    long long Lpad = 0; int cmp_ctas = 0;
    uint8_t* d_refmask = nullptr; uint8_t* d_seq = nullptr; uint8_t* d_mchars = nullptr;
    int32_t* d_cpos = nullptr; uint32_t* d_ccounts = nullptr; uint32_t* d_ctotals = nullptr;
    uint8_t* h_seq = nullptr; uint32_t* h_ctotals = nullptr;   // pinned
    int32_t* h_cpos = nullptr; size_t h_cpos_cap = 0; uint8_t* h_mchars = nullptr; size_t h_mchars_cap = 0;   // pinned

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

#define FN3_OUT_HDR 32
#define FN3_FIRST_HITS 256      // records fetched together with the counter

static int ensure_out(fn3_engine* e, long long cap) {
    if (cap <= e->out_cap) return FN3_OK;
    long long want = std::max<long long>(cap, e->out_cap * 2);
    want = std::max<long long>(want, 4096);
    if (e->d_outbuf) { cudaStreamSynchronize(e->q_stream); cudaFree(e->d_outbuf); }
    e->d_outbuf = nullptr; e->out_cap = 0;
    CU(e, cudaMalloc(&e->d_outbuf, FN3_OUT_HDR + (size_t)want * sizeof(EdgeRec)));
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
    if ((s = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", s);
    cudaDeviceProp prop;
    if ((s = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", s);
    e->sm_count = prop.multiProcessorCount;
    int pb = 1; while ((1ll << pb) < genome_length) ++pb;
    e->pos_bits = pb;
    e->max_run = (pb >= 32) ? 1u : (uint32_t)std::min<long long>(1ll << (32 - pb), 1ll << 30);
    e->chunk_words = (size_t)env_ll("FN3_CHUNK_MB", 256) * (1u << 20) / 4;
    if (e->chunk_words < 1024) e->chunk_words = 1024;
    e->hdr_shift = (int)env_ll("FN3_HDR_SHIFT", 20);
    if (e->hdr_shift < 2) e->hdr_shift = 2;
    if (e->hdr_shift > 26) e->hdr_shift = 26;
    e->n_slots = (int)env_ll("FN3_SLOTS", 8);
    if (e->n_slots < 1) e->n_slots = 1;
    if (e->n_slots > FN3_MAX_SLOTS) e->n_slots = FN3_MAX_SLOTS;

    if ((s = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", s);
    if ((s = cudaStreamCreateWithFlags(&e->q_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", s);
    for (int i = 0; i < 4; ++i)
        if ((s = cudaEventCreate(&e->ev[i])) != cudaSuccess) return fail("cudaEventCreate", s);

    // query state (compact form); the last block is the end sentinel for rank(L)
    {
        QueryState& q = e->qstate;
        memset(&q, 0, sizeof q);
        q.n_blk = (int)((genome_length + 31) / 32 + 1);
        q.nw = (q.n_blk + 31) / 32;
        q.nw = (q.nw + 3) & ~3;                                  // entries start 32-byte aligned behind bmpref
        q.ent_stride = (long long)q.n_blk + 1;
        if ((s = cudaMalloc(&q.bmpref, (size_t)q.nw * sizeof(uint2) * e->n_slots)) != cudaSuccess) return fail("cudaMalloc(bmpref)", s);
        if ((s = cudaMalloc(&q.ent, (size_t)q.ent_stride * sizeof(QEntry) * e->n_slots)) != cudaSuccess) return fail("cudaMalloc(entries)", s);
        if ((s = cudaMalloc(&q.info, sizeof(uint4) * FN3_MAX_SLOTS)) != cudaSuccess) return fail("cudaMalloc(info)", s);
        if ((s = cudaMalloc(&q.bm_scratch, (size_t)q.nw * 2 * sizeof(uint32_t) * e->n_slots)) != cudaSuccess) return fail("cudaMalloc(bm)", s);
        e->max_smem = (int)prop.sharedMemPerBlockOptin;
        e->stage_limit = env_ll("FN3_FORCE_GLOBAL_QUERY", 0) ? 0 : e->max_smem;   // tests: exercise the non-staged variants
        // opt in to the full shared memory for the kernels that stage the query (minus their static part)
        const void* staged[5] = {(const void*)k_query<false, 1, false>, (const void*)k_query<false, 1, true>,
                                 (const void*)k_query<true, 1, false>, (const void*)k_query<true, 1, true>,
                                 (const void*)k_build_query<true>};
        e->q_static_smem = 0;
        for (int i = 0; i < 5; ++i) {
            cudaFuncAttributes fa;
            if ((s = cudaFuncGetAttributes(&fa, staged[i])) != cudaSuccess) return fail("cudaFuncGetAttributes", s);
            if (i < 4) e->q_static_smem = std::max(e->q_static_smem, (int)fa.sharedSizeBytes);
            if ((s = cudaFuncSetAttribute(staged[i], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          e->max_smem - (int)fa.sharedSizeBytes)) != cudaSuccess)
                return fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", s);
        }
    }
    if ((s = cudaMalloc(&e->d_slots, sizeof(QuerySlot) * FN3_MAX_SLOTS)) != cudaSuccess) return fail("cudaMalloc(slots)", s);
    if ((s = cudaHostAlloc(&e->h_slots, sizeof(QuerySlot) * FN3_MAX_SLOTS, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc(slots)", s);
    if (ensure_out(e, 4096) != FN3_OK || ensure_hout(e, 4096) != FN3_OK) {
        set_err(nullptr, FN3_ENOMEM, "fn3_create: survivor buffers: %s", e->err.c_str());
        fn3_destroy(e);
        return nullptr;
    }
    if ((s = cudaMalloc(&e->d_scratch_hdr, 2 * sizeof(RowHead))) != cudaSuccess) return fail("cudaMalloc(scratch hdr)", s);
    e->store_bytes = 0;
    return e;
}

extern "C" int fn3_destroy(fn3_engine* e) {
    if (!e) return FN3_OK;
    cudaSetDevice(e->device);
    if (e->q_stream) cudaStreamSynchronize(e->q_stream);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    for (auto p : e->data_chunks) cudaFree(p);
    for (auto p : e->hdr_chunks) cudaFree(p);
    cudaFree(e->qstate.bmpref); cudaFree(e->qstate.ent); cudaFree(e->qstate.info); cudaFree(e->qstate.bm_scratch);
    cudaFree(e->d_slots);
    cudaFree(e->d_outbuf); cudaFree(e->d_rows); cudaFree(e->d_scratch);
    cudaFree(e->d_scratch_hdr); cudaFree(e->d_flush);
    cudaFree(e->d_refmask); cudaFree(e->d_seq); cudaFree(e->d_mchars); cudaFree(e->d_cpos);
    cudaFree(e->d_ccounts); cudaFree(e->d_ctotals);
    if (e->h_seq) cudaFreeHost(e->h_seq);
    if (e->h_ctotals) cudaFreeHost(e->h_ctotals);
    if (e->h_cpos) cudaFreeHost(e->h_cpos);
    if (e->h_mchars) cudaFreeHost(e->h_mchars);
    if (e->in_stream) cudaStreamDestroy(e->in_stream);
    if (e->h_slots) cudaFreeHost(e->h_slots);
    if (e->h_outbuf) cudaFreeHost(e->h_outbuf);
    if (e->h_rows) cudaFreeHost(e->h_rows);
    if (e->stage) cudaFreeHost(e->stage);
    if (e->stage_hdr) cudaFreeHost(e->stage_hdr);
    for (int i = 0; i < 4; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    if (e->q_stream) cudaStreamDestroy(e->q_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
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

// header + copy of the first V words (the early-exit filter's contiguous read)
static void fill_head(RowHead& hd, const RowHdr& h, const uint32_t* words) {
    hd.hdr = h;
    const uint32_t nv = h.data ? std::min<uint32_t>(h.end[3], FN3_HEAD_WORDS) : 0u;
    for (uint32_t i = 0; i < nv; ++i) hd.first[i] = words[i];
    for (uint32_t i = nv; i < FN3_HEAD_WORDS; ++i) hd.first[i] = 0u;
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
        RowHead* p = nullptr;
        size_t bytes = sizeof(RowHead) << e->hdr_shift;
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

static int append_impl(fn3_engine* e, long long n, const int32_t* pos, const long long* off6, const uint8_t* invalid,
                       long long* out_first_row) {
    std::lock_guard<std::mutex> g(e->store_mu);
    CU(e, cudaSetDevice(e->device));
    // pass 1: validate + size
    std::vector<long long> words((size_t)n);
    std::vector<uint32_t> nblks((size_t)n);
    for (long long i = 0; i < n; ++i) {
        long long o[7];
        for (int b = 0; b < 7; ++b) o[b] = off6[6 * i + b];
        long long w = row_words_needed(e, pos, o, invalid ? invalid[i] : 0, &nblks[(size_t)i]);
        if (w < 0) return FN3_EARG;
        words[(size_t)i] = w;
    }
    const long long first = e->n_rows;
    // pass 2: pack into pinned staging in batches (each batch is one contiguous arena allocation), copy
    const size_t batch_limit = std::min<size_t>((size_t)16 << 20, e->chunk_words);   // words
    long long i = 0;
    while (i < n) {
        size_t batch_words = 0; long long j = i;
        while (j < n) {
            size_t need = ((size_t)words[(size_t)j] + 3) & ~(size_t)3;
            if (need == 0) need = 4;
            if (j > i && batch_words + need > batch_limit) break;
            batch_words += need; ++j;
        }
        int rc = ensure_stage(e, batch_words, (size_t)(j - i));
        if (rc) return rc;
        uint32_t* dbase = nullptr;
        rc = arena_alloc(e, batch_words, &dbase);
        if (rc) return rc;
        size_t w = 0;
        for (long long r = i; r < j; ++r) {
            long long o[7];
            for (int b = 0; b < 7; ++b) o[b] = off6[6 * r + b];
            RowHdr h;
            const bool inv = invalid && invalid[r];
            size_t need = ((size_t)words[(size_t)r] + 3) & ~(size_t)3;
            if (need == 0) need = 4;
            if (inv) {
                h.data = 0; for (int b = 0; b < 6; ++b) h.end[b] = 0;
                memset(e->stage + w, 0, need * 4);
            } else {
                pack_row(e, pos, o, e->stage + w, h.end);
                for (size_t z = (size_t)words[(size_t)r]; z < need; ++z) e->stage[w + z] = 0;
                h.data = (unsigned long long)(uintptr_t)(dbase + w);
            }
            fill_head(e->stage_hdr[r - i], h, e->stage + w);
            w += need;
        }
        CU(e, cudaMemcpyAsync(dbase, e->stage, batch_words * 4, cudaMemcpyHostToDevice, e->copy_stream));
        e->h2d_bytes += (long long)batch_words * 4 + (j - i) * (long long)sizeof(RowHead);
        // headers: may straddle header chunks
        long long r = i;
        while (r < j) {
            const long long row = first + r;
            rc = ensure_hdr_chunk(e, row);
            if (rc) return rc;
            const long long in_chunk = row & ((1ll << e->hdr_shift) - 1);
            const long long take = std::min<long long>(j - r, (1ll << e->hdr_shift) - in_chunk);
            CU(e, cudaMemcpyAsync(e->hdr_chunks[(size_t)(row >> e->hdr_shift)] + in_chunk, e->stage_hdr + (r - i),
                                  (size_t)take * sizeof(RowHead), cudaMemcpyHostToDevice, e->copy_stream));
            r += take;
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
    CU(e, cudaSetDevice(e->device));
    RowHdr dead; memset(&dead, 0, sizeof dead);
    // queries may be in flight on q_stream: a header flips atomically per 32-byte sector store at worst
    // between two comparisons, which is the reference's own semantics for remove-during-mcompare.
    CU(e, cudaMemcpyAsync(e->hdr_chunks[(size_t)(row >> e->hdr_shift)] + (row & ((1ll << e->hdr_shift) - 1)), &dead,
                          sizeof dead, cudaMemcpyHostToDevice, e->copy_stream));
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
    CU(e, cudaSetDevice(e->device));
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
}

// the early-exit filter looks at this many leading V words of each candidate's row head (0 = STREAM variant)
static int filter_len_for(int ceiling) {
    if (ceiling < 0) return 4;
    if (ceiling > 24) return 0;                  // 32 head words cannot push S1 past the ceiling reliably
    return std::min(FN3_HEAD_WORDS, ((ceiling + 1 + 7 + 3) / 4) * 4);
}

template <bool COUNT, int QMODE>
static void launch_query(bool filter, dim3 g, size_t smem, cudaStream_t st, const CompareParams& p) {
    if (filter) k_query<COUNT, QMODE, true><<<g, FN3_Q_THREADS, smem, st>>>(p);
    else k_query<COUNT, QMODE, false><<<g, FN3_Q_THREADS, smem, st>>>(p);
}

// enqueue one query tile: `ns` slots (h_slots + slot_blocks filled by the caller, or `resident` slots that
// already live on the device — measurement hook) against candidates rows[0..n) / row_begin..row_begin+n-1.
// Normally ONE kernel (k_query builds the query structure in shared memory itself); queries too large for
// shared memory take k_build_query + the global-memory variant.  `split` (measurement hook) receives an event
// recorded between the optional build kernel and k_query.
static int enqueue_query(fn3_engine* e, const HdrTable& ht, const long long* d_rows, long long row_begin, long long n,
                         int ns, int ceiling, bool count_bytes, const QuerySlot* resident = nullptr,
                         cudaEvent_t split = nullptr) {
    cudaStream_t st = e->q_stream;
    const QuerySlot* d_slots = resident ? resident : e->d_slots;
    if (!resident) {
        CU(e, cudaMemcpyAsync(e->d_slots, e->h_slots, sizeof(QuerySlot) * ns, cudaMemcpyHostToDevice, st));
        e->h2d_bytes += (long long)sizeof(QuerySlot) * ns;
    }
    uint32_t blocks = 0;
    for (int s = 0; s < ns; ++s) blocks = std::max(blocks, e->slot_blocks[s]);
    blocks = std::min<uint32_t>(blocks, (uint32_t)e->qstate.n_blk);
    const size_t smem = (size_t)e->qstate.nw * sizeof(uint2) + ((size_t)blocks + 1) * sizeof(QEntry);
    const bool in_smem = smem + (size_t)e->q_static_smem + 256 <= (size_t)e->stage_limit;
    if (!in_smem) {
        const size_t bm_bytes = (size_t)e->qstate.nw * 2 * sizeof(uint32_t);
        if (bm_bytes + 1024 <= (size_t)e->stage_limit)
            k_build_query<true><<<ns, FN3_BUILD_THREADS, bm_bytes, st>>>(d_slots, e->qstate, e->pos_bits);
        else
            k_build_query<false><<<ns, FN3_BUILD_THREADS, 0, st>>>(d_slots, e->qstate, e->pos_bits);
        e->launches += 1;
        CU(e, cudaGetLastError());
    }
    if (split) CU(e, cudaEventRecord(split, st));
    if (n <= 0) return FN3_OK;
    CompareParams p;
    p.hdr = ht; p.qs = e->qstate; p.rows = d_rows; p.row_begin = row_begin; p.n = n;
    p.slots = d_slots;
    p.out = e->d_out; p.out_count = e->d_count; p.touched = e->d_count + 1; p.out_cap = e->out_cap;
    p.ceiling = ceiling; p.pos_bits = e->pos_bits; p.filter_len = filter_len_for(ceiling);
    const bool filter = p.filter_len > 0;
    const long long warps_per_cta = FN3_Q_THREADS / 32;
    const long long units = filter ? (n + 31) / 32 : n;             // warp-sized work units
    long long ctas = (units + warps_per_cta - 1) / warps_per_cta;
    long long per_sm = 2;                                           // __launch_bounds__(512, 2)
    if (in_smem && 2 * (smem + (size_t)e->q_static_smem + 1024) > (size_t)228 * 1024) per_sm = 1;
    const long long max_ctas = (long long)e->sm_count * per_sm;
    if (ctas > max_ctas) ctas = max_ctas;
    dim3 g((unsigned)ctas, (unsigned)ns);
    if (in_smem) {
        if (count_bytes) launch_query<true, 1>(filter, g, smem, st, p);
        else launch_query<false, 1>(filter, g, smem, st, p);
    } else {
        if (count_bytes) launch_query<true, 0>(filter, g, 0, st, p);
        else launch_query<false, 0>(filter, g, 0, st, p);
    }
    e->launches += 1;
    CU(e, cudaGetLastError());
    return FN3_OK;
}

// upload a non-stored sample into scratch slot `which` (0/1); returns its header
static int upload_scratch(fn3_engine* e, int which, const int32_t* pos, const int32_t off[7], int invalid, RowHdr* out,
                          uint32_t* nblk = nullptr, RowHead* head_out = nullptr) {
    RowHdr h; memset(&h, 0, sizeof h);
    if (nblk) *nblk = 0;
    if (invalid) { *out = h; if (head_out) fill_head(*head_out, h, nullptr); return FN3_OK; }
    long long o[7];
    for (int b = 0; b < 7; ++b) o[b] = off[b];
    long long w = row_words_needed(e, pos, o, 0, nblk);
    if (w < 0) return FN3_EARG;
    size_t need = ((size_t)w + 3) & ~(size_t)3; if (need == 0) need = 4;
    // scratch holds two samples: [0, half) and [half, 2*half)
    if (need * 2 > e->scratch_words) {
        size_t want = std::max(need * 2, e->scratch_words * 2);
        want = std::max<size_t>(want, 1u << 16);
        CU(e, cudaStreamSynchronize(e->q_stream));
        if (e->d_scratch) cudaFree(e->d_scratch);
        e->d_scratch = nullptr; e->scratch_words = 0;
        CU(e, cudaMalloc(&e->d_scratch, want * 4));
        e->scratch_words = want;
    }
    {
        std::lock_guard<std::mutex> g(e->store_mu);        // staging buffer is shared with append
        int rc = ensure_stage(e, need, 1);
        if (rc) return rc;
        pack_row(e, pos, o, e->stage, h.end);
        for (size_t z = (size_t)w; z < need; ++z) e->stage[z] = 0;
        uint32_t* dst = e->d_scratch + (which ? e->scratch_words / 2 : 0);
        h.data = (unsigned long long)(uintptr_t)dst;
        if (head_out) fill_head(*head_out, h, e->stage);
        CU(e, cudaMemcpyAsync(dst, e->stage, need * 4, cudaMemcpyHostToDevice, e->q_stream));
        e->h2d_bytes += (long long)need * 4;
        CU(e, cudaStreamSynchronize(e->q_stream));
    }
    *out = h;
    return FN3_OK;
}

// shared tail of one-vs-all style calls: slot 0 prepared in h_slots[0]
static int run_one_vs_all(fn3_engine* e, int ceiling, const int64_t* rows, int64_t n_rows_arg, long long n_store,
                          const HdrTable& ht, fn3_hit* out, int64_t cap, int64_t* n_out) {
    long long n = rows ? n_rows_arg : n_store;
    *n_out = 0;
    if (e->h_slots[0].hdr.data == 0ull || n <= 0) return FN3_OK;   // invalid query: nothing is a neighbour
    int rc;
    const long long* d_rows = nullptr;
    if (rows) {
        if ((rc = ensure_rows(e, n))) return rc;
        for (long long i = 0; i < n; ++i) {
            if (rows[i] < 0 || rows[i] >= n_store)
                return set_err(e, FN3_ENOTFOUND, "candidate row %lld does not exist", (long long)rows[i]);
            e->h_rows[i] = rows[i];
        }
        CU(e, cudaMemcpyAsync(e->d_rows, e->h_rows, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, e->q_stream));
        e->h2d_bytes += n * (long long)sizeof(long long);
        d_rows = e->d_rows;
    }
    if ((rc = ensure_out(e, n))) return rc;
    CU(e, cudaMemsetAsync(e->d_count, 0, FN3_OUT_HDR, e->q_stream));
    if ((rc = enqueue_query(e, ht, d_rows, 0, n, 1, ceiling, false))) return rc;
    // ONE device-to-host copy brings the counter and the first records; a second one only for long lists
    const long long first = std::min<long long>(n, FN3_FIRST_HITS);
    CU(e, cudaMemcpyAsync(e->h_outbuf, e->d_outbuf, FN3_OUT_HDR + (size_t)first * sizeof(EdgeRec), cudaMemcpyDeviceToHost,
                          e->q_stream));
    CU(e, cudaStreamSynchronize(e->q_stream));
    e->d2h_bytes += FN3_OUT_HDR + first * (long long)sizeof(EdgeRec);
    const long long cnt = (long long)e->h_count[0];
    *n_out = cnt;
    if (cnt > cap) return set_err(e, FN3_ECAPACITY, "output buffer holds %lld hits, %lld needed", (long long)cap, cnt);
    if (cnt > first) {
        if ((rc = ensure_hout(e, cnt))) return rc;
        CU(e, cudaMemcpyAsync(e->h_out, e->d_out, (size_t)cnt * sizeof(EdgeRec), cudaMemcpyDeviceToHost, e->q_stream));
        CU(e, cudaStreamSynchronize(e->q_stream));
        e->d2h_bytes += cnt * (long long)sizeof(EdgeRec);
    }
    for (long long i = 0; i < cnt; ++i) {
        const EdgeRec& r = e->h_out[i];
        out[i].row = r.row2; out[i].dist = r.dist; out[i].n1 = r.n1; out[i].n2 = r.n2; out[i].nboth = r.nboth;
    }
    return FN3_OK;
}

static int snapshot_store(fn3_engine* e, HdrTable& ht, long long& n_store) {
    std::lock_guard<std::mutex> g(e->store_mu);
    fill_hdr_table(e, ht);
    n_store = e->n_rows;
    return FN3_OK;
}

static int stored_hdr(fn3_engine* e, long long row, RowHdr* h, const char* who, uint32_t* nblk = nullptr) {
    std::lock_guard<std::mutex> g(e->store_mu);
    if (row < 0 || row >= e->n_rows) return set_err(e, FN3_ENOTFOUND, "%s: row %lld does not exist", who, row);
    if (e->h_flags[(size_t)row] & 2) return set_err(e, FN3_ENOTFOUND, "%s: row %lld was removed", who, row);
    *h = e->h_hdr[(size_t)row];
    if (nblk) *nblk = e->h_nblk[(size_t)row];
    return FN3_OK;
}

extern "C" int fn3_one_vs_all(fn3_engine* e, int64_t query_row, int32_t ceiling, const int64_t* rows, int64_t n_rows,
                              fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || (rows && n_rows < 0)) return set_err(e, FN3_EARG, "fn3_one_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    RowHdr qh; int rc;
    if ((rc = stored_hdr(e, query_row, &qh, "fn3_one_vs_all", &e->slot_blocks[0]))) return rc;
    HdrTable ht; long long n_store;
    snapshot_store(e, ht, n_store);
    QuerySlot& s = e->h_slots[0];
    s.hdr = qh; s.row = query_row; s.skip_le = -1; s.exclude = query_row; s.pad = 0;
    return run_one_vs_all(e, ceiling, rows, n_rows, n_store, ht, out, cap, n_out);
}

// insert() of the reference's server (findNeighbour3-server.py:516 persist + :530 mcompare) as ONE call on ONE
// stream with ONE host synchronisation: row data + row head + query slot go host->device, k_query runs against
// every earlier row, the neighbour records come back.
extern "C" int fn3_insert_query(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                                int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || !out_row || (cap > 0 && !out) || (!invalid && !off))
        return set_err(e, FN3_EARG, "fn3_insert_query: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    HdrTable ht; long long row; RowHdr h; memset(&h, 0, sizeof h);
    uint32_t nblk = 0;
    {
        std::lock_guard<std::mutex> g(e->store_mu);
        long long o[7] = {0, 0, 0, 0, 0, 0, 0};
        if (!invalid) for (int b = 0; b < 7; ++b) o[b] = off[b];
        const long long w = row_words_needed(e, pos, o, invalid, &nblk);
        if (w < 0) return FN3_EARG;
        size_t need = ((size_t)w + 3) & ~(size_t)3; if (need == 0) need = 4;
        int rc = ensure_stage(e, need, 1);
        if (rc) return rc;
        uint32_t* dbase = nullptr;
        if ((rc = arena_alloc(e, need, &dbase))) return rc;
        row = e->n_rows;
        if ((rc = ensure_hdr_chunk(e, row))) return rc;
        if (invalid) memset(e->stage, 0, need * 4);
        else {
            pack_row(e, pos, o, e->stage, h.end);
            for (size_t z = (size_t)w; z < need; ++z) e->stage[z] = 0;
            h.data = (unsigned long long)(uintptr_t)dbase;
        }
        fill_head(e->stage_hdr[0], h, e->stage);
        // the header chunk may have just been created (zeroed on copy_stream): order q_stream behind it
        CU(e, cudaStreamSynchronize(e->copy_stream));
        CU(e, cudaMemcpyAsync(dbase, e->stage, need * 4, cudaMemcpyHostToDevice, e->q_stream));
        CU(e, cudaMemcpyAsync(e->hdr_chunks[(size_t)(row >> e->hdr_shift)] + (row & ((1ll << e->hdr_shift) - 1)),
                              e->stage_hdr, sizeof(RowHead), cudaMemcpyHostToDevice, e->q_stream));
        e->h2d_bytes += (long long)need * 4 + (long long)sizeof(RowHead);
        // the pinned staging buffers are reused by the next append: wait until the copies have left them.
        // (q_stream is otherwise idle here; this is the copy latency only, the kernel is queued below.)
        e->h_hdr.push_back(h);
        e->h_flags.push_back(invalid ? 1 : 0);
        const uint32_t np = invalid ? 0u : (uint32_t)(o[6] - o[0]);
        e->h_npos.push_back(np);
        e->h_nblk.push_back(nblk);
        e->positions += np;
        e->n_live++; if (invalid) e->n_invalid++;
        e->n_rows = row + 1;
        fill_hdr_table(e, ht);
        QuerySlot& s = e->h_slots[0];
        s.hdr = h; s.row = row; s.skip_le = -1; s.exclude = row; s.pad = 0;
        e->slot_blocks[0] = nblk;
        *out_row = row;
        // run the query while still holding store_mu: the staging buffers must not be repacked by a concurrent
        // append until the stream has consumed them (the synchronise inside run_one_vs_all covers that)
        rc = run_one_vs_all(e, ceiling, nullptr, 0, row, ht, out, cap, n_out);
        if (invalid || row == 0) CU(e, cudaStreamSynchronize(e->q_stream));   // paths that return without a sync
        return rc;
    }
}

extern "C" int fn3_lists_vs_all(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                                const int64_t* rows, int64_t n_rows, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || (!invalid && !off)) return set_err(e, FN3_EARG, "fn3_lists_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    RowHdr qh; int rc;
    if ((rc = upload_scratch(e, 0, pos, off, invalid, &qh, &e->slot_blocks[0]))) return rc;
    HdrTable ht; long long n_store;
    snapshot_store(e, ht, n_store);
    QuerySlot& s = e->h_slots[0];
    s.hdr = qh; s.row = -1; s.skip_le = -1; s.exclude = -1; s.pad = 0;
    return run_one_vs_all(e, ceiling, rows, n_rows, n_store, ht, out, cap, n_out);
}

extern "C" int fn3_pair(fn3_engine* e, int64_t row1, int64_t row2, int32_t ceiling, fn3_hit* out, int32_t* has_dist) {
    if (!e || !out || !has_dist) return set_err(e, FN3_EARG, "fn3_pair: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    RowHdr qh, ch; int rc;
    if ((rc = stored_hdr(e, row1, &qh, "fn3_pair", &e->slot_blocks[0]))) return rc;
    if ((rc = stored_hdr(e, row2, &ch, "fn3_pair"))) return rc;
    HdrTable ht; long long n_store;
    snapshot_store(e, ht, n_store);
    QuerySlot& s = e->h_slots[0];
    s.hdr = qh; s.row = row1; s.skip_le = -1; s.exclude = -1; s.pad = 0;   // the self pair IS computed here (:382-421)
    int64_t n = 0; int64_t r2 = row2;
    rc = run_one_vs_all(e, ceiling, &r2, 1, n_store, ht, out, 1, &n);
    *has_dist = (rc == FN3_OK && n == 1) ? 1 : 0;
    return rc;
}

extern "C" int fn3_pair_lists(fn3_engine* e, const int32_t* pos1, const int32_t off1[7], int32_t invalid1,
                              const int32_t* pos2, const int32_t off2[7], int32_t invalid2, int32_t ceiling,
                              fn3_hit* out, int32_t* has_dist) {
    if (!e || !out || !has_dist) return set_err(e, FN3_EARG, "fn3_pair_lists: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    *has_dist = 0;
    RowHdr qh, ch; int rc;
    if ((rc = upload_scratch(e, 0, pos1, off1, invalid1, &qh, &e->slot_blocks[0]))) return rc;
    RowHead chead;
    if ((rc = upload_scratch(e, 1, pos2, off2, invalid2, &ch, nullptr, &chead))) return rc;
    if (qh.data == 0ull || ch.data == 0ull) return FN3_OK;   // either invalid: None (:438-442)
    // scratch header table: one chunk holding the candidate header at row 0
    CU(e, cudaMemcpyAsync(e->d_scratch_hdr, &chead, sizeof chead, cudaMemcpyHostToDevice, e->q_stream));
    HdrTable ht; memset(&ht, 0, sizeof ht);
    ht.chunk[0] = e->d_scratch_hdr; ht.shift = 1;
    QuerySlot& s = e->h_slots[0];
    s.hdr = qh; s.row = -1; s.skip_le = -1; s.exclude = -1; s.pad = 0;
    int64_t n = 0;
    rc = run_one_vs_all(e, ceiling, nullptr, 0, 1, ht, out, 1, &n);
    if (rc == FN3_OK && n == 1) { *has_dist = 1; out->row = -1; }
    return rc;
}

extern "C" int fn3_all_vs_all(fn3_engine* e, int32_t ceiling, int32_t upper_only, int64_t row_begin, int64_t row_end,
                              int64_t stride, int64_t phase, fn3_edge* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || stride < 1 || phase < 0 || phase >= stride)
        return set_err(e, FN3_EARG, "fn3_all_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    HdrTable ht; long long n_store;
    snapshot_store(e, ht, n_store);
    if (row_begin < 0) row_begin = 0;
    if (row_end > n_store) row_end = n_store;
    *n_out = 0;
    long long total = 0;
    int rc;
    const int T = e->n_slots;
    if ((rc = ensure_out(e, (long long)T * n_store))) return rc;
    std::vector<long long> qrows;
    {
        std::lock_guard<std::mutex> g(e->store_mu);
        for (long long r = row_begin; r < row_end; ++r)
            if ((r % stride) == phase && e->h_flags[(size_t)r] == 0) qrows.push_back(r);
    }
    for (size_t t0 = 0; t0 < qrows.size(); t0 += (size_t)T) {
        const int ns = (int)std::min<size_t>((size_t)T, qrows.size() - t0);
        long long min_row = n_store;
        {
            std::lock_guard<std::mutex> g(e->store_mu);
            for (int s = 0; s < ns; ++s) {
                const long long r = qrows[t0 + s];
                QuerySlot& qs = e->h_slots[s];
                qs.hdr = e->h_hdr[(size_t)r]; qs.row = r; qs.exclude = r; qs.skip_le = upper_only ? r : -1; qs.pad = 0;
                e->slot_blocks[s] = e->h_nblk[(size_t)r];
                min_row = std::min(min_row, r);
            }
        }
        const long long cbegin = upper_only ? min_row + 1 : 0;
        CU(e, cudaMemsetAsync(e->d_count, 0, FN3_OUT_HDR, e->q_stream));
        if ((rc = enqueue_query(e, ht, nullptr, cbegin, n_store - cbegin, ns, ceiling, false))) return rc;
        CU(e, cudaMemcpyAsync(e->h_count, e->d_count, FN3_OUT_HDR, cudaMemcpyDeviceToHost, e->q_stream));
        CU(e, cudaStreamSynchronize(e->q_stream));
        e->d2h_bytes += FN3_OUT_HDR;
        const long long cnt = (long long)e->h_count[0];
        if (cnt > 0) {
            if (total + cnt <= cap)
                CU(e, cudaMemcpy(out + total, e->d_out, (size_t)cnt * sizeof(EdgeRec), cudaMemcpyDeviceToHost));
            e->d2h_bytes += cnt * (long long)sizeof(EdgeRec);
            total += cnt;
        }
    }
    *n_out = total;
    if (total > cap) return set_err(e, FN3_ECAPACITY, "output buffer holds %lld edges, %lld needed", (long long)cap, total);
    return FN3_OK;
}


// ---- ingest ----------------------------------------------------------------------------------

extern "C" int fn3_set_reference(fn3_engine* e, const char* reference, const int32_t* excluded, int64_t n_excluded) {
    if (!e || !reference || n_excluded < 0 || (n_excluded > 0 && !excluded))
        return set_err(e, FN3_EARG, "fn3_set_reference: bad arguments");
    std::lock_guard<std::mutex> g(e->ingest_mu);
    CU(e, cudaSetDevice(e->device));
    const long long per_cta = (long long)FN3_CMP_THREADS * FN3_CMP_PER_THREAD;
    e->cmp_ctas = (int)((e->L + per_cta - 1) / per_cta);
    e->Lpad = (long long)e->cmp_ctas * per_cta;
    std::vector<uint8_t> rm((size_t)e->Lpad, 0);
    for (long long i = 0; i < e->L; ++i) {
        const char c = reference[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T')
            return set_err(e, FN3_EARG, "fn3_set_reference: reference holds a character other than ACGT at %lld", i);
        rm[(size_t)i] = (uint8_t)c;
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
    CU(e, cudaMemsetAsync(e->d_seq, 0, (size_t)e->Lpad, e->in_stream));
    CU(e, cudaStreamSynchronize(e->in_stream));
    return FN3_OK;
}

extern "C" int fn3_compress(fn3_engine* e, const char* sequence, int32_t* pos, int64_t pos_cap, int32_t off[7],
                            char* m_chars, int64_t m_cap, int64_t* n_pos) {
    if (!e || !sequence || !off || !n_pos) return set_err(e, FN3_EARG, "fn3_compress: bad arguments");
    std::lock_guard<std::mutex> g(e->ingest_mu);
    if (!e->d_refmask) return set_err(e, FN3_EARG, "fn3_compress: call fn3_set_reference first");
    CU(e, cudaSetDevice(e->device));
    cudaStream_t st = e->in_stream;
    memcpy(e->h_seq, sequence, (size_t)e->L);
    CU(e, cudaMemcpyAsync(e->d_seq, e->h_seq, (size_t)e->L, cudaMemcpyHostToDevice, st));
    k_compress_count<<<e->cmp_ctas, FN3_CMP_THREADS, 0, st>>>(e->d_seq, e->d_refmask, e->d_ccounts);
    k_compress_scan<<<1, 32, 0, st>>>(e->d_ccounts, e->d_ctotals, e->cmp_ctas);
    k_compress_write<<<e->cmp_ctas, FN3_CMP_THREADS, 0, st>>>(e->d_seq, e->d_refmask, e->d_ccounts, e->d_ctotals,
                                                               e->d_cpos, e->d_mchars);
    e->launches += 3;
    CU(e, cudaGetLastError());
    CU(e, cudaMemcpyAsync(e->h_ctotals, e->d_ctotals, 6 * 4, cudaMemcpyDeviceToHost, st));
    CU(e, cudaStreamSynchronize(st));
    long long total = 0;
    off[0] = 0;
    for (int c = 0; c < 6; ++c) { total += e->h_ctotals[c]; off[c + 1] = (int32_t)total; }
    const long long n_m = e->h_ctotals[5];
    *n_pos = total;
    if (total > pos_cap || n_m > m_cap || (total > 0 && !pos) || (n_m > 0 && !m_chars))
        return set_err(e, FN3_ECAPACITY, "fn3_compress: need room for %lld positions and %lld mixed characters", total, n_m);
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
    if (total) memcpy(pos, e->h_cpos, (size_t)total * 4);
    if (n_m) memcpy(m_chars, e->h_mchars, (size_t)n_m);
    return FN3_OK;
}

// ---- measurement hook ----------------------------------------------------------------------

extern "C" int fn3_profile_queries(fn3_engine* e, const int64_t* query_rows, int32_t n, int32_t ceiling, int32_t flush_l2,
                                   float* ms_build, float* ms_compare, int64_t* n_hits, int64_t* bytes_touched) {
    if (!e || !query_rows || n < 1) return set_err(e, FN3_EARG, "fn3_profile_queries: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    HdrTable ht; long long n_store;
    snapshot_store(e, ht, n_store);
    int rc;
    if ((rc = ensure_out(e, n_store))) return rc;
    if (flush_l2 && !e->d_flush) {
        e->flush_n = (256ll << 20) / 16;
        CU(e, cudaMalloc(&e->d_flush, (size_t)e->flush_n * 16));
    }
    cudaStream_t st = e->q_stream;
    std::vector<cudaEvent_t> evs((size_t)n * 3, nullptr);
    for (auto& ev : evs) CU(e, cudaEventCreate(&ev));
    unsigned long long* d_pc = nullptr;                       // per-step counters: [hits, touched, -, -]
    CU(e, cudaMalloc(&d_pc, (size_t)n * FN3_OUT_HDR));
    CU(e, cudaMemsetAsync(d_pc, 0, (size_t)n * FN3_OUT_HDR, st));
    unsigned long long* saved = e->d_count;
    // all query slots go to the device in one copy; then the whole batch is queued without any host
    // synchronisation, events bracketing each step's pieces
    std::vector<QuerySlot> hs((size_t)n);
    uint32_t maxw = 1;
    for (int it = 0; it < n; ++it) {
        RowHdr qh;
        uint32_t nb = 0;
        if ((rc = stored_hdr(e, query_rows[it], &qh, "fn3_profile_queries", &nb))) break;
        QuerySlot& s = hs[(size_t)it];
        s.hdr = qh; s.row = query_rows[it]; s.skip_le = -1; s.exclude = query_rows[it]; s.pad = 0;
        maxw = std::max(maxw, nb);
    }
    QuerySlot* d_ps = nullptr;
    if (rc == FN3_OK) {
        CU(e, cudaMalloc(&d_ps, (size_t)n * sizeof(QuerySlot)));
        CU(e, cudaMemcpyAsync(d_ps, hs.data(), (size_t)n * sizeof(QuerySlot), cudaMemcpyHostToDevice, st));
        CU(e, cudaStreamSynchronize(st));
        e->slot_blocks[0] = maxw;             // sizes the shared-memory staging of every step
    }
    for (int it = 0; it < n && rc == FN3_OK; ++it) {
        if (flush_l2) { k_fill<<<e->sm_count * 4, 512, 0, st>>>(e->d_flush, e->flush_n, (uint32_t)it); e->launches += 1; }
        e->d_count = d_pc + (size_t)it * (FN3_OUT_HDR / sizeof(unsigned long long));
        cudaEventRecord(evs[(size_t)it * 3 + 0], st);
        if ((rc = enqueue_query(e, ht, nullptr, 0, n_store, 1, ceiling, false, d_ps + it, evs[(size_t)it * 3 + 1]))) break;
        cudaEventRecord(evs[(size_t)it * 3 + 2], st);
    }
    e->d_count = saved;
    cudaError_t sync = cudaStreamSynchronize(st);
    std::vector<unsigned long long> pc((size_t)n * 4, 0);
    if (rc == FN3_OK && sync == cudaSuccess) {
        for (int it = 0; it < n; ++it) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, evs[(size_t)it * 3 + 0], evs[(size_t)it * 3 + 1]);
            cudaEventElapsedTime(&b, evs[(size_t)it * 3 + 1], evs[(size_t)it * 3 + 2]);
            if (ms_build) ms_build[it] = a;
            if (ms_compare) ms_compare[it] = b;
        }
        cudaMemcpy(pc.data(), d_pc, (size_t)n * FN3_OUT_HDR, cudaMemcpyDeviceToHost);
        if (n_hits) for (int it = 0; it < n; ++it) n_hits[it] = (int64_t)pc[(size_t)it * 4];
    }
    if (rc == FN3_OK && sync == cudaSuccess && bytes_touched) {
        // instrumented, untimed pass: bytes of headers + row data each compare launch actually loads
        cudaMemsetAsync(d_pc, 0, (size_t)n * FN3_OUT_HDR, st);
        for (int it = 0; it < n && rc == FN3_OK; ++it) {
            e->d_count = d_pc + (size_t)it * (FN3_OUT_HDR / sizeof(unsigned long long));
            if ((rc = enqueue_query(e, ht, nullptr, 0, n_store, 1, ceiling, true, d_ps + it))) break;
        }
        e->d_count = saved;
        cudaStreamSynchronize(st);
        cudaMemcpy(pc.data(), d_pc, (size_t)n * FN3_OUT_HDR, cudaMemcpyDeviceToHost);
        for (int it = 0; it < n; ++it) bytes_touched[it] = (int64_t)pc[(size_t)it * 4 + 1];
    }
    cudaFree(d_ps);
    for (auto ev : evs) cudaEventDestroy(ev);
    cudaFree(d_pc);
    if (rc != FN3_OK) return rc;
    if (sync != cudaSuccess) return set_err(e, FN3_ECUDA, "fn3_profile_queries: %s", cudaGetErrorString(sync));
    return FN3_OK;
}
