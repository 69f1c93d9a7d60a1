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
//   query map  : uint8 code per genome position (0 ref,1..4 ACGT,5 N,6 M), L2-resident
//   query rank : per 32 positions {maskDef, maskU, maskM, cumDef, cumU, cumM, 0, 0} (32 B)
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

struct __align__(32) RankBlk {
    uint32_t mDef, mU, mM;     // bit i: position 32*b+i of the query is ACGT-nonref / N-or-M / M
    uint32_t cDef, cU, cM;     // number of such positions strictly below 32*b
    uint32_t pad0, pad1;
};
static_assert(sizeof(RankBlk) == 32, "RankBlk must be one 32-byte sector");

struct __align__(32) EdgeRec {   // device-side survivor record (== fn3_edge)
    long long row1, row2;
    int dist, n1, n2, nboth;
};
static_assert(sizeof(EdgeRec) == sizeof(fn3_edge), "EdgeRec/fn3_edge mismatch");

#define FN3_MAX_HDR_CHUNKS 128
#define FN3_MAX_SLOTS 16

struct HdrTable {
    const RowHdr* chunk[FN3_MAX_HDR_CHUNKS];
    int shift;                 // rows per chunk = 1 << shift
};

struct QuerySlot {             // one query of a tile
    RowHdr hdr;                // where the query's own row lives (stored row or scratch)
    long long row;             // store row of the query, -1 if not stored
    long long skip_le;         // candidates with row <= skip_le are skipped (upper triangle), -1 = none
    long long exclude;         // candidate row to skip (self), -1 = none
    long long pad;
};

struct CompareParams {
    HdrTable hdr;
    const long long* rows;     // optional explicit candidate list
    long long row_begin;       // else candidates row_begin .. row_begin+n-1
    long long n;
    const uint8_t* qmap;       // slot s at qmap + s*map_stride
    const RankBlk* qrank;      // slot s at qrank + s*rank_stride
    const QuerySlot* slots;    // device array [n_slots]
    long long map_stride;      // bytes
    long long rank_stride;     // RankBlk entries
    long long rank_last;       // index of the totals entry inside a slot's rank array
    EdgeRec* out;
    unsigned long long* out_count;
    unsigned long long* touched;   // instrumented variant only
    long long out_cap;
    int ceiling;
    int pos_bits;
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint4 ld_stream_v4(const uint32_t* p) {
    // candidate row data is read exactly once per comparison: keep it out of L1 so that the
    // query map / rank gathers own the L1.
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void ld_rank(const RankBlk* p, uint4& a, uint2& b) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    a = __ldg(q);
    b = __ldg(reinterpret_cast<const uint2*>(q + 1));
}

// number of set bits of m at bit positions < k   (k in 0..31)
__device__ __forceinline__ uint32_t below(uint32_t m, uint32_t k) { return __popc(m & ((1u << k) - 1u)); }

// ------------------------------------------------------------------------------------------
// query build: scatter the query's row into the code map, then derive the rank structure
// ------------------------------------------------------------------------------------------

__global__ void k_scatter_query(const QuerySlot* __restrict__ slots, uint8_t* __restrict__ qmap,
                                long long map_stride, int pos_bits) {
    const QuerySlot& qs = slots[blockIdx.y];
    const uint32_t* d = reinterpret_cast<const uint32_t*>(qs.hdr.data);
    if (d == nullptr) return;
    uint8_t* map = qmap + (long long)blockIdx.y * map_stride;
    const uint32_t e0 = qs.hdr.end[0], e1 = qs.hdr.end[1], e2 = qs.hdr.end[2], e3 = qs.hdr.end[3],
                   e4 = qs.hdr.end[4], e5 = qs.hdr.end[5];
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < e5; i += gridDim.x * blockDim.x) {
        uint32_t w = d[i];
        if (i < e3) {
            uint8_t code = 1 + (i >= e0) + (i >= e1) + (i >= e2);
            map[w] = code;
        } else {
            uint8_t code = (i < e4) ? 5 : 6;
            uint32_t s = w & pmask, len = (pos_bits >= 32) ? 1u : ((w >> pos_bits) + 1u);
            for (uint32_t j = 0; j < len; ++j) map[s + j] = code;
        }
    }
}

// pass A: 32 positions per thread -> masks + CTA-local exclusive counts; CTA totals to `partial`
__global__ void __launch_bounds__(1024) k_rank_a(const uint8_t* __restrict__ qmap, RankBlk* __restrict__ qrank,
                                                 uint3* __restrict__ partial, long long map_stride,
                                                 long long rank_stride, long long n_blk, int ctas_per_slot) {
    const int slot = blockIdx.y;
    const uint8_t* map = qmap + (long long)slot * map_stride;
    RankBlk* rank = qrank + (long long)slot * rank_stride;
    const long long b = (long long)blockIdx.x * 1024 + threadIdx.x;
    uint32_t mD = 0, mU = 0, mM = 0;
    if (b < n_blk) {
        const uint4* src = reinterpret_cast<const uint4*>(map + b * 32);
        uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
        uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t c = (w[k] >> (8 * j)) & 0xffu;
                uint32_t bit = 1u << (k * 4 + j);
                if (c - 1u < 4u) mD |= bit;
                if (c >= 5u) mU |= bit;
                if (c == 6u) mM |= bit;
            }
        }
    }
    uint32_t cD = __popc(mD), cU = __popc(mU), cM = __popc(mM);
    // inclusive warp scan
    uint32_t sD = cD, sU = cU, sM = cM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t tD = __shfl_up_sync(0xffffffffu, sD, o), tU = __shfl_up_sync(0xffffffffu, sU, o),
                 tM = __shfl_up_sync(0xffffffffu, sM, o);
        if (lane >= o) { sD += tD; sU += tU; sM += tM; }
    }
    __shared__ uint32_t wD[32], wU[32], wM[32];
    if (lane == 31) { wD[warp] = sD; wU[warp] = sU; wM[warp] = sM; }
    __syncthreads();
    if (warp == 0) {
        uint32_t aD = wD[lane], aU = wU[lane], aM = wM[lane];
        uint32_t iD = aD, iU = aU, iM = aM;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t tD = __shfl_up_sync(0xffffffffu, iD, o), tU = __shfl_up_sync(0xffffffffu, iU, o),
                     tM = __shfl_up_sync(0xffffffffu, iM, o);
            if (lane >= o) { iD += tD; iU += tU; iM += tM; }
        }
        wD[lane] = iD - aD; wU[lane] = iU - aU; wM[lane] = iM - aM;   // exclusive warp offsets
        if (lane == 31) partial[slot * ctas_per_slot + blockIdx.x] = make_uint3(iD, iU, iM);
    }
    __syncthreads();
    if (b < n_blk) {
        uint4 a = make_uint4(mD, mU, mM, wD[warp] + sD - cD);
        uint4 c = make_uint4(wU[warp] + sU - cU, wM[warp] + sM - cM, 0u, 0u);
        uint4* dst = reinterpret_cast<uint4*>(rank + b);
        dst[0] = a; dst[1] = c;
    }
}

// pass B: add the sum of all preceding CTAs' totals
__global__ void __launch_bounds__(1024) k_rank_b(RankBlk* __restrict__ qrank, const uint3* __restrict__ partial,
                                                 long long rank_stride, long long n_blk, int ctas_per_slot) {
    const int slot = blockIdx.y;
    if (blockIdx.x == 0) return;
    RankBlk* rank = qrank + (long long)slot * rank_stride;
    const uint3* part = partial + slot * ctas_per_slot;
    uint32_t aD = 0, aU = 0, aM = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += 1024) { uint3 p = part[i]; aD += p.x; aU += p.y; aM += p.z; }
    __shared__ uint32_t sD[32], sU[32], sM[32];
    aD = __reduce_add_sync(0xffffffffu, aD); aU = __reduce_add_sync(0xffffffffu, aU); aM = __reduce_add_sync(0xffffffffu, aM);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sD[warp] = aD; sU[warp] = aU; sM[warp] = aM; }
    __syncthreads();
    uint32_t tD = __reduce_add_sync(0xffffffffu, sD[lane]), tU = __reduce_add_sync(0xffffffffu, sU[lane]),
             tM = __reduce_add_sync(0xffffffffu, sM[lane]);
    const long long b = (long long)blockIdx.x * 1024 + threadIdx.x;
    if (b < n_blk) { rank[b].cDef += tD; rank[b].cU += tU; rank[b].cM += tM; }
}

// ------------------------------------------------------------------------------------------
// the compare kernel: one warp per (query slot, candidate row)
// ------------------------------------------------------------------------------------------
//
// Streaming identity (SURVEY.md §8 "Spec"), with V = definite non-reference positions,
// U = N u M positions, q = query (seq1), c = candidate (seq2):
//   S1    = #{p in V_c : code_q[p] is ref, or ACGT and != code_c[p]}         (monotone)
//   dist  = S1 + |V_q| - |V_c n V_q| - |U_c n V_q|
//   n1    = |U_q|,  n2 = |N_c| + |M_q| - |N_c n M_q|,  nboth = n1 + |N_c| - |N_c n U_q|
// S1 alone is a lower bound of dist, so a pair leaves as soon as S1 > ceiling.

template <bool COUNT>
__global__ void __launch_bounds__(256) k_compare(const CompareParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int slot = blockIdx.y;
    const QuerySlot qs = p.slots[slot];
    if (qs.hdr.data == 0ull) return;                       // invalid query: no distances at all
    const uint8_t* __restrict__ qmap = p.qmap + (long long)slot * p.map_stride;
    const RankBlk* __restrict__ qrank = p.qrank + (long long)slot * p.rank_stride;
    const uint32_t nVq = qrank[p.rank_last].cDef, nUq = qrank[p.rank_last].cU, nMq = qrank[p.rank_last].cM;
    const int k = p.ceiling;
    const int pos_bits = p.pos_bits;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    const long long hmask = (1ll << p.hdr.shift) - 1;
    unsigned long long touched = 0;

    for (long long i = warp0; i < p.n; i += nwarps) {
        const long long row = p.rows ? p.rows[i] : p.row_begin + i;
        if (row == qs.exclude || row <= qs.skip_le) continue;
        const uint4* hp = reinterpret_cast<const uint4*>(p.hdr.chunk[row >> p.hdr.shift] + (row & hmask));
        const uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        if (COUNT) touched += 32;
        const unsigned long long daddr = ((unsigned long long)h0.y << 32) | h0.x;
        if (daddr == 0ull) continue;                       // invalid or removed candidate
        const uint32_t* __restrict__ d = reinterpret_cast<const uint32_t*>(daddr);
        const uint32_t e0 = h0.z, e1 = h0.w, e2 = h1.x, e3 = h1.y, e4 = h1.z, e5 = h1.w;

        int S1 = 0;                                        // warp-uniform running mismatch count
        int qdefV = 0, qdefU = 0, nNc = 0, nNU = 0, nNM = 0;   // per-lane partials
        bool dead = false;
        for (uint32_t base = 0; base < e5; base += 128) {
            const uint32_t idx = base + lane * 4;
            uint4 w4 = make_uint4(0, 0, 0, 0);
            if (idx < e5) { w4 = ld_stream_v4(d + idx); if (COUNT) touched += 16; }
            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
            int mism = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ii = idx + j;
                if (ii < e3) {
                    const uint32_t b = 1u + (ii >= e0) + (ii >= e1) + (ii >= e2);
                    const uint32_t q = __ldg(qmap + w[j]);
                    mism += (q < 5u) & (q != b);
                    qdefV += (q - 1u < 4u);
                } else if (ii < e5) {
                    const uint32_t s = w[j] & pmask;
                    const uint32_t len = (pos_bits >= 32) ? 1u : ((w[j] >> pos_bits) + 1u);
                    const uint32_t e = s + len;
                    uint4 a0, a1; uint2 b0, b1;
                    ld_rank(qrank + (s >> 5), a0, b0);
                    ld_rank(qrank + (e >> 5), a1, b1);
                    const uint32_t ks = s & 31u, ke = e & 31u;
                    qdefU += (int)(a1.w + below(a1.x, ke)) - (int)(a0.w + below(a0.x, ks));
                    if (ii < e4) {                          // an N run of the candidate
                        nNc += (int)len;
                        nNU += (int)(b1.x + below(a1.y, ke)) - (int)(b0.x + below(a0.y, ks));
                        nNM += (int)(b1.y + below(a1.z, ke)) - (int)(b0.y + below(a0.z, ks));
                    }
                }
            }
            if (base < e3) {                                // warp-uniform: this step touched V words
                S1 += __reduce_add_sync(0xffffffffu, mism);
                if (S1 > k) { dead = true; break; }
            }
        }
        if (dead) continue;
        qdefV = __reduce_add_sync(0xffffffffu, qdefV);
        qdefU = __reduce_add_sync(0xffffffffu, qdefU);
        const int dist = S1 + (int)nVq - qdefV - qdefU;
        if (dist > k) continue;
        nNc = __reduce_add_sync(0xffffffffu, nNc);
        nNU = __reduce_add_sync(0xffffffffu, nNU);
        nNM = __reduce_add_sync(0xffffffffu, nNM);
        if (lane == 0) {
            const unsigned long long at = atomicAdd(p.out_count, 1ull);
            if ((long long)at < p.out_cap) {
                EdgeRec r;
                r.row1 = qs.row; r.row2 = row; r.dist = dist;
                r.n1 = (int)nUq; r.n2 = nNc + (int)nMq - nNM; r.nboth = (int)nUq + nNc - nNU;
                p.out[at] = r;
            }
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
    std::vector<RowHdr*> hdr_chunks;        // device
    std::vector<RowHdr> h_hdr;              // host mirror (published rows)
    std::vector<uint8_t> h_flags;           // bit0 invalid, bit1 removed
    std::vector<uint32_t> h_npos;           // canonical position count per row
    long long n_rows = 0;                   // published
    long long n_live = 0, n_invalid = 0;
    long long store_words = 0, store_bytes = 0, positions = 0;

    // staging
    uint32_t* stage = nullptr; size_t stage_words = 0;     // pinned
    RowHdr* stage_hdr = nullptr; size_t stage_hdr_n = 0;   // pinned

    // query state
    int n_slots = 8;
    long long map_stride = 0, rank_stride = 0, n_blk = 0; int rank_ctas = 0;
    uint8_t* qmap = nullptr; RankBlk* qrank = nullptr; uint3* partial = nullptr;
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
    RowHdr* d_scratch_hdr = nullptr;         // [2]
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
        CU(e, cudaHostAlloc(&e->stage_hdr, want * sizeof(RowHdr), cudaHostAllocDefault));
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

    // query state: map padded so that rank block n_blk-1 (the totals entry) reads zeros
    e->n_blk = (genome_length + 31) / 32 + 1;
    e->map_stride = ((e->n_blk * 32 + 255) / 256) * 256;
    e->rank_stride = e->n_blk;
    e->rank_ctas = (int)((e->n_blk + 1023) / 1024);
    if ((s = cudaMalloc(&e->qmap, (size_t)e->map_stride * e->n_slots)) != cudaSuccess) return fail("cudaMalloc(qmap)", s);
    if ((s = cudaMalloc(&e->qrank, (size_t)e->rank_stride * sizeof(RankBlk) * e->n_slots)) != cudaSuccess) return fail("cudaMalloc(qrank)", s);
    if ((s = cudaMalloc(&e->partial, (size_t)e->rank_ctas * e->n_slots * sizeof(uint3))) != cudaSuccess) return fail("cudaMalloc(partial)", s);
    if ((s = cudaMalloc(&e->d_slots, sizeof(QuerySlot) * FN3_MAX_SLOTS)) != cudaSuccess) return fail("cudaMalloc(slots)", s);
    if ((s = cudaHostAlloc(&e->h_slots, sizeof(QuerySlot) * FN3_MAX_SLOTS, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc(slots)", s);
    if (ensure_out(e, 4096) != FN3_OK || ensure_hout(e, 4096) != FN3_OK) {
        set_err(nullptr, FN3_ENOMEM, "fn3_create: survivor buffers: %s", e->err.c_str());
        fn3_destroy(e);
        return nullptr;
    }
    if ((s = cudaMalloc(&e->d_scratch_hdr, 2 * sizeof(RowHdr))) != cudaSuccess) return fail("cudaMalloc(scratch hdr)", s);
    if ((s = cudaMemset(e->qmap, 0, (size_t)e->map_stride * e->n_slots)) != cudaSuccess) return fail("cudaMemset", s);
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
    cudaFree(e->qmap); cudaFree(e->qrank); cudaFree(e->partial); cudaFree(e->d_slots);
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
static long long row_words_needed(fn3_engine* e, const int32_t* pos, const long long off[7], int invalid) {
    if (invalid) return 0;
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
            if (b >= 4) {
                if (i && p[i] == p[i - 1] + 1 && run_len < (long long)e->max_run) ++run_len;
                else { ++runs; run_len = 1; }
            }
        }
        if (b >= 4) words += runs;
    }
    if (words >= (1ll << 31)) { set_err(e, FN3_EARG, "sample too large"); return -1; }
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
        RowHdr* p = nullptr;
        size_t bytes = sizeof(RowHdr) << e->hdr_shift;
        cudaError_t s = cudaMalloc(&p, bytes);
        if (s != cudaSuccess) return set_err(e, FN3_ENOMEM, "cudaMalloc of a header chunk failed: %s", cudaGetErrorString(s));
        cudaMemset(p, 0, bytes);
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
    for (long long i = 0; i < n; ++i) {
        long long o[7];
        for (int b = 0; b < 7; ++b) o[b] = off6[6 * i + b];
        long long w = row_words_needed(e, pos, o, invalid ? invalid[i] : 0);
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
            RowHdr& h = e->stage_hdr[r - i];
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
            w += need;
        }
        CU(e, cudaMemcpyAsync(dbase, e->stage, batch_words * 4, cudaMemcpyHostToDevice, e->copy_stream));
        e->h2d_bytes += (long long)batch_words * 4 + (j - i) * (long long)sizeof(RowHdr);
        // headers: may straddle header chunks
        long long r = i;
        while (r < j) {
            const long long row = first + r;
            rc = ensure_hdr_chunk(e, row);
            if (rc) return rc;
            const long long in_chunk = row & ((1ll << e->hdr_shift) - 1);
            const long long take = std::min<long long>(j - r, (1ll << e->hdr_shift) - in_chunk);
            CU(e, cudaMemcpyAsync(e->hdr_chunks[(size_t)(row >> e->hdr_shift)] + in_chunk, e->stage_hdr + (r - i),
                                  (size_t)take * sizeof(RowHdr), cudaMemcpyHostToDevice, e->copy_stream));
            r += take;
        }
        CU(e, cudaStreamSynchronize(e->copy_stream));
        // publish host mirror
        for (long long r2 = i; r2 < j; ++r2) {
            const bool inv = invalid && invalid[r2];
            e->h_hdr.push_back(e->stage_hdr[r2 - i]);
            e->h_flags.push_back(inv ? 1 : 0);
            const uint32_t np = inv ? 0u : (uint32_t)(off6[6 * r2 + 6] - off6[6 * r2]);
            e->h_npos.push_back(np);
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

// enqueue: slots upload + map clear + scatter + rank build for `ns` slots (h_slots filled by caller)
// (`resident` != nullptr: the slots are already on the device — measurement hook)
static int enqueue_query_build(fn3_engine* e, int ns, const QuerySlot* resident = nullptr) {
    cudaStream_t st = e->q_stream;
    const QuerySlot* d_slots = resident ? resident : e->d_slots;
    if (!resident) {
        CU(e, cudaMemcpyAsync(e->d_slots, e->h_slots, sizeof(QuerySlot) * ns, cudaMemcpyHostToDevice, st));
        e->h2d_bytes += (long long)sizeof(QuerySlot) * ns;
    }
    CU(e, cudaMemsetAsync(e->qmap, 0, (size_t)e->map_stride * ns, st));
    uint32_t maxw = 1;
    for (int s = 0; s < ns; ++s) maxw = std::max(maxw, e->h_slots[s].hdr.end[5]);
    dim3 gs((unsigned)std::min<uint32_t>((maxw + 255) / 256, 1024u), (unsigned)ns);
    k_scatter_query<<<gs, 256, 0, st>>>(d_slots, e->qmap, e->map_stride, e->pos_bits);
    dim3 gr((unsigned)e->rank_ctas, (unsigned)ns);
    k_rank_a<<<gr, 1024, 0, st>>>(e->qmap, e->qrank, e->partial, e->map_stride, e->rank_stride, e->n_blk, e->rank_ctas);
    k_rank_b<<<gr, 1024, 0, st>>>(e->qrank, e->partial, e->rank_stride, e->n_blk, e->rank_ctas);
    e->launches += 3;
    CU(e, cudaGetLastError());
    return FN3_OK;
}

static int enqueue_compare(fn3_engine* e, const HdrTable& ht, const long long* d_rows, long long row_begin, long long n,
                           int ns, int ceiling, bool count_bytes, const QuerySlot* resident = nullptr) {
    if (n <= 0) return FN3_OK;
    CompareParams p;
    p.hdr = ht; p.rows = d_rows; p.row_begin = row_begin; p.n = n;
    p.qmap = e->qmap; p.qrank = e->qrank; p.slots = resident ? resident : e->d_slots;
    p.map_stride = e->map_stride; p.rank_stride = e->rank_stride; p.rank_last = e->n_blk - 1;
    p.out = e->d_out; p.out_count = e->d_count; p.touched = e->d_count + 1; p.out_cap = e->out_cap;
    p.ceiling = ceiling; p.pos_bits = e->pos_bits;
    const long long warps_per_cta = 8;
    long long ctas = (n + warps_per_cta - 1) / warps_per_cta;
    const long long max_ctas = (long long)e->sm_count * 8;
    if (ctas > max_ctas) ctas = max_ctas;
    dim3 g((unsigned)ctas, (unsigned)ns);
    if (count_bytes) k_compare<true><<<g, 256, 0, e->q_stream>>>(p);
    else k_compare<false><<<g, 256, 0, e->q_stream>>>(p);
    e->launches += 1;
    CU(e, cudaGetLastError());
    return FN3_OK;
}

// upload a non-stored sample into scratch slot `which` (0/1); returns its header
static int upload_scratch(fn3_engine* e, int which, const int32_t* pos, const int32_t off[7], int invalid, RowHdr* out) {
    RowHdr h; memset(&h, 0, sizeof h);
    if (invalid) { *out = h; return FN3_OK; }
    long long o[7];
    for (int b = 0; b < 7; ++b) o[b] = off[b];
    long long w = row_words_needed(e, pos, o, 0);
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
    if ((rc = enqueue_query_build(e, 1))) return rc;
    if ((rc = enqueue_compare(e, ht, d_rows, 0, n, 1, ceiling, false))) return rc;
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

static int stored_hdr(fn3_engine* e, long long row, RowHdr* h, const char* who) {
    std::lock_guard<std::mutex> g(e->store_mu);
    if (row < 0 || row >= e->n_rows) return set_err(e, FN3_ENOTFOUND, "%s: row %lld does not exist", who, row);
    if (e->h_flags[(size_t)row] & 2) return set_err(e, FN3_ENOTFOUND, "%s: row %lld was removed", who, row);
    *h = e->h_hdr[(size_t)row];
    return FN3_OK;
}

extern "C" int fn3_one_vs_all(fn3_engine* e, int64_t query_row, int32_t ceiling, const int64_t* rows, int64_t n_rows,
                              fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || (rows && n_rows < 0)) return set_err(e, FN3_EARG, "fn3_one_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    RowHdr qh; int rc;
    if ((rc = stored_hdr(e, query_row, &qh, "fn3_one_vs_all"))) return rc;
    HdrTable ht; long long n_store;
    snapshot_store(e, ht, n_store);
    QuerySlot& s = e->h_slots[0];
    s.hdr = qh; s.row = query_row; s.skip_le = -1; s.exclude = query_row; s.pad = 0;
    return run_one_vs_all(e, ceiling, rows, n_rows, n_store, ht, out, cap, n_out);
}

extern "C" int fn3_lists_vs_all(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                                const int64_t* rows, int64_t n_rows, fn3_hit* out, int64_t cap, int64_t* n_out) {
    if (!e || !n_out || (cap > 0 && !out) || (!invalid && !off)) return set_err(e, FN3_EARG, "fn3_lists_vs_all: bad arguments");
    std::lock_guard<std::mutex> q(e->query_mu);
    CU(e, cudaSetDevice(e->device));
    RowHdr qh; int rc;
    if ((rc = upload_scratch(e, 0, pos, off, invalid, &qh))) return rc;
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
    if ((rc = stored_hdr(e, row1, &qh, "fn3_pair"))) return rc;
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
    if ((rc = upload_scratch(e, 0, pos1, off1, invalid1, &qh))) return rc;
    if ((rc = upload_scratch(e, 1, pos2, off2, invalid2, &ch))) return rc;
    if (qh.data == 0ull || ch.data == 0ull) return FN3_OK;   // either invalid: None (:438-442)
    // scratch header table: one chunk holding the candidate header at row 0
    CU(e, cudaMemcpyAsync(e->d_scratch_hdr, &ch, sizeof ch, cudaMemcpyHostToDevice, e->q_stream));
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
                min_row = std::min(min_row, r);
            }
        }
        const long long cbegin = upper_only ? min_row + 1 : 0;
        CU(e, cudaMemsetAsync(e->d_count, 0, FN3_OUT_HDR, e->q_stream));
        if ((rc = enqueue_query_build(e, ns))) return rc;
        if ((rc = enqueue_compare(e, ht, nullptr, cbegin, n_store - cbegin, ns, ceiling, false))) return rc;
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
    CU(e, cudaMemset(e->d_seq, 0, (size_t)e->Lpad));
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
        if ((rc = stored_hdr(e, query_rows[it], &qh, "fn3_profile_queries"))) break;
        QuerySlot& s = hs[(size_t)it];
        s.hdr = qh; s.row = query_rows[it]; s.skip_le = -1; s.exclude = query_rows[it]; s.pad = 0;
        maxw = std::max(maxw, qh.end[5]);
    }
    QuerySlot* d_ps = nullptr;
    if (rc == FN3_OK) {
        CU(e, cudaMalloc(&d_ps, (size_t)n * sizeof(QuerySlot)));
        CU(e, cudaMemcpyAsync(d_ps, hs.data(), (size_t)n * sizeof(QuerySlot), cudaMemcpyHostToDevice, st));
        CU(e, cudaStreamSynchronize(st));
        e->h_slots[0].hdr.end[5] = maxw;      // sizes the scatter grid
    }
    for (int it = 0; it < n && rc == FN3_OK; ++it) {
        if (flush_l2) { k_fill<<<e->sm_count * 4, 512, 0, st>>>(e->d_flush, e->flush_n, (uint32_t)it); e->launches += 1; }
        e->d_count = d_pc + (size_t)it * (FN3_OUT_HDR / sizeof(unsigned long long));
        cudaEventRecord(evs[(size_t)it * 3 + 0], st);
        if ((rc = enqueue_query_build(e, 1, d_ps + it))) break;
        cudaEventRecord(evs[(size_t)it * 3 + 1], st);
        if ((rc = enqueue_compare(e, ht, nullptr, 0, n_store, 1, ceiling, false, d_ps + it))) break;
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
            if ((rc = enqueue_query_build(e, 1, d_ps + it))) break;
            if ((rc = enqueue_compare(e, ht, nullptr, 0, n_store, 1, ceiling, true, d_ps + it))) break;
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
