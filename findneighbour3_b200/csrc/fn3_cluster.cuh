// fn3_cluster.cuh — per-position call counts over a set of rows, and what the reference derives from them:
//   * consensus()   (src/seqComparer.py:460-501): positions carried by at least `min_votes` of the samples, per list
//   * _msa steps 1-4 (src/seqComparer.py:895-950): the variant sites of a cluster, the aligned characters of every
//     member at those sites, and the N / M counts in and out of the alignment
//   * estimate_expected_unk_sites (:664-689): N / M counts of other samples at the same sites
// Included by fn3_engine.cu only.  All integer work; bound by HBM (the vote planes are genome-sized).
#pragma once
#include "fn3_types.h"

#define FN3_SEL_THREADS 256
#define FN3_SEL_PER_THREAD 16          // a CTA covers 4096 positions

// ---- votes ----------------------------------------------------------------------------------------------
// planes: cnt[b * Lpad + p] = number of given samples that carry position p in list b (0..3 = A,C,G,T; 4 = N).
// A sample given twice votes twice (the reference counts the objects of its list one by one, :483-490).

// stored rows: one CTA per row, threads stride the row's words; N run words are expanded (want_n).
// (One warp per row — 16 dependent load->atomic rounds per lane, 7 CTAs for a 50-row cluster — took 14.7 us: latency.)
#define FN3_VOTE_THREADS 256
__global__ void __launch_bounds__(FN3_VOTE_THREADS) k_votes_rows(const long long* __restrict__ rows, long long n, const HdrTable ht,
                                                                 uint32_t* __restrict__ cnt, long long Lpad, int want_n, int pos_bits) {
    const uint32_t tid = threadIdx.x;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {
        int tl;
        const unsigned char* tile = tile_of(ht, rows[i], tl);
        const uint4* hp = reinterpret_cast<const uint4*>(tile + tl * 32);
        const uint4 h0 = __ldcg(hp), h1 = __ldcg(hp + 1);          // every thread, one address: a broadcast
        if ((h0.x | h0.y) == 0u) continue;                       // invalid sample: it has no lists (:485)
        const uint32_t* __restrict__ d = reinterpret_cast<const uint32_t*>(((unsigned long long)h0.y << 32) | h0.x);
        const uint32_t e0 = h0.z, e1 = h0.w, e2 = h1.x, e3 = h1.y, e4 = h1.z;
        for (uint32_t w = tid; w < e3; w += FN3_VOTE_THREADS) {
            const uint32_t b = (w >= e0) + (w >= e1) + (w >= e2);
            atomicAdd(&cnt[(long long)b * Lpad + d[w]], 1u);
        }
        if (want_n)
            for (uint32_t w = e3 + tid; w < e4; w += FN3_VOTE_THREADS) {
                const uint32_t x = d[w], s = x & pmask, len = (pos_bits >= 32) ? 1u : ((x >> pos_bits) + 1u);
                for (uint32_t j = 0; j < len; ++j) atomicAdd(&cnt[4ll * Lpad + s + j], 1u);
            }
    }
}

// samples that are not stored: plain lists as in fn3_append_bulk (pos concatenated, off[6n+1]); one thread per position
__global__ void __launch_bounds__(256) k_votes_lists(const int32_t* __restrict__ pos, const long long* __restrict__ off,
                                                     long long n_lists, long long total, uint32_t* __restrict__ cnt,
                                                     long long Lpad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = n_lists;                          // last list whose start is <= i
        while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
        const int b = (int)(lo % 6);
        if (b < 5) atomicAdd(&cnt[(long long)b * Lpad + pos[i]], 1u);
    }
}

// ---- ordered selection of positions from the vote planes ------------------------------------------------
// MODE 0 (_msa steps 1-3): ONE list — p is a variant site iff the samples show more than one base there, the reference
//   base (refcode[p]: 0..3 = A,C,G,T) included whenever some of the `thr` samples has no A/C/G/T call at p (:916-931).
// MODE 1 (consensus, :492-500): FIVE lists — p enters list b iff cnt[b][p] >= thr (thr >= 1).
// MODE 2 (seqComparer_mt._msa step 2, src/seqComparer_mt.py:1127-1137): MODE 0, except that a sample's N at p also
//   "accounts for" p — the reference base is added only when some sample has neither a base call nor an N there
//   (needs the N plane: k_votes_rows with want_n).
template <int MODE> struct SelLists { static const int N = MODE == 1 ? 5 : 1; };

template <int MODE>
__device__ __forceinline__ void sel_flags(const uint32_t* __restrict__ cnt, const uint8_t* __restrict__ refcode, long long Lpad,
                                          long long p0, uint32_t thr, uint32_t flags[SelLists<MODE>::N]) {
    if (MODE != 1) {
        uint32_t seen[FN3_SEL_PER_THREAD], tot[FN3_SEL_PER_THREAD];
#pragma unroll
        for (int i = 0; i < FN3_SEL_PER_THREAD; ++i) { seen[i] = 0u; tot[i] = 0u; }
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int k = 0; k < FN3_SEL_PER_THREAD / 4; ++k) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(cnt + (long long)b * Lpad + p0) + k);
                const uint32_t c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { seen[4 * k + j] |= (uint32_t)(c[j] != 0u) << b; tot[4 * k + j] += c[j]; }
            }
        if (MODE == 2)
#pragma unroll
            for (int k = 0; k < FN3_SEL_PER_THREAD / 4; ++k) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(cnt + 4ll * Lpad + p0) + k);
                tot[4 * k] += v.x; tot[4 * k + 1] += v.y; tot[4 * k + 2] += v.z; tot[4 * k + 3] += v.w;
            }
        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(refcode + p0));
        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t f = 0u;
#pragma unroll
        for (int i = 0; i < FN3_SEL_PER_THREAD; ++i) {
            uint32_t m = seen[i];
            if (tot[i] < thr) m |= 1u << ((rw[i >> 2] >> (8 * (i & 3))) & 3u);     // somebody shows the reference base
            if (__popc(m) > 1) f |= 1u << i;
        }
        flags[0] = f;
    } else {
#pragma unroll
        for (int b = 0; b < SelLists<MODE>::N; ++b) {
            uint32_t f = 0u;
#pragma unroll
            for (int k = 0; k < FN3_SEL_PER_THREAD / 4; ++k) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(cnt + (long long)b * Lpad + p0) + k);
                const uint32_t c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c[j] >= thr) f |= 1u << (4 * k + j);
            }
            flags[b] = f;
        }
    }
}

// counts[cta * NL + l] = number of selected positions of list l in the CTA's 4096 positions
template <int MODE>
__global__ void __launch_bounds__(FN3_SEL_THREADS) k_sel_count(const uint32_t* __restrict__ cnt, const uint8_t* __restrict__ refcode,
                                                                long long Lpad, uint32_t thr, uint32_t* __restrict__ counts) {
    const int NL = SelLists<MODE>::N;
    const long long p0 = ((long long)blockIdx.x * FN3_SEL_THREADS + threadIdx.x) * FN3_SEL_PER_THREAD;
    uint32_t flags[NL];
    sel_flags<MODE>(cnt, refcode, Lpad, p0, thr, flags);
    __shared__ uint32_t sm[NL][FN3_SEL_THREADS / 32];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const uint32_t s = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(flags[l]));
        if ((threadIdx.x & 31) == 0) sm[l][threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x < NL) {
        uint32_t t = 0;
        for (int w = 0; w < FN3_SEL_THREADS / 32; ++w) t += sm[threadIdx.x][w];
        counts[(long long)blockIdx.x * NL + threadIdx.x] = t;
    }
}

// one CTA: per list, exclusive scan of the CTA counts (in place); totals[l] = list length
__global__ void __launch_bounds__(1024) k_sel_scan(uint32_t* __restrict__ counts, uint32_t* __restrict__ totals, int ncta, int NL) {
    __shared__ uint32_t ws[33];
    const int chunk = (ncta + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = min(ncta, (int)threadIdx.x * chunk), hi = min(ncta, lo + chunk);
    for (int l = 0; l < NL; ++l) {
        uint32_t local = 0;
        for (int i = lo; i < hi; ++i) local += counts[(long long)i * NL + l];
        uint32_t total;
        uint32_t run = block_exscan(local, ws, &total);
        for (int i = lo; i < hi; ++i) { const uint32_t v = counts[(long long)i * NL + l]; counts[(long long)i * NL + l] = run; run += v; }
        if (threadIdx.x == 0) totals[l] = total;
    }
}

// out: the NL lists one behind the other (list l starts at totals[0] + .. + totals[l-1]), each in position order
template <int MODE>
__global__ void __launch_bounds__(FN3_SEL_THREADS) k_sel_write(const uint32_t* __restrict__ cnt, const uint8_t* __restrict__ refcode,
                                                                long long Lpad, uint32_t thr, const uint32_t* __restrict__ counts,
                                                                const uint32_t* __restrict__ totals, int32_t* __restrict__ out) {
    const int NL = SelLists<MODE>::N;
    __shared__ uint32_t ws[33];
    const long long p0 = ((long long)blockIdx.x * FN3_SEL_THREADS + threadIdx.x) * FN3_SEL_PER_THREAD;
    uint32_t flags[NL];
    sel_flags<MODE>(cnt, refcode, Lpad, p0, thr, flags);
    uint32_t start = 0;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        uint32_t total;
        const uint32_t within = block_exscan((uint32_t)__popc(flags[l]), ws, &total);
        uint32_t at = start + counts[(long long)blockIdx.x * NL + l] + within;
        uint32_t f = flags[l];
        while (f) {
            const int j = __ffs(f) - 1;
            f &= f - 1u;
            out[at++] = (int32_t)(p0 + j);
        }
        start += totals[l];
    }
}

// ---- alignment at the variant sites ---------------------------------------------------------------------
// first index k with sites[k] >= p
__device__ __forceinline__ long long site_lower_bound(const int32_t* __restrict__ sites, long long ns, uint32_t p) {
    long long lo = 0, hi = ns;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if ((uint32_t)__ldg(sites + mid) < p) lo = mid + 1; else hi = mid; }
    return lo;
}

// One warp per row.  codes (optional): codes[i * ns + k] = what row i shows at sites[k]: 0 reference base, 1..4 A,C,G,T,
// 5 N, 6 M — written in the reference's scan order A,C,T,G,N,M so that a later list wins (:941-948).
// counts[i*4..] = { |N|, |M|, N at the sites, M at the sites }.  With codes the last two are counted on the final
// characters (aligned.count('N'), :975-976); without, they are the plain intersections (:678-681).
__global__ void __launch_bounds__(256) k_align(const long long* __restrict__ rows, long long n, const HdrTable ht,
                                               const int32_t* __restrict__ sites, long long ns, uint8_t* __restrict__ codes,
                                               int32_t* __restrict__ counts, int pos_bits) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
        int tl;
        const unsigned char* tile = tile_of(ht, rows[i], tl);
        const uint4* hp = reinterpret_cast<const uint4*>(tile + tl * 32);
        const uint4 h0 = __ldcg(hp), h1 = __ldcg(hp + 1);
        int allN = 0, allM = 0, atN = 0, atM = 0;
        if ((h0.x | h0.y) != 0u) {
            const uint32_t* __restrict__ d = reinterpret_cast<const uint32_t*>(((unsigned long long)h0.y << 32) | h0.x);
            const uint32_t end[6] = {h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            uint8_t* crow = codes ? codes + i * ns : nullptr;
            if (crow) {
                const int order[4] = {0, 1, 3, 2};                          // A, C, T, G
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int li = order[o];
                    const uint32_t lo = li ? end[li - 1] : 0u, hi = end[li];
                    for (uint32_t w = lo + lane; w < hi; w += 32) {
                        const uint32_t p = d[w];
                        const long long k = site_lower_bound(sites, ns, p);
                        if (k < ns && (uint32_t)__ldg(sites + k) == p) crow[k] = (uint8_t)(li + 1);
                    }
                    __syncwarp();
                }
            }
#pragma unroll
            for (int li = 4; li < 6; ++li) {                                // N runs, then M runs
                int all = 0, at = 0;
                for (uint32_t w = end[li - 1] + lane; w < end[li]; w += 32) {
                    const uint32_t x = d[w], s = x & pmask, len = (pos_bits >= 32) ? 1u : ((x >> pos_bits) + 1u);
                    all += (int)len;
                    const long long k0 = site_lower_bound(sites, ns, s), k1 = site_lower_bound(sites, ns, s + len);
                    at += (int)(k1 - k0);
                    if (crow) for (long long k = k0; k < k1; ++k) crow[k] = (uint8_t)(li + 1);
                }
                __syncwarp();
                if (li == 4) { allN = all; atN = at; } else { allM = all; atM = at; }
            }
            if (crow) {                                                     // count what the aligned string shows
                __threadfence_block();
                atN = 0; atM = 0;
                for (long long k = lane; k < ns; k += 32) {
                    const uint32_t c = __ldcg(crow + k);
                    atN += c == 5u; atM += c == 6u;
                }
            }
        }
        allN = __reduce_add_sync(0xffffffffu, allN); allM = __reduce_add_sync(0xffffffffu, allM);
        atN = __reduce_add_sync(0xffffffffu, atN); atM = __reduce_add_sync(0xffffffffu, atM);
        if (lane == 0) { counts[i * 4 + 0] = allN; counts[i * 4 + 1] = allM; counts[i * 4 + 2] = atN; counts[i * 4 + 3] = atM; }
    }
}
