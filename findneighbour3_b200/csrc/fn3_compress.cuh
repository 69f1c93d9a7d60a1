// fn3_compress.cuh — device compress(): byte compare against the reference under the exclusion mask and ordered
// compaction into the six sorted lists (replaces the Python loop of seqComparer.compress, seqComparer.py:274-307).
#pragma once
#include <stdint.h>

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

// (the exclusive scan of the per-CTA counts, per class, is k_sel_scan of fn3_cluster.cuh: same counts[cta * NL + l] layout)

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


// ------------------------------------------------------------------------------------------
// composition counts: NucleicAcid.examine (src/NucleicAcid.py:50-58) counts 18 characters with one bytes.count() pass
// each; here one pass over the sequence yields the whole 256-bin byte histogram.
// ------------------------------------------------------------------------------------------
// Sequences are ~25 % each of A,C,G,T plus N and '-': those six are counted in registers (byte-wise SIMD compare +
// popcount; a shared-memory atomic per byte would serialise on four addresses), everything else — rare — goes to a
// shared 256-bin histogram.  hist: 256 x u64, zeroed by the caller.
#define FN3_HIST_THREADS 256

__global__ void __launch_bounds__(FN3_HIST_THREADS) k_byte_hist(const uint8_t* __restrict__ bytes, long long n,
                                                                unsigned long long* __restrict__ hist) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0u;                                   // FN3_HIST_THREADS == 256
    __syncthreads();
    const uint32_t key[6] = {0x41414141u, 0x43434343u, 0x47474747u, 0x54545454u, 0x4e4e4e4eu, 0x2d2d2d2du};   // A C G T N -
    uint32_t cnt[6] = {0u, 0u, 0u, 0u, 0u, 0u};
    const long long n16 = n >> 4;
    const uint4* v = reinterpret_cast<const uint4*>(bytes);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(v + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t known = 0u;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const uint32_t m = __vcmpeq4(w[k], key[c]);  // 0xff in every byte that matches
                cnt[c] += (uint32_t)__popc(m) >> 3;
                known |= m;
            }
            if (known != 0xffffffffu) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (((known >> (8 * j)) & 0xffu) == 0u) atomicAdd(&sh[(w[k] >> (8 * j)) & 0xffu], 1u);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)                // the last n % 16 bytes
        for (long long i = n16 << 4; i < n; ++i) atomicAdd(&sh[bytes[i]], 1u);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const uint32_t t = __reduce_add_sync(0xffffffffu, cnt[c]);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(&sh[key[c] & 0xffu], t);
    }
    __syncthreads();
    const uint32_t mine = sh[threadIdx.x];
    if (mine) atomicAdd(&hist[threadIdx.x], (unsigned long long)mine);
}
