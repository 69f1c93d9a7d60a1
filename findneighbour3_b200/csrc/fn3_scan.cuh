// fn3_scan.cuh — k_scan: the STREAM variant of the compare kernel (ceilings above the head filter's reach: the
// recompression scan at snpCompressionCeiling = 250, distmat at that ceiling, "no ceiling"), written for sm_100a.
//
// Reference being replaced (davidhwyllie/findNeighbour3, src/seqComparer.py): the N x countDifferences_byKey loop of
// compress_relative_to_consensus :584-592 and of mcompare :165-169 when the cut-off is large.
//
// Data movement: whole candidate rows travel HBM -> shared memory as 1-D BULK ASYNC COPIES
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, SASS UBLKCP) into a per-warp ring of
// FN3_SCAN_D slots of FN3_SCAN_CH words; every slot has its own mbarrier (SASS SYNCS).  Lane 0 of a warp is the
// producer of its own ring, the whole warp its consumer, so no CTA-wide barrier sits in the loop and no thread spends
// issue slots on global loads, address arithmetic or bounds tests: row i+1 and i+2 land while row i is scanned.
// Rows longer than a slot are streamed as a sequence of chunks through the same ring.
//
// Work distribution: a work unit is a head tile of 32 rows (or 32 entries of an explicit candidate list), split into
// 1 << split_shift sub-units; the warps of a CTA draw sub-units from a shared-memory counter, so that the rows of a
// small store are spread over all warps of the GPU (one 32-row tile per warp leaves a quarter of the warps idle at
// 50 000 rows) and a warp that meets long rows does not hold the others up.
//
// Per candidate word the scan costs one LDS into the query's touched-block bitmap (fn3_kernels.cuh: QView); the rare
// words that fall in a block the query touches are remembered in a per-lane bit mask and resolved afterwards, lanes
// working in parallel, by look_touched().  Run words (N, M) take two rank lookups each, as before.
#pragma once
#include "fn3_kernels.cuh"

#ifndef FN3_SCAN_THREADS
#define FN3_SCAN_THREADS 768     // 24 warps, one CTA per SM (the query structure + the rings fill its shared memory)
#endif
#ifndef FN3_SCAN_CH
#define FN3_SCAN_CH 640          // words per ring slot (2.5 KB: a typical TB row is ~500 words); multiple of 4, <= 1024
#endif
#ifndef FN3_SCAN_D
#define FN3_SCAN_D 2             // ring slots per warp: one being scanned, one in flight (3 and 4 measured: no gain)
#endif
#ifndef FN3_SCAN_R
#define FN3_SCAN_R 16            // neighbour records a warp collects before it appends them with ONE atomic
#endif
#define FN3_SCAN_WARPS (FN3_SCAN_THREADS / 32)
#define FN3_SCAN_META_WORDS 12   // per slot: header (8 words), row, first word, words, last-chunk flag
#define FN3_SCAN_RING_BYTES ((size_t)FN3_SCAN_WARPS * (FN3_SCAN_D * (FN3_SCAN_CH * 4 + FN3_SCAN_META_WORDS * 4 + 8) + FN3_SCAN_R * 32))

static_assert(FN3_SCAN_CH % 4 == 0 && FN3_SCAN_CH <= 1024, "ring slot: 16-byte multiples, at most 32 mask bits per lane");
static_assert(FN3_SCAN_R <= 32, "one lane per staged record");

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FN3_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FN3_DONE_%=;\n"
        "bra FN3_WAIT_%=;\n"
        "FN3_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared; completion is signalled on the mbarrier as `bytes` of transaction count
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- one chunk of a row, already in shared memory ----------------------------------------------------

// sm[0..n) = row words [off, off+n).  Adds to the row's tally; sets t.dead once S1 exceeds the ceiling.
__device__ __forceinline__ void scan_chunk(const QView& qv, const uint32_t* __restrict__ sm, uint32_t off, uint32_t n,
                                           uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t e4, int k,
                                           int pos_bits, uint32_t pmask, int lane, Tally& t, int& lane_S1) {
    const bool early = k < 0x3fffffff;                           // warp-uniform: is there a ceiling a row could exceed?
    const uint32_t v_n = e3 > off ? min(e3 - off, n) : 0u;       // V words (A,C,G,T positions) in this chunk
    uint32_t tm = 0u;                                            // this lane's words whose block the query touches
    // 256 words per warp step: lane l takes words 4l..4l+3 of each half (two conflict-free 128-bit LDS)
    for (uint32_t sb = 0, sh = 0; sb < v_n; sb += 256, sh += 8) {
        const uint32_t li = sb + lane * 4;
        uint4 w0 = *reinterpret_cast<const uint4*>(sm + li);
        uint4 w1 = make_uint4(0u, 0u, 0u, 0u);
        uint32_t vm = 0xffu;
        if (sb + 256u <= v_n) {                                  // warp-uniform: a full step
            w1 = *reinterpret_cast<const uint4*>(sm + li + 128);
        } else {                                                 // the last, partial step
            const uint32_t a = li >= v_n ? 0u : min(v_n - li, 4u);
            const uint32_t b = li + 128u >= v_n ? 0u : min(v_n - li - 128u, 4u);
            vm = lowmask(a) | (lowmask(b) << 4);
            if (b) w1 = *reinterpret_cast<const uint4*>(sm + li + 128);
            // what lies behind the V words is not a position
            if (a < 4u) { w0.w = 0u; if (a < 3u) { w0.z = 0u; if (a < 2u) { w0.y = 0u; if (a < 1u) w0.x = 0u; } } }
            if (b < 4u) { w1.w = 0u; if (b < 3u) { w1.z = 0u; if (b < 2u) { w1.y = 0u; if (b < 1u) w1.x = 0u; } } }
        }
        // eight independent bitmap lookups: all loads first, then the bit tests; the common answer (no word of this lane
        // in a block the query touches) costs one OR per word, the exact bits are formed only when some test fired
        const uint32_t pw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        uint32_t bw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bw[j] = qv.bm[pw[j] >> 10];
        uint32_t any = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) { bw[j] >>= ((pw[j] >> 5) & 31u); any |= bw[j]; }
        uint32_t t8 = 0u;
        if (any & 1u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) t8 |= (bw[j] & 1u) << j;
            t8 &= vm;
        }
        tm |= t8 << sh;
        // untouched block: the query is reference there, the candidate is not — a certain mismatch
        if (early) {
            t.S1 += __reduce_add_sync(0xffffffffu, __popc(vm) - __popc(t8));
            if (t.S1 > k) { t.dead = true; return; }
        } else {
            lane_S1 += __popc(vm) - __popc(t8);           // no ceiling to exceed: one reduction per row, at the end
        }
    }
    if (__any_sync(0xffffffffu, tm != 0u)) {
        int m2 = 0;
        do {
            if (tm) {
                const uint32_t b = (uint32_t)__ffs(tm) - 1u;
                tm &= tm - 1u;
                const uint32_t li = (b >> 3) * 256u + ((b >> 2) & 1u) * 128u + lane * 4 + (b & 3u);
                const uint32_t r = look_touched(qv, sm[li], off + li, e0, e1, e2);
                m2 += r & 1u;
                t.qdefV += r >> 1;
            }
        } while (__any_sync(0xffffffffu, tm != 0u));
        if (early) {
            t.S1 += __reduce_add_sync(0xffffffffu, m2);
            if (t.S1 > k) { t.dead = true; return; }
        } else {
            lane_S1 += m2;
        }
    }
    // run words: [start, start+len) of N (row index below e4) or M
    for (uint32_t i = (e3 > off ? e3 - off : 0u) + lane; i < n; i += 32) {
        const uint32_t x = sm[i];
        const uint32_t s = x & pmask;
        const uint32_t len = (pos_bits >= 32) ? 1u : ((x >> pos_bits) + 1u);
        const bool isN = off + i < e4;
        uint32_t aD, aU, aM, bD, bU, bM;
        rank3(qv, s, aD, aU, aM, isN);
        rank3(qv, s + len, bD, bU, bM, isN);
        t.qdefU += (int)(bD - aD);
        if (isN) { t.nNc += (int)len; t.nNU += (int)(bU - aU); t.nNM += (int)(bM - aM); }
    }
}

// ---- the kernel ------------------------------------------------------------------------------------

template <bool COUNT, int QMODE>
__global__ void __launch_bounds__(FN3_SCAN_THREADS, 1) k_scan(const CompareParams p) {
    extern __shared__ __align__(128) uint4 qsm[];
    __shared__ uint32_t ws[FN3_WS_WORDS];
    __shared__ unsigned int next_unit;
    if (p.debug_stop == 3) return;
    FN3_TR(13); FN3_TRG(22);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.y;
    const QuerySlot qs = p.slot_inline ? p.slot0 : p.slots[slot];
    if (p.commit_row >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
        put_head(p.hdr, p.commit_row, p.commit_head);
    if (qs.hdr.data == 0ull) {                             // invalid query: no distances at all
        if (p.host_flag) signal_done(p);
        return;
    }
#if FN3_PF
    if (QMODE == 1 && threadIdx.x * 32u < qs.hdr.end[5])   // the query's own row, wanted by the build below: on its way into L2
        prefetch_l2(reinterpret_cast<const uint32_t*>(qs.hdr.data) + threadIdx.x * 32u);   // while the first rows are requested
#endif
    // ring carve-out: [slots: WARPS x D x CH words | meta: WARPS x D x 12 words | staged records: WARPS x R x 32 bytes |
    //                  mbarriers: WARPS x D x 8 bytes]
    unsigned char* ring_base = reinterpret_cast<unsigned char*>(qsm) + p.ring_off;
    uint32_t* my_ring = reinterpret_cast<uint32_t*>(ring_base) + (size_t)warp * FN3_SCAN_D * FN3_SCAN_CH;
    unsigned char* meta_base = ring_base + (size_t)FN3_SCAN_WARPS * FN3_SCAN_D * FN3_SCAN_CH * 4;
    uint32_t* my_meta = reinterpret_cast<uint32_t*>(meta_base) + warp * FN3_SCAN_D * FN3_SCAN_META_WORDS;
    unsigned char* rec_base = meta_base + (size_t)FN3_SCAN_WARPS * FN3_SCAN_D * FN3_SCAN_META_WORDS * 4;
    EdgeRec* my_recs = reinterpret_cast<EdgeRec*>(rec_base) + warp * FN3_SCAN_R;
    const uint32_t my_bar = smem_addr(rec_base + (size_t)FN3_SCAN_WARPS * FN3_SCAN_R * 32) + warp * FN3_SCAN_D * 8;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < FN3_SCAN_D; ++s) mbar_init(my_bar + s * 8, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == 0) next_unit = 0u;
    __syncthreads();                                       // the unit counter is initialised (each warp's mbarriers are its own)
    FN3_TR(14);
    unsigned long long touched = 0;

    // ---- work units ---------------------------------------------------------------------------------
    const int exclude = (int)qs.exclude, skip_le = (int)qs.skip_le;
    const int tshift = p.hdr.shift - 5, sb_shift = p.hdr.sb_shift, gs = sb_shift - 5;
    const uint32_t gmask = (1u << gs) - 1u, tmask = (1u << tshift) - 1u, n_rows32 = (uint32_t)p.n_rows;
    const uint32_t units = p.rows ? (uint32_t)((p.n_list + 31) >> 5)
                                  : (uint32_t)(((p.n_rows + (1ll << sb_shift) - 1) >> sb_shift) << gs);
    // upper triangle: superblocks that hold only rows <= skip_le cannot contain a candidate
    const uint32_t ufirst = (p.rows || skip_le < 0) ? 0u : (((uint32_t)skip_le + 1u) >> sb_shift) << gs;
    const int ss = p.split_shift;                          // a unit is drawn as 1 << ss sub-units of 32 >> ss rows
    const uint32_t sub_total = (units - ufirst) << ss;
    uint4 cur0 = make_uint4(0, 0, 0, 0), cur1 = cur0, nxt0 = cur0, nxt1 = cur0;     // headers of the current / next sub-unit
    int cur_row = -1, nxt_row = -1;
    bool nxt_ok = false;

    // draw the next sub-unit of this CTA and request its headers (one per lane; not waited for here)
    auto fetch_next = [&]() {
        unsigned int i = 0;
        if (lane == 0) i = atomicAdd(&next_unit, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        const unsigned long long su = (unsigned long long)blockIdx.x + (unsigned long long)i * gridDim.x;
        nxt0 = make_uint4(0, 0, 0, 0); nxt1 = nxt0; nxt_row = -1;
        nxt_ok = su < (unsigned long long)sub_total;
        if (!nxt_ok) return;
        // sub-unit -> (tile, part of the tile) with the TILE in the low digits: a CTA's sub-units b, b + grid, b + 2 grid, ...
        // then cover all parts of the tiles evenly (with the part in the low bits and a grid that is a multiple of four,
        // CTA b would always draw part b & 3 — and the rows of a half-filled last superblock all sit in part 0)
        const uint32_t n_units = units - ufirst;
        const uint32_t uu = ufirst + (uint32_t)(su % n_units), part = (uint32_t)(su / n_units);
        if (((uint32_t)lane >> (5 - ss)) != part) return;  // this lane's row belongs to another sub-unit of the tile
        const unsigned char* tile;
        int tl = lane, row = -1;
        if (p.rows) {
            const long long ci = (long long)uu * 32 + lane;
            if (ci >= p.n_list) return;
            row = (int)p.rows[ci];
            tile = tile_of(p.hdr, row, tl);
        } else {
            const uint32_t r = ((uu >> gs) << sb_shift) + ((uint32_t)lane << gs) + (uu & gmask);
            if (r >= n_rows32) return;
            row = (int)r;
            tile = p.hdr.chunk[uu >> tshift] + (size_t)(uu & tmask) * FN3_TILE_BYTES;
        }
        if (row == exclude || row <= skip_le) return;
        const uint4* hp = reinterpret_cast<const uint4*>(tile + tl * 32);
        nxt0 = __ldg(hp); nxt1 = __ldg(hp + 1);
        nxt_row = row;
        if (COUNT) touched += 32;
    };

    // ---- producer: the lane that holds a row's header issues the bulk copies of that row ------------------
    uint32_t p_mask = 0u;                                  // rows of the current sub-unit not yet issued
    int p_lane = -1;                                       // lane holding the row being issued, -1 = none
    uint32_t p_off = 0u, p_e5 = 0u, p_slot = 0u, outstanding = 0u;
    int p_row = -1;

    auto issue_one = [&]() -> bool {
        if (p_lane < 0) {
            while (p_mask == 0u) {
                if (!nxt_ok) return false;
                cur0 = nxt0; cur1 = nxt1; cur_row = nxt_row;
                fetch_next();
                p_mask = __ballot_sync(0xffffffffu, cur_row >= 0 && (cur0.x | cur0.y) != 0u);
            }
            p_lane = __ffs(p_mask) - 1;
            p_mask &= p_mask - 1u;
            p_off = 0u;
            p_e5 = __shfl_sync(0xffffffffu, cur1.w, p_lane);
            p_row = __shfl_sync(0xffffffffu, cur_row, p_lane);
        }
        const uint32_t n = min((uint32_t)FN3_SCAN_CH, p_e5 - p_off);
        const uint32_t last = p_off + n >= p_e5 ? 1u : 0u;
        if (lane == p_lane) {
            uint32_t* m = my_meta + p_slot * FN3_SCAN_META_WORDS;
            reinterpret_cast<uint4*>(m)[0] = cur0;
            reinterpret_cast<uint4*>(m)[1] = cur1;
            reinterpret_cast<uint4*>(m)[2] = make_uint4((uint32_t)cur_row, p_off, n, last);
            const uint32_t bytes = ((n + 3u) & ~3u) * 4u;   // rows are 16-byte aligned and padded to 4 words
            const uint32_t bar = my_bar + p_slot * 8;
            mbar_arrive_expect_tx(bar, bytes);
            if (bytes) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(((unsigned long long)cur0.y << 32) | cur0.x) + p_off;
                bulk_g2s(smem_addr(my_ring + p_slot * FN3_SCAN_CH), src, bytes, bar);
            }
            if (COUNT) touched += bytes;
        }
        p_off += n;
        if (last) p_lane = -1;
        p_slot = p_slot + 1u == FN3_SCAN_D ? 0u : p_slot + 1u;
        ++outstanding;
        return true;
    };

    // the first rows are requested BEFORE the query structure is built: their flight time hides behind the build
    fetch_next();                                          // headers of the first sub-unit ...
    if (nxt_ok && p.debug_stop != 1) {
        cur0 = nxt0; cur1 = nxt1; cur_row = nxt_row;
        fetch_next();                                      // ... and of the second
        p_mask = __ballot_sync(0xffffffffu, cur_row >= 0 && (cur0.x | cur0.y) != 0u);
        FN3_TR(15);
#pragma unroll 1
        for (int i = 0; i < FN3_SCAN_D; ++i)
            if (!issue_one()) break;
    }
    FN3_TR(16);

    const int nw = p.nw, ncw = p.ncw;
    uint32_t nVq, nUq, nMq;
    QView qv;
    if (QMODE == 1) {
        uint32_t* sm_bm = reinterpret_cast<uint32_t*>(qsm);
        uint32_t* sm_co = reinterpret_cast<uint32_t*>(qsm + nw / 2);
        QEntry* sm_en = reinterpret_cast<QEntry*>(qsm + nw / 2 + ncw / 4);
        uint32_t tot[3];
        build_query<false, FN3_SCAN_THREADS>(qs, nw, ncw, p.fshift, p.pos_bits, sm_bm, sm_co, sm_en, ws, tot);
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
    const int k = p.ceiling;
    const int pos_bits = p.pos_bits;
    const uint32_t pmask = (pos_bits >= 32) ? 0xffffffffu : ((1u << pos_bits) - 1u);
    FN3_TR(17);

    // ---- neighbour records: collected per warp, appended FN3_SCAN_R at a time with one atomic ---------------
    // (one atomic per record on the single counter bounds a scan in which every row is a neighbour — an outbreak, or
    //  no ceiling at all — at the L2's per-address atomic rate: 50 000 records took ~30 us)
    uint32_t n_staged = 0u;
    auto flush = [&]() {
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.out_count, (unsigned long long)n_staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((uint32_t)lane < n_staged && (long long)(base + lane) < p.out_cap) {
            const EdgeRec r = my_recs[lane];
            p.out[base + lane] = r;
            if (p.host_out && (long long)(base + lane) < p.host_cap) p.host_out[base + lane] = r;   // posted write over PCIe
        }
        n_staged = 0u;
        __syncwarp();
    };
    auto emit = [&](const Tally& t0, long long crow) {
        const int qdefV = __reduce_add_sync(0xffffffffu, t0.qdefV);
        const int qdefU = __reduce_add_sync(0xffffffffu, t0.qdefU);
        const int dist = t0.S1 + (int)nVq - qdefV - qdefU;
        if (dist > k) return;
        const int nNc = __reduce_add_sync(0xffffffffu, t0.nNc);
        const int nNU = __reduce_add_sync(0xffffffffu, t0.nNU);
        const int nNM = __reduce_add_sync(0xffffffffu, t0.nNM);
        if (lane == 0) {
            EdgeRec r;
            r.row1 = qs.row; r.row2 = crow; r.dist = dist;
            r.n1 = (int)nUq; r.n2 = nNc + (int)nMq - nNM; r.nboth = (int)nUq + nNc - nNU;
            my_recs[n_staged] = r;
        }
        if (++n_staged == FN3_SCAN_R) flush();
    };

    // ---- consumer ------------------------------------------------------------------------------------------
    uint32_t c_slot = 0u, phases = 0u;
    Tally t; t.S1 = 0; t.qdefV = 0; t.qdefU = 0; t.nNc = 0; t.nNU = 0; t.nNM = 0; t.dead = false;
    int lane_S1 = 0;                                       // per-lane share of S1 when no ceiling asks for it step by step
    while (outstanding) {
        __syncwarp();                                      // the issuing lane's meta stores are visible to the warp
        mbar_wait(my_bar + c_slot * 8, (phases >> c_slot) & 1u);
        phases ^= 1u << c_slot;
        const uint4* m = reinterpret_cast<const uint4*>(my_meta + c_slot * FN3_SCAN_META_WORDS);
        const uint4 h0 = m[0], h1 = m[1], mi = m[2];       // mi = {row, first word, words, last chunk}
        if (mi.y == 0u) { t.S1 = 0; t.qdefV = 0; t.qdefU = 0; t.nNc = 0; t.nNU = 0; t.nNM = 0; t.dead = false; lane_S1 = 0; }
        if (!t.dead) {
            scan_chunk(qv, my_ring + c_slot * FN3_SCAN_CH, mi.y, mi.z, h0.z, h0.w, h1.x, h1.y, h1.z, k, pos_bits, pmask, lane, t, lane_S1);
            if (t.dead) {
                if (p_lane >= 0 && p_row == (int)mi.x) p_lane = -1;      // do not fetch the rest of a row that is out
            } else if (mi.w) {
                if (k >= 0x3fffffff) t.S1 += __reduce_add_sync(0xffffffffu, lane_S1);
                emit(t, (long long)(int)mi.x);
            }
        }
        __syncwarp();                                      // every lane is done with the slot before it is refilled
        c_slot = c_slot + 1u == FN3_SCAN_D ? 0u : c_slot + 1u;
        --outstanding;
        issue_one();
    }
    FN3_TR(19);
    if (n_staged) flush();
    FN3_TR(20); FN3_TRG(23);
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, o);
        if (lane == 0) atomicAdd(p.touched, touched);
    }
    if (p.host_flag) signal_done(p);
}
