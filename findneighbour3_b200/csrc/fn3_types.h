// fn3_types.h — data layout shared by the device kernels and the host engine.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FN3_HD __host__ __device__ __forceinline__
#else
#define FN3_HD inline
#endif

// ---- row data ---------------------------------------------------------------------------------
// One row per sample: 32-bit words, 16-byte aligned:  [A | C | G | T | N run words | M run words]
//   A..T      sorted int32 positions (the reference's per-base sets, seqComparer.py:290)
//   run word  (len-1) << pos_bits | start      (pos_bits = ceil(log2 genome_length))

struct __align__(32) RowHdr {      // 32 B = one sector
    unsigned long long data;       // device address of the row's first word; 0 = invalid sample or removed row
    uint32_t end[6];               // cumulative ends (in words) of A,C,G,T,Nruns,Mruns
};

// ---- row heads: what the early-exit filter reads -----------------------------------------------
// Per row: the header and a 16-bit FINGERPRINT of each of its first 32 V positions, fp = position >> fshift
// (fshift chosen so that the genome has at most ~16 K coarse blocks: 512-position blocks for a 4.4 Mb genome).
// Heads live in TILES of 32 rows laid out for coalesced warp access (lane l reads hdr[l] and fp[j][l]):
//     tile = [ RowHdr hdr[32] | uint4 fp[4][32] ]        (1024 + 2048 = 3072 bytes)
// and consecutive rows are spread over different tiles: inside a superblock of 32*G0 rows, row r sits in
// tile r % G0, lane r / G0 — so a family of consecutively inserted relatives lands in different warps.
#define FN3_HEAD_FP 32
#define FN3_TILE_BYTES 3072
#define FN3_TILE_FP_OFF 1024

struct __align__(32) RowHead {     // row-order staging format (96 B); k_scatter_heads / k_query place it in its tile
    RowHdr hdr;
    uint16_t fp[FN3_HEAD_FP];      // beyond min(end[3], 32): the engine's pad id (a coarse id no position maps to)
};

#define FN3_MAX_HDR_CHUNKS 32    /* x 2^20 rows each by default: 33 M rows per GPU; keeps the kernel parameters small */
struct HdrTable {
    const unsigned char* chunk[FN3_MAX_HDR_CHUNKS];   // chunk c holds the tiles of rows [c << shift, (c+1) << shift)
    int shift;                     // rows per chunk = 1 << shift  (>= sb_shift)
    int sb_shift;                  // rows per superblock = 1 << sb_shift (>= 5); G0 = 1 << (sb_shift - 5)
};

// (tile, lane) of a row inside its chunk, and back
FN3_HD void fn3_slot_of(long long local_row, int sb_shift, long long& tile, int& lane) {
    const int gs = sb_shift - 5;
    const long long sb = local_row >> sb_shift, within = local_row & ((1ll << sb_shift) - 1);
    lane = (int)(within >> gs);
    tile = (sb << gs) + (within & ((1ll << gs) - 1));
}
FN3_HD long long fn3_row_of(long long tile, int lane, int sb_shift) {
    const int gs = sb_shift - 5;
    return ((tile >> gs) << sb_shift) + ((long long)lane << gs) + (tile & ((1ll << gs) - 1));
}

// ---- the query, compact --------------------------------------------------------
// bm[w]     : bit per 32-position block 32w..32w+31 "the query is non-reference/N/M somewhere in it" (touched)
// pref[w]   : number of touched blocks below block 32w
// coarse    : one BYTE per coarse block (position >> fshift): 1 = touched — what the fingerprints are tested against
//             (one LDS.U8 per fingerprint; the last byte is always 1 and is the id padded fingerprints carry)
// entry[i]  : the i-th touched block; entry[nE] is a sentinel whose counts are the totals |V_q|,|U_q|,|M_q|
struct __align__(32) QEntry {
    uint32_t mDef, mU, mB0, mB1;   // per position: ACGT-nonref / N-or-M / base code bit 0 / bit 1 (A0 C1 G2 T3)
    uint32_t mM, cDef, cU, cM;     // per position: M ; number of Def / U / M positions strictly below this block
};

struct QueryState {                // global-memory copies (only queries too large for shared memory use them)
    uint32_t* bm;                  // slot s at bm + s*2*nw : [touched-block bitmap nw | prefix counts nw]
    uint32_t* coarse;              // slot s at coarse + s*ncw (a byte map of 4*ncw entries)
    QEntry* ent;                   // slot s at ent + s*ent_stride
    uint4* info;                   // slot s: {nE, |V_q|, |U_q|, |M_q|}
    long long ent_stride;
};

struct QuerySlot {                 // one query of a tile of queries
    RowHdr hdr;                    // the query's own row (stored row or scratch)
    long long row;                 // store row of the query, -1 if not stored
    long long skip_le;             // candidates with row <= skip_le are skipped (upper triangle), -1 = none
    long long exclude;             // candidate row to skip (self), -1 = none
    long long pad;
};

struct __align__(32) EdgeRec {     // device-side survivor record (== fn3_edge)
    long long row1, row2;
    int dist, n1, n2, nboth;
};

#define FN3_MAX_SLOTS 256

struct CompareParams {
    HdrTable hdr;
    QueryState qs;
    const long long* rows;         // optional explicit candidate list (n_list entries); else all rows < n_rows
    long long n_list;
    long long n_rows;
    const QuerySlot* slots;        // device array [gridDim.y] unless slot_inline
    QuerySlot slot0;               // the single query, by value (one-vs-all): no copy, no dependent load
    RowHead commit_head;           // fn3_insert_query: head of the row being inserted, placed by the kernel
    long long commit_row;          // -1 = nothing to commit
    EdgeRec* out;
    unsigned long long* out_count;
    unsigned long long* touched;   // instrumented variant only
    long long out_cap;
    // zero-copy completion (one-vs-all): survivors are ALSO written to mapped pinned host memory; the last CTA to
    // finish publishes the count and the sequence number there and re-arms the device counters for the next query
    EdgeRec* host_out;             // mapped host buffer (nullptr = off)
    volatile unsigned long long* host_flag;   // mapped: [0] sequence number, [1] survivor count
    unsigned int* done_ctas;       // device counter of finished CTAs
    long long host_cap;
    unsigned long long seq;
    int slot_inline;
    int ceiling;
    int pos_bits;
    int use_filter;
    int fshift;
    int nw;                        // bmpref words (multiple of 4)
    int ncw;                       // coarse byte map size in 32-bit words (multiple of 8)
    int debug_stop;                // profiling aid: 1 stop after build, 2 after phase A, 3 at entry, 4 after slot load
    unsigned int ring_off;         // k_scan: byte offset of the row rings in dynamic shared memory (multiple of 128)
    int split_shift;               // k_scan: a 32-row work unit is drawn as 1 << split_shift sub-units
};
