/*
 * fn3.h — C-ABI of the B200-native SNV-distance engine (libfn3.so).
 *
 * This is the drop-in boundary for ONE path of davidhwyllie/findNeighbour3: the masked
 * pairwise SNV distance computation of class `seqComparer`.  The reference has no FFI
 * (it is pure Python); every entry point below names the reference method whose work it
 * takes over (paths relative to the reference checkout, `src/seqComparer.py` unless
 * stated).  The Python class `findneighbour3_b200.seqComparer.seqComparer` binds these
 * symbols with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; caller allocates every output buffer;
 *   - return value: FN3_OK (0) or a negative FN3_E* class, text via fn3_last_error();
 *   - a sample is handed over as SIX sorted, mutually disjoint, 0-based int32 position
 *     lists concatenated in the order A,C,G,T,N,M with 7 CSR offsets `off[0..6]`
 *     (off[0]=0, list b is pos[off[b] .. off[b+1]) ).  That is the reference's
 *     `compressed_sequence` dict (seqComparer.py:290-305) with sets sorted and the M
 *     dict reduced to its keys (the compare path never reads the M characters,
 *     seqComparer.py:449-450, 416-417);
 *   - `invalid != 0` reproduces `{'invalid':1}` (seqComparer.py:301-303): nothing is
 *     stored and every pair involving the row yields "no distance";
 *   - "no distance" (the reference's None) is never emitted: only pairs with
 *     dist <= ceiling come back (seqComparer.py:455, server.py:532-538);
 *   - all calls are thread-safe; appends may run concurrently with queries
 *     (server.py:513-516 runs compress+persist outside the write semaphore).
 *
 * There is no CPU implementation behind this ABI: fn3_create fails (NULL) when no
 * CUDA device is usable.
 */
#ifndef FN3_H
#define FN3_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FN3_ABI_VERSION 4   /* 4: + fn3_group_* (engine groups: one process, several GPUs); 2: + fn3_consensus, fn3_consensus_lists, fn3_msa_sites, fn3_msa_align, fn3_unpickle_samples; 3: + fn3_msa_sites_rule, fn3_ingest_query, fn3_byte_histogram */

/* status codes */
#define FN3_OK          0
#define FN3_EARG       -1   /* bad argument (unsorted / out of range / overlapping lists, bad row) */
#define FN3_ENOTFOUND  -2   /* row does not exist or was removed */
#define FN3_ECAPACITY  -3   /* caller's output buffer too small; *n_out holds the required count */
#define FN3_ECUDA      -4   /* CUDA runtime failure (text in fn3_last_error) */
#define FN3_ENOMEM     -5   /* host or device allocation failed */

typedef struct fn3_engine fn3_engine;

/* One surviving neighbour: what mcompare()/countDifferences_byKey() report for a pair
 * whose distance is <= the ceiling (seqComparer.py:415-419).
 *   dist  = |{p : code_q[p] != code_c[p], p not in N_q u M_q, p not in N_c u M_c}|   (:446-453)
 *   n1    = |N_q u M_q|            (:416)
 *   n2    = |N_c u M_q|            (:417 — sic: the reference uses seq1's M keys)
 *   nboth = |N_q u M_q u N_c|      (:418, :378)                                        */
typedef struct fn3_hit {
    int64_t row;      /* candidate row (seq2) */
    int32_t dist;
    int32_t n1;
    int32_t n2;
    int32_t nboth;
} fn3_hit;

/* One edge of the all-vs-all sparse matrix (distmat(), seqComparer.py:174-197):
 * row1 is seq1 (the query), row2 is seq2. */
typedef struct fn3_edge {
    int64_t row1;
    int64_t row2;
    int32_t dist;
    int32_t n1;
    int32_t n2;
    int32_t nboth;
} fn3_edge;

/* Store statistics (summarise_stored_items(), seqComparer.py:199-219, engine side). */
typedef struct fn3_stats {
    int64_t n_rows;          /* rows ever appended (including removed ones) */
    int64_t n_live;          /* rows not removed */
    int64_t n_invalid;       /* live rows stored as invalid */
    int64_t store_words;     /* 32-bit words of row data resident in HBM (live + dead) */
    int64_t store_bytes;     /* device bytes reserved for the store (chunks + headers) */
    int64_t positions;       /* sum over live rows of |A|+|C|+|G|+|T|+|N|+|M| (canonical B_pair positions) */
    int64_t h2d_bytes;       /* bytes copied host->device by this engine since creation */
    int64_t d2h_bytes;       /* bytes copied device->host by this engine since creation */
    int64_t t_pack_ns;       /* fn3_insert_query host-time split (cumulative): validate + pack the sample */
    int64_t t_enqueue_ns;    /*   ... enqueue the copies and the kernel */
    int64_t t_wait_ns;       /*   ... wait for the result */
    int64_t n_timed;         /*   ... number of fn3_insert_query calls timed */
    int64_t head_refreshes;  /* times the heads were rebuilt from a new background bitmap (fn3_refresh_heads, automatic or not) */
    int64_t scan_mode;       /* 1 while full low-ceiling scans are steered to the streaming kernel (most candidates pass the head filter) */
} fn3_stats;

/* ---- life cycle ------------------------------------------------------------------ */

/* Replaces seqComparer.__init__'s in-RAM store (`self.seqProfile = {}`, :108).
 * genome_length = len(reference) (:93); device = CUDA ordinal.  NULL on failure
 * (no device, bad length); fn3_last_error(NULL) then holds the reason. */
fn3_engine* fn3_create(int64_t genome_length, int32_t device);
int         fn3_destroy(fn3_engine* e);
int         fn3_abi_version(void);

/* Text of the last error on this engine (or of the last failed fn3_create when e==NULL).
 * The pointer stays valid until the next failing call on the same thread. */
const char* fn3_last_error(fn3_engine* e);

/* ---- store: persist / remove (seqComparer.py:119-134) --------------------------- */

/* persist(obj, guid): append one sample; *out_row receives its row index (rows are
 * dense, 0-based, never reused).  The double-delta form (patch + consensus, :503-548)
 * is expanded by the caller before the call — the HBM store always holds expanded rows. */
int fn3_append(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid,
               int64_t* out_row);

/* Bulk persist (server start-up reload, server.py:358-362): n samples, lists
 * concatenated, off has 6*n+1 entries (sample i, list b = pos[off[6i+b] .. off[6i+b+1]) ),
 * invalid[i] as above (may be NULL = none invalid).  Rows first_row .. first_row+n-1. */
int fn3_append_bulk(fn3_engine* e, int64_t n, const int32_t* pos, const int64_t* off,
                    const uint8_t* invalid, int64_t* out_first_row);

/* remove(guid) (:125-134): tombstone the row; it never matches again. */
int fn3_remove(fn3_engine* e, int64_t row);

int fn3_stats_get(fn3_engine* e, fn3_stats* out);

/* Reclaim the tombstones remove() leaves behind (its words stay in HBM, its head slot is still scanned): rebuild the store
 * ON THE DEVICE — live rows copied into fresh chunks in row order, renumbered densely, heads rebuilt from the row data; no
 * row travels through the host.  new_row[r] (r < *n_before, cap >= the current row count) receives the new number of old
 * row r, or -1 if it was removed; the caller re-keys its guid <-> row maps.  Needs room for a second copy of the live rows. */
int fn3_compact(fn3_engine* e, int64_t* new_row, int64_t cap, int64_t* n_before, int64_t* n_after);

/* Rebuild the row heads the low-ceiling filter reads.  A head carries 32 of a row's variant positions; which 32 is free (any
 * variant position of the candidate in a block the query does not touch is a certain mismatch) but decides the filter's
 * power: isolates of one lineage share most of their low-coordinate variants, so the engine keeps a bitmap of BACKGROUND
 * positions — non-reference in at least min_ppm / 1e6 of the live valid rows (and in at least 8) — and a head takes the row's
 * first 32 positions that are not background.  Results never change; only how many candidates are ruled out from their
 * head.  The engine calls this itself whenever the store has doubled (FN3_BG_AUTO=0 turns that off, FN3_BG_PPM sets the
 * fraction: default 10 000 = 1 %); min_ppm <= 0 clears the background (heads = the first 32 positions, as before). */
int fn3_refresh_heads(fn3_engine* e, int64_t min_ppm);

/* Copy a stored row back as the six plain lists (load(), :136-141, used by tests and by
 * the sharded host layer to ship a query to other ranks).  pos_cap in int32 elements;
 * FN3_ECAPACITY with *n_pos = required if too small.  *invalid: 1 invalid, 0 valid. */
int fn3_row_get(fn3_engine* e, int64_t row, int32_t* pos, int64_t pos_cap, int32_t off[7],
                int32_t* invalid, int64_t* n_pos);

/* ---- the hot path --------------------------------------------------------------- */

/* mcompare(guid, guids) (:149-172): stored row `query_row` (seq1) against the rows in
 * `rows[0..n_rows)` (NULL = every row 0..n_rows_in_store-1), skipping the query itself
 * (:166) and removed rows.  Survivors (dist <= ceiling) are written to out[0..*n_out) in
 * no particular order (the reference's order is arbitrary too, :161).  An invalid
 * query yields zero survivors (:438-442). */
int fn3_one_vs_all(fn3_engine* e, int64_t query_row, int32_t ceiling,
                   const int64_t* rows, int64_t n_rows,
                   fn3_hit* out, int64_t cap, int64_t* n_out);

/* insert() of the reference's server — persist(obj, guid) immediately followed by mcompare(guid)
 * (findNeighbour3-server.py:516, :530) — as ONE call: the sample is appended (*out_row receives its row) and
 * compared against every EARLIER row; one stream, one host synchronisation.  The row stays inserted even if
 * the call then fails with FN3_ECAPACITY (repeat the query with fn3_one_vs_all and a larger buffer). */
int fn3_insert_query(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                     int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out);

/* Same, for a query that is NOT in the store (lists as in fn3_append): the sharded
 * host layer uses it on the ranks that do not own the query row. */
int fn3_lists_vs_all(fn3_engine* e, const int32_t* pos, const int32_t off[7], int32_t invalid,
                     int32_t ceiling, const int64_t* rows, int64_t n_rows,
                     fn3_hit* out, int64_t cap, int64_t* n_out);

/* distmat(half, diagonal=False) (:174-197): every ordered pair (row1,row2), row1 != row2,
 * of live rows in [row_begin,row_end) x [0,n_rows) — query rows are the block row this
 * caller owns (multi-GPU: one block row set per rank).  upper_only != 0 keeps only
 * row1 < row2 (`half=True` with guids ordered as rows).  `stride`/`phase`: only query
 * rows r with r % stride == phase are taken (round-robin block-row ownership);
 * stride=1, phase=0 = all. */
int fn3_all_vs_all(fn3_engine* e, int32_t ceiling, int32_t upper_only,
                   int64_t row_begin, int64_t row_end, int64_t stride, int64_t phase,
                   fn3_edge* out, int64_t cap, int64_t* n_out);

/* When fn3_all_vs_all returns FN3_ECAPACITY the edges stay on the device: fetch them into a buffer of at least
 * *n_out entries without recomputing (valid until the next query call on this engine). */
int fn3_all_vs_all_fetch(fn3_engine* e, fn3_edge* out, int64_t cap, int64_t* n_out);

/* countDifferences_byKey((k1,k2), cutoff) (:382-421) on two stored rows.
 * *has_dist = 0 reproduces the reference's None distance (invalid sample or > ceiling). */
int fn3_pair(fn3_engine* e, int64_t row1, int64_t row2, int32_t ceiling,
             fn3_hit* out, int32_t* has_dist);

/* setComparator1/2 + countDifferences (:338-352, :424-458) on two samples that are not
 * in the store (the reference's tests assign _seq1/_seq2 directly, :1710-1711). */
int fn3_pair_lists(fn3_engine* e,
                   const int32_t* pos1, const int32_t off1[7], int32_t invalid1,
                   const int32_t* pos2, const int32_t off2[7], int32_t invalid2,
                   int32_t ceiling, fn3_hit* out, int32_t* has_dist);

/* ---- ingest: compress() on the device (seqComparer.py:274-307) ---- */

/* Upload the reference sequence (len = genome_length, bytes 'A','C','G','T') and the
 * excluded positions (sorted, may be NULL/0) once; needed only by fn3_compress*. */
int fn3_set_reference(fn3_engine* e, const char* reference, const int32_t* excluded,
                      int64_t n_excluded);

/* compress(sequence): byte-compare `sequence` (len = genome_length, any case-sensitive
 * chars; '-' == 'N', :285) against the reference under the mask and return the six sorted
 * lists plus the M characters (m_chars[i] belongs to M list entry i).  pos_cap in
 * elements; FN3_ECAPACITY with *n_pos = required if too small.  No maxNs decision is
 * taken here (the caller applies :301). */
int fn3_compress(fn3_engine* e, const char* sequence, int32_t* pos, int64_t pos_cap,
                 int32_t off[7], char* m_chars, int64_t m_cap, int64_t* n_pos);

/* findNeighbour3.insert()'s data path in ONE call (src/findNeighbour3-server.py:511-530): the
 * character counts of NucleicAcid.examine (src/NucleicAcid.py:50-58), compress (:274-307), the maxNs decision (:301),
 * persist (:119-124) and mcompare (:149-172) from one host->device copy of `sequence` (len = genome_length, already
 * upper-cased as examine leaves it, NucleicAcid.py:35).
 *   hist     256 counts, one per byte value of the sequence, or NULL (the caller derives composition['A'] ... from it)
 *   pos/off/m_chars/n_pos  the compressed sample as fn3_compress returns it (the server stores it in MongoDB, :547)
 *   *invalid 1 if len(N)+len(M) > max_ns: the sample is stored as {'invalid':1} and has no neighbours
 *   *out_row, out[0..*n_out)  as fn3_insert_query
 * FN3_ECAPACITY from the compress step (pos / m_chars too small; *n_pos = required) stores NOTHING: call again with
 * room.  FN3_ECAPACITY from the query step (out too small; *n_out = required) has stored the row: fetch the
 * neighbours with fn3_one_vs_all(*out_row). */
int fn3_ingest_query(fn3_engine* e, const char* sequence, int64_t max_ns, int32_t ceiling, int64_t* hist,
                     int32_t* pos, int64_t pos_cap, int32_t off[7], char* m_chars, int64_t m_cap, int64_t* n_pos,
                     int32_t* invalid, int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out);

/* The character counts of NucleicAcid.examine (src/NucleicAcid.py:50-58: eighteen bytes.count() passes) as one device
 * pass: hist[b] = number of bytes of value b in bytes[0..n).  No engine needed (the reference's NucleicAcid() takes no
 * arguments); FN3_ECUDA without a usable device — there is no CPU path. */
int fn3_byte_histogram(int32_t device, const uint8_t* bytes, int64_t n, int64_t hist[256]);

/* ---- cluster statistics on the device ------------------- */

/* consensus(compressed_sequences, cutoff_proportion) (:460-501) over STORED rows: position p enters output list b
 * (order A,C,G,T,N; off[0..5] are the 6 CSR offsets into pos) iff at least min_votes of rows[0..n_rows) carry p in
 * their list b.  A row given twice votes twice (the reference counts the objects of its list one by one, :483-490);
 * invalid rows vote for nothing.  The caller turns the reference's float threshold `len(list)*cutoff_proportion`
 * (:493) into min_votes = max(1, smallest integer >= it) and leaves out the objects that the reference's vote
 * ignores (those held in patch form, :485).  FN3_ECAPACITY with *n_pos = required if pos is too small. */
int fn3_consensus(fn3_engine* e, const int64_t* rows, int64_t n_rows, int64_t min_votes,
                  int32_t* pos, int64_t pos_cap, int32_t off[6], int64_t* n_pos);

/* The same vote over samples that are NOT stored: n samples as in fn3_append_bulk (off_in has 6*n+1 entries;
 * the M lists are ignored, as in the reference). */
int fn3_consensus_lists(fn3_engine* e, int64_t n, const int32_t* pos_in, const int64_t* off_in, int64_t min_votes,
                        int32_t* pos, int64_t pos_cap, int32_t off[6], int64_t* n_pos);

/* _msa steps 1-3 (:895-931): the ordered variant positions of the valid stored rows rows[0..n_rows): positions at
 * which the samples show more than one base, a sample without an A/C/G/T call there (N and M included) counting as
 * the reference base.  Invalid rows are rejected (FN3_EARG; the reference raises ValueError, :320).
 * FN3_ECAPACITY with *n_sites = required if sites is too small. */
int fn3_msa_sites(fn3_engine* e, const int64_t* rows, int64_t n_rows, int32_t* sites, int64_t sites_cap,
                  int64_t* n_sites);

/* The same with the rule of the caller's class stated: FN3_SITES_CALLS is fn3_msa_sites (seqComparer._msa,
 * src/seqComparer.py:916-931); FN3_SITES_CALLS_OR_N is the older seqComparer_mt._msa (src/seqComparer_mt.py:1127-1137),
 * where a sample's N at a position also accounts for it: the reference base joins the bases seen at p only when
 * some sample has neither an A/C/G/T call nor an N at p. */
#define FN3_SITES_CALLS       0
#define FN3_SITES_CALLS_OR_N  1
int fn3_msa_sites_rule(fn3_engine* e, const int64_t* rows, int64_t n_rows, int32_t rule, int32_t* sites,
                       int64_t sites_cap, int64_t* n_sites);

/* _msa step 4 and the N/M tallies (:933-950, :905-908, :975-977), and estimate_expected_unk_sites (:664-689):
 * for every row i and every site k (sites strictly increasing), codes[i*n_sites + k] = 0 reference base, 1..4
 * A,C,G,T, 5 N, 6 M (a later list of the reference's scan order A,C,T,G,N,M wins, :941-948); counts[4*i..] =
 * { |N_i|, |M_i|, N characters at the sites, M characters at the sites }.  codes may be NULL (counts only: the
 * last two are then the plain intersections |N_i n sites|, |M_i n sites| of :678-681). */
int fn3_msa_align(fn3_engine* e, const int64_t* rows, int64_t n_rows, const int32_t* sites, int64_t n_sites,
                  uint8_t* codes, int32_t* counts);

/* ---- wire format: the reference's stored samples --------------------- */

/* findNeighbour3 keeps every sample in GridFS as pickle.dumps(obj, protocol=2) of its compressed_sequence dict
 * (mongoStore.py:344-363) and reloads all of them at start-up (findNeighbour3-server.py:358-362).  Turn n such pickles
 * (blob i = blobs[blob_off[i] .. blob_off[i+1]) ) straight into the packed form fn3_append_bulk takes, on host threads,
 * without creating Python objects: pos / off[6n+1] as there; status[i] = 0 parsed, valid sample; 1 parsed,
 * {'invalid': 1}; 2 not a plain compressed_sequence pickle (patch_and_consensus form, numpy scalars, other opcodes) —
 * its six lists are left empty and the caller unpickles it itself.  No engine is involved (pure host function).
 * pos_cap in elements (sum of blob sizes / 2 always suffices); FN3_ECAPACITY with *n_pos = required if too small. */
int fn3_unpickle_samples(int64_t n, const uint8_t* blobs, const int64_t* blob_off, int32_t* pos, int64_t pos_cap,
                         int64_t* off, uint8_t* status, int64_t* n_pos);

/* ---- engine groups: one process, every GPU of the box ------------------------------ */

/* The reference's server constructs ONE seqComparer (findNeighbour3-server.py:340-345) and calls sc.mcompare from that
 * one process (:530).  A group is that object's store spread over n_devices GPUs: global row g lives on device
 * device_ids[g % n_devices] as local row g / n_devices (candidates sharded by insertion order), and every
 * call below fans out to one host thread per device inside the library — each enqueues on its own device's stream and
 * waits for its own completion flag — and merges the thresholded neighbour records.  All row arguments and results are
 * GLOBAL rows; semantics, status codes and capacity protocol are those of the single-engine call of the same name.
 * Where a call needs a row that lives on another device (a stored query against the other shards, all-vs-all of the
 * sharded store, consensus / MSA over rows of several shards) the kernels read it where it lives, over NVLink, through
 * peer access; fn3_group_peer_access() tells whether that is available (without it only the query calls work, through a
 * staging copy; the all-vs-all and cluster calls return FN3_ECUDA).  The same device may be listed more than once
 * (several shards on one GPU: how the group logic is tested on a one-GPU box).  NULL on failure (fn3_last_error(NULL)). */
typedef struct fn3_group fn3_group;
fn3_group*  fn3_group_create(int64_t genome_length, int32_t n_devices, const int32_t* device_ids);
int         fn3_group_destroy(fn3_group* g);
const char* fn3_group_last_error(fn3_group* g);
int32_t     fn3_group_n_devices(fn3_group* g);
int32_t     fn3_group_peer_access(fn3_group* g);
/* the i-th member engine (borrowed): for device-local calls — fn3_compress, fn3_pair_lists, fn3_profile_queries */
fn3_engine* fn3_group_engine(fn3_group* g, int32_t i);
int64_t     fn3_group_launch_count(fn3_group* g);

int fn3_group_append(fn3_group* g, const int32_t* pos, const int32_t off[7], int32_t invalid, int64_t* out_row);
int fn3_group_append_bulk(fn3_group* g, int64_t n, const int32_t* pos, const int64_t* off, const uint8_t* invalid,
                          int64_t* out_first_row);
int fn3_group_remove(fn3_group* g, int64_t row);
int fn3_group_row_get(fn3_group* g, int64_t row, int32_t* pos, int64_t pos_cap, int32_t off[7], int32_t* invalid,
                      int64_t* n_pos);
int fn3_group_stats_get(fn3_group* g, fn3_stats* out);
int fn3_group_set_reference(fn3_group* g, const char* reference, const int32_t* excluded, int64_t n_excluded);

/* persist + mcompare (server.py:516, :530): the owner device appends the sample and compares it with its earlier rows
 * (fn3_insert_query), every other device compares it with its whole shard (fn3_lists_vs_all), all at the same time */
int fn3_group_insert_query(fn3_group* g, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                           int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out);
int fn3_group_ingest_query(fn3_group* g, const char* sequence, int64_t max_ns, int32_t ceiling, int64_t* hist,
                           int32_t* pos, int64_t pos_cap, int32_t off[7], char* m_chars, int64_t m_cap, int64_t* n_pos,
                           int32_t* invalid, int64_t* out_row, fn3_hit* out, int64_t cap, int64_t* n_out);
int fn3_group_one_vs_all(fn3_group* g, int64_t query_row, int32_t ceiling, const int64_t* rows, int64_t n_rows,
                         fn3_hit* out, int64_t cap, int64_t* n_out);
int fn3_group_lists_vs_all(fn3_group* g, const int32_t* pos, const int32_t off[7], int32_t invalid, int32_t ceiling,
                           const int64_t* rows, int64_t n_rows, fn3_hit* out, int64_t cap, int64_t* n_out);
int fn3_group_pair(fn3_group* g, int64_t row1, int64_t row2, int32_t ceiling, fn3_hit* out, int32_t* has_dist);
/* distmat over the sharded store (:174-197): every device takes every live row of the group as a query against its own shard */
int fn3_group_all_vs_all(fn3_group* g, int32_t ceiling, int32_t upper_only, fn3_edge* out, int64_t cap, int64_t* n_out);
int fn3_group_consensus(fn3_group* g, const int64_t* rows, int64_t n_rows, int64_t min_votes, int32_t* pos,
                        int64_t pos_cap, int32_t off[6], int64_t* n_pos);
int fn3_group_consensus_lists(fn3_group* g, int64_t n, const int32_t* pos_in, const int64_t* off_in, int64_t min_votes,
                              int32_t* pos, int64_t pos_cap, int32_t off[6], int64_t* n_pos);
int fn3_group_msa_sites_rule(fn3_group* g, const int64_t* rows, int64_t n_rows, int32_t rule, int32_t* sites,
                             int64_t sites_cap, int64_t* n_sites);
int fn3_group_msa_align(fn3_group* g, const int64_t* rows, int64_t n_rows, const int32_t* sites, int64_t n_sites,
                        uint8_t* codes, int32_t* counts);

/* ---- measurement hooks (bench.py only; no effect on results) -------------------- */

/* Run the one-vs-all DEVICE work (query-map build + compare kernel, inputs already resident in HBM) once
 * for each of the n stored rows in query_rows, queued back to back on the engine's stream with no host
 * synchronisation in between.  ms_compare[i] receives the CUDA-event time of step i's compare kernel (the query
 * structure is built inside it); ms_build[i] that of the separate build kernel a query too large for shared memory needs
 * (0 otherwise: no event is recorded between the step's first event and the kernel); n_hits[i] its survivor count;
 * bytes_touched[i] (may be NULL) the bytes of headers + row data the compare kernel of step i actually
 * loaded, counted by a separate instrumented pass that is NOT timed; n_passed[i] (may be NULL; filled with
 * bytes_touched) the candidates the head filter could not rule out (0 for ceilings that do not use it).  flush_l2 != 0
 * writes a 256 MiB scratch buffer before each step (outside the timed events) so that every step starts with a cold L2. */
int fn3_profile_queries(fn3_engine* e, const int64_t* query_rows, int32_t n, int32_t ceiling, int32_t flush_l2,
                        float* ms_build, float* ms_compare, int64_t* n_hits, int64_t* bytes_touched, int64_t* n_passed);

/* Number of kernel launches issued by this engine since creation (bench.py's gpu_launches). */
int64_t fn3_launch_count(fn3_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* FN3_H */
