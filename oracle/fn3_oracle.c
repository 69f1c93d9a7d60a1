/*
 * fn3_oracle.c — TEST INFRASTRUCTURE: plain-C CPU restatement of the reference's masked
 * pairwise SNV distance (davidhwyllie/findNeighbour3, src/seqComparer.py).
 *
 * It restates the reference's SET ALGEBRA on sorted int32 arrays (merge-based union /
 * difference / symmetric difference), step by step as the reference does it:
 *   countDifferences        seqComparer.py:424-458
 *   countDifferences_byKey  seqComparer.py:382-421  (incl. the n2 quirk at :417)
 *   _setStats               seqComparer.py:354-380
 *   mcompare                seqComparer.py:149-172  (fn3o_one_vs_all; threaded like
 *                           seqComparer_mt.py:363-385: contiguous candidate ranges per thread)
 * It does NOT use the streaming identity of the CUDA kernel, so the two implementations
 * are independent.  Pinned against the unmodified reference through tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
 * load this library.  Build: gcc -O2 -shared -fPIC -pthread (see oracle/build_oracle.py).
 *
 * Sample format = the C-ABI's: six sorted disjoint lists A,C,G,T,N,M concatenated, 7 offsets.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const int32_t* p; int64_t n; } set_t;      /* a sorted set of positions */

static int64_t set_union(set_t a, set_t b, int32_t* out) {
    int64_t i = 0, j = 0, k = 0;
    while (i < a.n && j < b.n) {
        if (a.p[i] < b.p[j]) out[k++] = a.p[i++];
        else if (a.p[i] > b.p[j]) out[k++] = b.p[j++];
        else { out[k++] = a.p[i]; ++i; ++j; }
    }
    while (i < a.n) out[k++] = a.p[i++];
    while (j < b.n) out[k++] = b.p[j++];
    return k;
}

static int64_t set_minus(set_t a, set_t b, int32_t* out) {      /* a - b */
    int64_t i = 0, j = 0, k = 0;
    while (i < a.n) {
        while (j < b.n && b.p[j] < a.p[i]) ++j;
        if (j < b.n && b.p[j] == a.p[i]) { ++i; continue; }
        out[k++] = a.p[i++];
    }
    return k;
}

static int64_t set_xor(set_t a, set_t b, int32_t* out) {        /* a ^ b */
    int64_t i = 0, j = 0, k = 0;
    while (i < a.n && j < b.n) {
        if (a.p[i] < b.p[j]) out[k++] = a.p[i++];
        else if (a.p[i] > b.p[j]) out[k++] = b.p[j++];
        else { ++i; ++j; }
    }
    while (i < a.n) out[k++] = a.p[i++];
    while (j < b.n) out[k++] = b.p[j++];
    return k;
}

static set_t list_of(const int32_t* pos, const int64_t* off, int b) {
    set_t s; s.p = pos + off[b]; s.n = off[b + 1] - off[b]; return s;
}

typedef struct { int32_t *u1, *u2, *t1, *t2, *x, *diff, *tmp; int64_t cap; } scratch_t;

static int scratch_fit(scratch_t* s, int64_t need) {
    if (need <= s->cap) return 0;
    int64_t cap = need * 2 + 64;
    int32_t** f[7] = {&s->u1, &s->u2, &s->t1, &s->t2, &s->x, &s->diff, &s->tmp};
    for (int i = 0; i < 7; ++i) {
        int32_t* p = (int32_t*)realloc(*f[i], (size_t)cap * sizeof(int32_t));
        if (!p) return -1;
        *f[i] = p;
    }
    s->cap = cap;
    return 0;
}

static void scratch_free(scratch_t* s) {
    free(s->u1); free(s->u2); free(s->t1); free(s->t2); free(s->x); free(s->diff); free(s->tmp);
    memset(s, 0, sizeof *s);
}

/* returns 1 and fills out[4] = {dist,n1,n2,nboth} if the pair has a distance <= cutoff, else 0; -1 on OOM */
static int pair_core(const int32_t* pos1, const int64_t* off1, int inv1,
                     const int32_t* pos2, const int64_t* off2, int inv2,
                     int64_t cutoff, int32_t out[4], scratch_t* s) {
    if (inv1 || inv2) return 0;                                   /* :438-442 */
    int64_t tot = (off1[6] - off1[0]) + (off2[6] - off2[0]) + 8;
    if (scratch_fit(s, tot)) return -1;
    /* uncertain positions of each sample: N | set(M.keys())           :449-450 */
    set_t N1 = list_of(pos1, off1, 4), M1 = list_of(pos1, off1, 5);
    set_t N2 = list_of(pos2, off2, 4), M2 = list_of(pos2, off2, 5);
    set_t U1, U2;
    U1.p = s->u1; U1.n = set_union(N1, M1, s->u1);
    U2.p = s->u2; U2.n = set_union(N2, M2, s->u2);
    /* differing_positions = union over nucleotides C,G,A,T of (S1[x]-U2) ^ (S2[x]-U1)   :446-451 */
    static const int order[4] = {1, 2, 0, 3};                     /* C,G,A,T in A,C,G,T list order */
    set_t diff; diff.p = s->diff; diff.n = 0;
    for (int k = 0; k < 4; ++k) {
        set_t a = list_of(pos1, off1, order[k]), b = list_of(pos2, off2, order[k]);
        set_t na, nb, x;
        na.p = s->t1; na.n = set_minus(a, U2, s->t1);
        nb.p = s->t2; nb.n = set_minus(b, U1, s->t2);
        x.p = s->x; x.n = set_xor(na, nb, s->x);
        int64_t n = set_union(diff, x, s->tmp);
        int32_t* sw = s->diff; s->diff = s->tmp; s->tmp = sw;
        diff.p = s->diff; diff.n = n;
    }
    if (diff.n > cutoff) return 0;                                /* :455 */
    /* statistics: seq1_uncertain = N1|M1keys ; seq2_uncertain = N2|M1keys (sic) ; union   :416-418, :378 */
    set_t s2u; s2u.p = s->t1; s2u.n = set_union(N2, M1, s->t1);
    int64_t nboth = set_union(U1, s2u, s->t2);
    out[0] = (int32_t)diff.n; out[1] = (int32_t)U1.n; out[2] = (int32_t)s2u.n; out[3] = (int32_t)nboth;
    return 1;
}

/* ---- exported ---------------------------------------------------------------------------- */

int fn3o_pair(const int32_t* pos1, const int32_t off1[7], int inv1,
              const int32_t* pos2, const int32_t off2[7], int inv2, int64_t cutoff, int32_t out[4]) {
    int64_t o1[7], o2[7];
    for (int i = 0; i < 7; ++i) { o1[i] = inv1 ? 0 : off1[i]; o2[i] = inv2 ? 0 : off2[i]; }
    scratch_t s; memset(&s, 0, sizeof s);
    int r = pair_core(pos1, o1, inv1, pos2, o2, inv2, cutoff, out, &s);
    scratch_free(&s);
    return r;
}

typedef struct {
    const int32_t* pos; const int64_t* off; const uint8_t* invalid;
    int64_t query, begin, end, cutoff;
    int64_t* out_row; int32_t* out_vals; int64_t n_out; int err;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    scratch_t s; memset(&s, 0, sizeof s);
    const int64_t* qo = j->off + 6 * j->query;
    int qinv = j->invalid ? j->invalid[j->query] : 0;
    for (int64_t c = j->begin; c < j->end; ++c) {
        if (c == j->query) continue;                              /* mcompare skips the self pair :166 */
        int32_t v[4];
        int r = pair_core(j->pos, qo, qinv, j->pos, j->off + 6 * c, j->invalid ? j->invalid[c] : 0, j->cutoff, v, &s);
        if (r < 0) { j->err = 1; break; }
        if (r == 1) {
            j->out_row[j->n_out] = c;
            memcpy(j->out_vals + 4 * j->n_out, v, sizeof v);
            j->n_out++;
        }
    }
    scratch_free(&s);
    return NULL;
}

/* one-vs-all over a store given as concatenated lists: off has 6n+1 entries.
 * Survivors (dist <= cutoff) are written as rows + 4 ints each; returns their count (or -1).
 * out_row / out_vals need room for n entries.  nthreads >= 1; candidate ranges are contiguous
 * per thread as in seqComparer_mt.mcompare (seqComparer_mt.py:365-377). */
int64_t fn3o_one_vs_all(const int32_t* pos, const int64_t* off, const uint8_t* invalid, int64_t n,
                        int64_t query, int64_t cutoff, int nthreads, int64_t* out_row, int32_t* out_vals) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    job_t jobs[256]; pthread_t th[256];
    int64_t per = (n + nthreads - 1) / nthreads;
    int64_t* rows = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    int32_t* vals = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * 4 * sizeof(int32_t));
    if (!rows || !vals) { free(rows); free(vals); return -1; }
    for (int t = 0; t < nthreads; ++t) {
        job_t* j = &jobs[t];
        j->pos = pos; j->off = off; j->invalid = invalid; j->query = query; j->cutoff = cutoff;
        j->begin = per * t < n ? per * t : n; j->end = per * (t + 1) < n ? per * (t + 1) : n;
        j->out_row = rows + j->begin; j->out_vals = vals + 4 * j->begin; j->n_out = 0; j->err = 0;
        if (nthreads == 1) worker(j); else pthread_create(&th[t], NULL, worker, j);
    }
    int64_t total = 0; int err = 0;
    for (int t = 0; t < nthreads; ++t) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        job_t* j = &jobs[t];
        if (j->err) err = 1;
        memcpy(out_row + total, j->out_row, (size_t)j->n_out * sizeof(int64_t));
        memcpy(out_vals + 4 * total, j->out_vals, (size_t)j->n_out * 4 * sizeof(int32_t));
        total += j->n_out;
    }
    free(rows); free(vals);
    return err ? -1 : total;
}
