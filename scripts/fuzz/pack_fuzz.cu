// ASan/UBSan differential fuzz of the host-side sample validation and row packing of fn3_engine.cu (no GPU needed: the
// engine struct is filled by hand the way fn3_create does):
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -g -I ../../include -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined pack_fuzz.cu -o /tmp/pack_fuzz
//   ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 /tmp/pack_fuzz
// Checks, on random valid and damaged samples over several genome lengths: row_words_needed + pack_row (bulk path) and
// pack_checked (insert path) give the same verdict, the same words and the same list ends; the packed row decodes back to the
// input lists; finish_row stays inside row_alloc_words (exact-size heap buffers); timing of the validation per sample.
#include "../../findneighbour3_b200/csrc/fn3_engine.cu"

#include <chrono>
#include <random>

static void init_engine(fn3_engine& e, long long L) {
    e.L = L;
    int pb = 1; while ((1ll << pb) < L) ++pb;
    e.pos_bits = pb;
    e.max_run = (pb >= 32) ? 1u : (uint32_t)std::min<long long>(1ll << (32 - pb), 1ll << 30);
    e.fshift = 5;
    while (((L - 1) >> e.fshift) + 2 > 16383) ++e.fshift;
    const long long n_coarse = ((L - 1) >> e.fshift) + 2;
    e.ncw = (int)((((n_coarse + 2) + 3) / 4 + 7) & ~7ll);
    e.fp_pad = (uint16_t)(e.ncw * 4 - 1);
    e.fp_dead = (uint16_t)(e.ncw * 4 - 2);
}

int main() {
    std::mt19937_64 rng(11);
    long long n_valid = 0, n_bad = 0;
    double t_checked = 0; long long n_timed = 0;
    for (long long L : {7ll, 31ll, 300ll, 70000ll, 4411532ll}) {
        fn3_engine& e = *new fn3_engine();
        init_engine(e, L);
        for (int it = 0; it < (L > 100000 ? 3000 : 20000); ++it) {
            // six disjoint sorted lists: V lists sparse, N/M in blocks
            std::vector<int32_t> lists[6];
            std::vector<uint8_t> used((size_t)L, 0);
            const int shape = (int)(rng() % 4);
            auto take = [&](int b, long long p) { if (p >= 0 && p < L && !used[(size_t)p]) { used[(size_t)p] = 1; lists[b].push_back((int32_t)p); } };
            const long long nv = shape == 0 ? 0 : (long long)(rng() % (L < 1000 ? (unsigned long long)L : 800ull));
            for (long long i = 0; i < nv; ++i) take((int)(rng() % 4), (long long)(rng() % (unsigned long long)L));
            const int nblocks = shape == 1 ? 0 : (int)(rng() % 30);
            for (int k = 0; k < nblocks; ++k) {
                const long long s = (long long)(rng() % (unsigned long long)L), len = 1 + (long long)(rng() % (shape == 3 ? 3000ull : 40ull));
                for (long long p = s; p < s + len; ++p) take(4 + (int)((rng() % 5) == 0), p);
            }
            // sometimes: the same position in two lists (a hand-made, malformed sample) — must be rejected by both paths
            bool overlap = false;
            if (rng() % 5 == 0) {
                int a = (int)(rng() % 6), b2 = (int)(rng() % 6);
                if (a != b2 && !lists[a].empty()) { lists[b2].push_back(lists[a][rng() % lists[a].size()]); overlap = true; }
            }
            for (int b = 0; b < 6; ++b) std::sort(lists[b].begin(), lists[b].end());
            std::vector<int32_t> pos; long long off[7] = {0};
            for (int b = 0; b < 6; ++b) { pos.insert(pos.end(), lists[b].begin(), lists[b].end()); off[b + 1] = (long long)pos.size(); }
            // damage some of them
            bool damaged = false;
            if (!overlap && !pos.empty() && rng() % 4 == 0) {
                damaged = true;
                switch (rng() % 5) {
                    case 0: pos[rng() % pos.size()] = (int32_t)(L + (long long)(rng() % 5)); break;             // out of range
                    case 1: pos[rng() % pos.size()] = -1 - (int32_t)(rng() % 3); break;                          // negative
                    case 2: { size_t i = rng() % pos.size(); pos[i] = pos[i ? i - 1 : 0]; if (!i) damaged = false; else {
                                  int b = 0; while (off[b + 1] <= (long long)i) ++b; if (off[b] == (long long)i) damaged = false; } break; }   // duplicate inside a list
                    case 3: std::swap(off[2], off[3]); if (off[2] == off[3]) damaged = false; break;             // decreasing offsets
                    default: { size_t i = rng() % pos.size(), j = rng() % pos.size(); if (pos[i] == pos[j]) damaged = false; std::swap(pos[i], pos[j]);
                               // a swap may keep every list sorted only if i == j
                               if (i == j) damaged = false; else damaged = true; }
                }
            }
            // exact-size copies so that ASan sees over-reads
            int32_t* ppos = (int32_t*)malloc(pos.size() * 4 + 1);
            if (!pos.empty()) memcpy(ppos, pos.data(), pos.size() * 4);
            uint32_t nblkA = 0, nblkB = 0;
            long long wordsA = row_words_needed(&e, ppos, off, 0, &nblkA);
            if (wordsA >= 0 && !raw_disjoint(ppos, off)) wordsA = -1;            // the bulk path's pass 1
            if (overlap && wordsA >= 0) { printf("OVERLAP accepted by the bulk path L=%lld it=%d\n", L, it); return 1; }
            const size_t npos = (size_t)std::max<long long>(off[6] - off[0], 0), nvw = (size_t)std::max<long long>(off[4] - off[0], 0);
            uint32_t* dstB = (uint32_t*)malloc((row_alloc_words(npos, nvw) + 8) * 4);
            uint32_t endsA[6], endsB[6];
            const auto t0 = std::chrono::steady_clock::now();
            const long long wordsB = pack_checked(&e, ppos, off, dstB, endsB, &nblkB);
            t_checked += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); ++n_timed;
            // ground truth by brute force: offsets non-decreasing, positions in range, strictly increasing inside a list, no
            // position in two lists
            bool truth = true;
            for (int b = 0; b < 6; ++b) if (off[b + 1] < off[b]) truth = false;
            if (truth) {
                std::vector<uint8_t> seen((size_t)L, 0);
                for (int b = 0; b < 6 && truth; ++b)
                    for (long long i = off[b]; i < off[b + 1] && truth; ++i) {
                        const long long v = pos[(size_t)i];
                        if (v < 0 || v >= L || (i > off[b] && v <= pos[(size_t)i - 1]) || seen[(size_t)v]) truth = false;
                        else seen[(size_t)v] = 1;
                    }
            }
            if ((wordsA >= 0) != truth || (wordsB >= 0) != truth) {
                printf("VERDICT L=%lld it=%d truth=%d bulk=%lld insert=%lld (overlap %d damaged %d)\n", L, it, (int)truth, wordsA, wordsB, (int)overlap, (int)damaged);
                return 1;
            }
            if (wordsA >= 0) {
                ++n_valid;
                uint32_t* dstA = (uint32_t*)malloc((row_alloc_words((size_t)wordsA, nvw)) * 4);
                pack_row(&e, ppos, off, dstA, endsA);
                if (wordsA != wordsB || memcmp(endsA, endsB, sizeof endsA) || memcmp(dstA, dstB, (size_t)wordsA * 4)) { printf("PACK differs L=%lld it=%d\n", L, it); return 1; }
                finish_row(&e, dstA, (size_t)wordsA, nvw);                      // must stay inside row_alloc_words(words, nv)
                // decode (as fn3_row_get does) and compare with the input
                std::vector<int32_t> back; const uint32_t pmask = (e.pos_bits >= 32) ? 0xffffffffu : ((1u << e.pos_bits) - 1u);
                uint32_t i = 0;
                for (int b = 0; b < 6; ++b) for (; i < endsA[b]; ++i) {
                    if (b < 4) back.push_back((int32_t)dstA[i]);
                    else { uint32_t s = dstA[i] & pmask, len = (e.pos_bits >= 32) ? 1u : ((dstA[i] >> e.pos_bits) + 1u); for (uint32_t j = 0; j < len; ++j) back.push_back((int32_t)(s + j)); }
                }
                if (back != pos) { printf("ROUNDTRIP differs L=%lld it=%d\n", L, it); return 1; }
                RowHdr h; memset(&h, 0, sizeof h); h.data = 1; for (int b = 0; b < 6; ++b) h.end[b] = endsA[b];
                RowHead hd; fill_head(&e, hd, h, dstA);
                free(dstA);
            } else ++n_bad;
            free(dstB); free(ppos);
        }
    }
    printf("valid %lld rejected %lld; pack_checked %.2f us per sample (all shapes)\n", n_valid, n_bad, 1e6 * t_checked / (double)n_timed);
    return 0;
}
