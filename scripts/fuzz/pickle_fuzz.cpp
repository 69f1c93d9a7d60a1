// ASan/UBSan fuzz harness for the native reader of the stored sample format (fn3_pickle.h; host code only):
//   cd scripts/fuzz && g++ -O1 -g -std=c++17 -fsanitize=address,undefined pickle_fuzz.cpp -o /tmp/pickle_fuzz && /tmp/pickle_fuzz seed*.bin
// seeds = pickle.dumps(sample, protocol=2) files; 400 000 mutated / random inputs, exact-size heap copies.  Last run: clean.
#include "../../findneighbour3_b200/csrc/fn3_pickle.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <fstream>
#include <iterator>
int main(int argc, char** argv) {
    // seeds: files given on the command line
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 1; i < argc; ++i) { std::ifstream f(argv[i], std::ios::binary); seeds.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
    std::mt19937_64 rng(7);
    long long n_ok = 0, n_inv = 0, n_no = 0;
    for (long long it = 0; it < 400000; ++it) {
        std::vector<uint8_t> b;
        if (!seeds.empty() && (it % 10) != 0) {
            b = seeds[rng() % seeds.size()];
            int k = 1 + rng() % 6;
            for (int j = 0; j < k && !b.empty(); ++j) {
                switch (rng() % 5) {
                    case 0: b[rng() % b.size()] = (uint8_t)rng(); break;
                    case 1: b.erase(b.begin() + rng() % b.size()); break;
                    case 2: b.insert(b.begin() + rng() % (b.size() + 1), (uint8_t)rng()); break;
                    case 3: b.resize(rng() % (b.size() + 1)); break;
                    default: { size_t a = rng() % b.size(); size_t l = rng() % 9; for (size_t q = a; q < b.size() && q < a + l; ++q) b[q] = (uint8_t)(rng() % 2 ? 0xff : 0x00); }
                }
            }
        } else {
            b.resize(rng() % 64);
            for (auto& x : b) x = (uint8_t)rng();
            if (b.size() > 2 && rng() % 2) { b[0] = 0x80; b[1] = 0x02; }
        }
        // exact-size heap copy so that ASan sees any over-read
        uint8_t* p = (uint8_t*)malloc(b.size() ? b.size() : 1);
        if (!b.empty()) memcpy(p, b.data(), b.size());
        fn3pickle::Parsed out;
        int st = fn3pickle::parse_sample(p, (int64_t)b.size(), out);
        free(p);
        if (st == 0) ++n_ok; else if (st == 1) ++n_inv; else ++n_no;
    }
    printf("ok %lld invalid %lld declined %lld\n", n_ok, n_inv, n_no);
    return 0;
}
