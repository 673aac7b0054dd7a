#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "matrix_reader.hpp"
int main(int argc, char **argv) {
    const long iters = argc > 1 ? atol(argv[1]) : 100000;
    std::mt19937_64 rng(12345);
    std::vector<std::string> seeds = {
        "%%MatrixMarket matrix coordinate real general\n% c\n3 3 5\n1 1 2.5\n2 1 -1e0\n3 3 1e-3\n1 2 +0.125E+1 0\n2 2 4\n",
        "%%MatrixMarket matrix coordinate complex symmetric\n4 4 4\n1 1 2.5 0.5\n2 1 -1e0 1\n3 3 1e-3 -2\n4 4 1 1\n",
        "%%MatrixMarket matrix coordinate real symmetric\n2 2 3\n1 1 1.0000000000000000001\n2 1 123456789012345678e-17\n2 2 0x1.8p1\n",
    };
    {   // a binary seed
        std::string b(sfio::BIN_MAGIC, 8);
        int64_t h[4] = {3, 3, 3, 2};
        b.append((const char *)h, 32);
        int64_t ti[2] = {0, 2}, tj[2] = {0, 1};
        double tx[4] = {1, 2, 3, 4};
        b.append((const char *)ti, 16); b.append((const char *)tj, 16); b.append((const char *)tx, 32);
        seeds.push_back(b);
    }
    long ok = 0, bad = 0;
    for (long it = 0; it < iters; ++it) {
        std::string s = seeds[rng() % seeds.size()];
        const int muts = 1 + rng() % 6;
        for (int m = 0; m < muts && !s.empty(); ++m) {
            const size_t pos = rng() % s.size();
            switch (rng() % 5) {
                case 0: s[pos] = (char)(rng() & 0xff); break;
                case 1: s.erase(pos, 1 + rng() % 4); break;
                case 2: s.insert(pos, 1 + rng() % 3, "0123456789-+.eE \n%"[rng() % 18]); break;
                case 3: s.resize(pos); break;
                default: s.insert(pos, s.substr(rng() % s.size(), rng() % 12)); break;
            }
        }
        std::vector<char> buf(s.begin(), s.end());          // exact size: an over-read trips ASan
        sfio::Triplets T;
        int rc = sfio::is_binary(buf.data(), buf.size()) ? sfio::parse_binary(buf.data(), buf.size(), T)
                                                         : sfio::parse_matrix_market(buf.data(), buf.size(), T, 1 + (int)(rng() % 3));
        if (rc == sfio::OK) {
            ++ok;
            const size_t vw = T.complex_field ? 2 : 1;
            if (T.tx.size() != T.ti.size() * vw || T.tj.size() != T.ti.size()) { printf("inconsistent sizes\n"); return 1; }
        } else ++bad;
    }
    printf("fuzz: %ld inputs, %ld parsed, %ld rejected, no crash\n", iters, ok, bad);
    return 0;
}
