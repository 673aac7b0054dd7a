#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" long sfcheck_factorize_complex(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cz,
                               unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int row_phase,
                               double *Lsz_out, long *stats, double *rhs);
extern "C" long sfcheck_factorize(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cx,
                       unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int ld_align,
                       double *Lsx_out, long *stats, long out_len);
extern "C" long sfcheck_factorize_complex_partitioned(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cz,
                                           unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int world,
                                           int row_phase, double *Lsz_out);
extern "C" int sfb200_set_ordering_processes(int);
namespace sfb200 { void set_ordering_processes(int); }
int main() {
    for (int kind = 0; kind < 2; ++kind)
    for (int g : {5, 9, 14}) {
        const long n = (long)g * g * g;
        std::vector<long> cp(n + 1, 0), ci; std::vector<std::complex<double>> cz; std::vector<double> cx;
        for (long v = 0; v < n; ++v) {
            const long i = v / (g * g), j = (v / g) % g, k = v % g;
            auto push = [&](long row, double re, double im) { ci.push_back(row); cz.emplace_back(re, im); cx.push_back(re); };
            push(v, 6.0, kind ? 0.7 : 0.0);
            if (k + 1 < g) push(v + 1, -1.0, 0.3);
            if (j + 1 < g) push(v + g, -1.0, -0.2);
            if (i + 1 < g) push(v + g * g, -1.0, 0.1);
            if (kind) {   // unsymmetric: upper entries too
                if (k > 0) push(v - 1, -0.8, 0.1);
                if (j > 0) push(v - g, -0.9, 0.2);
            }
            cp[v + 1] = (long)ci.size();
        }
        // sort rows within columns is not required
        long stats[8] = {0};
        long rc = sfcheck_factorize_complex(kind, kind ? 0 : 1, n, cp.data(), ci.data(), (const double *)cz.data(), 5380000000ull, 0, 0, -1, -1, kind, nullptr, stats, nullptr);
        std::vector<double> out(2 * (size_t)stats[1] + 2), rhs(2 * (size_t)n, 1.0);
        rc = sfcheck_factorize_complex(kind, kind ? 0 : 1, n, cp.data(), ci.data(), (const double *)cz.data(), 5380000000ull, 8, 16, 0, 0, kind, out.data(), stats, rhs.data());
        printf("kind %d g %d: rc %ld nsuper %ld xsize %ld launches %ld\n", kind, g, rc, stats[0], stats[1], stats[3]);
        if (g == 9) {   // the same matrix on 3 ranks (sfb200_zgroup's composition), small thresholds so that fronts are shared work
            setenv("SFCHECK_DIST_PANEL_MIN", "32", 1); setenv("SFCHECK_DIST_MIN_M", "16", 1);
            std::vector<double> out3(out.size());
            rc = sfcheck_factorize_complex_partitioned(kind, kind ? 0 : 1, n, cp.data(), ci.data(), (const double *)cz.data(), 5380000000ull, 8, 16, 0, 0, 3, kind, out3.data());
            double dmax = 0;
            for (size_t q = 0; q < 2 * (size_t)stats[1]; ++q) dmax = std::max(dmax, std::abs(out3[q] - out[q]));
            printf("kind %d g %d on 3 ranks: rc %ld, max |difference to one rank| %.2e\n", kind, g, rc, dmax);
        }
    }
    // parallel ordering under the sanitizer (forks)
    sfb200::set_ordering_processes(4);
    {
        const int g = 30; const long n = (long)g * g * g;
        std::vector<long> cp(n + 1, 0), ci; std::vector<double> cx;
        for (long v = 0; v < n; ++v) {
            const long i = v / (g * g), j = (v / g) % g, k = v % g;
            ci.push_back(v); cx.push_back(6.0);
            if (k + 1 < g) { ci.push_back(v + 1); cx.push_back(-1.0); }
            if (j + 1 < g) { ci.push_back(v + g); cx.push_back(-1.0); }
            if (i + 1 < g) { ci.push_back(v + g * g); cx.push_back(-1.0); }
            cp[v + 1] = (long)ci.size();
        }
        long stats[8] = {0};
        long rc = sfcheck_factorize(0, 1, n, cp.data(), ci.data(), cx.data(), 5380000000ull, 0, 0, -1, -1, 0, nullptr, stats, 0);
        printf("parallel order lap3d 30: rc %ld xsize %ld\n", rc, stats[6]);
    }
    return 0;
}
