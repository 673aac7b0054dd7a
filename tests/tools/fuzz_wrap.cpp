// A structure handed to sfb200_symbolic_wrap by somebody else's analysis, damaged in one or a few places: the call must
// answer SFB200_ERR_ARG or hand back something build_plan / embed_complex can digest -- never index out of bounds (ASan).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../include/sfb200.h"
#include "capi_internal.hpp"
#include "complex_embed.hpp"
#include "plan.hpp"
using namespace sfb200;
int main(int argc, char **argv) {
    const long iters = argc > 1 ? atol(argv[1]) : 20000;
    std::mt19937_64 rng(777);
    long refused = 0, planned = 0, plan_refused = 0;
    for (int kind = 0; kind < 2; ++kind) {
        const int g = 7; const i64 n = (i64)g * g * g;
        std::vector<i64> cp(n + 1, 0), ci; std::vector<double> cx;
        for (i64 v = 0; v < n; ++v) {
            const i64 i = v / (g * g), j = (v / g) % g, k = v % g;
            auto push = [&](i64 row, double x) { ci.push_back(row); cx.push_back(x); };
            push(v, 6.0);
            if (k + 1 < g) push(v + 1, -1.0);
            if (j + 1 < g) push(v + g, -1.0);
            if (i + 1 < g) push(v + g * g, -1.0);
            if (kind) { if (k > 0) push(v - 1, -0.8); if (j > 0) push(v - g, -0.9); }
            cp[v + 1] = (i64)ci.size();
        }
        sfb200_symbolic *s0 = nullptr;
        if (sfb200_analyze(kind, kind ? 0 : 1, n, cp.data(), ci.data(), cx.data(), 5380000000ull, &s0)) return 2;
        const Symbolic &S = s0->S;
        for (long it = 0; it < iters; ++it) {
            std::vector<i64> Super = S.Super, Lsip = S.Lsip, Lsxp = S.Lsxp, Lsi = S.Lsi, Lp = S.L.ptr, Li = S.L.idx, Up = S.U.ptr, Ui = S.U.idx;
            std::vector<i64> *arrs[] = {&Super, &Lsip, &Lsxp, &Lsi, &Lp, &Li, &Up, &Ui};
            const int muts = 1 + (int)(rng() % 3);
            for (int m = 0; m < muts; ++m) {
                std::vector<i64> &a = *arrs[rng() % (kind ? 8 : 6)];
                if (a.empty()) continue;
                i64 &x = a[rng() % a.size()];
                switch (rng() % 5) {
                    case 0: x += 1; break;
                    case 1: x -= 1; break;
                    case 2: x = (i64)(rng() % (2 * n)) - n / 2; break;
                    case 3: x = -x; break;
                    default: x = (i64)1 << (rng() % 62); break;
                }
            }
            // the API's contract: the index arrays are as long as the last pointer entries say (Lp[n], Up[n], Lsip[nsuper]) --
            // a count the callee cannot verify from plain pointers; the drop-in cross-checks them with nzmax / isize
            if (Lp[n] > S.L.ptr[n]) Lp[n] = S.L.ptr[n];
            if (kind && Up[n] > S.U.ptr[n]) Up[n] = S.U.ptr[n];
            if (Lsip[S.nsuper] > S.Lsip[S.nsuper]) Lsip[S.nsuper] = S.Lsip[S.nsuper];
            sfb200_symbolic *w = nullptr;
            const bool cplx = (rng() & 1) != 0;
            const int rc = (cplx ? sfb200_symbolic_wrap_complex : sfb200_symbolic_wrap)(
                kind, n, S.nsuper, Super.data(), Lsip.data(), Lsxp.data(), Lsi.data(), Lp.data(), Li.data(),
                kind ? Up.data() : nullptr, kind ? Ui.data() : nullptr, &w);
            if (rc != SFB200_OK) { ++refused; continue; }
            PlanParams prm; HostPlan P;
            if (cplx) {
                ComplexEmbedding E;
                if (embed_complex(w->S, E) != 0 || build_plan(E.R, prm, P) != 0) ++plan_refused; else ++planned;
            } else {
                if (build_plan(w->S, prm, P) != 0) ++plan_refused; else ++planned;
            }
            sfb200_symbolic_free(w);
        }
        sfb200_symbolic_free(s0);
    }
    printf("fuzz wrap: %ld refused by sfb200_symbolic_wrap, %ld by the plan builder, %ld accepted; no crash\n", refused, plan_refused, planned);
    return 0;
}
