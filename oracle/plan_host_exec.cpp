/*
 * oracle/plan_host_exec.cpp -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * A CPU interpreter of the engine's static schedule (csrc/plan.hpp).  It walks the same launch list the GPU
 * replays and applies every op with plain loops, so that the schedule itself -- levels, blocked panel
 * factorization, ancestor segments, relative-index tables, assembly map, padded layout, copy-out -- can be
 * checked against the compiled reference on a box without a GPU.  The arithmetic each op stands for is the
 * reference's: loadA (CH.c:1998-2028 / LU.c:2478-2536), SYRK/GEMM + mapped subtract (CH.c:2055-2086 /
 * LU.c:2564-2604), POTRF + TRSM (CH.c:2134-2145), GETRF-nopiv + TRSM (LU.c:2652-2663).
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "complex_embed.hpp"
#include "plan.hpp"
#include "symbolic.hpp"

using namespace sfb200;

namespace {

int diag_cholesky(double *A, i64 ld, int nb) {
    for (int j = 0; j < nb; ++j) {
        double d = A[j + j * ld];
        if (!(d > 0.0)) return 1;
        d = std::sqrt(d);
        A[j + j * ld] = d;
        for (int i = j + 1; i < nb; ++i) A[i + j * ld] /= d;
        for (int c = j + 1; c < nb; ++c)
            for (int i = c; i < nb; ++i) A[i + c * ld] -= A[i + j * ld] * A[c + j * ld];
    }
    return 0;
}

int diag_lu(double *A, i64 ld, int nb) {
    for (int k = 0; k < nb; ++k) {
        const double piv = A[k + k * ld];
        if (piv == 0.0) return 1;
        for (int i = k + 1; i < nb; ++i) A[i + k * ld] /= piv;
        for (int c = k + 1; c < nb; ++c)
            for (int i = k + 1; i < nb; ++i) A[i + c * ld] -= A[i + k * ld] * A[k + c * ld];
    }
    return 0;
}

// leaf look-ahead: the update a G_SKIP_HEAD GEMM left to this diagonal block
void diag_pending_update(double *X, const DiagOp &d, bool lu) {
    if (d.upd_k <= 0) return;
    double *A = X + d.a_off;
    const double *L = X + d.l_off, *U = X + d.u_off;
    for (int c = 0; c < d.nb; ++c)
        for (int r = lu ? 0 : c; r < d.nb; ++r) {
            double acc = 0.0;
            for (int k = 0; k < d.upd_k; ++k) acc += L[r + (i64)k * d.ld] * (lu ? U[k + (i64)c * d.ld] : L[c + (i64)k * d.ld]);
            A[r + (i64)c * d.ld] -= acc;
        }
}

void trsm(double *X, const TrsmOp &t) {
    const double *D = X + t.t_off;
    for (int v = 0; v < t.nvec; ++v) {
        double *b = X + t.b_off + (i64)v * t.vstride;
        for (int j = 0; j < t.nb; ++j) {
            double acc = b[(i64)j * t.estride];
            for (int p = 0; p < j; ++p) {
                const double tjp = (t.flags & T_TRANS) ? D[p + (i64)j * t.ldt] : D[j + (i64)p * t.ldt];
                acc -= b[(i64)p * t.estride] * tjp;
            }
            if (!(t.flags & T_UNIT)) acc /= D[j + (i64)j * t.ldt];
            b[(i64)j * t.estride] = acc;
            if (t.flags & T_COPY_T) X[t.w_off + v + (i64)j * t.ldw] = acc;
        }
    }
}

// The scatter epilogues of the GPU kernels add with red.global.add.f64: several ranks (processes of the gloo test tier,
// working in each other's shared-memory images) may hit the same entry of a shared panel at the same time.
static inline void atomic_sub(double *p, double v) {
    uint64_t *u = reinterpret_cast<uint64_t *>(p);
    uint64_t old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
    do {
        double d;
        std::memcpy(&d, &old, sizeof(d));
        d -= v;
        std::memcpy(&neu, &d, sizeof(d));
    } while (!__atomic_compare_exchange_n(u, &old, neu, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

// one op; B = tile size of its launch (tiles matter only when the op is shared between ranks: tile_mod > 1).
// img[r] = image of rank r: a scattered entry goes into the image of the rank its upd_coloff names (the peer-memory adds of
// the GPU kernels); X = this rank's own image.
void gemm(double *X, double *const *img, const GemmOp &g, const HostPlan &P, int B = 128) {
    const double *A = X + g.a_off, *Bm = X + g.b_off;
    const bool lower = (g.flags & G_LOWER) != 0;
    const i64 ntile = gemm_tiles(g.M, g.N, B, lower);
    for (i64 t = g.tile_rem; t < ntile; t += g.tile_mod) {
        int ti, tj;
        decode_tile(t, g.M, g.N, B, lower, ti, tj);
        for (int j = tj * B; j < std::min(g.N, (tj + 1) * B); ++j) {
            int lim = 0;
            i64 col = 0, relbase = 0;
            double *D = X;
            if (g.flags & G_SCATTER) {
                const i64 q = g.meta + j;
                col = P.upd_coloff[q] & kOffsetMask;
                D = img[P.upd_coloff[q] >> kOwnerShift];
                relbase = P.upd_relbase[q];
                lim = (g.flags & G_LIM_SEGB) ? P.upd_segb[q] : (g.flags & G_LIM_SEGE) ? P.upd_sege[q] : j;
                if (g.flags & G_LIM_SEGE) col += P.upd_uoff[q];
            } else if (lower) {
                lim = j;
            }
            if ((g.flags & G_SKIP_HEAD) && j < (int)g.meta) lim = std::max(lim, (int)g.meta);   // the next diagonal block's own
            for (int i = std::max(lim, ti * B); i < std::min(g.M, (ti + 1) * B); ++i) {
                double acc = 0.0;
                for (int k = 0; k < g.K; ++k) {
                    const double b = (g.flags & G_B_KMAJOR) ? Bm[k + (i64)j * g.ldb] : Bm[j + (i64)k * g.ldb];
                    acc += A[i + (i64)k * g.lda] * b;
                }
                if (g.flags & G_SCATTER) atomic_sub(D + col + P.relmap[relbase + i], acc);
                else X[g.c_off + i + (i64)j * g.ldc] -= acc;
            }
        }
    }
}

int small_front(double *X, double *const *img, const HostPlan &P, i64 s) {
    const SnodeDev &d = P.snode[s];
    const int n = d.nscol, R = d.nsrow, m = R - n;
    double *A = X + d.xoff;
    const i64 ld = d.ld;
    GemmOp g{};
    g.lda = g.ldb = d.ld; g.M = g.N = m; g.K = n; g.meta = d.lsip + n; g.tile_mod = 1; g.tile_rem = 0;
    if (P.kind == Kind::Cholesky) {
        for (int j = 0; j < n; ++j) {
            double dj = A[j + j * ld];
            if (!(dj > 0.0)) return 1;
            dj = std::sqrt(dj);
            A[j + j * ld] = dj;
            for (int i = j + 1; i < R; ++i) A[i + j * ld] /= dj;
            for (int c = j + 1; c < n; ++c)
                for (int i = c; i < R; ++i) A[i + c * ld] -= A[i + j * ld] * A[c + j * ld];
        }
        if (m > 0) {
            g.a_off = g.b_off = d.xoff + n; g.flags = G_LOWER | G_SCATTER;
            gemm(X, img, g, P);
        }
    } else {
        const int rows = R + m;      // L11\U11, L21, U12^T
        for (int k = 0; k < n; ++k) {
            const double piv = A[k + k * ld];
            if (piv == 0.0) return 1;
            for (int i = k + 1; i < R; ++i) A[i + k * ld] /= piv;                      // L column
            for (int c = k + 1; c < n; ++c) {                                          // trailing columns
                const double u = A[k + c * ld];
                for (int i = k + 1; i < R; ++i) A[i + c * ld] -= A[i + k * ld] * u;
                for (int i = R; i < rows; ++i) A[i + c * ld] -= A[i + k * ld] * A[c + k * ld];   // U12^T -= U12^T[:,k] L[c,k]
            }
        }
        if (m > 0) {
            g.a_off = d.xoff + n; g.b_off = d.xoff + R; g.flags = G_SCATTER | G_LIM_SEGB;
            gemm(X, img, g, P);
            g.a_off = d.xoff + R; g.b_off = d.xoff + n; g.flags = G_SCATTER | G_LIM_SEGE;
            gemm(X, img, g, P);
        }
    }
    return 0;
}

struct CheckHandle {
    Symbolic S;
    HostPlan P;
};

// a complex matrix on the interpreter: Z = its analysis, the base handle holds the real embedding (complex_embed.hpp)
struct ZCheckHandle {
    Symbolic Z;
    ComplexEmbedding E;
    CheckHandle base;
    std::vector<double> phase;      // empty: rows not rotated
    std::vector<ZExtractTask> tasks;
};

void assemble(const CheckHandle &h, double *X) {
    const Symbolic &S = h.S;
    const HostPlan &P = h.P;
    for (i64 p = 0; p < P.nnz; ++p)
        if (P.asm_dst_l[p] >= 0) X[P.asm_dst_l[p]] = S.L.val[p];
    if (P.kind == Kind::LU) {
        const Tri &U = P.split ? S.U : S.L;
        for (i64 p = 0; p < P.unnz; ++p)
            if (P.asm_dst_u[p] >= 0) X[P.asm_dst_u[p]] = U.val[p];
    }
}

long run_one(const HostPlan &P, double *X, double *const *img, const Launch &L) {
    long bad = 0;
    switch (L.kind) {
        case L_SMALL_A: case L_SMALL_B:
            for (int k = 0; k < L.count; ++k) bad += small_front(X, img, P, P.small_tasks[L.first + k]);
            break;
        case L_POTRF:
            for (int k = 0; k < L.count; ++k) { const DiagOp &d = P.diag_ops[L.first + k]; diag_pending_update(X, d, false); bad += diag_cholesky(X + d.a_off, d.ld, d.nb); }
            break;
        case L_GETRF:
            for (int k = 0; k < L.count; ++k) { const DiagOp &d = P.diag_ops[L.first + k]; diag_pending_update(X, d, true); bad += diag_lu(X + d.a_off, d.ld, d.nb); }
            break;
        case L_TRSM: case L_TRSM_MMA:
            for (int k = 0; k < L.count; ++k) trsm(X, P.trsm_ops[L.first + k]);
            break;
        case L_GEMM128: case L_GEMM64: case L_GEMM128WS: {
            // the tile prefix must be consistent with the geometry the kernels decode
            i64 tiles = 0;
            const int B = L.kind == L_GEMM64 ? 64 : 128;
            for (int k = 0; k < L.count; ++k) {
                const GemmOp &g = P.gemm_ops[L.first + k];
                if (g.tile0 != tiles) return -101;
                tiles += share_of_tiles(gemm_tiles(g.M, g.N, B, (g.flags & G_LOWER) != 0), g.tile_mod, g.tile_rem);
                gemm(X, img, g, P, B);
            }
            if (tiles != L.grid) return -102;
            break;
        }
        default: return -103;
    }
    return bad;
}

/* The engine puts LF_SIDE launches on a second stream: they may run any time between their issue point and the next
 * LF_WAIT_SIDE launch (or the end of the range, where the engine joins the side stream); LF_DIAG_AHEAD launches likewise
 * up to the next LF_WAIT_DIAG launch.  The default order here is
 * "immediately"; SFCHECK_SIDE_LATE=1 runs them as LATE as that contract allows, so that a side launch that overlaps
 * something the main stream does in between shows up as a wrong factor in the CPU test tier. */
long run_launches(const HostPlan &P, double *X, double *const *img, int first, int count) {
    const bool late = getenv("SFCHECK_SIDE_LATE") != nullptr;
    std::vector<int> deferred, deferred_diag;
    long bad = 0;
    auto flush = [&](std::vector<int> &q) -> long {
        for (int li : q) { const long b = run_one(P, X, img, P.launches[li]); if (b < 0) return b; bad += b; }
        q.clear();
        return 0;
    };
    for (int li = first; li < first + count; ++li) {
        const Launch &L = P.launches[li];
        if (late && (L.flags & LF_SIDE)) { deferred.push_back(li); continue; }
        if (late && (L.flags & LF_DIAG_AHEAD)) { deferred_diag.push_back(li); continue; }
        if (L.flags & LF_WAIT_SIDE) { const long b = flush(deferred); if (b < 0) return b; }
        if (L.flags & LF_WAIT_DIAG) { const long b = flush(deferred_diag); if (b < 0) return b; }
        const long b = run_one(P, X, img, L);
        if (b < 0) return b;
        bad += b;
    }
    long b = flush(deferred);
    if (b < 0) return b;
    b = flush(deferred_diag);
    return b < 0 ? b : bad;
}

void pack(const HostPlan &P, const double *X, double *Lsx_out) {
    for (const PackTask &t : P.pack_tasks)
        for (int c = 0; c < t.cols; ++c)
            std::memcpy(Lsx_out + t.dst + (i64)c * t.rows, X + t.src + (i64)c * t.ld, sizeof(double) * t.rows);
}

}  // namespace

extern "C" {

/* ---- handle API: one rank of a `world`-way etree partition (world = 1: the whole factorization) ------------------ */
void *sfcheck_open(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cx,
                   unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int ld_align, int rank,
                   int world) {
    CscView A;
    A.n = n; A.colptr = (const i64 *)Cp; A.rowind = (const i64 *)Ci; A.val = Cx; A.symmetric = is_symmetric != 0;
    auto *h = new CheckHandle();
    if (analyze(A, kind ? Kind::LU : Kind::Cholesky, dev_slot, h->S) != 0) { delete h; return nullptr; }
    PlanParams prm;
    if (nb > 0) prm.nb = nb;
    if (outer > 0) prm.outer = outer % 100000;
    if (outer >= 100000) prm.outer2 = outer / 100000;     // test hook: outer = outer2 * 100000 + outer
    if (small_a >= 0) prm.small_a_bytes = small_a;
    if (small_b >= 0) prm.small_b_bytes = small_b;
    if (ld_align > 0) prm.ld_align = ld_align;
    if (const char *e = getenv("SFCHECK_DIST_MIN_M")) prm.dist_scatter_min_m = atoi(e);
    if (const char *e = getenv("SFCHECK_DIST_PANEL_MIN")) prm.dist_panel_min_cols = atoi(e);
    if (build_plan(h->S, prm, h->P, rank, world) != 0) { delete h; return nullptr; }
    return h;
}
void sfcheck_close(void *hv) { delete (CheckHandle *)hv; }

/* launch i of the schedule: out[8] = {kind, level, grid, ops, flags, algorithmic flops of its GEMM / diag / trsm ops, max K,
 * min K} (K = 0 for non-GEMM launches) -- lets a host-only tool join an ncu launch list with the plan. */
long sfcheck_launch_info(void *hv, int i, double *out) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    if (i < 0 || i >= (int)P.launches.size()) return -1;
    const Launch &L = P.launches[i];
    double flops = 0, kmax = 0, kmin = 0;
    if (L.kind == L_GEMM128 || L.kind == L_GEMM64 || L.kind == L_GEMM128WS) {
        kmin = 1e30;
        for (int q = L.first; q < L.first + L.count; ++q) {
            const GemmOp &g = P.gemm_ops[q];
            const double pairs = (g.flags & G_LOWER) ? (double)g.N * (g.N + 1) / 2 + (double)(g.M - g.N) * g.N : (double)g.M * g.N;
            flops += 2.0 * pairs * g.K / g.tile_mod;
            kmax = std::max(kmax, (double)g.K); kmin = std::min(kmin, (double)g.K);
        }
    } else if (L.kind == L_POTRF || L.kind == L_GETRF) {
        for (int q = L.first; q < L.first + L.count; ++q) { const double nb = P.diag_ops[q].nb; flops += (L.kind == L_GETRF ? 2.0 : 1.0) * nb * nb * nb / 3.0; }
    } else if (L.kind == L_TRSM || L.kind == L_TRSM_MMA) {
        for (int q = L.first; q < L.first + L.count; ++q) flops += (double)P.trsm_ops[q].nb * P.trsm_ops[q].nb * P.trsm_ops[q].nvec;
    }
    out[0] = L.kind; out[1] = L.level; out[2] = L.grid; out[3] = L.count; out[4] = L.flags; out[5] = flops; out[6] = kmax; out[7] = kmin;
    return 0;
}

long sfcheck_int(void *hv, const char *name) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    const std::string k(name);
    if (k == "nlevels") return P.nlevels;
    if (k == "nlaunches") return (long)P.launches.size();
    if (k == "nsmall") return P.nsmall;
    if (k == "nlarge") return P.nlarge;
    if (k == "relmap") return (long)P.relmap.size();
    if (k == "xsize_pad") return P.xsize_pad;
    if (k == "xsize_work") return P.xsize_pad + P.wsize;       // image + this rank's scratch (LU: U11^T copies)
    if (k == "xsize") return P.xsize;
    if (k == "nphases") return (long)P.phases.size();
    if (k == "nsegments") return (long)P.segments.size();
    if (k == "ncolls") return (long)P.colls.size();
    if (k == "nsuper") return P.nsuper;
    return -1;
}
double sfcheck_double(void *hv, const char *name) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    const std::string k(name);
    if (k == "flops") return P.flops;
    if (k == "local_flops") return P.local_flops;
    return -1;
}
/* out[5] = {coll_first, coll_count, launch_first, launch_count, level} */
void sfcheck_segment(void *hv, int k, long *out) {
    const Segment &sg = ((CheckHandle *)hv)->P.segments[k];
    out[0] = sg.coll_first; out[1] = sg.coll_count; out[2] = sg.launch_first; out[3] = sg.launch_count; out[4] = sg.level;
}
/* out[8] = {type, root, lo, hi, offset in this rank's image, rows, cols, ld}; dst (may be NULL) receives, on the root of a
 * C_BCAST, the block's offset in the image of every rank lo..hi-1 */
void sfcheck_coll(void *hv, int i, long *out, long *dst) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    const Coll &c = P.colls[i];
    out[0] = c.type; out[1] = c.root; out[2] = c.lo; out[3] = c.hi; out[4] = c.offset; out[5] = c.rows; out[6] = c.cols; out[7] = c.ld;
    if (dst && c.dst_first >= 0) for (int q = 0; q < c.hi - c.lo; ++q) dst[q] = P.coll_dst[c.dst_first + q];
}
/* image_elems[r] for r < world: elements of every rank's image (without the rank's private scratch) */
void sfcheck_image_elems(void *hv, long *out) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    for (int r = 0; r < P.world; ++r) out[r] = P.image_elems[r];
}
void sfcheck_owner(void *hv, int *owner) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    for (i64 s = 0; s < P.nsuper; ++s) owner[s] = P.owner[s];
}
/* X: this rank's image, xsize_pad + wsize doubles ("xsize_work"), zeroed by the caller */
void sfcheck_begin(void *hv, double *X) { assemble(*(CheckHandle *)hv, X); }
/* img[r] = base address of rank r's image as THIS process sees it (shared memory between the test's processes, the
 * stand-in for CUDA IPC mappings); remote update entries are subtracted there directly */
long sfcheck_run_segment(void *hv, int k, double *const *img) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    const Segment &sg = P.segments[k];
    return run_launches(P, img[P.rank], img, sg.launch_first, sg.launch_count);
}
/* root side of exchange step i (C_BCAST): copy the block from the root's image into every other image of the group */
long sfcheck_push(void *hv, int i, double *const *img) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    const Coll &c = P.colls[i];
    if (c.type != C_BCAST || c.root != P.rank || c.dst_first < 0) return -1;
    for (int q = c.lo; q < c.hi; ++q) {
        if (q == c.root) continue;
        const i64 dst = P.coll_dst[c.dst_first + (q - c.lo)];
        for (int col = 0; col < c.cols; ++col)
            std::memcpy(img[q] + dst + (i64)col * c.ld, img[c.root] + c.offset + (i64)col * c.ld, sizeof(double) * c.rows);
    }
    return 0;
}
/* copy only the panels this rank owns into Lsx_out (reference layout); others are left untouched */
void sfcheck_pack_owned(void *hv, const double *X, double *Lsx_out) {
    const HostPlan &P = ((CheckHandle *)hv)->P;
    const bool lu = P.kind == Kind::LU;
    for (i64 s = 0; s < P.nsuper; ++s) {
        if (P.owner[s] != P.rank) continue;
        const SnodeDev &d = P.snode[s];
        const int rows = lu ? 2 * d.nsrow - d.nscol : d.nsrow;
        for (int c = 0; c < d.nscol; ++c)
            std::memcpy(Lsx_out + P.lsxp[s] + (i64)c * rows, X + d.xoff + (i64)c * d.ld, sizeof(double) * rows);
    }
}

/* ---- interpreter of the triangular-solve schedule (same task lists the solve kernels consume) ---------------------- */
static void solve_task(const HostPlan &P, const SolveTask &t, int kind, const double *X, double *b) {
    const bool lu = P.kind == Kind::LU;
    auto gidx = [&](int i) { return i < t.nscol ? (i64)t.col0 + i : (i64)P.lsi32[t.lsip + i]; };
    const double *D = X + t.xoff + t.j0 + (i64)t.j0 * t.ld;
    double *xb = b + t.col0 + t.j0;
    switch (kind) {
        case S_DIAG_F:
            for (int j = 0; j < t.w; ++j) {
                if (!lu) xb[j] /= D[j + (i64)j * t.ld];
                for (int i = j + 1; i < t.w; ++i) xb[i] -= D[i + (i64)j * t.ld] * xb[j];
            }
            break;
        case S_DIAG_B:
            for (int j = t.w - 1; j >= 0; --j) {
                xb[j] /= D[j + (i64)j * t.ld];
                for (int i = 0; i < j; ++i) xb[i] -= (lu ? D[i + (i64)j * t.ld] : D[j + (i64)i * t.ld]) * xb[j];
            }
            break;
        case S_UPD_F:
            for (int i = t.r0; i < t.r1; ++i) {
                double acc = 0;
                for (int jj = 0; jj < t.w; ++jj) acc += X[t.xoff + i + (i64)(t.j0 + jj) * t.ld] * xb[jj];
                b[gidx(i)] -= acc;
            }
            break;
        case S_UPD_B: {
            const int shift = lu ? t.nsrow - t.nscol : 0;
            for (int jj = 0; jj < t.w; ++jj) {
                double acc = 0;
                for (int i = t.r0; i < t.r1; ++i) acc += X[t.xoff + shift + i + (i64)(t.j0 + jj) * t.ld] * b[gidx(i)];
                xb[jj] -= acc;
            }
            break;
        }
        case S_UPD_BU:
            for (int jj = 0; jj < t.w; ++jj) {
                double acc = 0;
                for (int c = t.r0; c < t.r1; ++c) acc += X[t.xoff + t.j0 + jj + (i64)c * t.ld] * b[t.col0 + c];
                xb[jj] -= acc;
            }
            break;
    }
}

/* factorize on the interpreter, then solve A x = b (permuted ordering) through the solve schedule; x overwrites b */
long sfcheck_factor_solve(void *hv, double *b) {
    CheckHandle &h = *(CheckHandle *)hv;
    const HostPlan &P = h.P;
    if (P.world != 1) return -1;
    std::vector<double> X(P.xsize_pad + P.wsize, 0.0);
    assemble(h, X.data());
    double *self = X.data();
    const long bad = run_launches(P, X.data(), &self, 0, (int)P.launches.size());
    for (const SolveLaunch &L : P.solve_fwd) for (int k = 0; k < L.count; ++k) solve_task(P, P.solve_tasks[L.first + k], L.kind, X.data(), b);
    for (const SolveLaunch &L : P.solve_bwd) for (int k = 0; k < L.count; ++k) solve_task(P, P.solve_tasks[L.first + k], L.kind, X.data(), b);
    return bad;
}

/* analyze + plan + interpret on one rank.  Lsx_out receives the factor in the reference's layout (xsize doubles).
 * Returns 0, or a negative code if analysis/plan failed, or a positive count of failed pivots. */
long sfcheck_factorize(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cx,
                       unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int ld_align,
                       double *Lsx_out, long *stats /* [nlevels, nlaunches, nsmall, nlarge, relmap, xsize_pad, xsize, -] */,
                       long out_len) {
    auto *h = (CheckHandle *)sfcheck_open(kind, is_symmetric, n, Cp, Ci, Cx, dev_slot, nb, outer, small_a, small_b, ld_align, 0, 1);
    if (!h) return -1;
    const HostPlan &P = h->P;
    if (stats) {
        stats[0] = P.nlevels; stats[1] = (long)P.launches.size(); stats[2] = P.nsmall; stats[3] = P.nlarge;
        stats[4] = (long)P.relmap.size(); stats[5] = P.xsize_pad; stats[6] = P.xsize;
    }
    if (!Lsx_out || out_len < P.xsize) { delete h; return 0; }     // sizing call
    std::vector<double> X(P.xsize_pad + P.wsize, 0.0);
    assemble(*h, X.data());
    int last_level = -1;
    for (const Launch &L : P.launches) {
        if (L.level < last_level) { delete h; return -100; }       // schedule must be level-monotone
        last_level = L.level;
    }
    double *self = X.data();
    const long bad = run_launches(P, X.data(), &self, 0, (int)P.launches.size());
    pack(P, X.data(), Lsx_out);
    delete h;
    return bad;
}

/* ---- complex matrices: analysis with the complex element size, real embedding, the ordinary schedule on the embedded
 * structure, complex panels read back (csrc/complex_embed.hpp).  Cz: interleaved (re, im) values; row_phase != 0 rotates every
 * row by the phase of its diagonal entry first (LU), as the engine does.  Lsz_out: xsize complex numbers (2 * xsize doubles);
 * stats = {nsuper, xsize (complex entries), embedded xsize_pad, nlaunches, embedded nnz, embedded unnz}.  With a null Lsz_out only
 * the sizes are returned.  The structure behind Lsz_out is the one sfb200_analyze_complex returns for the same input.
 * rhs (may be NULL): n complex numbers, solved in place with the embedded substitution.  Return value like sfcheck_factorize. */
long sfcheck_factorize_complex(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cz,
                               unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int row_phase,
                               double *Lsz_out, long *stats, double *rhs /* may be NULL: n complex, solved in place */) {
    ZCheckHandle h;
    CscView A;
    A.n = n; A.colptr = (const i64 *)Cp; A.rowind = (const i64 *)Ci; A.val = Cz; A.symmetric = is_symmetric != 0; A.is_complex = true;
    if (analyze(A, kind ? Kind::LU : Kind::Cholesky, dev_slot, h.Z) != 0) return -1;
    if (embed_complex(h.Z, h.E) != 0) return -2;
    if (row_phase && kind) {
        std::vector<i64> pos;
        diagonal_positions(h.Z, pos);
        h.phase.resize(2 * (size_t)n);
        const bool from_u = !h.Z.symmetric;
        row_phases(pos, (from_u ? h.Z.U.val : h.Z.L.val).data(), h.phase.data());
    }
    if (expand_embedding(h.Z, h.E, h.phase.empty() ? nullptr : h.phase.data()) != 0) return -3;
    h.base.S = h.E.R;
    PlanParams prm;
    if (nb > 0) prm.nb = nb;
    if (outer > 0) prm.outer = outer % 100000;
    if (outer >= 100000) prm.outer2 = outer / 100000;
    if (small_a >= 0) prm.small_a_bytes = small_a;
    if (small_b >= 0) prm.small_b_bytes = small_b;
    if (build_plan(h.base.S, prm, h.base.P, 0, 1) != 0) return -4;
    const HostPlan &P = h.base.P;
    if (stats) {
        stats[0] = h.Z.nsuper; stats[1] = h.Z.xsize; stats[2] = P.xsize_pad; stats[3] = (long)P.launches.size();
        stats[4] = P.nnz; stats[5] = P.unnz;
    }
    if (!Lsz_out) return 0;
    std::vector<double> X(P.xsize_pad + P.wsize, 0.0);
    assemble(h.base, X.data());
    double *self = X.data();
    const long bad = run_launches(P, X.data(), &self, 0, (int)P.launches.size());
    build_zextract_tasks(P, h.Z, h.tasks);
    const double *ph = h.phase.empty() ? nullptr : h.phase.data();
    for (const ZExtractTask &t : h.tasks) zextract_host(t, X.data(), Lsz_out, ph, h.Z.Lsi.data());
    if (rhs) {
        // the complex solve IS the real solve of the embedded system on the interleaved right-hand side; rotated rows:
        // (D^H A) x = D^H b
        if (ph)
            for (long k = 0; k < n; ++k) {
                const double re = rhs[2 * k], im = rhs[2 * k + 1], dr = ph[2 * k], di = ph[2 * k + 1];
                rhs[2 * k] = dr * re + di * im; rhs[2 * k + 1] = dr * im - di * re;
            }
        for (const SolveLaunch &L : P.solve_fwd) for (int k = 0; k < L.count; ++k) solve_task(P, P.solve_tasks[L.first + k], L.kind, X.data(), rhs);
        for (const SolveLaunch &L : P.solve_bwd) for (int k = 0; k < L.count; ++k) solve_task(P, P.solve_tasks[L.first + k], L.kind, X.data(), rhs);
    }
    return bad;
}

/* All `world` ranks simulated in one process: every rank has its own zeroed image (its compact layout), segments run in
 * lock step -- in rank order, or in reverse rank order with SFCHECK_RANKS_REVERSED=1: the result must not depend on which
 * rank of a segment runs first -- remote update entries land directly in the destination's image, a C_BCAST copies the
 * block from the root's image to the group's.  Lsx_out = union of the owned panels.  balance[r] receives the flops owned by
 * rank r; ncoll_out[2] counts the exchange steps met by type; image_out[r] (may be NULL) the elements of rank r's image. */
static long run_partition(std::vector<CheckHandle *> &H, double *Lsx_out, double *balance, long *ncoll_out, long *image_out);

long sfcheck_factorize_partitioned(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cx,
                                   unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int world,
                                   double *Lsx_out, double *balance, long *ncoll_out, long *image_out) {
    std::vector<CheckHandle *> H(world);
    for (int r = 0; r < world; ++r) {
        H[r] = (CheckHandle *)sfcheck_open(kind, is_symmetric, n, Cp, Ci, Cx, dev_slot, nb, outer, small_a, small_b, 0, r, world);
        if (!H[r]) return -1;
    }
    return run_partition(H, Lsx_out, balance, ncoll_out, image_out);
}

/* A complex matrix on `world` ranks, the way sfb200_zgroup_factorize composes it (csrc/zengine.cu): analysis with the complex
 * element size, real embedding, values expanded on the host (rows rotated first if row_phase), the partition of the EMBEDDED
 * structure simulated in lock step, the embedded factor packed in the reference's layout, complex panels read out of it. */
long sfcheck_factorize_complex_partitioned(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cz,
                                           unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int world,
                                           int row_phase, double *Lsz_out) {
    Symbolic Z;
    CscView A;
    A.n = n; A.colptr = (const i64 *)Cp; A.rowind = (const i64 *)Ci; A.val = Cz; A.symmetric = is_symmetric != 0; A.is_complex = true;
    if (analyze(A, kind ? Kind::LU : Kind::Cholesky, dev_slot, Z) != 0) return -1;
    ComplexEmbedding E;
    if (embed_complex(Z, E) != 0) return -2;
    std::vector<double> phase;
    if (row_phase && kind) {
        std::vector<i64> pos;
        diagonal_positions(Z, pos);
        phase.resize(2 * (size_t)n);
        row_phases(pos, (!Z.symmetric ? Z.U.val : Z.L.val).data(), phase.data());
    }
    if (expand_embedding(Z, E, phase.empty() ? nullptr : phase.data()) != 0) return -3;
    PlanParams prm;
    if (nb > 0) prm.nb = nb;
    if (outer > 0) prm.outer = outer % 100000;
    if (outer >= 100000) prm.outer2 = outer / 100000;
    if (small_a >= 0) prm.small_a_bytes = small_a;
    if (small_b >= 0) prm.small_b_bytes = small_b;
    if (const char *e = getenv("SFCHECK_DIST_MIN_M")) prm.dist_scatter_min_m = atoi(e);
    if (const char *e = getenv("SFCHECK_DIST_PANEL_MIN")) prm.dist_panel_min_cols = atoi(e);
    std::vector<CheckHandle *> H(world);
    for (int r = 0; r < world; ++r) {
        H[r] = new CheckHandle();
        H[r]->S = E.R;
        if (build_plan(H[r]->S, prm, H[r]->P, r, world) != 0) return -4;
    }
    std::vector<double> real_factor((size_t)E.R.xsize, 0.0);
    const long bad = run_partition(H, real_factor.data(), nullptr, nullptr, nullptr);      // frees the handles
    if (bad < 0) return bad;
    const bool lu = kind != 0;
    for (i64 s = 0; s < Z.nsuper; ++s) {
        ZExtractTask t{};
        t.nscol = (int32_t)(Z.Super[s + 1] - Z.Super[s]);
        t.nsrow = (int32_t)(Z.Lsip[s + 1] - Z.Lsip[s]);
        t.src = E.R.Lsxp[s];
        t.dst = Z.Lsxp[s];
        t.lsip = Z.Lsip[s];
        t.c0 = 0; t.ncols = t.nscol;
        t.ld = lu ? 4 * t.nsrow - 2 * t.nscol : 2 * t.nsrow;
        t.mode = lu ? 1 : 0;
        zextract_host(t, real_factor.data(), Lsz_out, phase.empty() ? nullptr : phase.data(), Z.Lsi.data());
    }
    return bad;
}

static long run_partition(std::vector<CheckHandle *> &H, double *Lsx_out, double *balance, long *ncoll_out, long *image_out) {
    const int world = (int)H.size();
    std::vector<std::vector<double>> X(world);
    std::vector<double *> img(world);
    for (int r = 0; r < world; ++r) {
        // every rank computes every rank's layout: they must agree
        for (int q = 0; q < world; ++q) if (H[r]->P.image_elems[q] != H[q]->P.xsize_pad) return -197;
        X[r].assign(H[r]->P.xsize_pad + H[r]->P.wsize, 0.0);     // image, then the rank's own scratch
        img[r] = X[r].data();
        if (image_out) image_out[r] = H[r]->P.xsize_pad;
    }
    for (int r = 0; r < world; ++r) { assemble(*H[r], X[r].data()); if (balance) balance[r] = H[r]->P.local_flops; }
    long bad = 0, ncoll[2] = {0, 0};
    const size_t nseg = H[0]->P.segments.size();
    for (int r = 1; r < world; ++r)
        if (H[r]->P.segments.size() != nseg || H[r]->P.colls.size() != H[0]->P.colls.size()) return -199;
    const bool reversed = getenv("SFCHECK_RANKS_REVERSED") != nullptr;
    for (size_t k = 0; k < nseg; ++k) {
        const Segment &sg = H[0]->P.segments[k];
        for (int i = sg.coll_first; i < sg.coll_first + sg.coll_count; ++i) {
            const Coll &c = H[0]->P.colls[i];
            // every rank must hold the identical step at the identical position (steps are matched by order)
            for (int q = 1; q < world; ++q) {
                const Coll &cq = H[q]->P.colls[i];
                const Segment &sq = H[q]->P.segments[k];
                if (sq.coll_first != sg.coll_first || sq.coll_count != sg.coll_count) return -198;
                if (cq.type != c.type || cq.root != c.root || cq.lo != c.lo || cq.hi != c.hi || cq.rows != c.rows || cq.cols != c.cols || cq.ld != c.ld) return -200;
            }
            if (c.type == C_BCAST) {
                const HostPlan &R = H[c.root]->P;
                const Coll &cr = R.colls[i];
                if (cr.dst_first < 0 || cr.offset < 0) return -202;
                for (int q = c.lo; q < c.hi; ++q) {
                    // the root's idea of where the block lives in q's image is q's own idea
                    if (R.coll_dst[cr.dst_first + (q - c.lo)] != H[q]->P.colls[i].offset) return -203;
                    if (q == c.root) continue;
                    for (int col = 0; col < c.cols; ++col)
                        std::memcpy(img[q] + H[q]->P.colls[i].offset + (i64)col * c.ld, img[c.root] + cr.offset + (i64)col * c.ld, sizeof(double) * c.rows);
                }
                for (int q = 0; q < world; ++q) if ((q < c.lo || q >= c.hi) && H[q]->P.colls[i].offset != -1) return -204;
            }
            ++ncoll[c.type];
        }
        for (int rr = 0; rr < world; ++rr) {
            const int r = reversed ? world - 1 - rr : rr;
            const Segment &sr = H[r]->P.segments[k];
            const long b = run_launches(H[r]->P, X[r].data(), img.data(), sr.launch_first, sr.launch_count);
            if (b < 0) return b;
            bad += b;
        }
    }
    for (int r = 0; r < world; ++r) sfcheck_pack_owned(H[r], X[r].data(), Lsx_out);
    if (ncoll_out) for (int t = 0; t < 2; ++t) ncoll_out[t] = ncoll[t];
    for (auto *h : H) delete h;
    return bad;
}

/* The exchange PROTOCOL under real concurrency: one host thread per rank, no lock step.  Every rank runs `epochs`
 * factorizations back to back exactly as the device does -- bump its epoch, zero and assemble its own image, then segment by
 * segment the exchange steps and launches -- and the steps use the plan's own flag sequence numbers (HostPlan::coll_sync):
 * a signal is a release store of epoch * 2^24 + seq into the peer's flag slot of its channel, a wait spins (acquire) until
 * the slot has reached its value, gives up after `timeout_s` and reports -300.  Remote update entries are atomic
 * subtractions in the peer's image, pushed blocks are copied by the root before it signals.  `jitter` > 0 makes every rank
 * sleep a rank- and step-dependent few microseconds before its exchange steps, so that repeated runs see different
 * interleavings.  What this catches that the lock-step simulation cannot: a missing or misplaced barrier (a remote add that
 * lands after its target was factorized, or before its target's owner zeroed the image of the next factorization), wrong
 * sequence numbers, a wait nobody signals (dead lock -> timeout).  Lsx_out = union of the owned panels after the last epoch. */
static long run_threaded(std::vector<CheckHandle *> &H, int epochs, int jitter, double timeout_s, double *Lsx_out);

long sfcheck_factorize_threaded(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cx,
                                unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int world, int epochs,
                                int jitter, double timeout_s, double *Lsx_out) {
    if (world < 1 || world > kMaxWorld) return -1;
    std::vector<CheckHandle *> H(world);
    for (int r = 0; r < world; ++r) {
        H[r] = (CheckHandle *)sfcheck_open(kind, is_symmetric, n, Cp, Ci, Cx, dev_slot, nb, outer, small_a, small_b, 0, r, world);
        if (!H[r]) return -1;
    }
    return run_threaded(H, epochs, jitter, timeout_s, Lsx_out);
}

long sfcheck_race_check(void *hv, int by_segment, int drop_flags);      // plan_race_check.cpp

/* Happens-before race check (plan_race_check.cpp) of the schedule(s) the engine replays for a COMPLEX matrix: the plan of the
 * embedded structure on one rank, or every rank's plan of its `world`-way partition.  Returns the number of hazards. */
long sfcheck_race_check_complex(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cz,
                                unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int world) {
    Symbolic Z;
    CscView A;
    A.n = n; A.colptr = (const i64 *)Cp; A.rowind = (const i64 *)Ci; A.val = Cz; A.symmetric = is_symmetric != 0; A.is_complex = true;
    if (analyze(A, kind ? Kind::LU : Kind::Cholesky, dev_slot, Z) != 0) return -1;
    ComplexEmbedding E;
    if (embed_complex(Z, E) != 0) return -2;
    PlanParams prm;
    if (nb > 0) prm.nb = nb;
    if (outer > 0) prm.outer = outer % 100000;
    if (outer >= 100000) prm.outer2 = outer / 100000;
    if (small_a >= 0) prm.small_a_bytes = small_a;
    if (small_b >= 0) prm.small_b_bytes = small_b;
    if (const char *e = getenv("SFCHECK_DIST_MIN_M")) prm.dist_scatter_min_m = atoi(e);
    if (const char *e = getenv("SFCHECK_DIST_PANEL_MIN")) prm.dist_panel_min_cols = atoi(e);
    long hazards = 0;
    for (int r = 0; r < world; ++r) {
        CheckHandle h;
        h.S = E.R;
        if (build_plan(h.S, prm, h.P, r, world) != 0) return -4;
        const long rc = sfcheck_race_check(&h, world > 1 ? 1 : 0, 0);
        if (rc < 0) return rc;
        hazards += rc;
    }
    return hazards;
}

/* The same protocol check on the partition sfb200_zgroup builds for a COMPLEX matrix: the plans of the embedded structure, one
 * free-running thread per rank, the plan's own flag sequence numbers; complex panels read out of the packed embedded factor. */
long sfcheck_factorize_complex_threaded(int kind, int is_symmetric, long n, const long *Cp, const long *Ci, const double *Cz,
                                        unsigned long long dev_slot, int nb, int outer, int small_a, int small_b, int world,
                                        int epochs, int jitter, double timeout_s, double *Lsz_out) {
    if (world < 1 || world > kMaxWorld) return -1;
    Symbolic Z;
    CscView A;
    A.n = n; A.colptr = (const i64 *)Cp; A.rowind = (const i64 *)Ci; A.val = Cz; A.symmetric = is_symmetric != 0; A.is_complex = true;
    if (analyze(A, kind ? Kind::LU : Kind::Cholesky, dev_slot, Z) != 0) return -1;
    ComplexEmbedding E;
    if (embed_complex(Z, E) != 0 || expand_embedding(Z, E, nullptr) != 0) return -2;
    PlanParams prm;
    if (nb > 0) prm.nb = nb;
    if (outer > 0) prm.outer = outer % 100000;
    if (outer >= 100000) prm.outer2 = outer / 100000;
    if (small_a >= 0) prm.small_a_bytes = small_a;
    if (small_b >= 0) prm.small_b_bytes = small_b;
    if (const char *e = getenv("SFCHECK_DIST_MIN_M")) prm.dist_scatter_min_m = atoi(e);
    if (const char *e = getenv("SFCHECK_DIST_PANEL_MIN")) prm.dist_panel_min_cols = atoi(e);
    std::vector<CheckHandle *> H(world);
    for (int r = 0; r < world; ++r) {
        H[r] = new CheckHandle();
        H[r]->S = E.R;
        if (build_plan(H[r]->S, prm, H[r]->P, r, world) != 0) return -4;
    }
    std::vector<double> real_factor((size_t)E.R.xsize, 0.0);
    const long rc = run_threaded(H, epochs, jitter, timeout_s, real_factor.data());        // frees the handles
    if (rc < 0) return rc;
    const bool lu = kind != 0;
    for (i64 s = 0; s < Z.nsuper; ++s) {
        ZExtractTask t{};
        t.nscol = (int32_t)(Z.Super[s + 1] - Z.Super[s]);
        t.nsrow = (int32_t)(Z.Lsip[s + 1] - Z.Lsip[s]);
        t.src = E.R.Lsxp[s];
        t.dst = Z.Lsxp[s];
        t.lsip = Z.Lsip[s];
        t.c0 = 0; t.ncols = t.nscol;
        t.ld = lu ? 4 * t.nsrow - 2 * t.nscol : 2 * t.nsrow;
        t.mode = lu ? 1 : 0;
        zextract_host(t, real_factor.data(), Lsz_out, nullptr, Z.Lsi.data());
    }
    return rc;
}

static long run_threaded(std::vector<CheckHandle *> &H, int epochs, int jitter, double timeout_s, double *Lsx_out) {
    const int world = (int)H.size();
    std::vector<std::vector<double>> X(world);
    std::vector<double *> img(world);
    for (int r = 0; r < world; ++r) { X[r].assign(H[r]->P.xsize_pad + H[r]->P.wsize, 0.0); img[r] = X[r].data(); }
    // flags[r][chan][src]
    std::vector<std::atomic<unsigned long long>> flags((size_t)world * 2 * kMaxWorld);
    for (auto &f : flags) f.store(0);
    std::atomic<long> failure{0}, bad{0};
    const bool no_wait = getenv("SFCHECK_THREADED_NO_WAIT") != nullptr;
    auto rank_main = [&](int r) {
        const HostPlan &P = H[r]->P;
        unsigned long long epoch = 0;
        auto slot = [&](int dst, int chan, int src) -> std::atomic<unsigned long long> & { return flags[((size_t)dst * 2 + chan) * kMaxWorld + src]; };
        for (int e = 0; e < epochs && failure.load() == 0; ++e) {
            ++epoch;
            std::fill(X[r].begin(), X[r].begin() + P.xsize_pad, 0.0);
            assemble(*H[r], X[r].data());
            for (size_t k = 0; k < P.segments.size() && failure.load() == 0; ++k) {
                const Segment &sg = P.segments[k];
                for (int i = sg.coll_first; i < sg.coll_first + sg.coll_count; ++i) {
                    const Coll &c = P.colls[i];
                    if (r < c.lo || r >= c.hi) continue;
                    const CollSync &cs = P.coll_sync[i];
                    if (jitter > 0) std::this_thread::sleep_for(std::chrono::microseconds((r * 7 + i * 13 + e * 3) % (jitter + 1)));
                    if (c.type == C_BCAST && r == c.root) {
                        for (int q = 0; q < cs.sig_count; ++q) {          // copy, then the destination's flag
                            const SyncPoint &sp = P.sync_points[cs.sig_first + q];
                            const i64 dst = P.coll_dst[c.dst_first + (sp.peer - c.lo)];
                            for (int col = 0; col < c.cols; ++col)
                                std::memcpy(img[sp.peer] + dst + (i64)col * c.ld, img[r] + c.offset + (i64)col * c.ld, sizeof(double) * c.rows);
                            slot(sp.peer, cs.chan, r).store(epoch * (1ull << 24) + sp.seq, std::memory_order_release);
                        }
                    } else {
                        for (int q = 0; q < cs.sig_count; ++q) {
                            const SyncPoint &sp = P.sync_points[cs.sig_first + q];
                            slot(sp.peer, cs.chan, r).store(epoch * (1ull << 24) + sp.seq, std::memory_order_release);
                        }
                    }
                    // self-test of this checker: SFCHECK_THREADED_NO_WAIT=1 drops every wait -- the factor must then come out wrong
                    for (int q = 0; q < (no_wait ? 0 : cs.wait_count); ++q) {
                        const SyncPoint &sp = P.sync_points[cs.wait_first + q];
                        const unsigned long long want = epoch * (1ull << 24) + sp.seq;
                        const auto t0 = std::chrono::steady_clock::now();
                        while (slot(r, cs.chan, sp.peer).load(std::memory_order_acquire) < want) {
                            if (failure.load() != 0) return;
                            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) { failure.store(-300); return; }
                            std::this_thread::yield();
                        }
                    }
                }
                const long b = run_launches(P, X[r].data(), img.data(), sg.launch_first, sg.launch_count);
                if (b < 0) { failure.store(b); return; }
                bad += b;
            }
        }
    };
    std::vector<std::thread> th;
    for (int r = 0; r < world; ++r) th.emplace_back(rank_main, r);
    for (auto &t : th) t.join();
    long rc = failure.load();
    if (rc == 0) {
        for (int r = 0; r < world; ++r) sfcheck_pack_owned(H[r], X[r].data(), Lsx_out);
        rc = bad.load();
    }
    for (auto *h : H) delete h;
    return rc;
}

}  // extern "C"
