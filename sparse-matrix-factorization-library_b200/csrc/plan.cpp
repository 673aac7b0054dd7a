// Host plan builder -- see plan.hpp.
#include "plan.hpp"

#include <algorithm>
#include <cassert>
#include <map>

namespace sfb200 {

namespace {

inline i64 round_up(i64 v, i64 a) { return (v + a - 1) / a * a; }

struct SlotOps {
    std::vector<DiagOp> diag;
    std::vector<TrsmOp> trsm;
    std::vector<GemmOp> gemm;
    std::vector<GemmOp> side;     // look-ahead: REST part of a top-level trailing update (side stream)
    bool next_of_split = false;   // gemm holds NEXT parts: they must wait for the previous block's REST
    bool diag_ahead = false;      // leaf look-ahead: the diagonal blocks apply the previous leaf's update themselves ...
    bool feeds_diag_ahead = false;   // ... which this slot's gemm ops therefore leave alone (G_SKIP_HEAD)
};

// Recursive blocked right-looking factorization of one large panel (replaces dpotrf_ + dtrsm_, CH.c:2135-2142, the
// blocked device POTRF of CH.c:2838-2940, and magma_dgetrf_nopiv + dtrsm_, LU.c:2653-2660).
//
// widths = {nb, outer, outer2...}: a block of level d is cut into sub-blocks of width widths[d-1]; after a sub-block
// [b0,b1) of the block [c0,c1) has been factorized, only the REST OF THAT BLOCK is updated (K = b1-b0), so the trailing
// updates of the big fronts run with K = the widest level.  Level-0 blocks (<= nb columns) are one k_diag + k_trsm.
// Every op gets a slot number that depends only on its position in this recursion, never on the front's size: op t of
// all large fronts of an etree level share one launch.
//
// LU panel rows: [0,R) = L11\U11 and L21, [R, R+m) = U12^T.  After sub-block b of block [c0,c1):
//   G1  rows [b1,R)   x cols [b1,c1)  -= L[.,b] U[b,.]        (B operand = U row block, k-major)
//   G2  U12^T rows    x cols [b1,c1)  -= U12^T[.,b] L[.,b]^T
//   G3  rows [b1,c1)  x cols [c1,n)   -= L[.,b] U[b,.]        (keeps the U entries right of the block current)
// Cholesky only has G1, lower-masked.
struct PanelEmitter {
    const SnodeDev &sn;
    const bool lu;
    const std::vector<int> &widths;
    std::map<int, SlotOps> &slots;
    std::vector<int> nslots;      // nslots[d] = slots used by one block of level d

    bool lookahead = false;
    bool leaf_lookahead = false;
    bool left_looking = false;
    i64 woff = -1;                // LU: offset of this front's U11^T scratch (ldw x nscol, element U11[k, j] at woff + j + k * ldw)
    int ldw = 0;

    PanelEmitter(const SnodeDev &s, bool is_lu, const std::vector<int> &w, std::map<int, SlotOps> &sl)
        : sn(s), lu(is_lu), widths(w), slots(sl) {
        nslots.assign(widths.size() + 1, 0);
        nslots[0] = 2;
        for (size_t d = 1; d < widths.size(); ++d) nslots[d] = (widths[d] / widths[d - 1]) * (nslots[d - 1] + 1);
    }

    // B = U11[k0 .. k0 + K, j0 .. j0 + N): k-major in the panel, row-major in the transposed scratch copy
    void gemm_u(int slot, i64 a, int k0, int j0, i64 c, int M, int N, int K, bool side = false, int skip_head = 0) {
        if (woff < 0) {
            gemm(slot, a, sn.xoff + k0 + (i64)j0 * sn.ld, c, M, N, K, G_B_KMAJOR, side, skip_head);
        } else {
            gemm(slot, a, woff + j0 + (i64)k0 * ldw, c, M, N, K, 0, side, skip_head, ldw);
        }
    }
    void gemm(int slot, i64 a, i64 b, i64 c, int M, int N, int K, int flags, bool side = false, int skip_head = 0, int ldb = 0) {
        if (M <= 0 || N <= 0 || K <= 0) return;
        GemmOp g{};
        g.a_off = a; g.b_off = b; g.c_off = c; g.lda = g.ldb = g.ldc = sn.ld;
        if (ldb > 0) g.ldb = ldb;
        g.M = M; g.N = N; g.K = K; g.flags = flags; g.tile_mod = 1; g.tile_rem = 0;
        if (skip_head > 0) { g.flags |= G_SKIP_HEAD; g.meta = skip_head; slots[slot].feeds_diag_ahead = true; }
        (side ? slots[slot].side : slots[slot].gemm).push_back(g);
    }
    void trsm(int slot, i64 d, int w, i64 b_off, int nvec, int es, int vs, int flags, i64 w_off = -1) {
        if (nvec <= 0) return;
        TrsmOp t{};
        t.t_off = d; t.b_off = b_off; t.ldt = sn.ld; t.nb = w; t.nvec = nvec; t.estride = es; t.vstride = vs; t.flags = flags;
        if (w_off >= 0) { t.flags |= T_COPY_T; t.w_off = w_off; t.ldw = ldw; }
        slots[slot].trsm.push_back(t);
    }

    // prev0 >= 0: the update of the previous leaf [prev0, j0) to this diagonal block is still pending (leaf look-ahead)
    void leaf(int j0, int j1, int slot, int prev0 = -1) {
        const i64 x = sn.xoff, ld = sn.ld;
        const int n = sn.nscol, R = sn.nsrow, m = R - n, w = j1 - j0;
        const i64 d = x + j0 + (i64)j0 * ld;
        DiagOp dop{};
        dop.a_off = d; dop.ld = sn.ld; dop.nb = w;
        if (prev0 >= 0) {
            dop.l_off = x + j0 + (i64)prev0 * ld;       // rows of this block in the previous leaf's columns
            dop.u_off = x + prev0 + (i64)j0 * ld;       // LU: the previous leaf's row block in this block's columns
            dop.upd_k = j0 - prev0;
            slots[slot].diag_ahead = true;
        }
        slots[slot].diag.push_back(dop);
        if (!lu) {
            trsm(slot + 1, d, w, x + j1 + (i64)j0 * ld, R - j1, sn.ld, 1, 0);
        } else {
            trsm(slot + 1, d, w, x + j1 + (i64)j0 * ld, R - j1, sn.ld, 1, T_TRANS);     // L below the block: X U11 = B
            trsm(slot + 1, d, w, x + R + (i64)j0 * ld, m, sn.ld, 1, T_UNIT);           // U12^T rows:        X L11^T = B
            // U11 row block, ALL later columns (vector v = column j1 + v); second copy into the U11^T scratch
            trsm(slot + 1, d, w, x + j0 + (i64)j1 * ld, n - j1, 1, sn.ld, T_UNIT, woff < 0 ? -1 : woff + j1 + (i64)j0 * ldw);
        }
    }

    void block(int c0, int c1, int depth, int slot) {
        if (depth == 0) { leaf(c0, c1, slot); return; }
        const i64 x = sn.xoff, ld = sn.ld;
        const int n = sn.nscol, R = sn.nsrow, m = R - n;
        const int w = widths[depth - 1], stride = nslots[depth - 1] + 1;
        // leaf look-ahead (sub-blocks are leaves): the update of leaf i leaves the next diagonal block alone; that block
        // subtracts the K = w product itself, so its factorization depends on leaf i's solves only, not on the update
        const bool leaf_la = leaf_lookahead && depth == 1;
        // Cholesky, every level below the top one: LEFT-looking inside the block.  Sub-block i first pulls the updates of
        // all earlier sub-blocks of this block (one GEMM, K = b0 - c0, on its own columns only) and is then factorized;
        // nothing is pushed to the right.  Same flops as the right-looking order, but every C entry of the block is
        // read-modified-written once instead of once per earlier sub-block, and K grows from w to (c1 - c0) - w: the rank-64
        // updates of the leaves were RMW-bound (4 flop/B).  The top level stays right-looking: its trailing updates are the
        // big K = outer2 GEMMs and carry the look-ahead split.
        const bool left = left_looking && !lu && depth < (int)widths.size();
        int i = 0;
        for (int b0 = c0; b0 < c1; b0 += w, ++i) {
            const int b1 = std::min(b0 + w, c1), K = b1 - b0;
            if (left && i > 0) {
                // issued in the slot the right-looking order would have used for the update after sub-block i - 1
                const int ps = slot + (i - 1) * stride + nslots[depth - 1];
                gemm(ps, x + b0 + (i64)c0 * ld, x + b0 + (i64)c0 * ld, x + b0 + (i64)b0 * ld, R - b0, b1 - b0, b0 - c0, G_LOWER, false,
                     leaf_la ? b1 - b0 : 0);
            }
            if (depth == 1) leaf(b0, b1, slot + i * stride, leaf_la && i > 0 ? (left ? c0 : b0 - w) : -1);
            else block(b0, b1, depth - 1, slot + i * stride);
            if (b1 >= c1 || left) continue;
            const int head = leaf_la ? std::min(b1 + w, c1) - b1 : 0;
            const int gs = slot + i * stride + nslots[depth - 1];
            // look-ahead split of the TOP-level update: [b1, s1) = the next block column (main stream), [s1, c1) = the rest
            const bool split = lookahead && depth == (int)widths.size() && b1 + w < c1;
            const int s1 = split ? b1 + w : c1;
            // every top-level update after the first one touches columns the side stream may still be updating (the REST of
            // the previous block column): its launches join the side stream first -- also the last, unsplit ones
            if (lookahead && depth == (int)widths.size() && i > 0) slots[gs].next_of_split = true;
            if (!lu) {
                // C rows [b1,R) x cols [b1,s1) is diagonal-aligned; the REST block (cols >= s1) only has rows >= s1 below
                // its own diagonal, so it is issued as its own diagonal-aligned op on rows [s1,R)
                gemm(gs, x + b1 + (i64)b0 * ld, x + b1 + (i64)b0 * ld, x + b1 + (i64)b1 * ld, R - b1, s1 - b1, K, G_LOWER, false, head);
                if (split)
                    gemm(gs, x + s1 + (i64)b0 * ld, x + s1 + (i64)b0 * ld, x + s1 + (i64)s1 * ld, R - s1, c1 - s1, K, G_LOWER, true);
            } else {
                gemm_u(gs, x + b1 + (i64)b0 * ld, b0, b1, x + b1 + (i64)b1 * ld, R - b1, s1 - b1, K, false, head);
                gemm(gs, x + R + (i64)b0 * ld, x + b1 + (i64)b0 * ld, x + R + (i64)b1 * ld, m, s1 - b1, K, 0);
                if (split) {
                    // LU: the next block column's factorization also works on its ROW block right of the diagonal (the U11 row
                    // solves and the row-block updates reach every later column), so rows [b1,s1) of the rest stay on the
                    // main stream; the side stream gets rows >= s1 only
                    gemm_u(gs, x + b1 + (i64)b0 * ld, b0, s1, x + b1 + (i64)s1 * ld, s1 - b1, c1 - s1, K);
                    gemm_u(gs, x + s1 + (i64)b0 * ld, b0, s1, x + s1 + (i64)s1 * ld, R - s1, c1 - s1, K, true);
                    gemm(gs, x + R + (i64)b0 * ld, x + s1 + (i64)b0 * ld, x + R + (i64)s1 * ld, m, c1 - s1, K, 0, true);
                }
                gemm_u(gs, x + b1 + (i64)b0 * ld, b0, c1, x + b1 + (i64)c1 * ld, c1 - b1, n - c1, K);
            }
        }
    }

    void run() { block(0, sn.nscol, (int)widths.size(), 0); }
};

}  // namespace

int build_plan(const Symbolic &S, const PlanParams &prm, HostPlan &P, int rank, int world) {
    P = HostPlan();
    if (world < 1 || rank < 0 || rank >= world) return 1;
    P.rank = rank;
    P.world = world;
    P.prm = prm;
    P.kind = S.kind;
    const bool lu = S.kind == Kind::LU;
    P.split = lu && !S.symmetric;
    const i64 n = S.n, ns = S.nsuper;
    P.n = n; P.nsuper = ns; P.nnz = S.L.ptr.empty() ? 0 : S.L.ptr[n];
    const Tri &Usrc = P.split ? S.U : S.L;     // LU on a symmetric matrix: U aliases L (LU.c:2718-2729)
    P.unnz = lu ? (Usrc.ptr.empty() ? 0 : Usrc.ptr[n]) : 0;
    P.isize = S.isize; P.xsize = S.xsize;
    P.lsxp = S.Lsxp;
    if (prm.nb < 8 || prm.nb > 64 || prm.outer % prm.nb != 0 || prm.ld_align < 1) return 1;
    if (prm.outer2 != 0 && (prm.outer2 % prm.outer != 0 || prm.outer2 <= prm.outer)) return 1;
    std::vector<int> widths = {prm.nb, prm.outer};
    if (prm.outer2 > 0) widths.push_back(prm.outer2);

    // ---- panels, levels -------------------------------------------------------------------------------------
    P.snode.resize(ns);
    std::vector<i64> sparent(ns, -1);
    for (i64 s = 0; s < ns; ++s) {
        SnodeDev &d = P.snode[s];
        d.nscol = (int32_t)(S.Super[s + 1] - S.Super[s]);
        d.nsrow = (int32_t)(S.Lsip[s + 1] - S.Lsip[s]);
        d.lsip = S.Lsip[s];
        const i64 rows = lu ? 2 * (i64)d.nsrow - d.nscol : d.nsrow;
        d.ld = (int32_t)round_up(rows, prm.ld_align);
        d.xoff = -1;               // assigned below, once it is known which panels this rank stores
        d.level = 0;
        if (d.nscol < d.nsrow) sparent[s] = S.SuperMap[S.Lsi[S.Lsip[s] + d.nscol]];
        const double c = d.nscol, m = d.nsrow - d.nscol;
        P.flops += lu ? 2.0 * c * c * c / 3.0 + 2.0 * m * c * c + 2.0 * m * m * c
                      : c * c * c / 3.0 + m * c * c + m * (m + 1.0) * c;
    }
    int maxlev = 0;
    for (i64 s = 0; s < ns; ++s) {     // parents have larger indices (postordered)
        const i64 p = sparent[s];
        if (p >= 0) {
            if (p <= s || p >= ns) return 5;      // parents follow their children (postordered structure)
            P.snode[p].level = std::max(P.snode[p].level, P.snode[s].level + 1);
        }
        maxlev = std::max(maxlev, (int)P.snode[s].level);
    }
    P.nlevels = ns > 0 ? maxlev + 1 : 0;

    // ---- ownership: etree subtrees onto `world` ranks ------------------------------------------------------
    // Start from the roots and keep opening the heaviest subtree (its root joins the "top" set, its children become
    // subtrees of their own) until the subtrees pack onto the ranks within 5 % of perfect balance (longest-processing-
    // time packing).  Whole subtrees are local to one rank.  Top supernodes receive partial updates from several ranks:
    // each goes to the rank with the least work so far, and the partial panel images are summed into that owner
    // (ReduceStep) right before its level is factorized.
    P.owner.assign(ns, 0);
    P.group_lo.assign(ns, 0);
    P.group_hi.assign(ns, 1);
    if (world > 1) {
        std::vector<double> work(ns), subtree(ns);
        for (i64 s = 0; s < ns; ++s) {
            const double c = P.snode[s].nscol, m = P.snode[s].nsrow - P.snode[s].nscol;
            work[s] = lu ? 2.0 * c * c * c / 3.0 + 2.0 * m * c * c + 2.0 * m * m * c : c * c * c / 3.0 + m * c * c + m * (m + 1.0) * c;
            subtree[s] = work[s];
        }
        std::vector<std::vector<i64>> kids(ns);
        std::vector<i64> frontier;
        for (i64 s = 0; s < ns; ++s) {
            if (sparent[s] >= 0) { subtree[sparent[s]] += subtree[s]; kids[sparent[s]].push_back(s); }   // children precede parents
            else frontier.push_back(s);
        }
        std::vector<char> is_top(ns, 0);
        std::vector<int32_t> where(ns, -1);
        std::vector<double> load(world, 0.0);
        auto pack = [&]() {
            std::sort(frontier.begin(), frontier.end(), [&](i64 a, i64 b) { return subtree[a] != subtree[b] ? subtree[a] > subtree[b] : a < b; });
            std::fill(load.begin(), load.end(), 0.0);
            double total = 0;
            for (i64 q : frontier) {
                const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                where[q] = r;
                load[r] += subtree[q];
                total += subtree[q];
            }
            return std::make_pair(*std::max_element(load.begin(), load.end()), total);
        };
        while (true) {
            const auto mt = pack();
            if (mt.first <= prm.balance_tol * mt.second / world || (i64)frontier.size() >= 64 * (i64)world) break;
            size_t pick = frontier.size();
            for (size_t k = 0; k < frontier.size(); ++k)        // frontier is sorted: first splittable = heaviest
                if (!kids[frontier[k]].empty()) { pick = k; break; }
            if (pick == frontier.size()) break;
            const i64 q = frontier[pick];
            is_top[q] = 1;
            frontier.erase(frontier.begin() + pick);
            frontier.insert(frontier.end(), kids[q].begin(), kids[q].end());
        }
        for (i64 s = ns - 1; s >= 0; --s) {                      // parents before children: local subtrees inherit
            if (is_top[s]) continue;
            if (where[s] < 0) where[s] = where[sparent[s]];
            P.owner[s] = where[s];
        }
        std::vector<i64> tops;
        for (i64 s = 0; s < ns; ++s) if (is_top[s]) tops.push_back(s);
        std::sort(tops.begin(), tops.end(), [&](i64 a, i64 b) { return work[a] != work[b] ? work[a] > work[b] : a < b; });
        for (i64 t : tops) {
            const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            P.owner[t] = r;
            load[r] += work[t];
        }
        // contributing rank range of every supernode = owners found in its subtree (bottom-up), plus its own owner
        for (i64 s = 0; s < ns; ++s) { P.group_lo[s] = P.owner[s]; P.group_hi[s] = P.owner[s] + 1; }
        for (i64 s = 0; s < ns; ++s)
            if (sparent[s] >= 0) {
                P.group_lo[sparent[s]] = std::min(P.group_lo[sparent[s]], P.group_lo[s]);
                P.group_hi[sparent[s]] = std::max(P.group_hi[sparent[s]], P.group_hi[s]);
            }
        for (i64 s = 0; s < ns; ++s) if (P.owner[s] == rank) P.local_flops += work[s];
    } else {
        P.local_flops = P.flops;
    }
    auto mine = [&](i64 s) { return P.owner[s] == rank; };

    // ---- which shared fronts the group works on together, and where every rank keeps what --------------------------------
    // dist_panel : a wide shared Cholesky front is factorized by its whole group: `outer`-wide block columns are owned
    //              in a weighted cyclic order (block_owner), every update aimed at a block column goes straight to that rank's image,
    //              each factored block column is pushed to the group
    // dist_update: the Schur update of a shared front with enough off-diagonal rows is split over the group (1/G of the
    //              tiles each); needs the factored panel in every image of the group
    // A rank STORES the panels it owns and the dist_panel / dist_update fronts of the groups it belongs to; its image holds
    // them back to back in supernode order.  All ranks compute all layouts (they are needed to aim remote updates).
    const int Wd = prm.outer;
    std::vector<char> dist_panel_of(ns, 0), dist_update_of(ns, 0);
    for (i64 s = 0; s < ns; ++s) {
        const SnodeDev &d = P.snode[s];
        const bool shared = P.group_hi[s] - P.group_lo[s] > 1;
        dist_panel_of[s] = shared && !lu && d.nscol >= prm.dist_panel_min_cols;
        dist_update_of[s] = shared && d.nsrow - d.nscol >= prm.dist_scatter_min_m;
    }
    // Block columns of a group-factorized front are dealt to the ranks of its group in a WEIGHTED cyclic order: the subtrees
    // of a nested dissection never weigh the same, and the work on these fronts (14 % of cfg5's flops) is what can be moved
    // freely -- with an even deal the heaviest rank of cfg5 executes 0.269 of the flops on 4 ranks (0.134 on 8) and every
    // other rank waits for it.  Loads are estimated from the structure (what a rank executes besides these fronts: its
    // subtrees, the fronts it owns, its even share of the split Schur updates); deepest fronts first, each front's
    // factorization work is poured over its group like water (lightest ranks first) and the shares become the weights of a
    // smooth weighted round-robin over its block columns.  Deterministic: every rank computes the same table.
    std::vector<std::vector<int32_t>> blk_owner(ns);
    if (world > 1) {
        std::vector<double> load(world, 0.0);
        auto sq = [](double v) { return v * v; };
        for (i64 s = 0; s < ns; ++s) {
            const double c = P.snode[s].nscol, m = P.snode[s].nsrow - P.snode[s].nscol;
            const double panel = lu ? 2.0 * c * c * c / 3.0 + 2.0 * m * c * c : c * c * c / 3.0 + m * c * c;
            const double upd = lu ? 2.0 * sq(m) * c : m * (m + 1.0) * c;
            const int glo = P.group_lo[s], ghi = P.group_hi[s];
            if (!dist_panel_of[s]) load[P.owner[s]] += panel;
            if (dist_update_of[s]) for (int r = glo; r < ghi; ++r) load[r] += upd / (ghi - glo);
            else load[P.owner[s]] += upd;
        }
        std::vector<i64> order;
        for (i64 s = 0; s < ns; ++s) if (dist_panel_of[s]) order.push_back(s);
        std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) { return P.snode[a].level < P.snode[b].level; });
        for (i64 s : order) {
            const double c = P.snode[s].nscol, m = P.snode[s].nsrow - P.snode[s].nscol;
            const double W = c * c * c / 3.0 + m * c * c;
            const int glo = P.group_lo[s], G = P.group_hi[s] - glo;
            // water level h with sum_r max(0, h - load_r) = W
            std::vector<double> l(load.begin() + glo, load.begin() + glo + G);
            std::vector<double> sorted = l;
            std::sort(sorted.begin(), sorted.end());
            double h = sorted[0] + W, acc = 0.0;
            for (int k = 0; k < G; ++k) {
                acc += sorted[k];
                const double cand = (W + acc) / (k + 1);
                if (k + 1 == G || cand <= sorted[k + 1]) { h = cand; break; }
            }
            std::vector<double> w(G);
            for (int r = 0; r < G; ++r) {
                w[r] = prm.weighted_blocks ? std::max(0.0, h - l[r]) / W : 1.0 / G;      // even deal = plain round-robin
                load[glo + r] += w[r] * W;
            }
            const int nblk = (P.snode[s].nscol + Wd - 1) / Wd;
            blk_owner[s].resize(nblk);
            std::vector<double> credit(G, 0.0);
            for (int b = 0; b < nblk; ++b) {
                int best = 0;
                for (int r = 0; r < G; ++r) { credit[r] += w[r]; if (credit[r] > credit[best] + 1e-12) best = r; }
                credit[best] -= 1.0;
                blk_owner[s][b] = glo + best;
            }
        }
    }
    auto block_owner = [&](i64 s, int b) { return (int)blk_owner[s][b]; };
    // destination of column c of supernode t: the rank whose image accumulates it (and assembles its matrix entries)
    auto dest_rank = [&](i64 t, i64 c) { return dist_panel_of[t] ? block_owner(t, (int)(c / Wd)) : P.owner[t]; };
    auto stores = [&](int r, i64 s) {
        return P.owner[s] == r || ((dist_panel_of[s] || dist_update_of[s]) && r >= P.group_lo[s] && r < P.group_hi[s]);
    };
    std::vector<std::vector<i64>> xoffs(world, std::vector<i64>(ns, -1));
    P.image_elems.assign(world, 0);
    for (int r = 0; r < world; ++r) {
        i64 xo = 0;
        for (i64 s = 0; s < ns; ++s) {
            if (!stores(r, s)) continue;
            xoffs[r][s] = xo;
            xo += (i64)P.snode[s].ld * P.snode[s].nscol;
        }
        P.image_elems[r] = xo;
        if (xo > kOffsetMask) return 2;
    }
    for (i64 s = 0; s < ns; ++s) P.snode[s].xoff = xoffs[rank][s];
    P.xsize_pad = P.image_elems[rank];

    // ---- ancestor segments and relative-index tables -----------------------------------------------------------
    struct Seg { i64 d; int32_t b, e; i64 rel_off; };
    std::vector<std::vector<Seg>> incoming(ns);
    P.upd_coloff.assign(S.isize, 0);
    P.upd_relbase.assign(S.isize, 0);
    P.upd_segb.assign(S.isize, 0);
    P.upd_sege.assign(S.isize, 0);
    P.upd_uoff.assign(lu ? S.isize : 0, 0);
    i64 rel_total = 0;
    for (i64 d = 0; d < ns; ++d) {
        const SnodeDev &sd = P.snode[d];
        const int m = sd.nsrow - sd.nscol;
        const i64 *rows = S.Lsi.data() + sd.lsip + sd.nscol;
        for (int b = 0; b < m;) {
            const i64 t = S.SuperMap[rows[b]];
            int e = b + 1;
            while (e < m && S.SuperMap[rows[e]] == t) ++e;
            incoming[t].push_back(Seg{d, b, e, rel_total});
            const SnodeDev &st = P.snode[t];
            for (int j = b; j < e; ++j) {
                const i64 q = sd.lsip + sd.nscol + j;
                const i64 tc = rows[j] - S.Super[t];
                const int dst = dest_rank(t, tc);
                P.upd_coloff[q] = ((i64)dst << kOwnerShift) | (xoffs[dst][t] + tc * (i64)st.ld);
                if (dst != rank) {        // entries of d's update this rank sends over NVLink (its share of d's update tiles)
                    const double cnt = lu ? (double)(m - b) + (double)(m - e) : (double)(m - j);
                    double share = 0.0;
                    if (dist_update_of[d]) { if (rank >= P.group_lo[d] && rank < P.group_hi[d]) share = 1.0 / (P.group_hi[d] - P.group_lo[d]); }
                    else if (P.owner[d] == rank) share = 1.0;
                    P.remote_update_entries += cnt * share;
                }
                P.upd_relbase[q] = rel_total - b;
                P.upd_segb[q] = b;
                P.upd_sege[q] = e;
                if (lu) P.upd_uoff[q] = st.nsrow - st.nscol;
            }
            rel_total += m - b;
            b = e;
        }
    }
    if (rel_total > (i64)0x7fffffff * 8) return 2;
    P.relmap.assign(rel_total, 0);

    // one sweep over the targets: position map of the target's rows, then (a) every segment aimed at it and
    // (b) the destinations of the matrix entries of its own columns (replaces Map, CH.c:2113-2114, and loadA's
    // addressing, CH.c:2010-2018 / LU.c:2490-2517)
    P.asm_dst_l.assign(P.nnz, 0);
    P.asm_dst_u.assign(P.unnz, 0);
    std::vector<int32_t> pos(n, -1);
    std::vector<i64> dup_stamp(n, 0);
    for (i64 t = 0; t < ns; ++t) {
        const SnodeDev &st = P.snode[t];
        const i64 *trow = S.Lsi.data() + st.lsip;
        for (int k = 0; k < st.nsrow; ++k) pos[trow[k]] = k;
        for (const Seg &g : incoming[t]) {
            const SnodeDev &sd = P.snode[g.d];
            const int m = sd.nsrow - sd.nscol;
            const i64 *rows = S.Lsi.data() + sd.lsip + sd.nscol;
            for (int i = g.b; i < m; ++i) {
                const int32_t r = pos[rows[i]];
                if (r < 0) return 3;              // containment property violated: not a valid supernodal structure
                P.relmap[g.rel_off + (i - g.b)] = r;
            }
        }
        for (i64 j = S.Super[t]; j < S.Super[t + 1]; ++j) {
            const i64 cj = j - S.Super[t];
            if (dest_rank(t, cj) != rank) {       // another rank assembles (and accumulates) this column
                for (i64 p = S.L.ptr[j]; p < S.L.ptr[j + 1]; ++p) P.asm_dst_l[p] = -1;
                if (lu) for (i64 p = Usrc.ptr[j]; p < Usrc.ptr[j + 1]; ++p) P.asm_dst_u[p] = -1;
                continue;
            }
            // Duplicate entries (the reader sums nothing, CH.c:526-587): loadA assigns in p order, so the LAST copy of an
            // (i, j) pair is the one that stays (CH.c:2018, LU.c:2498/2513).  The assembly kernel scatters all entries at
            // once, so the earlier copies are switched off here: walk each list backwards, first visit wins.
            for (i64 p = S.L.ptr[j + 1] - 1; p >= S.L.ptr[j]; --p) {
                const i64 i = S.L.idx[p];
                if (pos[i] < 0) return 4;
                // LU takes the diagonal from U (LU.c:2496): the L copy of a diagonal entry is skipped
                P.asm_dst_l[p] = ((lu && i == j) || dup_stamp[i] == 2 * j + 1) ? -1 : st.xoff + cj * st.ld + pos[i];
                dup_stamp[i] = 2 * j + 1;
            }
            if (lu) {
                for (i64 p = Usrc.ptr[j + 1] - 1; p >= Usrc.ptr[j]; --p) {
                    const i64 i = Usrc.idx[p];   // column index of an entry of row j of U
                    if (pos[i] < 0) return 4;
                    P.asm_dst_u[p] = dup_stamp[i] == 2 * j + 2 ? -1
                                     : pos[i] < st.nscol ? st.xoff + (i64)pos[i] * st.ld + cj
                                                         : st.xoff + cj * st.ld + (st.nsrow - st.nscol) + pos[i];
                    dup_stamp[i] = 2 * j + 2;
                }
            }
        }
        for (int k = 0; k < st.nsrow; ++k) pos[trow[k]] = -1;
    }

    // ---- the launch schedule, level by level --------------------------------------------------------------------
    std::vector<std::vector<i64>> by_level(P.nlevels);
    for (i64 s = 0; s < ns; ++s) by_level[P.snode[s].level].push_back(s);

    auto panel_bytes = [&](const SnodeDev &d) {
        const i64 rows = lu ? 2 * (i64)d.nsrow - d.nscol : d.nsrow;
        return rows * d.nscol * (i64)sizeof(double);
    };
    const int small_limits[4] = {std::min(2048, prm.small_a_bytes), prm.small_a_bytes, std::min(16384, prm.small_b_bytes),
                                 prm.small_b_bytes};

    // top-level stride of the slot numbering (see PanelEmitter): slots of one top-level block column are consecutive
    int top_stride = 2 + 1;
    for (size_t d = 1; d < widths.size(); ++d) top_stride = (widths[d] / widths[d - 1]) * top_stride + 1;
    const std::vector<int> dist_widths = {prm.nb, prm.outer};       // blocking inside group-factorized panels
    // a rows x cols block of supernode s starting at panel element (r0, c0): root's copy goes to the images of [lo,hi)
    auto make_bcast = [&](i64 s, int root, int r0, int c0, int rows, int cols) {
        const SnodeDev &d = P.snode[s];
        const int lo = P.group_lo[s], hi = P.group_hi[s];
        const i64 rel = r0 + (i64)c0 * d.ld;
        Coll c{};
        c.type = C_BCAST; c.root = root; c.lo = lo; c.hi = hi;
        c.offset = (rank >= lo && rank < hi) ? xoffs[rank][s] + rel : -1;
        c.rows = rows; c.cols = cols; c.ld = d.ld; c.dst_first = -1;
        if (rank == root) {
            c.dst_first = (int32_t)P.coll_dst.size();
            for (int q = lo; q < hi; ++q) P.coll_dst.push_back(xoffs[q][s] + rel);
        }
        return c;
    };
    auto make_barrier = [&](int lo, int hi) {
        Coll c{};
        c.type = C_BARRIER; c.root = lo; c.lo = lo; c.hi = hi; c.offset = -1; c.dst_first = -1;
        return c;
    };

    const int min_k128 = prm.gemm128_min_k >= 0 ? prm.gemm128_min_k : (world == 1 ? 257 : 0);
    for (int lev = 0; lev < P.nlevels; ++lev) {
        auto emit_gemm = [&](std::vector<GemmOp> &ops, int top_block, int lflags = 0) {
            // longest tiles first: the CTAs of a launch are dispatched in index order, so the deep-K ops must not be
            // the ones left running alone at the end of the launch
            std::stable_sort(ops.begin(), ops.end(), [](const GemmOp &a, const GemmOp &b) { return a.K > b.K; });
            for (int pass = 0; pass < 3; ++pass) {
                // pass 0: 128x128 tiles, warp-specialized persistent kernel (every op whose B operand is row-major);
                // pass 1: 128x128 tiles, classic kernel (k-major B: LU's U row blocks); pass 2: 64x64 tiles
                const int B = pass < 2 ? 128 : 64;
                Launch L{};
                L.kind = pass == 0 ? L_GEMM128WS : pass == 1 ? L_GEMM128 : L_GEMM64;
                L.level = lev;
                L.top_block = top_block;
                L.flags = lflags;
                L.first = (int32_t)P.gemm_ops.size();
                i64 tiles = 0;
                for (GemmOp g : ops) {
                    const bool big = g.M >= prm.gemm128_min_m && g.N > 64 && g.K >= min_k128;
                    // k-major B: one 128-byte bulk copy per column and k-step (kernels.cu: k_gemm_ws) needs aligned
                    // sources and whole k-steps
                    const bool kmajor_ok = prm.ws_kmajor && (g.b_off & 1) == 0 && (g.ldb & 1) == 0 && g.K % 16 == 0;
                    const bool ws_ok = prm.use_ws && (!(g.flags & G_B_KMAJOR) || kmajor_ok);
                    const int want = !big ? 2 : (ws_ok ? 0 : 1);
                    if (want != pass) continue;
                    g.tile0 = (int32_t)tiles;
                    tiles += share_of_tiles(gemm_tiles(g.M, g.N, B, (g.flags & G_LOWER) != 0), g.tile_mod, g.tile_rem);
                    P.gemm_ops.push_back(g);
                }
                L.count = (int32_t)P.gemm_ops.size() - L.first;
                L.grid = (int32_t)tiles;
                if (L.count > 0) P.launches.push_back(L);
            }
        };
        auto emit_slots = [&](std::map<int, SlotOps> &slots, int stride) {
            bool wait_diag = false;              // the next main-stream launch must join the diagonal-block stream
            auto emit_diag = [&](SlotOps &so, int top_block, int lflags) {
                Launch L{};
                L.kind = lu ? L_GETRF : L_POTRF; L.level = lev; L.top_block = top_block; L.flags = lflags;
                L.first = (int32_t)P.diag_ops.size();
                L.count = L.grid = (int32_t)so.diag.size();
                P.diag_ops.insert(P.diag_ops.end(), so.diag.begin(), so.diag.end());
                P.launches.push_back(L);
                so.diag.clear();
            };
            for (auto it = slots.begin(); it != slots.end(); ++it) {
                SlotOps &so = it->second;
                const int top_block = stride > 0 ? it->first / stride : -1;
                const size_t first_launch = P.launches.size();
                if (!so.diag.empty()) emit_diag(so, top_block, 0);
                if (!so.trsm.empty()) {
                    // Cholesky's solves (vectors = panel rows, contiguous in every column) go to the tensor-pipe kernel, 128 rows
                    // per CTA; LU's three flavours stay on the vector-pipe kernel, 64 vectors per CTA
                    for (int pass = 0; pass < 2; ++pass) {
                        const int per_cta = pass == 0 ? 128 : 64;
                        Launch L{};
                        L.kind = pass == 0 ? L_TRSM_MMA : L_TRSM; L.level = lev; L.top_block = top_block;
                        L.first = (int32_t)P.trsm_ops.size();
                        i64 ctas = 0;
                        for (TrsmOp t : so.trsm) {
                            const bool mma = prm.trsm_mma && !lu && t.flags == 0 && t.vstride == 1;
                            if (mma != (pass == 0)) continue;
                            t.tile0 = (int32_t)ctas;
                            ctas += (t.nvec + per_cta - 1) / per_cta;
                            P.trsm_ops.push_back(t);
                        }
                        L.count = (int32_t)P.trsm_ops.size() - L.first;
                        L.grid = (int32_t)ctas;
                        if (L.count > 0) P.launches.push_back(L);
                    }
                }
                // whatever this slot issued first on the main stream is the first consumer of a block factorized ahead
                if (wait_diag && P.launches.size() > first_launch) { P.launches[first_launch].flags |= LF_WAIT_DIAG; wait_diag = false; }
                // leaf look-ahead: the next slot's diagonal blocks only need this slot's predecessors (the solves of the
                // leaf), so they are issued BEFORE this slot's update and run next to it on their own stream
                auto nx = std::next(it);
                const bool ahead = so.feeds_diag_ahead && nx != slots.end() && nx->first == it->first + 1 && nx->second.diag_ahead &&
                                   !nx->second.diag.empty();
                if (ahead) emit_diag(nx->second, top_block, LF_DIAG_AHEAD);
                const size_t first_gemm = P.launches.size();
                if (!so.gemm.empty()) emit_gemm(so.gemm, top_block, (so.next_of_split ? LF_WAIT_SIDE : 0) | (ahead ? LF_LEAVE_SMS : 0));
                if (!so.side.empty()) emit_gemm(so.side, top_block, LF_SIDE);
                if (wait_diag && P.launches.size() > first_gemm) { P.launches[first_gemm].flags |= LF_WAIT_DIAG; wait_diag = false; }
                if (ahead) wait_diag = true;
            }
        };
        // a segment = the collectives gathered so far, then the launches emitted since the previous segment ended
        Phase ph{};
        ph.level = lev;
        ph.launch_first = (int32_t)P.launches.size();
        ph.seg_first = (int32_t)P.segments.size();
        int seg_launch0 = (int32_t)P.launches.size(), seg_coll0 = (int32_t)P.colls.size();
        auto end_segment = [&]() {
            Segment sg{};
            sg.level = lev;
            sg.coll_first = seg_coll0; sg.coll_count = (int32_t)P.colls.size() - seg_coll0;
            sg.launch_first = seg_launch0; sg.launch_count = (int32_t)P.launches.size() - seg_launch0;
            P.segments.push_back(sg);         // also when empty: segment k means the same step on every rank
            seg_launch0 = (int32_t)P.launches.size();
            seg_coll0 = (int32_t)P.colls.size();
        };

        std::vector<i64> cls[4];
        std::map<int, SlotOps> slots;
        std::vector<GemmOp> scatter;                // Schur updates computed entirely by the front's owner
        std::vector<GemmOp> scatter_dist;           // this rank's share of the Schur updates of shared fronts
        std::vector<Coll> panel_bcasts;             // factored shared panels, owner -> group (before the shared updates)
        std::vector<i64> dist_panels;               // shared fronts factorized by their whole group
        if (world > 1) {
            // first step of a factorization: nobody adds into a peer's image before that peer has zeroed and assembled it
            if (lev == 0) P.colls.push_back(make_barrier(0, world));
            // the shared fronts of this level are about to be factorized: the ranks of each group meet first -- everything
            // they have aimed at those fronts (remote adds of all lower levels) is then in place.  One barrier per distinct
            // group; a group inside another one of this level is covered by it.
            std::vector<std::pair<int, int>> groups;
            for (i64 s : by_level[lev]) {
                const std::pair<int, int> g{P.group_lo[s], P.group_hi[s]};
                if (g.second - g.first > 1 && std::find(groups.begin(), groups.end(), g) == groups.end()) groups.push_back(g);
            }
            for (size_t a = 0; a < groups.size(); ++a) {
                bool covered = lev == 0;             // the start barrier already spans everybody
                for (size_t b = 0; b < groups.size() && !covered; ++b)
                    covered = b != a && groups[b].first <= groups[a].first && groups[a].second <= groups[b].second &&
                              (groups[b] != groups[a]);
                if (!covered) P.colls.push_back(make_barrier(groups[a].first, groups[a].second));
            }
        }
        for (i64 s : by_level[lev]) {
            const SnodeDev &d = P.snode[s];
            const int m = d.nsrow - d.nscol;
            const int glo = P.group_lo[s], ghi = P.group_hi[s];
            const bool in_group = rank >= glo && rank < ghi;
            // (1) a wide shared Cholesky front is factorized by its whole group: block columns are owned round-robin and
            //     collect their updates directly, each one is pushed to the group once it is factorized
            const bool dist_panel = dist_panel_of[s] != 0;
            // (2) Schur update of a shared front: every rank of the group computes 1/G of the update tiles and adds them
            //     straight into the images of the ancestors' owners; needs the factored panel in every image of the group
            const bool dist_update = dist_update_of[s] != 0;
            auto scatter_ops = [&](std::vector<GemmOp> &out, int mod, int rem) {
                if (m <= 0) return;
                GemmOp g{};
                g.lda = g.ldb = d.ld; g.ldc = 0; g.M = g.N = m; g.K = d.nscol;
                g.meta = d.lsip + d.nscol;
                g.tile_mod = mod; g.tile_rem = rem;
                if (!lu) {
                    g.a_off = g.b_off = d.xoff + d.nscol;
                    g.flags = G_LOWER | G_SCATTER;
                    out.push_back(g);
                } else {
                    g.a_off = d.xoff + d.nscol; g.b_off = d.xoff + d.nsrow;      // L21 * (U12^T)^T, rows >= seg begin
                    g.flags = G_SCATTER | G_LIM_SEGB;
                    out.push_back(g);
                    g.a_off = d.xoff + d.nsrow; g.b_off = d.xoff + d.nscol;      // U12^T * L21^T, rows >= seg end
                    g.flags = G_SCATTER | G_LIM_SEGE;
                    out.push_back(g);
                }
            };
            if (dist_update) {
                // the group only reads the rows below the diagonal block (L21; LU: L21 and U12^T)
                if (!dist_panel) panel_bcasts.push_back(make_bcast(s, P.owner[s], d.nscol, 0, (lu ? 2 * d.nsrow - d.nscol : d.nsrow) - d.nscol, d.nscol));
                if (in_group) scatter_ops(scatter_dist, ghi - glo, rank - glo);
            }
            if (dist_panel) {
                dist_panels.push_back(s);
                if (!dist_update && mine(s)) scatter_ops(scatter, 1, 0);
                if (mine(s)) ++P.nlarge;
                continue;
            }
            if (!mine(s)) continue;
            const i64 pb = panel_bytes(d);
            int c = -1;
            if (d.nsrow - d.nscol <= prm.small_max_m)
                for (int k = 0; k < 4; ++k) if (pb <= small_limits[k]) { c = k; break; }
            if (c >= 0 && !dist_update) { cls[c].push_back(s); ++P.nsmall; continue; }
            ++P.nlarge;
            {
                PanelEmitter pe(d, lu, widths, slots);
                pe.lookahead = prm.lookahead != 0;
                pe.leaf_lookahead = prm.leaf_lookahead != 0;
                pe.left_looking = prm.left_looking != 0;
                if (lu && prm.lu_ut_copy && d.nscol > prm.nb) {      // a front with more than one leaf has in-panel updates
                    pe.ldw = (int)round_up(d.nscol, 2);
                    pe.woff = P.xsize_pad + P.wsize;
                    P.wsize += (i64)pe.ldw * d.nscol;
                }
                pe.run();
            }
            if (!dist_update) scatter_ops(scatter, 1, 0);
        }
        for (int c = 0; c < 4; ++c) {
            if (cls[c].empty()) continue;
            Launch L{};
            L.kind = c < 2 ? L_SMALL_A : L_SMALL_B;
            L.level = lev;
            L.top_block = -1;
            L.first = (int32_t)P.small_tasks.size();
            L.count = L.grid = (int32_t)cls[c].size();
            i64 smem = 0;
            for (i64 s : cls[c]) { P.small_tasks.push_back((int32_t)s); smem = std::max(smem, panel_bytes(P.snode[s])); }
            L.smem = (int32_t)smem;
            P.launches.push_back(L);
        }
        emit_slots(slots, top_stride);
        // the owner-computed Schur updates of fronts factorized by a group must wait for the group's work: they go last
        std::vector<GemmOp> scatter_late;
        if (!dist_panels.empty()) scatter_late.swap(scatter);
        if (!scatter.empty()) emit_gemm(scatter, -1, LF_WAIT_SIDE);

        // ---- group-factorized panels: block column b is owned by rank lo + b % G -----------------------------------------
        if (!dist_panels.empty()) {
            int nsteps = 0;
            for (i64 s : dist_panels) nsteps = std::max(nsteps, (P.snode[s].nscol + Wd - 1) / Wd);
            auto factor_block = [&](int b) {         // F(b): diag / trsm / inner updates of block column b, on its owner
                std::map<int, SlotOps> fs;
                for (i64 s : dist_panels) {
                    const SnodeDev &d = P.snode[s];
                    const int b0 = b * Wd;
                    if (b0 >= d.nscol || block_owner(s, b) != rank) continue;
                    PanelEmitter pe(d, false, dist_widths, fs);
                    pe.leaf_lookahead = prm.leaf_lookahead != 0;
                    pe.left_looking = prm.left_looking != 0;
                    pe.block(b0, std::min(b0 + Wd, (int)d.nscol), 1, 0);
                }
                emit_slots(fs, 0);
            };
            // Look-ahead across ranks: the owner of block column b+1 applies step b to that block column only, factorizes
            // it and lets the rest of its step-b updates wait until the broadcast of b+1 has been issued -- the other
            // ranks meanwhile apply step b to everything they own, so nobody waits for the owner's bulk update.
            std::vector<GemmOp> deferred;
            factor_block(0);
            for (int b = 0; b < nsteps; ++b) {
                end_segment();
                // the factorized block column travels to the group (whole columns: a contiguous run of the image) ...
                for (i64 s : dist_panels) {
                    const SnodeDev &d = P.snode[s];
                    const int b0 = b * Wd, b1 = std::min(b0 + Wd, (int)d.nscol);
                    // rows above the block column's diagonal block are structurally zero: they stay home
                    if (b0 < d.nscol) P.colls.push_back(make_bcast(s, block_owner(s, b), b0, b0, d.nsrow - b0, b1 - b0));
                }
                if (!deferred.empty()) { emit_gemm(deferred, -1, 0); deferred.clear(); }
                // ... every rank applies it to the later block columns it owns, then the owner of the next one factorizes it
                std::vector<GemmOp> next, rest;
                for (i64 s : dist_panels) {
                    const SnodeDev &d = P.snode[s];
                    if (rank < P.group_lo[s] || rank >= P.group_hi[s]) continue;
                    const int b0 = b * Wd, b1 = std::min(b0 + Wd, (int)d.nscol), nblk = (d.nscol + Wd - 1) / Wd;
                    if (b0 >= d.nscol) continue;
                    for (int c = b + 1; c < nblk; ++c) {
                        if (block_owner(s, c) != rank) continue;
                        const int c0 = c * Wd, c1 = std::min(c0 + Wd, (int)d.nscol);
                        GemmOp g{};
                        g.a_off = g.b_off = d.xoff + c0 + (i64)b0 * d.ld;
                        g.c_off = d.xoff + c0 + (i64)c0 * d.ld;
                        g.lda = g.ldb = g.ldc = d.ld;
                        g.M = d.nsrow - c0; g.N = c1 - c0; g.K = b1 - b0;
                        g.flags = G_LOWER; g.tile_mod = 1; g.tile_rem = 0;
                        (c == b + 1 ? next : rest).push_back(g);
                    }
                }
                if (!next.empty()) {
                    emit_gemm(next, -1, 0);
                    factor_block(b + 1);
                    deferred.swap(rest);
                } else if (!rest.empty()) {
                    emit_gemm(rest, -1, 0);
                }
            }
            if (!deferred.empty()) emit_gemm(deferred, -1, 0);
            if (!scatter_late.empty()) emit_gemm(scatter_late, -1, LF_WAIT_SIDE);
        }
        end_segment();
        // ---- shared Schur updates: factored panels to the group, then everybody's share of the tiles ------------------
        P.colls.insert(P.colls.end(), panel_bcasts.begin(), panel_bcasts.end());
        if (!scatter_dist.empty()) emit_gemm(scatter_dist, -1, LF_WAIT_SIDE);
        end_segment();
        ph.launch_count = (int32_t)P.launches.size() - ph.launch_first;
        ph.seg_count = (int32_t)P.segments.size() - ph.seg_first;
        P.phases.push_back(ph);
    }

    // ---- flags of the exchange steps (see plan_types.h: CollSync) ------------------------------------------------------------
    if (world > 1) {
        std::vector<uint32_t> sent(world, 0), recv(world, 0), psent(world, 0), precv(world, 0);
        P.coll_sync.assign(P.colls.size(), CollSync{0, 0, 0, 0, 0, 0});
        for (size_t i = 0; i < P.colls.size(); ++i) {
            const Coll &c = P.colls[i];
            CollSync &cs = P.coll_sync[i];
            cs.chan = c.type == C_BARRIER ? 0 : 1;
            if (rank < c.lo || rank >= c.hi) continue;
            cs.sig_first = (int32_t)P.sync_points.size();
            if (c.type == C_BARRIER) {
                for (int q = c.lo; q < c.hi; ++q) if (q != rank) P.sync_points.push_back(SyncPoint{q, ++sent[q]});
            } else if (rank == c.root) {
                for (int k = 1; k < c.hi - c.lo; ++k) {
                    const int q = c.lo + (c.root - c.lo + k) % (c.hi - c.lo);
                    P.sync_points.push_back(SyncPoint{q, ++psent[q]});
                }
            }
            cs.sig_count = (int32_t)P.sync_points.size() - cs.sig_first;
            cs.wait_first = (int32_t)P.sync_points.size();
            if (c.type == C_BARRIER) {
                for (int q = c.lo; q < c.hi; ++q) if (q != rank) P.sync_points.push_back(SyncPoint{q, ++recv[q]});
            } else if (rank != c.root) {
                P.sync_points.push_back(SyncPoint{c.root, ++precv[c.root]});
            }
            cs.wait_count = (int32_t)P.sync_points.size() - cs.wait_first;
        }
    }

    // ---- triangular-solve schedule (device-resident factor; single-GPU plans only) -----------------------------------
    if (world == 1) {
        P.lsi32.assign(S.Lsi.begin(), S.Lsi.end());
        P.nz_col.resize(P.nnz);
        for (i64 j = 0; j < n; ++j) for (i64 p = S.L.ptr[j]; p < S.L.ptr[j + 1]; ++p) P.nz_col[p] = (int32_t)j;
        P.nz_row.assign(S.L.idx.begin(), S.L.idx.begin() + P.nnz);
        if (P.split) {
            P.unz_col.assign(S.U.idx.begin(), S.U.idx.begin() + P.unnz);
            P.unz_row.resize(P.unnz);
            for (i64 i = 0; i < n; ++i) for (i64 p = S.U.ptr[i]; p < S.U.ptr[i + 1]; ++p) P.unz_row[p] = (int32_t)i;
        }
        const int SB = 64, SLAB = 256;       // one row per thread of the 256-thread update kernels
        auto base_task = [&](i64 s, int j0) {
            const SnodeDev &d = P.snode[s];
            SolveTask t{};
            t.xoff = d.xoff; t.lsip = d.lsip; t.ld = d.ld; t.nsrow = d.nsrow; t.nscol = d.nscol;
            t.col0 = (int32_t)S.Super[s]; t.j0 = j0; t.w = std::min(SB, d.nscol - j0);
            return t;
        };
        auto push_launch = [&](std::vector<SolveLaunch> &dst, int kind, const std::vector<SolveTask> &tasks) {
            if (tasks.empty()) return;
            dst.push_back(SolveLaunch{kind, (int32_t)P.solve_tasks.size(), (int32_t)tasks.size(), 0});
            P.solve_tasks.insert(P.solve_tasks.end(), tasks.begin(), tasks.end());
        };
        for (int pass = 0; pass < 2; ++pass) {           // 0 forward, 1 backward
            for (int li = 0; li < P.nlevels; ++li) {
                const int lev = pass == 0 ? li : P.nlevels - 1 - li;
                int maxblk = 0;
                for (i64 s : by_level[lev]) maxblk = std::max(maxblk, (P.snode[s].nscol + SB - 1) / SB);
                for (int bi = 0; bi < maxblk; ++bi) {
                    const int blk = pass == 0 ? bi : maxblk - 1 - bi;
                    std::vector<SolveTask> diag, upd, updu;
                    for (i64 s : by_level[lev]) {
                        const SnodeDev &d = P.snode[s];
                        const int j0 = blk * SB;
                        if (j0 >= d.nscol) continue;
                        SolveTask t = base_task(s, j0);
                        diag.push_back(t);
                        const int j1 = j0 + t.w;
                        if (pass == 0 || !lu) {
                            for (int r = j1; r < d.nsrow; r += SLAB) { t.r0 = r; t.r1 = std::min(r + SLAB, (int)d.nsrow); upd.push_back(t); }
                        } else {
                            // LU backward: U12 through the U12^T rows (off-diagonal rows), U11 through later panel columns
                            for (int r = d.nscol; r < d.nsrow; r += SLAB) { t.r0 = r; t.r1 = std::min(r + SLAB, (int)d.nsrow); upd.push_back(t); }
                            for (int c = j1; c < d.nscol; c += SLAB) { t.r0 = c; t.r1 = std::min(c + SLAB, (int)d.nscol); updu.push_back(t); }
                        }
                    }
                    if (pass == 0) {
                        push_launch(P.solve_fwd, S_DIAG_F, diag);
                        push_launch(P.solve_fwd, S_UPD_F, upd);
                    } else {
                        push_launch(P.solve_bwd, S_UPD_B, upd);
                        push_launch(P.solve_bwd, S_UPD_BU, updu);
                        push_launch(P.solve_bwd, S_DIAG_B, diag);
                    }
                }
            }
        }
    }

    // ---- copy-out tasks ------------------------------------------------------------------------------------------
    for (i64 s = 0; s < ns; ++s) {
        const SnodeDev &d = P.snode[s];
        if (!mine(s)) continue;              // every panel leaves through its owner
        const int rows = lu ? 2 * d.nsrow - d.nscol : d.nsrow;
        const int chunk = std::max(1, 65536 / std::max(rows, 1));
        for (int c0 = 0; c0 < d.nscol; c0 += chunk) {
            PackTask t{};
            t.src = d.xoff + (i64)c0 * d.ld;
            t.dst = S.Lsxp[s] + (i64)c0 * rows;
            t.rows = rows; t.ld = d.ld; t.cols = std::min(chunk, d.nscol - c0);
            P.pack_tasks.push_back(t);
        }
    }
    return 0;
}

}  // namespace sfb200
