/*
 * oracle/plan_race_check.cpp -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * Happens-before race detector for the engine's static schedule (csrc/plan.hpp).  The numeric phase replays a fixed
 * launch list on three CUDA streams -- main, "side" (LF_SIDE: the REST part of the top-level trailing updates) and "diag"
 * (LF_DIAG_AHEAD: diagonal blocks factorized next to the previous leaf's update) -- and inside one launch every op is a
 * set of CTAs that run concurrently.  The interpreter of plan_host_exec.cpp executes ONE legal order (or, with
 * SFCHECK_SIDE_LATE, a second one); this pass checks ALL of them at once: it walks the launch list in issue order,
 * enumerates the elements of the factor every op reads, writes and atomically adds to, and reports every pair of accesses
 * to the same element, at least one of them a plain write (or a write against an atomic add), that the schedule's
 * synchronisation does not order.  Atomic adds (the scatter epilogues) commute with each other.
 *
 * Ordering rules, as engine.cu: issue_launch_la / run_launches implement them:
 *   - launches of one stream run in issue order; ops of one launch are unordered among themselves;
 *   - a side / diag launch starts after everything issued before it on the main stream;
 *   - a main launch flagged LF_WAIT_SIDE (LF_WAIT_DIAG) waits for every side (diag) launch issued before it;
 *   - a range of launches (an etree level on one GPU, a segment on several) ends with all streams joined.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "plan.hpp"
#include "symbolic.hpp"

using namespace sfb200;

namespace {

enum Stream { MAIN = 0, SIDE = 1, DIAG = 2, COPY = 3 };

struct Checker {
    const HostPlan &P;
    // one access id per op; id -> launch, stream.  Copy-out pieces (COPY stream) carry the index of the launch they were
    // issued in front of, and whether they were issued behind a join of all streams.
    std::vector<int32_t> aid_launch;
    std::vector<int8_t> aid_stream, aid_after_join;
    // per launch: stream, range (join) index, index of the latest LF_WAIT_SIDE / LF_WAIT_DIAG main launch at or before it
    std::vector<int8_t> stream;
    std::vector<int32_t> range, wait_side, wait_diag;
    // shadow memory: last plain writer, last atomic adder, last two distinct readers per stream
    std::vector<int32_t> lastW, lastA, lastR[4], prevR[4];
    long hazards = 0;
    int printed = 0;

    explicit Checker(const HostPlan &p) : P(p) {
        const size_t n = (size_t)(P.xsize_pad + P.wsize);
        lastW.assign(n, -1); lastA.assign(n, -1);
        lean = getenv("SFCHECK_RACE_LEAN") != nullptr;      // full-size runs: one reader slot per stream (24 instead of 40 B / element)
        for (int s = 0; s < 4; ++s) { lastR[s].assign(n, -1); if (!lean) prevR[s].assign(n, -1); }
    }
    bool lean = false;

    bool launch_hb(int li, int lj) const {            // li issued before lj: does li complete before lj starts?
        if (range[li] != range[lj]) return true;      // ranges end with all streams joined
        const int si = stream[li], sj = stream[lj];
        if (si == sj) return true;
        if (si == MAIN) return true;                  // side / diag launches wait for the main stream's issue point
        // si is SIDE or DIAG: somebody on the main stream must have waited for it in between
        const std::vector<int32_t> &w = si == SIDE ? wait_side : wait_diag;
        if (sj == MAIN) return w[lj] > li;            // a waiting main launch k with li < k <= lj
        return lj > 0 && w[lj - 1] > li;              // other helper stream: waits for main's issue point, so k < lj
    }
    bool hb(int a, int b) const {
        if (a < 0 || a == b) return true;
        // the copy stream only rejoins the others when the whole factorization is over: nothing issued later is ordered
        // behind a copy; a copy is ordered behind what the main stream had completed or waited for at its issue point
        if (aid_stream[a] == COPY) return aid_stream[b] == COPY;
        if (aid_stream[b] == COPY) {
            if (aid_after_join[b] || aid_stream[a] == MAIN) return true;
            const int pos = aid_launch[b];            // issued in front of launch `pos`
            const std::vector<int32_t> &w = aid_stream[a] == SIDE ? wait_side : wait_diag;
            return pos > 0 && w[pos - 1] > aid_launch[a];
        }
        const int li = aid_launch[a], lj = aid_launch[b];
        if (li == lj) return false;                   // two ops of one launch
        return launch_hb(li, lj);
    }
    void report(const char *what, int a, int b, i64 e) {
        ++hazards;
        if (printed++ < 8)
            fprintf(stderr, "race: %s, element %lld: launch %d (kind %d, level %d, flags %d, stream %d) vs launch %d (kind %d, flags %d, stream %d)\n",
                    what, (long long)e, aid_launch[a], P.launches[aid_launch[a]].kind, P.launches[aid_launch[a]].level,
                    P.launches[aid_launch[a]].flags, aid_stream[a], aid_launch[b],
                    P.launches[std::min<int>(aid_launch[b], (int)P.launches.size() - 1)].kind,
                    P.launches[std::min<int>(aid_launch[b], (int)P.launches.size() - 1)].flags, aid_stream[b]);
    }
    void rd(int id, i64 e) {
        if (!hb(lastW[e], id)) report("read after unordered write", lastW[e], id, e);
        if (!hb(lastA[e], id)) report("read after unordered atomic add", lastA[e], id, e);
        const int s = aid_stream[id];
        if (lastR[s][e] != id) { if (!lean) prevR[s][e] = lastR[s][e]; lastR[s][e] = id; }
    }
    void check_readers(int id, i64 e, const char *what) {
        for (int s = 0; s < 4; ++s) {
            if (!hb(lastR[s][e], id)) report(what, lastR[s][e], id, e);
            if (!lean && !hb(prevR[s][e], id)) report(what, prevR[s][e], id, e);
        }
    }
    void wr(int id, i64 e) {
        if (!hb(lastW[e], id)) report("write after unordered write", lastW[e], id, e);
        if (!hb(lastA[e], id)) report("write after unordered atomic add", lastA[e], id, e);
        check_readers(id, e, "write after unordered read");
        lastW[e] = id;
    }
    void add(int id, i64 e) {
        if (!hb(lastW[e], id)) report("atomic add after unordered write", lastW[e], id, e);
        check_readers(id, e, "atomic add after unordered read");
        lastA[e] = id;
    }

    // ---- footprints ---------------------------------------------------------------------------------------------------
    void scatter_targets(int id, const GemmOp &g, int B) {
        const bool lower = (g.flags & G_LOWER) != 0;
        const i64 ntile = gemm_tiles(g.M, g.N, B, lower);
        for (i64 t = g.tile_rem; t < ntile; t += g.tile_mod) {
            int ti, tj;
            decode_tile(t, g.M, g.N, B, lower, ti, tj);
            for (int j = tj * B; j < std::min(g.N, (tj + 1) * B); ++j) {
                const i64 q = g.meta + j;
                // adds into another rank's image are ordered by the exchange steps (the group barrier in front of the target's
                // level), not by this rank's streams: this checker covers one rank's image
                if ((int)(P.upd_coloff[q] >> kOwnerShift) != P.rank) continue;
                i64 col = P.upd_coloff[q] & kOffsetMask;
                const i64 relbase = P.upd_relbase[q];
                int lim = (g.flags & G_LIM_SEGB) ? P.upd_segb[q] : (g.flags & G_LIM_SEGE) ? P.upd_sege[q] : j;
                if (g.flags & G_LIM_SEGE) col += P.upd_uoff[q];
                for (int i = std::max(lim, ti * B); i < std::min(g.M, (ti + 1) * B); ++i) add(id, col + P.relmap[relbase + i]);
            }
        }
    }
    void gemm(int id, const GemmOp &g, int B) {
        for (int k = 0; k < g.K; ++k) {
            for (int i = 0; i < g.M; ++i) rd(id, g.a_off + i + (i64)k * g.lda);
            for (int j = 0; j < g.N; ++j) rd(id, (g.flags & G_B_KMAJOR) ? g.b_off + k + (i64)j * g.ldb : g.b_off + j + (i64)k * g.ldb);
        }
        if (g.flags & G_SCATTER) { scatter_targets(id, g, B); return; }
        const bool lower = (g.flags & G_LOWER) != 0;
        for (int j = 0; j < g.N; ++j) {
            int lim = lower ? j : 0;
            if ((g.flags & G_SKIP_HEAD) && j < (int)g.meta) lim = std::max(lim, (int)g.meta);
            for (int i = lim; i < g.M; ++i) wr(id, g.c_off + i + (i64)j * g.ldc);
        }
    }
    void diag(int id, const DiagOp &d, bool lu) {
        if (d.upd_k > 0) {
            for (int k = 0; k < d.upd_k; ++k)
                for (int r = 0; r < d.nb; ++r) {
                    rd(id, d.l_off + r + (i64)k * d.ld);
                    if (lu) rd(id, d.u_off + k + (i64)r * d.ld);
                }
        }
        for (int c = 0; c < d.nb; ++c)                 // k_diag rewrites the whole nb x nb block (zeros above the diagonal)
            for (int r = 0; r < d.nb; ++r) wr(id, d.a_off + r + (i64)c * d.ld);
    }
    void trsm(int id, const TrsmOp &t) {
        for (int c = 0; c < t.nb; ++c)
            for (int r = 0; r < t.nb; ++r) rd(id, t.t_off + r + (i64)c * t.ldt);
        for (int v = 0; v < t.nvec; ++v)
            for (int j = 0; j < t.nb; ++j) {
                wr(id, t.b_off + (i64)v * t.vstride + (i64)j * t.estride);
                if (t.flags & T_COPY_T) wr(id, t.w_off + v + (i64)j * t.ldw);
            }
    }
    void small(int id, i64 s) {
        const SnodeDev &d = P.snode[s];
        const int n = d.nscol, R = d.nsrow, m = R - n;
        const bool lu = P.kind == Kind::LU;
        const int rows = lu ? R + m : R;
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < rows; ++r) wr(id, d.xoff + r + (i64)c * d.ld);
        if (m <= 0) return;
        GemmOp g{};
        g.M = g.N = m; g.K = n; g.meta = d.lsip + n; g.tile_mod = 1; g.tile_rem = 0;
        if (!lu) { g.flags = G_LOWER | G_SCATTER; scatter_targets(id, g, 128); }
        else {
            g.flags = G_SCATTER | G_LIM_SEGB; scatter_targets(id, g, 128);
            g.flags = G_SCATTER | G_LIM_SEGE; scatter_targets(id, g, 128);
        }
    }

    int new_id(int launch, int strm, bool after_join = false) {
        aid_launch.push_back(launch); aid_stream.push_back((int8_t)strm); aid_after_join.push_back(after_join);
        return (int)aid_launch.size() - 1;
    }
    // streamed copy-out of columns [c0, c1) of front s (engine.cu: copy_columns / k_pack): reads on the copy stream
    void copy_out(int id, i64 s, int c0, int c1) {
        const SnodeDev &d = P.snode[s];
        const int rows = P.kind == Kind::LU ? 2 * d.nsrow - d.nscol : d.nsrow;
        for (int c = c0; c < c1; ++c)
            for (int r = 0; r < rows; ++r) rd(id, d.xoff + r + (i64)c * d.ld);
    }

    // model_copy: the host-buffer call streams the factor out behind the factorization (engine.cu:
    // factorize_streamed_issue) -- whenever the schedule of a level moves on to top-level block column t, the columns left
    // of t * top_width of the level's fronts leave on the copy stream behind an event recorded on the MAIN stream only; the
    // rest of the level follows its join.  (Superset of the engine's behaviour: it sends only fronts of >= 1 MiB early.)
    long run(const std::vector<std::pair<int, int>> &ranges, bool model_copy) {
        const int nl = (int)P.launches.size();
        const int top_width = P.prm.outer2 > 0 ? P.prm.outer2 : P.prm.outer;
        // self-test hook: pretend the engine sent `ahead` block columns too many (must be reported)
        const int ahead = getenv("SFCHECK_RACE_COPY_AHEAD") ? atoi(getenv("SFCHECK_RACE_COPY_AHEAD")) : 0;
        std::vector<std::vector<i64>> by_level(P.nlevels);
        for (i64 s = 0; s < P.nsuper; ++s) by_level[P.snode[s].level].push_back(s);
        std::vector<int> sent(P.nsuper, 0);
        stream.assign(nl, MAIN); range.assign(nl, 0); wait_side.assign(nl, -1); wait_diag.assign(nl, -1);
        for (size_t r = 0; r < ranges.size(); ++r)
            for (int li = ranges[r].first; li < ranges[r].first + ranges[r].second; ++li) range[li] = (int)r;
        int ws = -1, wd = -1;
        for (int li = 0; li < nl; ++li) {
            const Launch &L = P.launches[li];
            stream[li] = (L.flags & LF_SIDE) ? SIDE : (L.flags & LF_DIAG_AHEAD) ? DIAG : MAIN;
            if (stream[li] == MAIN && (L.flags & LF_WAIT_SIDE)) ws = li;
            if (stream[li] == MAIN && (L.flags & LF_WAIT_DIAG)) wd = li;
            wait_side[li] = ws; wait_diag[li] = wd;
        }
        int cur_block = -1, cur_range = -1;
        auto finish_level = [&](int lev, int pos) {       // behind the join that ends the level
            const int id = new_id(pos, COPY, true);
            for (i64 s : by_level[lev]) { copy_out(id, s, sent[s], P.snode[s].nscol); sent[s] = P.snode[s].nscol; }
        };
        for (int li = 0; li < nl; ++li) {
            const Launch &L = P.launches[li];
            if (model_copy) {
                if (range[li] != cur_range) {
                    if (cur_range >= 0) finish_level(P.launches[li - 1].level, li);
                    cur_range = range[li]; cur_block = -1;
                }
                if (L.top_block > cur_block && cur_block >= 0) {
                    const int id = new_id(li, COPY);
                    for (i64 s : by_level[L.level]) {
                        const int c1 = std::min<int>(P.snode[s].nscol, (L.top_block + ahead) * top_width);
                        if (c1 > sent[s]) { copy_out(id, s, sent[s], c1); sent[s] = c1; }
                    }
                }
                if (L.top_block >= 0) cur_block = std::max(cur_block, (int)L.top_block);
            }
            for (int k = 0; k < L.count; ++k) {
                const int id = new_id(li, stream[li]);
                switch (L.kind) {
                    case L_SMALL_A: case L_SMALL_B: small(id, P.small_tasks[L.first + k]); break;
                    case L_POTRF: diag(id, P.diag_ops[L.first + k], false); break;
                    case L_GETRF: diag(id, P.diag_ops[L.first + k], true); break;
                    case L_TRSM: case L_TRSM_MMA: trsm(id, P.trsm_ops[L.first + k]); break;
                    case L_GEMM128: case L_GEMM128WS: gemm(id, P.gemm_ops[L.first + k], 128); break;
                    case L_GEMM64: gemm(id, P.gemm_ops[L.first + k], 64); break;
                    default: return -103;
                }
            }
        }
        if (model_copy && nl > 0) finish_level(P.launches[nl - 1].level, nl);
        return hazards;
    }
};

}  // namespace

struct CheckHandle {         // same layout as in plan_host_exec.cpp
    Symbolic S;
    HostPlan P;
};

extern "C" {

/* Number of unordered conflicting access pairs in the schedule of this handle (0 = race-free); the first few are printed
 * to stderr.  by_segment != 0: the ranges that end with a join are the plan's segments (multi-GPU replay), otherwise its
 * etree levels (single-GPU replay; this also covers the copy stream of the host-buffer call, see Checker::run).  drop_flags: LaunchFlags to ignore -- a test hook that must make the count positive
 * (e.g. LF_WAIT_SIDE: without the joins the look-ahead split is a race). */
long sfcheck_race_check(void *hv, int by_segment, int drop_flags) {
    CheckHandle *h = (CheckHandle *)hv;
    HostPlan P = h->P;                         // private copy: the hook edits flags
    for (Launch &L : P.launches) L.flags &= ~drop_flags;
    std::vector<std::pair<int, int>> ranges;
    if (by_segment) for (const Segment &s : P.segments) ranges.emplace_back(s.launch_first, s.launch_count);
    else for (const Phase &ph : P.phases) ranges.emplace_back(ph.launch_first, ph.launch_count);
    // the streamed copy-out exists on the single-GPU host-buffer path only; SFCHECK_RACE_NO_COPY leaves it out
    Checker c(P);
    return c.run(ranges, !by_segment && !getenv("SFCHECK_RACE_NO_COPY"));
}

/* The solve sweeps run on one stream, launch after launch: the only concurrency is between the tasks of one launch
 * (k_solve_diag: plain read-modify-write of its 64 unknowns; the update kernels: reads of the right-hand side, atomic adds
 * into it).  Returns the number of right-hand-side entries on which two DIFFERENT tasks of one launch conflict (plain write
 * against anything, atomic add against a read). */
long sfcheck_solve_race_check(void *hv) {
    CheckHandle *h = (CheckHandle *)hv;
    const HostPlan &P = h->P;
    if (P.world != 1) return -1;
    const size_t n = (size_t)P.n;
    // per entry, for the launch being scanned: task that wrote / added / read it (-1 none, -2 several tasks)
    std::vector<int32_t> w(n), a(n), r(n);
    long hazards = 0;
    auto scan = [&](const std::vector<SolveLaunch> &sweep) {
        for (const SolveLaunch &L : sweep) {
            std::fill(w.begin(), w.end(), -1); std::fill(a.begin(), a.end(), -1); std::fill(r.begin(), r.end(), -1);
            auto mark = [](std::vector<int32_t> &v, i64 e, int task) { if (v[e] == -1) v[e] = task; else if (v[e] != task) v[e] = -2; };
            for (int k = 0; k < L.count; ++k) {
                const SolveTask &t = P.solve_tasks[L.first + k];
                auto gidx = [&](int i) { return i < t.nscol ? (i64)t.col0 + i : (i64)P.lsi32[t.lsip + i]; };
                const i64 blk = (i64)t.col0 + t.j0;
                switch (L.kind) {
                    case S_DIAG_F: case S_DIAG_B: for (int j = 0; j < t.w; ++j) mark(w, blk + j, k); break;
                    case S_UPD_F:
                        for (int j = 0; j < t.w; ++j) mark(r, blk + j, k);
                        for (int i = t.r0; i < t.r1; ++i) mark(a, gidx(i), k);
                        break;
                    case S_UPD_B:
                        for (int i = t.r0; i < t.r1; ++i) mark(r, gidx(i), k);
                        for (int j = 0; j < t.w; ++j) mark(a, blk + j, k);
                        break;
                    case S_UPD_BU:
                        for (int c = t.r0; c < t.r1; ++c) mark(r, (i64)t.col0 + c, k);
                        for (int j = 0; j < t.w; ++j) mark(a, blk + j, k);
                        break;
                    default: break;
                }
            }
            for (size_t e = 0; e < n; ++e) {
                const bool wrote = w[e] != -1, added = a[e] != -1, read = r[e] != -1;
                bool bad = false;
                if (wrote && (w[e] == -2 || (added && a[e] != w[e]) || (read && r[e] != w[e]))) bad = true;   // plain write vs another task
                if (added && read && (a[e] == -2 || r[e] == -2 || a[e] != r[e])) bad = true;                    // add vs another task's read
                if (bad && hazards++ < 8)
                    fprintf(stderr, "solve race: launch kind %d, right-hand-side entry %zu (writer %d, adder %d, reader %d)\n", L.kind, e, w[e], a[e], r[e]);
            }
        }
    };
    scan(P.solve_fwd);
    scan(P.solve_bwd);
    return hazards;
}

}  // extern "C"
