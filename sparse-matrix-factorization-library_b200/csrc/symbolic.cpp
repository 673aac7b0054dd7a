// Host symbolic analysis -- see symbolic.hpp.  Each stage cites the reference lines whose OUTPUT it must match
// bit for bit; the code itself is an independent restatement built on std::vector.
#include "symbolic.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <thread>

extern "C" int METIS_NodeND(int64_t *nvtxs, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *options,
                            int64_t *perm, int64_t *iperm);

namespace sfb200 {

// parallel_order.cpp
bool parallel_nested_dissection(i64 n, const std::vector<i64> &xadj, const std::vector<i64> &adj, int procs, std::vector<i64> &perm);

// processes used by the ordering: 1 = one METIS_NodeND call, the reference's ordering (default); > 1 = the opt-in parallel
// nested dissection of parallel_order.cpp (a different, equally valid permutation).  -1: not set, SFB200_PARALLEL_ORDER decides.
static std::atomic<int> g_ordering_processes{-1};
void set_ordering_processes(int procs) { g_ordering_processes.store(procs < 1 ? 1 : procs); }
int ordering_processes() {
    int p = g_ordering_processes.load();
    if (p < 0) {
        const char *e = getenv("SFB200_PARALLEL_ORDER");
        p = e ? std::max(1, atoi(e)) : 1;
    }
    return p;
}

// ---------------------------------------------------------------------------------------------------------
// Ordering.  Reference: SparseFrame_metis, CH.c:864-954 (LU.c:965-1066 feeds it lower(C+C^T) instead).
// The adjacency lists must be filled in the reference's order -- METIS is deterministic for a given input,
// but a different neighbour order gives a different (equally valid) permutation.
// ---------------------------------------------------------------------------------------------------------
void fill_reducing_order(i64 n, const i64 *colptr, const i64 *rowind, std::vector<i64> &perm) {
    perm.assign(n, 0);
    std::vector<i64> xadj(n + 1, 0);
    for (i64 j = 0; j < n; ++j)
        for (i64 p = colptr[j]; p < colptr[j + 1]; ++p) {
            const i64 i = rowind[p];
            if (j < i) { ++xadj[j + 1]; ++xadj[i + 1]; }
        }
    std::partial_sum(xadj.begin(), xadj.end(), xadj.begin());
    const i64 nadj = xadj[n];
    if (nadj == 0) {                       // diagonal matrix: identity (CH.c:933-939)
        std::iota(perm.begin(), perm.end(), i64(0));
        return;
    }
    std::vector<i64> cursor(xadj.begin(), xadj.end() - 1), adj(nadj);
    for (i64 j = 0; j < n; ++j)
        for (i64 p = colptr[j]; p < colptr[j + 1]; ++p) {
            const i64 i = rowind[p];
            if (j < i) { adj[cursor[j]++] = i; adj[cursor[i]++] = j; }
        }
    // Duplicate (i, j) entries -- the reader sums nothing -- would give METIS adjacency lists with repeated neighbours, which
    // it does not allow: the reference does exactly that (CH.c:920-931), and a 17-vertex band matrix with duplicates corrupts
    // its heap inside METIS_NodeND (found by tests/tools/differential_fuzz.py; the reference segfaults on it).  Repeated
    // neighbours are therefore removed first, keeping the first occurrence.  A matrix without duplicates -- every list is
    // already a set -- goes to METIS exactly as the reference sends it, so its ordering stays bit-identical.
    if (!getenv("SFB200_METIS_KEEP_DUPLICATES")) {
        std::vector<i64> stamp(n, -1);
        bool any = false;
        for (i64 v = 0; v < n && !any; ++v)
            for (i64 p = xadj[v]; p < xadj[v + 1]; ++p) {
                if (stamp[adj[p]] == v) { any = true; break; }
                stamp[adj[p]] = v;
            }
        if (any) {
            std::fill(stamp.begin(), stamp.end(), i64(-1));
            std::vector<i64> xadj2(n + 1, 0), adj2;
            adj2.reserve(nadj);
            for (i64 v = 0; v < n; ++v) {
                for (i64 p = xadj[v]; p < xadj[v + 1]; ++p)
                    if (stamp[adj[p]] != v) { stamp[adj[p]] = v; adj2.push_back(adj[p]); }
                xadj2[v + 1] = (i64)adj2.size();
            }
            xadj.swap(xadj2);
            adj.swap(adj2);
        }
    }
    std::vector<i64> iperm(n);
    i64 nv = n;
    // One ordering at a time per process: this METIS build keeps its allocator core and its random-number state in plain
    // globals (`gkmcore` is a bss symbol of libmetis_static.a, not thread-local), so two concurrent calls -- the
    // reference's driver analyses MATRIX_THREAD_NUM matrices in parallel, CH.c:3375 -- race and return orderings that
    // differ from run to run.  Serialised, every call starts from the same seed and is deterministic.
    static std::mutex metis_mutex;
    std::lock_guard<std::mutex> lock(metis_mutex);
    const int procs = ordering_processes();
    if (procs > 1 && parallel_nested_dissection(n, xadj, adj, procs, perm)) return;
    METIS_NodeND(&nv, xadj.data(), adj.data(), nullptr, nullptr, perm.data(), iperm.data());
}

// ---------------------------------------------------------------------------------------------------------
// Symmetric permutation into triangular halves.  Reference: SparseFrame_perm, CH.c:956-1066 / LU.c:1068-1288.
// Entry order inside each column/row follows the traversal (new column ascending, old entries in file order).
// ---------------------------------------------------------------------------------------------------------
void permute_triangles(const CscView &A, Kind kind, const std::vector<i64> &perm, Tri &L, Tri &LT, Tri &U, Tri &UT) {
    const i64 n = A.n;
    const bool split = (kind == Kind::LU) && !A.symmetric;   // keep lower and upper parts apart
    const i64 nnz = A.colptr[n];
    std::vector<i64> inv(n, -1);
    for (i64 j = 0; j < n; ++j) if (perm[j] >= 0) inv[perm[j]] = j;
    const i64 vw = A.is_complex ? 2 : 1;      // doubles per value (CH.c:1050-1060: Complex entries are moved whole)

    // Every output is a stable counting sort of the same traversal (new column ascending, old entries in file order) on its
    // own key, so the four of them are built independently -- and concurrently: this stage is a third of the analysis once
    // the ordering is cached (cfg2: 0.27 s per call, two calls).  which: 0 = L, 1 = LT, 2 = U, 3 = UT.
    auto build = [&](int which, Tri &t) {
        // destination list of entry (i, j), or -1 if this output does not take it
        auto key = [&](i64 i, i64 j, i64 &index) -> i64 {
            if (!split) {
                const i64 lo = std::min(i, j), hi = std::max(i, j);
                if (which == 0) { index = hi; return lo; }
                index = lo; return hi;
            }
            switch (which) {
                case 0: index = i; return j <= i ? j : -1;      // L by column
                case 1: index = j; return j <= i ? i : -1;      // LT by row
                case 2: index = j; return j >= i ? i : -1;      // U by row
                default: index = i; return j >= i ? j : -1;     // UT by column
            }
        };
        t.ptr.assign(n + 1, 0);
        for (i64 j = 0; j < n; ++j) {
            const i64 jo = perm[j];
            if (jo < 0) continue;
            for (i64 p = A.colptr[jo]; p < A.colptr[jo + 1]; ++p) {
                i64 index;
                const i64 k = key(inv[A.rowind[p]], j, index);
                if (k >= 0) ++t.ptr[k + 1];
            }
        }
        std::partial_sum(t.ptr.begin(), t.ptr.end(), t.ptr.begin());
        t.idx.assign(t.ptr[n], 0);
        t.val.assign(t.ptr[n] * vw, 0.0);
        std::vector<i64> cur(t.ptr.begin(), t.ptr.end() - 1);
        for (i64 j = 0; j < n; ++j) {
            const i64 jo = perm[j];
            if (jo < 0) continue;
            for (i64 p = A.colptr[jo]; p < A.colptr[jo + 1]; ++p) {
                i64 index;
                const i64 k = key(inv[A.rowind[p]], j, index);
                if (k < 0) continue;
                // a value moves as it is: an entry of a complex "symmetric" matrix that the permutation carries across the
                // diagonal is NOT conjugated (CH.c:1048-1060)
                const i64 q = cur[k]++;
                t.idx[q] = index;
                for (i64 w = 0; w < vw; ++w) t.val[q * vw + w] = A.val ? A.val[p * vw + w] : 0.0;
            }
        }
    };
    auto clear = [](Tri &t) { t.ptr.clear(); t.idx.clear(); t.val.clear(); };
    if (!split) { clear(U); clear(UT); }
    if (nnz < 200000) {                       // small: not worth a thread
        build(0, L); build(1, LT);
        if (split) { build(2, U); build(3, UT); }
        return;
    }
    std::vector<std::thread> th;
    th.emplace_back(build, 1, std::ref(LT));
    if (split) { th.emplace_back(build, 2, std::ref(U)); th.emplace_back(build, 3, std::ref(UT)); }
    build(0, L);
    for (auto &x : th) x.join();
}

// ---------------------------------------------------------------------------------------------------------
// Elimination tree (Liu, with path compression through an "ancestor" shortcut array).
// Reference: SparseFrame_etree, CH.c:1068-1127; LU also walks the strictly-upper entries (LU.c:1360-1386).
// ---------------------------------------------------------------------------------------------------------
void elimination_tree(i64 n, const Tri &LT, const Tri *UT, std::vector<i64> &parent) {
    parent.assign(n, -1);
    std::vector<i64> shortcut(n, -1);
    auto climb = [&](i64 i, i64 j) {
        while (true) {
            const i64 a = shortcut[i];
            if (a < 0) { parent[i] = j; shortcut[i] = j; return; }
            if (a == j) return;
            shortcut[i] = j;
            i = a;
        }
    };
    for (i64 j = 0; j < n; ++j) {
        for (i64 p = LT.ptr[j]; p < LT.ptr[j + 1]; ++p) if (LT.idx[p] < j) climb(LT.idx[p], j);
        if (UT) for (i64 p = UT->ptr[j]; p < UT->ptr[j + 1]; ++p) if (UT->idx[p] < j) climb(UT->idx[p], j);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Depth-first postorder.  Reference: SparseFrame_postorder, CH.c:1129-1236.  Without weights children are
// visited in ascending label order; with weights, in ascending (weight, label) order (bucket sort, 1173-1200).
// Roots are visited in ascending label order.
// ---------------------------------------------------------------------------------------------------------
void tree_postorder(i64 n, const std::vector<i64> &parent, const std::vector<i64> *weight, std::vector<i64> &post) {
    post.assign(n, 0);
    std::vector<i64> first_child(n, -1), sibling(n, -1);
    if (!weight) {
        for (i64 j = n - 1; j >= 0; --j) {
            const i64 p = parent[j];
            if (p >= 0 && p < n) { sibling[j] = first_child[p]; first_child[p] = j; }
        }
    } else {
        std::vector<i64> bucket(n + 1, -1);
        for (i64 j = 0; j < n; ++j)
            if (parent[j] >= 0) { const i64 w = (*weight)[j]; sibling[j] = bucket[w]; bucket[w] = j; }
        // w = n included: a column whose count is n (the first column of a dense matrix) is a legal child.  The reference's
        // bucket array has n entries and its scan stops at n - 1 (CH.c:1177-1200): such a column is written out of bounds and
        // dropped from the tree there, leaving Post incomplete -- the one place where this restatement deliberately differs.
        for (i64 w = n; w >= 0; --w)
            for (i64 j = bucket[w]; j >= 0;) {
                const i64 nxt = sibling[j], p = parent[j];
                sibling[j] = first_child[p];
                first_child[p] = j;
                j = nxt;
            }
    }
    std::vector<i64> stack;
    stack.reserve(n);
    for (i64 j = n - 1; j >= 0; --j) if (parent[j] < 0) stack.push_back(j);
    i64 k = 0;
    while (!stack.empty()) {
        const i64 j = stack.back();
        const i64 c = first_child[j];
        if (c >= 0 && c < n) { stack.push_back(c); first_child[j] = sibling[c]; }
        else { stack.pop_back(); post[k++] = j; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Column counts of the Cholesky factor (Gilbert-Ng-Peyton skeleton/leaf counting with a disjoint-set forest).
// Reference: SparseFrame_colcount, CH.c:1238-1352; LU also scans row j of U (LU.c:1603-1626).
// ---------------------------------------------------------------------------------------------------------
void column_counts(i64 n, const Tri &L, const Tri *U, const std::vector<i64> &parent, const std::vector<i64> &post,
                   std::vector<i64> &colcount) {
    colcount.assign(n, 0);
    std::vector<i64> first_desc(n, -1), set(n), prev_leaf(n), prev_nbr(n, -1);
    for (i64 k = 0; k < n; ++k) {
        for (i64 v = post[k]; v >= 0 && first_desc[v] < 0; v = parent[v]) first_desc[v] = k;
    }
    std::iota(set.begin(), set.end(), i64(0));
    std::iota(prev_leaf.begin(), prev_leaf.end(), i64(0));

    auto visit = [&](i64 j, i64 k, i64 i) {
        if (i <= j) return;
        if (first_desc[j] > prev_nbr[i]) {           // j is a new leaf of row i's subtree
            const i64 pl = prev_leaf[i];
            i64 root = pl;
            while (root != set[root]) root = set[root];
            for (i64 v = pl; v != root;) { const i64 nx = set[v]; set[v] = root; v = nx; }
            ++colcount[j];
            --colcount[root];
            prev_leaf[i] = j;
        }
        prev_nbr[i] = k;
    };
    for (i64 k = 0; k < n; ++k) {
        const i64 j = post[k];
        prev_nbr[j] = k;
        for (i64 p = L.ptr[j]; p < L.ptr[j + 1]; ++p) visit(j, k, L.idx[p]);
        if (U) for (i64 p = U->ptr[j]; p < U->ptr[j + 1]; ++p) visit(j, k, U->idx[p]);
        set[j] = parent[j];
    }
    for (i64 k = 0; k < n; ++k) {
        const i64 j = post[k];
        if (parent[j] >= 0) colcount[parent[j]] += colcount[j];
    }
    for (i64 j = 0; j < n; ++j) ++colcount[j];
}

// ---------------------------------------------------------------------------------------------------------
// Relaxed-amalgamation policy.  Reference: should_relax, Include/parameter.h:31-46.
// ---------------------------------------------------------------------------------------------------------
static bool merge_allowed(i64 cols, double zero_rate) {
    static const i64 col_limit[3] = {16, 64, 256};
    static const double rate_limit[3] = {0.8, 0.1, 0.05};
    for (int k = 2; k >= 0; --k)
        if (cols > col_limit[k] && zero_rate > rate_limit[k]) return false;
    return true;
}

// LU: pattern of lower(C + C^T) with sorted columns (LU.c:2254-2396)
static void symmetrized_lower_pattern(const CscView &A, std::vector<i64> &outp, std::vector<i64> &outi) {
    const i64 n = A.n, nnz = A.colptr[n];
    std::vector<i64> tp(n + 1, 0);
    for (i64 p = 0; p < nnz; ++p) ++tp[A.rowind[p] + 1];
    std::partial_sum(tp.begin(), tp.end(), tp.begin());
    std::vector<i64> ti(nnz), cur(tp.begin(), tp.end() - 1), ci(A.rowind, A.rowind + nnz);
    for (i64 j = 0; j < n; ++j)
        for (i64 p = A.colptr[j]; p < A.colptr[j + 1]; ++p) ti[cur[A.rowind[p]]++] = j;
    for (i64 j = 0; j < n; ++j) {
        std::sort(ci.begin() + A.colptr[j], ci.begin() + A.colptr[j + 1]);
        std::sort(ti.begin() + tp[j], ti.begin() + tp[j + 1]);
    }
    outp.assign(n + 1, 0);
    outi.clear();
    outi.reserve(nnz);
    for (i64 j = 0; j < n; ++j) {
        i64 p = A.colptr[j], pe = A.colptr[j + 1], q = tp[j], qe = tp[j + 1];
        while (p < pe && ci[p] < j) ++p;
        while (q < qe && ti[q] < j) ++q;
        while (p < pe || q < qe) {
            if (q >= qe) outi.push_back(ci[p++]);
            else if (p >= pe) outi.push_back(ti[q++]);
            else if (ci[p] < ti[q]) outi.push_back(ci[p++]);
            else if (ci[p] > ti[q]) outi.push_back(ti[q++]);
            else { outi.push_back(ci[p++]); ++q; }
        }
        outp[j + 1] = (i64)outi.size();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Supernodes, row structure, stages.  Reference: SparseFrame_analyze_supernodal, CH.c:1354-1914 / LU.c:1647-2231.
// ---------------------------------------------------------------------------------------------------------
static void supernodal_structure(const CscView &A, Symbolic &S) {
    const i64 n = S.n;
    const bool lu = S.kind == Kind::LU;
    const bool split = lu && !S.symmetric;
    const uint64_t slot = S.devSlotSize;
    // bytes per panel entry: Float or Complex (CH.c:1480-1496, 1558-1574, 1743-1790, 1893-1896)
    const uint64_t FB = S.is_complex ? 2 * sizeof(double) : sizeof(double), IB = sizeof(i64);

    // relabel by the weighted postorder and rebuild the triangles in the final ordering (CH.c:1429-1447)
    {
        std::vector<i64> inv_post(n), perm2(n), parent2(n), cc2(n);
        for (i64 k = 0; k < n; ++k) inv_post[S.Post[k]] = k;
        for (i64 k = 0; k < n; ++k) {
            const i64 v = S.Post[k], p = S.Parent[v];
            perm2[k] = S.Perm[v];
            parent2[k] = p < 0 ? -1 : inv_post[p];
            cc2[k] = S.ColCount[v];
        }
        S.Perm.swap(perm2); S.Parent.swap(parent2); S.ColCount.swap(cc2);
        permute_triangles(A, S.kind, S.Perm, S.L, S.LT, S.U, S.UT);
    }
    const std::vector<i64> &Parent = S.Parent, &CC = S.ColCount;

    // fundamental supernodes, capped so that one panel (+ its row list) fits a device slot (CH.c:1471-1502)
    std::vector<i64> nchild(n, 0);
    for (i64 j = 0; j < n; ++j) if (Parent[j] >= 0 && Parent[j] < n) ++nchild[Parent[j]];
    std::vector<i64> fstart;
    fstart.reserve(n + 1);
    if (n > 0) fstart.push_back(0);
    for (i64 j = 1; j < n; ++j) {
        const i64 head = fstart.back();
        const uint64_t width = uint64_t(j - head + 1), cc = uint64_t(CC[head]);
        const uint64_t entries = lu ? width * (2 * cc - width) : width * cc;
        const bool chain_breaks = Parent[j - 1] != j || CC[j - 1] != CC[j] + 1 || nchild[j] > 1;
        if (chain_breaks || entries * FB + cc * IB > slot) fstart.push_back(j);
    }
    const i64 nf = (i64)fstart.size();
    fstart.push_back(n);
    S.nfsuper = nf;

    std::vector<i64> width(nf), rows(nf), fmap(n), fparent(nf);
    for (i64 s = 0; s < nf; ++s) {
        width[s] = fstart[s + 1] - fstart[s];
        rows[s] = CC[fstart[s]];
        std::fill(fmap.begin() + fstart[s], fmap.begin() + fstart[s + 1], s);
    }
    for (i64 s = 0; s < nf; ++s) {
        const i64 p = Parent[fstart[s + 1] - 1];
        fparent[s] = p < 0 ? -1 : fmap[p];
    }

    // relaxed amalgamation: a supernode may merge into the group of its parent when the parent's group is the
    // group of the next supernode, the merged panel fits the slot, and the explicit-zero rate is acceptable
    // (CH.c:1524-1591, policy parameter.h:31-46)
    std::vector<i64> group(nf), zeros(nf, 0);
    std::iota(group.begin(), group.end(), i64(0));
    for (i64 s = nf - 2; s >= 0; --s) {
        const i64 p = fparent[s];
        if (p < 0 || p >= nf || group[s + 1] != group[p]) continue;
        const i64 g = group[p];
        const i64 sn = width[s], pn = width[g], sc = rows[s], pc = rows[g];
        const uint64_t cols = uint64_t(sn + pn);
        const uint64_t entries = lu ? cols * uint64_t(sn - pn + 2 * pc) : cols * uint64_t(sn + pc);
        if (entries * FB + uint64_t(sn + pc) * IB > slot) continue;
        const i64 total = zeros[s] + zeros[g] + sn * (sn + pc - sc);
        const i64 dense = (sn + pn) * (sn + pn + 1) / 2 + (sn + pn) * (pc - pn);
        if (!merge_allowed(sn + pn, double(total) / double(dense))) continue;
        width[g] = sn + pn;
        rows[g] = sn + pc;
        zeros[g] = total;
        group[s] = g;
    }
    S.Super.assign(1, 0);
    std::vector<i64> ncol, nrow_;
    for (i64 s = 0; s < nf; ++s)
        if (group[s] == s) { S.Super.push_back(fstart[s + 1]); ncol.push_back(width[s]); nrow_.push_back(rows[s]); }
    const i64 ns = (i64)ncol.size();
    S.nsuper = ns;
    S.Super[ns] = n;
    S.SuperMap.assign(n, 0);
    for (i64 s = 0; s < ns; ++s) std::fill(S.SuperMap.begin() + S.Super[s], S.SuperMap.begin() + S.Super[s + 1], s);
    S.Sparent.assign(ns, -1);
    for (i64 s = 0; s < ns; ++s) {
        const i64 p = Parent[S.Super[s + 1] - 1];
        S.Sparent[s] = p < 0 ? -1 : S.SuperMap[p];
    }

    // panel offsets (CH.c:1632-1645; LU panel = nscol x (2 nsrow - nscol), LU.c:1946)
    S.Lsip.assign(ns + 1, 0);
    S.Lsxp.assign(ns + 1, 0);
    for (i64 s = 0; s < ns; ++s) {
        S.Lsip[s + 1] = S.Lsip[s] + nrow_[s];
        S.Lsxp[s + 1] = S.Lsxp[s] + S.panel_entries(ncol[s], nrow_[s]);
    }
    S.isize = S.Lsip[ns];
    S.xsize = S.Lsxp[ns];

    // row lists: own columns first, then every row j whose row-subtree passes through the supernode, found by
    // climbing the supernodal tree from each nonzero of row j until a supernode already stamped with j
    // (CH.c:1660-1692; LU also climbs from the nonzeros of column j of U, LU.c:1996-2008)
    S.Lsi.assign(S.isize, 0);
    {
        std::vector<i64> fill(S.Lsip.begin(), S.Lsip.end() - 1), stamp(ns);
        for (i64 s = 0; s < ns; ++s) {
            stamp[s] = S.Super[s + 1];
            for (i64 k = S.Super[s]; k < S.Super[s + 1]; ++k) S.Lsi[fill[s]++] = k;
        }
        auto climb = [&](i64 i, i64 j) {
            for (i64 d = S.SuperMap[i]; d >= 0 && stamp[d] <= j; d = S.Sparent[d]) {
                S.Lsi[fill[d]++] = j;
                stamp[d] = j + 1;
            }
        };
        for (i64 j = 0; j < n; ++j) {
            for (i64 p = S.LT.ptr[j]; p < S.LT.ptr[j + 1]; ++p) climb(S.LT.idx[p], j);
            if (split) for (i64 p = S.UT.ptr[j]; p < S.UT.ptr[j + 1]; ++p) climb(S.UT.idx[p], j);
        }
    }

    // largest single update block any descendant sends to one ancestor (CH.c:1694-1719, LU.c:2010-2035)
    S.csize = 0;
    for (i64 s = 0; s < ns; ++s) {
        const i64 nscol = S.Super[s + 1] - S.Super[s], nsrow = S.Lsip[s + 1] - S.Lsip[s];
        if (nscol >= nsrow) continue;
        const i64 *r = S.Lsi.data() + S.Lsip[s];
        i64 seg0 = nscol, owner = S.SuperMap[r[nscol]];
        for (i64 k = nscol; k < nsrow; ++k) {
            const i64 o = S.SuperMap[r[k]];
            if (o == owner) continue;
            const i64 blk = lu ? (k - seg0) * (2 * nsrow - k - seg0) : (k - seg0) * (nsrow - seg0);
            S.csize = std::max(S.csize, blk);
            seg0 = k;
            owner = o;
        }
        S.csize = std::max(S.csize, (nsrow - seg0) * (nsrow - seg0));
    }

    // "stages": greedy packing of supernodal subtrees into device slots, scanned from the root down
    // (CH.c:1721-1846).  Stage ids are reversed at the end so that leaves come first.
    {
        std::vector<i64> head(ns, -1), next(ns, -1), asz(ns, 0), msz(ns, 0), stage(ns, -1);
        i64 nst = ns > 0 ? 1 : 0;
        auto fits = [&](i64 st, i64 pe, i64 nr) {
            return uint64_t(asz[st] + pe) * FB + uint64_t(msz[st] + nr) * IB <= slot;
        };
        for (i64 s = ns - 1; s >= 0; --s) {
            const i64 nr = S.Lsip[s + 1] - S.Lsip[s];
            const i64 pe = S.panel_entries(S.Super[s + 1] - S.Super[s], nr);
            const i64 par = S.Sparent[s];
            i64 st;
            if (par >= 0) {
                st = stage[par];
                if (fits(st, pe, nr)) { stage[s] = st; asz[st] += pe; msz[st] += nr; continue; }
                st = head[stage[par]];
            } else {
                st = 0;
            }
            for (; st >= 0; st = next[st])
                if (fits(st, pe, nr)) { stage[s] = st; asz[st] += pe; msz[st] += nr; break; }
            if (st >= 0) continue;
            stage[s] = nst;
            asz[nst] = pe;
            msz[nst] = nr;
            if (par >= 0) { next[nst] = head[stage[par]]; head[stage[par]] = nst; }
            else { next[nst] = next[0]; next[0] = nst; }
            ++nst;
        }
        S.nstage = nst;
        S.ST_Map.assign(ns, 0);
        for (i64 s = 0; s < ns; ++s) S.ST_Map[s] = nst - 1 - stage[s];
        S.ST_Pointer.assign(nst + 1, 0);
        for (i64 s = 0; s < ns; ++s) ++S.ST_Pointer[S.ST_Map[s] + 1];
        std::partial_sum(S.ST_Pointer.begin(), S.ST_Pointer.end(), S.ST_Pointer.begin());
        S.ST_Index.assign(ns, 0);
        std::vector<i64> cur(S.ST_Pointer.begin(), S.ST_Pointer.end() - 1);
        for (i64 s = 0; s < ns; ++s) S.ST_Index[cur[S.ST_Map[s]]++] = s;
    }

    // initial ready queue: supernodes nobody updates first, in stage order (CH.c:1848-1873)
    {
        std::vector<char> has_child(ns, 0);
        for (i64 s = 0; s < ns; ++s) {
            const i64 nscol = S.Super[s + 1] - S.Super[s], nsrow = S.Lsip[s + 1] - S.Lsip[s];
            if (nscol < nsrow) has_child[S.SuperMap[S.Lsi[S.Lsip[s] + nscol]]] = 1;
        }
        S.LeafQueue.assign(ns, -1);
        S.nsleaf = 0;
        for (i64 k = 0; k < ns; ++k) {
            const i64 s = S.ST_Index[k];
            if (!has_child[s]) S.LeafQueue[S.nsleaf++] = s;
        }
    }

    // byte offsets of each panel / row list inside its stage's slot (CH.c:1875-1904)
    S.Aoffset.assign(ns, 0);
    S.Moffset.assign(ns, 0);
    for (i64 st = 0; st < S.nstage; ++st) {
        size_t a = 0, m = 0;
        for (i64 k = S.ST_Pointer[st]; k < S.ST_Pointer[st + 1]; ++k) {
            const i64 s = S.ST_Index[k];
            const i64 nsrow = S.Lsip[s + 1] - S.Lsip[s];
            S.Aoffset[s] = a;
            a += size_t(S.panel_entries(S.Super[s + 1] - S.Super[s], nsrow)) * FB;
            S.Moffset[s] = m;
            m += size_t(nsrow) * IB;
        }
        for (i64 k = S.ST_Pointer[st]; k < S.ST_Pointer[st + 1]; ++k) S.Moffset[S.ST_Index[k]] += a;
    }
}

static double wall() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int analyze(const CscView &A, Kind kind, uint64_t devSlotSize, Symbolic &S, const i64 *given_perm) {
    if (A.n < 0 || !A.colptr || (A.colptr[A.n] > 0 && !A.rowind)) return 1;
    for (i64 j = 0; j < A.n; ++j) if (A.colptr[j + 1] < A.colptr[j]) return 1;
    for (i64 p = 0; p < A.colptr[A.n]; ++p) if (A.rowind[p] < 0 || A.rowind[p] >= A.n) return 1;
    const double t0 = wall();
    S = Symbolic();
    S.kind = kind;
    S.symmetric = (kind == Kind::Cholesky) ? true : A.symmetric;
    S.is_complex = A.is_complex;
    S.n = A.n;
    S.nnz = A.colptr[A.n];
    S.devSlotSize = devSlotSize;
    const bool split = kind == Kind::LU && !A.symmetric;

    if (split) symmetrized_lower_pattern(A, S.CPCTp, S.CPCTi);       // part of matrix_info either way
    if (given_perm) {
        S.Perm.assign(given_perm, given_perm + A.n);
    } else if (split) {
        fill_reducing_order(A.n, S.CPCTp.data(), S.CPCTi.data(), S.Perm);
    } else {
        fill_reducing_order(A.n, A.colptr, A.rowind, S.Perm);
    }
    {   // METIS must hand back a permutation
        std::vector<char> seen(A.n, 0);
        for (i64 j = 0; j < A.n; ++j) {
            if (S.Perm[j] < 0 || S.Perm[j] >= A.n || seen[S.Perm[j]]) return 2;
            seen[S.Perm[j]] = 1;
        }
    }
    S.Order = S.Perm;
    const double t1 = wall();
    S.t_order = given_perm ? 0.0 : t1 - t0;
    CscView Av = A;
    Av.symmetric = S.symmetric;
    permute_triangles(Av, kind, S.Perm, S.L, S.LT, S.U, S.UT);
    elimination_tree(S.n, S.LT, split ? &S.UT : nullptr, S.Parent);
    tree_postorder(S.n, S.Parent, nullptr, S.Post);
    column_counts(S.n, S.L, split ? &S.U : nullptr, S.Parent, S.Post, S.ColCount);
    tree_postorder(S.n, S.Parent, &S.ColCount, S.Post);
    const double t2 = wall();
    S.t_tree = t2 - t1;
    supernodal_structure(Av, S);
    S.t_structure = wall() - t2;
    return 0;
}

}  // namespace sfb200
