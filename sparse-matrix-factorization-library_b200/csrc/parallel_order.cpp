// Opt-in parallel fill-reducing ordering (METIS_NodeND is 93-96 % of the analysis time and single-threaded).
//
// Nested dissection with the top levels of the separator tree computed here (METIS_ComputeVertexSeparator, the multilevel
// bisection METIS_NodeND itself starts with) and the subgraphs below them ordered by METIS_NodeND in FORKED PROCESSES, one per
// subgraph: this METIS build keeps its allocator and random-number state in plain globals (symbolic.cpp), so threads are out,
// while a forked child has its own copy of them.  Children write their piece of the ordering into an anonymous shared mapping
// and leave with _exit; they touch nothing but METIS and that mapping (no CUDA, no stdio, no locks).
//
// The result is a nested-dissection ordering of the same kind and quality as METIS_NodeND's, but NOT the same permutation:
// the symbolic structure that follows is self-consistent, not bit-identical to the reference's.  That is why it is opt-in
// (sfb200_set_ordering_processes / SF_PARALLEL_ORDER) and the default stays the reference's single METIS_NodeND call.
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "symbolic.hpp"

extern "C" int METIS_NodeND(int64_t *nvtxs, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *options, int64_t *perm,
                            int64_t *iperm);
extern "C" int METIS_ComputeVertexSeparator(int64_t *nvtxs, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *options,
                                            int64_t *sepsize, int64_t *part);

namespace sfb200 {

namespace {

struct Graph {
    std::vector<i64> xadj, adj;     // local numbering 0..n-1
    std::vector<i64> gid;           // global vertex of every local one
    i64 n() const { return (i64)gid.size(); }
};

constexpr i64 kMinSplit = 20000;    // below this a subgraph is ordered in one piece: a separator costs more than it saves

// out[k] = global vertex eliminated k-th among this graph's vertices
void node_nd(Graph &g, i64 *out) {
    i64 nv = g.n();
    if (nv == 0) return;
    if (g.xadj[nv] == 0) { std::copy(g.gid.begin(), g.gid.end(), out); return; }
    std::vector<i64> perm(nv), iperm(nv);
    METIS_NodeND(&nv, g.xadj.data(), g.adj.data(), nullptr, nullptr, perm.data(), iperm.data());
    for (i64 k = 0; k < nv; ++k) out[k] = g.gid[perm[k]];
}

// induced subgraph on the vertices with part[v] == which
void induced(const Graph &g, const std::vector<i64> &part, i64 which, Graph &sub) {
    const i64 n = g.n();
    std::vector<i64> local(n, -1);
    sub.gid.clear();
    for (i64 v = 0; v < n; ++v) if (part[v] == which) { local[v] = (i64)sub.gid.size(); sub.gid.push_back(g.gid[v]); }
    sub.xadj.assign(sub.gid.size() + 1, 0);
    sub.adj.clear();
    i64 k = 0;
    for (i64 v = 0; v < n; ++v) {
        if (part[v] != which) continue;
        for (i64 p = g.xadj[v]; p < g.xadj[v + 1]; ++p) if (local[g.adj[p]] >= 0) sub.adj.push_back(local[g.adj[p]]);
        sub.xadj[++k] = (i64)sub.adj.size();
    }
}

// `out` lies in a MAP_SHARED mapping: a forked child fills its slice.  levels = separator levels still to be opened here.
void order_graph(Graph &g, int levels, i64 *out) {
    i64 nv = g.n();
    if (levels <= 0 || nv < kMinSplit) { node_nd(g, out); return; }
    std::vector<i64> part(nv, 0);
    i64 sepsize = 0;
    if (METIS_ComputeVertexSeparator(&nv, g.xadj.data(), g.adj.data(), nullptr, nullptr, &sepsize, part.data()) != 1 /* METIS_OK */) {
        node_nd(g, out);
        return;
    }
    Graph g0, g1;
    induced(g, part, 0, g0);
    induced(g, part, 1, g1);
    if (g0.n() == 0 || g1.n() == 0) { node_nd(g, out); return; }
    // elimination order: part 0, part 1, then the separator (the order METIS_NodeND gives its own separators)
    i64 *out0 = out, *out1 = out + g0.n(), *outs = out + g0.n() + g1.n();
    for (i64 v = 0, k = 0; v < nv; ++v) if (part[v] == 2) outs[k++] = g.gid[v];
    { Graph drop; std::swap(drop, g); }                    // the parent graph is no longer needed: free it before recursing
    const pid_t child = fork();
    if (child == 0) {
        order_graph(g0, levels - 1, out0);
        _exit(0);
    }
    order_graph(g1, levels - 1, out1);
    bool child_ok = false;
    if (child > 0) {
        int status = 0;
        child_ok = waitpid(child, &status, 0) == child && WIFEXITED(status) && WEXITSTATUS(status) == 0;
    }
    if (!child_ok) order_graph(g0, 0, out0);               // no process to be had (or it died): order that half here
}

}  // namespace

// perm[k] = vertex eliminated k-th (the meaning of METIS_NodeND's perm).  procs: processes to use (rounded down to a power
// of two).  Returns false if the shared mapping cannot be created (the caller then orders with one METIS_NodeND call).
bool parallel_nested_dissection(i64 n, const std::vector<i64> &xadj, const std::vector<i64> &adj, int procs, std::vector<i64> &perm) {
    int levels = 0;
    while ((2 << levels) <= procs) ++levels;
    if (levels == 0 || n < kMinSplit) return false;
    const size_t bytes = sizeof(i64) * (size_t)n;
    void *shared = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shared == MAP_FAILED) return false;
    i64 *out = (i64 *)shared;
    std::fill(out, out + n, i64(-1));
    Graph g;
    g.xadj = xadj;
    g.adj = adj;
    g.gid.resize(n);
    std::iota(g.gid.begin(), g.gid.end(), i64(0));
    order_graph(g, levels, out);
    // every vertex exactly once, or the caller orders with the single METIS_NodeND call (a child that was killed, a METIS
    // failure inside one: its slice stays -1)
    std::vector<char> seen((size_t)n, 0);
    bool ok = true;
    for (i64 k = 0; k < n && ok; ++k) {
        const i64 v = out[k];
        if (v < 0 || v >= n || seen[v]) ok = false;
        else seen[v] = 1;
    }
    if (ok) perm.assign(out, out + n);
    munmap(shared, bytes);
    return ok;
}

}  // namespace sfb200
