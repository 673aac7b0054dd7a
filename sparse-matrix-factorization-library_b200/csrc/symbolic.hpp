// Host-side symbolic analysis for the B200 numeric factorization engine.
//
// The north star keeps symbolic analysis in host C/C++ and demands that its result be bit-exact with the
// reference (SparseFrame_analyze, Cholesky/Source/SparseFrame.c:1916-1978 and LU/Source/SparseFrame.c:2233-2458).
// This is an independent re-statement: one parameterised pipeline covers Cholesky and LU (the LU reference is a
// copy of the Cholesky one with a different panel-size formula and an extra pass over the upper triangle).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sfb200 {

using i64 = int64_t;

enum class Kind : int { Cholesky = 0, LU = 1 };

// Compressed-column input exactly as SparseFrame_compress leaves it (CH.c:526-587): column pointers, row
// indices in file order inside each column, no de-duplication, no sorting.  Symmetric inputs hold one triangle.
struct CscView {
    i64 n = 0;
    const i64 *colptr = nullptr;
    const i64 *rowind = nullptr;
    const double *val = nullptr;
    bool symmetric = true;
    bool is_complex = false;     // val holds interleaved (re, im) pairs: the reference's Complex (Include/arch.h:9-16)
};

// One triangular half in compressed form (pointer / index / value), reference field names in comments.
struct Tri {
    std::vector<i64> ptr, idx;
    std::vector<double> val;     // one double per entry; two (re, im) for a complex matrix
};

struct Symbolic {
    Kind kind = Kind::Cholesky;
    bool symmetric = true;          // LU only: unsymmetric input keeps a separate U
    bool is_complex = false;        // matrix_info->isComplex: values are (re, im) pairs, panels hold Complex (CH.c:1648-1651)
    i64 n = 0;
    i64 nnz = 0;
    uint64_t devSlotSize = 0;       // common_info->devSlotSize; the structure depends on it (CH.c:1476-1496, 1556-1574)

    std::vector<i64> Order;         // fill-reducing ordering as METIS_NodeND returned it (before the postorder relabelling)
    std::vector<i64> Perm;          // final ordering: METIS composed with the weighted postorder (CH.c:1438-1443)
    std::vector<i64> Parent;        // etree of the permuted matrix, relabelled (CH.c:1439)
    std::vector<i64> Post;          // weighted postorder in pre-relabel labels (matrix_info->Post)
    std::vector<i64> ColCount;      // relabelled column counts

    Tri L, LT;                      // Lp/Li/Lx : lower(P A P^T) by column ; LTp/LTi/LTx : by row
    Tri U, UT;                      // LU + unsymmetric only: Up/Ui/Ux upper part by row ; UTp/UTi/UTx by column

    i64 nfsuper = 0, nsuper = 0;
    std::vector<i64> Super, SuperMap, Sparent;
    i64 isize = 0, xsize = 0, csize = 0;
    std::vector<i64> Lsip, Lsxp, Lsi;

    // The reference's device-cache bookkeeping.  The B200 engine keeps the whole factor resident and never reads
    // these, but they are part of matrix_info and of "bit-exact symbolic structure", so they are reproduced.
    i64 nstage = 0;
    std::vector<i64> ST_Map, ST_Pointer, ST_Index;
    i64 nsleaf = 0;
    std::vector<i64> LeafQueue;
    std::vector<size_t> Aoffset, Moffset;

    // LU + unsymmetric only: pattern of lower(C + C^T) fed to METIS (LU.c:2254-2396)
    std::vector<i64> CPCTp, CPCTi;

    // wall time of the stages in seconds (where the analysis time goes)
    double t_order = 0.0;           // adjacency + METIS_NodeND (0 when the caller supplied the ordering)
    double t_tree = 0.0;            // permutation into triangles, etree, postorders, column counts
    double t_structure = 0.0;       // supernodes, row lists, stages

    i64 panel_entries(i64 nscol, i64 nsrow) const {
        return kind == Kind::LU ? nscol * (2 * nsrow - nscol) : nscol * nsrow;
    }
};

// Full analysis.  Returns 0 on success, non-zero on a malformed input or a METIS failure (2: not a permutation).
// given_perm != nullptr: use this fill-reducing ordering (new position k holds old index given_perm[k], the meaning of
// METIS_NodeND's perm / matrix_info->Perm before the postorder relabelling) instead of calling METIS -- the ordering is
// 96 % of the analysis time on the 3-D configs, and an ordering saved from an earlier analysis of the same pattern
// reproduces that analysis bit for bit.
int analyze(const CscView &A, Kind kind, uint64_t devSlotSize, Symbolic &out, const i64 *given_perm = nullptr);

// Processes the fill-reducing ordering may use: 1 (default) = one METIS_NodeND call, bit-identical to the reference's ordering;
// > 1 = nested dissection whose top separators are computed here and whose subgraphs are ordered by METIS_NodeND in forked
// processes (parallel_order.cpp) -- same kind of ordering, not the same permutation.
void set_ordering_processes(int procs);
int ordering_processes();

// Individual stages, exposed for the unit tests.
void fill_reducing_order(i64 n, const i64 *colptr, const i64 *rowind, std::vector<i64> &perm);
void permute_triangles(const CscView &A, Kind kind, const std::vector<i64> &perm, Tri &L, Tri &LT, Tri &U, Tri &UT);
void elimination_tree(i64 n, const Tri &LT, const Tri *UT, std::vector<i64> &parent);
void tree_postorder(i64 n, const std::vector<i64> &parent, const std::vector<i64> *weight, std::vector<i64> &post);
void column_counts(i64 n, const Tri &L, const Tri *U, const std::vector<i64> &parent, const std::vector<i64> &post,
                   std::vector<i64> &colcount);

}  // namespace sfb200
