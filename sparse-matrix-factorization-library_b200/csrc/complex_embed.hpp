// Complex matrices (matrix_info->isComplex) on the FP64 engine: the real embedding.
//
// The reference factorizes a complex matrix with zherk / zgemm / zpotrf / ztrsm (CH.c:2060-2070, 2134-2145) and
// zgemm / magma_zgetrf_nopiv / ztrsm (LU.c:2569-2581, 2652-2663).  Here every complex entry z = a + ib becomes the real
// 2x2 block phi(z) = [a -b; b a] at rows (2i, 2i+1), columns (2j, 2j+1).  phi is a ring homomorphism (phi(xy) = phi(x) phi(y),
// phi(conj z) = phi(z)^T), so
//   * Cholesky: phi(A) is symmetric positive definite iff A is Hermitian positive definite, and because a Cholesky factor is
//     unique, chol(phi(A)) = phi(L) with A = L L^H: complex column j of L is real column 2j of the embedded factor read as
//     (re, im) pairs -- byte for byte, since the rows interleave the same way;
//   * LU without pivoting: lu(phi(A)) = (phi(L) D, D^-1 phi(U)) with D block diagonal, D_k = [1 0; Im u_kk / Re u_kk 1]
//     (the 2x2 diagonal blocks of phi(U) are not triangular; D is what makes them so).  Column 2k+1 of phi(L) D and row 2k of
//     D^-1 phi(U) are untouched by D, so L and U are read off without arithmetic:
//         l_ik = F[2i+1, 2k+1] - i F[2i, 2k+1]  (i > k),      u_kj = F[2k, 2j] - i F[2k, 2j+1]  (j >= k).
//     The real elimination divides by Re u_kk where the complex one divides by u_kk: it needs Re u_kk != 0 and loses accuracy
//     when |Re u_kk| << |u_kk|.  zengine.cu therefore rotates every row by the phase of its diagonal entry first (a unitary
//     diagonal scaling, undone exactly when the factors are read back).  For a matrix that is strictly diagonally dominant by
//     columns the pivots stay in the discs |u_kk - a_kk| <= sum_{i != k} |a_ik|, so after the rotation
//     Re u'_kk >= |a_kk| - sum_{i != k} |a_ik| > 0: the bound the complex elimination has for |u_kk|.
//   * the supernodal structure of phi(A) is the structure of A with every index doubled, so no second analysis is needed;
//   * forward / backward substitution with the embedded factor on the interleaved right-hand side IS the complex solve.
// The price: the embedded factorization executes twice the real flops of a native complex kernel (8/3 n^3 instead of 4/3 n^3
// for a dense front) and holds the factor twice.  What it buys: every kernel on the path is the FP64 engine's own.
#pragma once
#include <cstdint>
#include <vector>

#include "plan.hpp"
#include "symbolic.hpp"

namespace sfb200 {

// Where one value of the embedding comes from: entry (code >> 2) of the complex value array, its imaginary (bit 0) or real
// part, negated (bit 1) or not.  kZSrcZero: structural zero.
constexpr i64 kZSrcZero = -1;

struct ComplexEmbedding {
    Symbolic R;                      // structure of phi(P A P^T): what build_plan reads (Super, Lsip, Lsxp, Lsi, SuperMap, L, U)
    std::vector<i64> src_l, src_u;   // per entry of R.L / R.U
    std::vector<int32_t> row_l, row_u;   // matrix row of every COMPLEX entry feeding R.L / R.U (row-phase scaling)
    bool u_from_l = false;           // LU of a complex symmetric matrix (A = A^T): R.U is fed from the L values (LU.c:2718-2729)
};

// Z: the analysis of the complex matrix (Symbolic::is_complex) or a wrapped foreign structure.  Structure only.
int embed_complex(const Symbolic &Z, ComplexEmbedding &E);

// out[q] = value of embedded entry q, from the interleaved complex values z; with `phase` (n unit complex numbers d_k) the
// value of an entry in row r is conj(d_r) z: the rows of the matrix are rotated so that its diagonal becomes real, >= 0
void expand_values(const std::vector<i64> &src, const double *z, const int32_t *rowz, const double *phase, double *out);

// E.R.L.val / E.R.U.val from the values Z carries (host path: CPU interpreter, tests)
int expand_embedding(const Symbolic &Z, ComplexEmbedding &E, const double *phase);

// position of the diagonal entry of every row in the values that feed the upper part (the lower part for Cholesky and for LU
// of a symmetric matrix): the LAST copy, the one loadA keeps (LU.c:2496-2517); -1 if the row has none
void diagonal_positions(const Symbolic &Z, std::vector<i64> &pos);
// d_k = a_kk / |a_kk| (1 where the diagonal entry is absent or zero), interleaved
void row_phases(const std::vector<i64> &pos, const double *z, double *phase);

// One chunk of columns of one complex panel, read back from the image of the embedded factor (k_zextract, and the CPU
// interpreter's restatement of it in oracle/plan_host_exec.cpp).  `mode`: 0 Cholesky, 1 LU.
struct ZExtractTask {
    i64 src;                 // offset of the embedded panel in the image (SnodeDev::xoff)
    i64 dst;                 // offset, in complex elements, of the panel in Lsx (Lsxp[s])
    i64 lsip;                // offset of the supernode's row list in Lsi (row phases)
    int32_t nsrow, nscol;    // of the complex supernode
    int32_t c0, ncols;       // complex columns [c0, c0 + ncols) of the panel
    int32_t ld;              // leading dimension of the embedded panel
    int32_t mode;
};

// P: single-GPU plan of E.R; Z: the complex structure.  One task per <= 32 K entries.
void build_zextract_tasks(const HostPlan &P, const Symbolic &Z, std::vector<ZExtractTask> &tasks);

// Reference semantics of one task (test infrastructure and documentation of the kernel): X = image, out = complex Lsx.
// `phase`: n unit complex numbers d_k (re, im) or nullptr -- rows of A were multiplied by conj(d_k) before the factorization
// (LU only); Lsi = the row lists of the complex structure, read only with `phase`.
void zextract_host(const ZExtractTask &t, const double *X, double *out, const double *phase, const i64 *Lsi);

}  // namespace sfb200
