// Real embedding of a complex supernodal structure -- see complex_embed.hpp.  Host only.
#include "complex_embed.hpp"

#include <algorithm>
#include <cmath>

namespace sfb200 {

namespace {

inline i64 zsrc(i64 entry, bool imag, bool neg) { return (entry << 2) | (neg ? 2 : 0) | (imag ? 1 : 0); }

// phi(z)[alpha][beta] for z = a + ib : [a -b; b a]
inline i64 phi_src(i64 entry, int alpha, int beta) {
    if (alpha == beta) return zsrc(entry, false, false);
    return alpha == 1 ? zsrc(entry, true, false) : zsrc(entry, true, true);
}

// One triangular half of phi(A).  T lists, per major index m (a column of the lower part / a row of the upper part), the minor
// indices of its entries; the embedding keeps, per real major index 2m + mu, the real entries on the same side of the REAL
// diagonal.  `hermitian_diag`: the imaginary part of a diagonal entry is taken as zero (zpotrf reads the real part only).
void embed_half(const Tri &T, i64 n, bool lower_by_col, bool hermitian_diag, Tri &out, std::vector<i64> &src) {
    out.ptr.assign(2 * n + 1, 0);
    // an off-diagonal entry gives 2 real entries to each of the two real majors; a diagonal one gives 2 + 1 (or 1 + 1)
    for (i64 m = 0; m < n; ++m)
        for (i64 p = T.ptr[m]; p < T.ptr[m + 1]; ++p) {
            const bool diag = T.idx[p] == m;
            if (!diag) { out.ptr[2 * m + 1] += 2; out.ptr[2 * m + 2] += 2; }
            else if (lower_by_col) { out.ptr[2 * m + 1] += hermitian_diag ? 1 : 2; out.ptr[2 * m + 2] += 1; }
            else { out.ptr[2 * m + 1] += 2; out.ptr[2 * m + 2] += 1; }
        }
    for (i64 k = 0; k < 2 * n; ++k) out.ptr[k + 1] += out.ptr[k];
    out.idx.assign(out.ptr[2 * n], 0);
    src.assign(out.ptr[2 * n], kZSrcZero);
    std::vector<i64> cur(out.ptr.begin(), out.ptr.end() - 1);
    for (i64 m = 0; m < n; ++m)
        for (i64 p = T.ptr[m]; p < T.ptr[m + 1]; ++p) {
            const i64 r = T.idx[p];                    // minor index: a row (lower part) or a column (upper part)
            for (int mu = 0; mu < 2; ++mu)
                for (int rho = 0; rho < 2; ++rho) {
                    const i64 major = 2 * m + mu, minor = 2 * r + rho;
                    if (r == m) {
                        if (minor < major) continue;                   // the other triangle's corner of a diagonal block
                        if (hermitian_diag && mu != rho) continue;     // Im a_jj of a Hermitian matrix: taken as zero
                    }
                    // lower part by columns: real entry (row 2r + rho, column 2m + mu); upper part by rows: (2m + mu, 2r + rho)
                    const int alpha = lower_by_col ? rho : mu, beta = lower_by_col ? mu : rho;
                    const i64 q = cur[major]++;
                    out.idx[q] = minor;
                    src[q] = phi_src(p, alpha, beta);
                }
        }
}

}  // namespace

void expand_values(const std::vector<i64> &src, const double *z, const int32_t *rowz, const double *phase, double *out) {
    for (size_t q = 0; q < src.size(); ++q) {
        const i64 c = src[q];
        if (c < 0) { out[q] = 0.0; continue; }
        const i64 p = c >> 2;
        double re = z[2 * p], im = z[2 * p + 1];
        if (phase) {                                   // conj(d_row) * z
            const double dr = phase[2 * (i64)rowz[p]], di = phase[2 * (i64)rowz[p] + 1];
            const double tr = dr * re + di * im;
            im = dr * im - di * re;
            re = tr;
        }
        const double v = (c & 1) ? im : re;
        out[q] = (c & 2) ? -v : v;
    }
}

void diagonal_positions(const Symbolic &Z, std::vector<i64> &pos) {
    const bool from_u = Z.kind == Kind::LU && !Z.symmetric;
    const Tri &T = from_u ? Z.U : Z.L;
    pos.assign(Z.n, -1);
    for (i64 m = 0; m < Z.n; ++m)
        for (i64 p = T.ptr[m]; p < T.ptr[m + 1]; ++p) if (T.idx[p] == m) pos[m] = p;      // the last copy stays
}

void row_phases(const std::vector<i64> &pos, const double *z, double *phase) {
    for (size_t k = 0; k < pos.size(); ++k) {
        double re = 1.0, im = 0.0;
        if (pos[k] >= 0) {
            const double a = z[2 * pos[k]], b = z[2 * pos[k] + 1];
            const double m = std::hypot(a, b);
            if (m > 0.0) { re = a / m; im = b / m; }
        }
        phase[2 * k] = re; phase[2 * k + 1] = im;
    }
}

int embed_complex(const Symbolic &Z, ComplexEmbedding &E) {
    const i64 n = Z.n, ns = Z.nsuper;
    const bool lu = Z.kind == Kind::LU;
    if ((i64)Z.Super.size() != ns + 1 || (i64)Z.Lsip.size() != ns + 1 || Z.L.ptr.size() != (size_t)n + 1) return 1;
    for (i64 j = 0; j < n; ++j)            // both halves hold their own side of the diagonal
        for (i64 p = Z.L.ptr[j]; p < Z.L.ptr[j + 1]; ++p) if (Z.L.idx[p] < j) return 1;
    const bool have_u = lu && !Z.symmetric;
    if (have_u) {
        if (Z.U.ptr.size() != (size_t)n + 1) return 1;
        for (i64 i = 0; i < n; ++i)
            for (i64 p = Z.U.ptr[i]; p < Z.U.ptr[i + 1]; ++p) if (Z.U.idx[p] < i) return 1;
    }
    E = ComplexEmbedding();
    Symbolic &R = E.R;
    R.kind = Z.kind;
    R.is_complex = false;
    R.symmetric = !lu;                      // phi(A)^T = phi(A^H): only the Hermitian (Cholesky) case stays symmetric
    R.n = 2 * n;
    R.nsuper = ns;
    R.devSlotSize = Z.devSlotSize;
    R.Super.resize(ns + 1);
    R.Lsip.resize(ns + 1);
    R.Lsxp.assign(ns + 1, 0);
    for (i64 s = 0; s <= ns; ++s) { R.Super[s] = 2 * Z.Super[s]; R.Lsip[s] = 2 * Z.Lsip[s]; }
    for (i64 s = 0; s < ns; ++s)
        R.Lsxp[s + 1] = R.Lsxp[s] + R.panel_entries(R.Super[s + 1] - R.Super[s], R.Lsip[s + 1] - R.Lsip[s]);
    R.isize = R.Lsip[ns];
    R.xsize = R.Lsxp[ns];
    R.Lsi.resize(R.isize);
    for (i64 q = 0; q < Z.isize; ++q) { R.Lsi[2 * q] = 2 * Z.Lsi[q]; R.Lsi[2 * q + 1] = 2 * Z.Lsi[q] + 1; }
    R.SuperMap.resize(2 * n);
    for (i64 s = 0; s < ns; ++s) std::fill(R.SuperMap.begin() + R.Super[s], R.SuperMap.begin() + R.Super[s + 1], s);
    R.Sparent = Z.Sparent;

    embed_half(Z.L, n, true, !lu, R.L, E.src_l);
    R.nnz = R.L.ptr[2 * n];
    E.row_l.assign(Z.L.idx.begin(), Z.L.idx.begin() + Z.L.ptr[n]);            // an entry of the lower part sits in row idx
    if (lu) {
        // LU of a complex symmetric matrix: row j of U is column j of L (LU.c:2718-2729), same values
        E.u_from_l = !have_u;
        const Tri &T = have_u ? Z.U : Z.L;
        embed_half(T, n, false, false, R.U, E.src_u);
        E.row_u.resize(T.ptr[n]);                                              // an entry of the upper part sits in its major row
        for (i64 m = 0; m < n; ++m) std::fill(E.row_u.begin() + T.ptr[m], E.row_u.begin() + T.ptr[m + 1], (int32_t)m);
    }
    return 0;
}

int expand_embedding(const Symbolic &Z, ComplexEmbedding &E, const double *phase) {
    const i64 n = Z.n;
    const bool lu = Z.kind == Kind::LU, have_u = lu && !Z.symmetric;
    if (Z.L.val.size() != (size_t)Z.L.ptr[n] * 2) return 1;
    E.R.L.val.resize(E.src_l.size());
    expand_values(E.src_l, Z.L.val.data(), E.row_l.data(), phase, E.R.L.val.data());
    if (lu) {
        const std::vector<double> &uz = have_u ? Z.U.val : Z.L.val;
        if (uz.size() != (size_t)((have_u ? Z.U : Z.L).ptr[n]) * 2) return 1;
        E.R.U.val.resize(E.src_u.size());
        expand_values(E.src_u, uz.data(), E.row_u.data(), phase, E.R.U.val.data());
    }
    return 0;
}

void build_zextract_tasks(const HostPlan &P, const Symbolic &Z, std::vector<ZExtractTask> &tasks) {
    tasks.clear();
    const bool lu = Z.kind == Kind::LU;
    for (i64 s = 0; s < Z.nsuper; ++s) {
        const SnodeDev &d = P.snode[s];
        const int nscol = (int)(Z.Super[s + 1] - Z.Super[s]), nsrow = (int)(Z.Lsip[s + 1] - Z.Lsip[s]);
        const int rows = lu ? 2 * nsrow - nscol : nsrow;
        const int chunk = std::max(1, 32768 / std::max(rows, 1));
        for (int c0 = 0; c0 < nscol; c0 += chunk) {
            ZExtractTask t{};
            t.src = d.xoff;
            t.dst = Z.Lsxp[s];
            t.lsip = Z.Lsip[s];
            t.nsrow = nsrow; t.nscol = nscol;
            t.c0 = c0; t.ncols = std::min(chunk, nscol - c0);
            t.ld = d.ld;
            t.mode = lu ? 1 : 0;
            tasks.push_back(t);
        }
    }
}

void zextract_host(const ZExtractTask &t, const double *X, double *out, const double *phase, const i64 *Lsi) {
    const double *F = X + t.src;
    const i64 *rows = phase ? Lsi + t.lsip : nullptr;
    const i64 ld = t.ld;
    const int slda = t.mode ? 2 * t.nsrow - t.nscol : t.nsrow;       // complex rows of the reference panel (LU.c:1946)
    for (int k = t.c0; k < t.c0 + t.ncols; ++k)
        for (int r = 0; r < slda; ++r) {
            double re, im;
            i64 prow = -1, pcol = -1;             // phase factors: value *= d[prow] * conj(d[pcol])
            if (!t.mode) {
                // column 2k of phi(L), rows (2r, 2r+1).  Above the diagonal the reference's panel holds the zeros of loadA's
                // memset (CH.c:2004-2007: zherk / zpotrf touch the lower part only), on it a real number (zpotrf)
                re = r < k ? 0.0 : F[2 * r + (2 * (i64)k) * ld];
                im = r <= k ? 0.0 : F[2 * r + 1 + (2 * (i64)k) * ld];
            } else if (r < t.nsrow && r > k) {    // L, strictly below the diagonal
                re = F[2 * r + 1 + (2 * (i64)k + 1) * ld];
                im = -F[2 * r + (2 * (i64)k + 1) * ld];
                if (phase) { prow = rows[r]; pcol = rows[k]; }
            } else if (r < t.nsrow) {             // U11 (r <= k): u_rk
                re = F[2 * r + (2 * (i64)k) * ld];
                im = -F[2 * r + (2 * (i64)k + 1) * ld];
                if (phase) prow = rows[r];
            } else {                              // U12^T: panel row nsrow + t, column k holds u_{k, rows[nscol + t]}
                const i64 tt = r - t.nsrow;
                re = F[2 * (i64)t.nsrow + 2 * tt + (2 * (i64)k) * ld];
                im = -F[2 * (i64)t.nsrow + 2 * tt + 1 + (2 * (i64)k) * ld];
                if (phase) prow = rows[k];
            }
            if (prow >= 0) {
                double dr = phase[2 * prow], di = phase[2 * prow + 1];
                if (pcol >= 0) {                  // d_row * conj(d_col)
                    const double cr = phase[2 * pcol], ci = phase[2 * pcol + 1];
                    const double tr = dr * cr + di * ci, ti = di * cr - dr * ci;
                    dr = tr; di = ti;
                }
                const double tr = re * dr - im * di, ti = re * di + im * dr;
                re = tr; im = ti;
            }
            double *o = out + 2 * (t.dst + (i64)k * slda + r);
            o[0] = re; o[1] = im;
        }
}

}  // namespace sfb200
