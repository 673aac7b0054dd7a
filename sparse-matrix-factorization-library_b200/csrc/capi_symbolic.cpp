// C ABI, symbolic half (include/sfb200.h).  Host only.
#include <cstring>
#include <string>

#include "../../include/sfb200.h"
#include "capi_internal.hpp"

namespace sfb200 {
thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
}  // namespace sfb200

using namespace sfb200;

extern "C" {

const char *sfb200_version(void) { return "sfb200 0.1 (sm_100a)"; }
const char *sfb200_last_error(void) { return g_last_error.c_str(); }

static int analyze_any(int kind, int is_symmetric, bool is_complex, int64_t n, const int64_t *Cp, const int64_t *Ci,
                       const double *Cx, uint64_t dev_slot_size, const int64_t *order, sfb200_symbolic **out) {
    if (!out || !Cp || n < 0 || (kind != SFB200_CHOLESKY && kind != SFB200_LU)) {
        set_error("sfb200_analyze: bad argument");
        return SFB200_ERR_ARG;
    }
    CscView A;
    A.n = n; A.colptr = Cp; A.rowind = Ci; A.val = Cx; A.symmetric = is_symmetric != 0; A.is_complex = is_complex;
    auto *s = new sfb200_symbolic();
    const int rc = analyze(A, kind == SFB200_LU ? Kind::LU : Kind::Cholesky, dev_slot_size, s->S, order);
    if (rc != 0) {
        delete s;
        set_error(rc == 2 ? (order ? "sfb200_analyze_ordered: `order` is not a permutation of 0..n-1"
                                   : "sfb200_analyze: METIS did not return a permutation")
                          : "sfb200_analyze: malformed matrix");
        return rc == 2 ? SFB200_ERR_ORDERING : SFB200_ERR_ARG;
    }
    *out = s;
    return SFB200_OK;
}

int sfb200_analyze_ordered(int kind, int is_symmetric, int64_t n, const int64_t *Cp, const int64_t *Ci, const double *Cx,
                           uint64_t dev_slot_size, const int64_t *order, sfb200_symbolic **out) {
    return analyze_any(kind, is_symmetric, false, n, Cp, Ci, Cx, dev_slot_size, order, out);
}

int sfb200_analyze_complex(int kind, int is_symmetric, int64_t n, const int64_t *Cp, const int64_t *Ci, const double *Cz,
                           uint64_t dev_slot_size, const int64_t *order, sfb200_symbolic **out) {
    return analyze_any(kind, is_symmetric, true, n, Cp, Ci, Cz, dev_slot_size, order, out);
}

int sfb200_symbolic_wrap_complex(int kind, int64_t n, int64_t nsuper, const int64_t *Super, const int64_t *Lsip,
                                 const int64_t *Lsxp, const int64_t *Lsi, const int64_t *Lp, const int64_t *Li,
                                 const int64_t *Up, const int64_t *Ui, sfb200_symbolic **out) {
    const int rc = sfb200_symbolic_wrap(kind, n, nsuper, Super, Lsip, Lsxp, Lsi, Lp, Li, Up, Ui, out);
    if (rc == SFB200_OK) (*out)->S.is_complex = true;
    return rc;
}

int sfb200_set_ordering_processes(int procs) {
    set_ordering_processes(procs);
    return ordering_processes();
}

int sfb200_analyze(int kind, int is_symmetric, int64_t n, const int64_t *Cp, const int64_t *Ci, const double *Cx,
                   uint64_t dev_slot_size, sfb200_symbolic **out) {
    return sfb200_analyze_ordered(kind, is_symmetric, n, Cp, Ci, Cx, dev_slot_size, nullptr, out);
}

double sfb200_symbolic_double(const sfb200_symbolic *s, const char *name) {
    if (!s || !name) return -1.0;
    const std::string k(name);
    if (k == "t_order") return s->S.t_order;
    if (k == "t_tree") return s->S.t_tree;
    if (k == "t_structure") return s->S.t_structure;
    if (k == "flops") return sfb200_symbolic_flops(s);
    return -1.0;
}

int sfb200_symbolic_wrap(int kind, int64_t n, int64_t nsuper, const int64_t *Super, const int64_t *Lsip,
                         const int64_t *Lsxp, const int64_t *Lsi, const int64_t *Lp, const int64_t *Li,
                         const int64_t *Up, const int64_t *Ui, sfb200_symbolic **out) {
    if (!out || !Super || !Lsip || !Lsxp || !Lsi || !Lp || !Li || n < 0 || nsuper < 0) {
        set_error("sfb200_symbolic_wrap: bad argument");
        return SFB200_ERR_ARG;
    }
    if (kind != SFB200_CHOLESKY && kind != SFB200_LU) { set_error("sfb200_symbolic_wrap: bad kind"); return SFB200_ERR_ARG; }
    // The plan builder indexes SuperMap[Lsi[..]], pos[Li[..]] and Super[t] with these values: everything it relies on is
    // checked here, so that a malformed structure is an error code and never an out-of-bounds access.
    {
        const bool lu = kind == SFB200_LU;
        auto bad = [&](const char *what) {
            set_error(std::string("sfb200_symbolic_wrap: invalid supernodal structure: ") + what);
            return SFB200_ERR_ARG;
        };
        if (Super[0] != 0 || Super[nsuper] != n) return bad("Super does not cover 0..n");
        if (Lsip[0] != 0 || Lsxp[0] != 0) return bad("Lsip / Lsxp do not start at 0");
        for (int64_t k = 0; k < nsuper; ++k) {
            const int64_t nscol = Super[k + 1] - Super[k], nsrow = Lsip[k + 1] - Lsip[k];
            if (nscol <= 0) return bad("Super is not strictly increasing");
            if (nsrow < nscol || nsrow > n) return bad("a supernode has fewer rows than columns (Lsip)");
            const int64_t panel = lu ? (2 * nsrow - nscol) * nscol : nsrow * nscol;
            if (Lsxp[k + 1] - Lsxp[k] != panel) return bad("Lsxp does not match the panel sizes");
            const int64_t *rows = Lsi + Lsip[k];
            for (int64_t i = 0; i < nscol; ++i)
                if (rows[i] != Super[k] + i) return bad("the leading rows of a supernode are not its own columns");
            for (int64_t i = nscol; i < nsrow; ++i)
                if (rows[i] <= rows[i - 1] || rows[i] >= n) return bad("off-diagonal rows not ascending inside 0..n");
        }
        auto check_csc = [&](const int64_t *ptr, const int64_t *idx) {
            if (ptr[0] != 0) return false;
            for (int64_t j = 0; j < n; ++j) if (ptr[j + 1] < ptr[j]) return false;
            for (int64_t p = 0; p < ptr[n]; ++p) if (idx[p] < 0 || idx[p] >= n) return false;
            return true;
        };
        if (!check_csc(Lp, Li)) return bad("Lp / Li");
        if (lu && Up && Ui && !check_csc(Up, Ui)) return bad("Up / Ui");
    }
    auto *s = new sfb200_symbolic();
    Symbolic &S = s->S;
    S.kind = kind == SFB200_LU ? Kind::LU : Kind::Cholesky;
    S.n = n;
    S.nsuper = nsuper;
    S.Super.assign(Super, Super + nsuper + 1);
    S.Lsip.assign(Lsip, Lsip + nsuper + 1);
    S.Lsxp.assign(Lsxp, Lsxp + nsuper + 1);
    S.isize = S.Lsip[nsuper];
    S.xsize = S.Lsxp[nsuper];
    S.Lsi.assign(Lsi, Lsi + S.isize);
    S.SuperMap.assign(n, 0);
    for (i64 k = 0; k < nsuper; ++k)
        for (i64 j = S.Super[k]; j < S.Super[k + 1]; ++j) S.SuperMap[j] = k;
    S.L.ptr.assign(Lp, Lp + n + 1);
    S.L.idx.assign(Li, Li + Lp[n]);
    S.nnz = Lp[n];
    S.symmetric = !(Up && Ui && S.kind == Kind::LU);
    if (!S.symmetric) {
        S.U.ptr.assign(Up, Up + n + 1);
        S.U.idx.assign(Ui, Ui + Up[n]);
    }
    *out = s;
    return SFB200_OK;
}

int64_t sfb200_symbolic_int(const sfb200_symbolic *s, const char *name) {
    if (!s || !name) return -1;
    const Symbolic &S = s->S;
    const std::string k(name);
    if (k == "n") return S.n;
    if (k == "nnz") return S.nnz;
    if (k == "unnz") return S.U.ptr.empty() ? 0 : S.U.ptr.back();
    if (k == "nsuper") return S.nsuper;
    if (k == "nfsuper") return S.nfsuper;
    if (k == "isize") return S.isize;
    if (k == "xsize") return S.xsize;
    if (k == "csize") return S.csize;
    if (k == "nstage") return S.nstage;
    if (k == "nsleaf") return S.nsleaf;
    if (k == "kind") return (int64_t)S.kind;
    if (k == "symmetric") return S.symmetric;
    if (k == "complex") return S.is_complex;
    return -1;
}

const void *sfb200_symbolic_array(const sfb200_symbolic *s, const char *name, int64_t *count) {
    if (count) *count = 0;
    if (!s || !name) return nullptr;
    const Symbolic &S = s->S;
    const std::string k(name);
    auto give = [&](const auto &v) -> const void * {
        if (count) *count = (int64_t)v.size();
        return v.empty() ? nullptr : (const void *)v.data();
    };
    if (k == "Perm") return give(S.Perm);
    if (k == "Order") return give(S.Order);
    if (k == "Parent") return give(S.Parent);
    if (k == "Post") return give(S.Post);
    if (k == "ColCount") return give(S.ColCount);
    if (k == "Lp") return give(S.L.ptr);
    if (k == "Li") return give(S.L.idx);
    if (k == "Lx") return give(S.L.val);
    if (k == "LTp") return give(S.LT.ptr);
    if (k == "LTi") return give(S.LT.idx);
    if (k == "LTx") return give(S.LT.val);
    if (k == "Up") return give(S.U.ptr);
    if (k == "Ui") return give(S.U.idx);
    if (k == "Ux") return give(S.U.val);
    if (k == "UTp") return give(S.UT.ptr);
    if (k == "UTi") return give(S.UT.idx);
    if (k == "UTx") return give(S.UT.val);
    if (k == "Super") return give(S.Super);
    if (k == "SuperMap") return give(S.SuperMap);
    if (k == "Sparent") return give(S.Sparent);
    if (k == "Lsip") return give(S.Lsip);
    if (k == "Lsxp") return give(S.Lsxp);
    if (k == "Lsi") return give(S.Lsi);
    if (k == "ST_Map") return give(S.ST_Map);
    if (k == "ST_Pointer") return give(S.ST_Pointer);
    if (k == "ST_Index") return give(S.ST_Index);
    if (k == "LeafQueue") return give(S.LeafQueue);
    if (k == "Aoffset") return give(S.Aoffset);
    if (k == "Moffset") return give(S.Moffset);
    if (k == "CPCTp") return give(S.CPCTp);
    if (k == "CPCTi") return give(S.CPCTi);
    return nullptr;
}

double sfb200_symbolic_flops(const sfb200_symbolic *s) {
    if (!s) return 0.0;
    const Symbolic &S = s->S;
    double f = 0.0;
    for (i64 k = 0; k < S.nsuper; ++k) {
        const double c = double(S.Super[k + 1] - S.Super[k]);
        const double m = double(S.Lsip[k + 1] - S.Lsip[k]) - c;
        if (S.kind == Kind::LU) f += 2.0 * c * c * c / 3.0 + 2.0 * m * c * c + 2.0 * m * m * c;
        else f += c * c * c / 3.0 + m * c * c + m * (m + 1.0) * c;
    }
    return f;
}

void sfb200_symbolic_free(sfb200_symbolic *s) { delete s; }

}  // extern "C"
