/*
 * oracle/sf_oracle.c -- TEST INFRASTRUCTURE (the CPU oracle), never linked into or called by the product.
 *
 * Plain-C restatement of the reference's NUMERIC factorization path for one CPU worker:
 *     SparseFrame_factorize_supernodal  Cholesky/Source/SparseFrame.c:2150-3017 (scheduler: 2221-2265, 2300-2331, 2960-2986)
 *     SparseFrame_cpuApplyFactorize     CH.c:2104-2148   (LU.c:2622-2666)
 *     SparseFrame_cpuApply              CH.c:2030-2102   (LU.c:2538-2620)
 *     SparseFrame_loadA                 CH.c:1998-2028   (LU.c:2478-2536)
 *     SparseFrame_solve_supernodal      CH.c:3036-3139   (LU.c:3592-3700)
 *     SparseFrame_validate              CH.c:3141-3266   (LU.c:3702-3858)
 * The BLAS/LAPACK calls of the reference (dsyrk_, dgemm_, dpotrf_, dtrsm_ -- un-vendored OpenBLAS, no pinned
 * version; magma_dgetrf_nopiv -- un-vendored MAGMA) are restated by their textbook definitions, so results agree
 * with the compiled reference (oracle/_ref/libsfref_*.so) to rounding, not bitwise.
 *
 * PINNING: tests/test_oracle.py checks this file against the compiled reference itself on every generator
 * (entrywise <= 1e-11 on the defined entries, identical zero pattern) and against the committed golden vectors
 * in tests/golden/ that were dumped from the compiled reference (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long Long;

/* ---- dense kernels, column-major ----------------------------------------------------------------------------- */

/* C(0:n,0:n) lower = A A^T, A is n x k (dsyrk_ "L","N", beta = 0; CH.c:2061) */
static void syrk_lower ( Long n, Long k, const double *A, Long lda, double *C, Long ldc )
{
    for ( Long j = 0; j < n; j++ )
        for ( Long i = j; i < n; i++ )
        {
            double s = 0;
            for ( Long p = 0; p < k; p++ ) s += A[i + p * lda] * A[j + p * lda];
            C[i + j * ldc] = s;
        }
}

/* C(m x n) = A B^T, A m x k, B n x k (dgemm_ "N","C"/"T", beta = 0; CH.c:2068, LU.c:2570-2577) */
static void gemm_nt ( Long m, Long n, Long k, const double *A, Long lda, const double *B, Long ldb, double *C, Long ldc )
{
    for ( Long j = 0; j < n; j++ )
        for ( Long i = 0; i < m; i++ )
        {
            double s = 0;
            for ( Long p = 0; p < k; p++ ) s += A[i + p * lda] * B[j + p * ldb];
            C[i + j * ldc] = s;
        }
}

/* dpotrf_ "L" (CH.c:2135): returns LAPACK-style info */
static int potrf_lower ( Long n, double *A, Long lda )
{
    for ( Long j = 0; j < n; j++ )
    {
        double d = A[j + j * lda];
        for ( Long p = 0; p < j; p++ ) d -= A[j + p * lda] * A[j + p * lda];
        if ( !( d > 0 ) ) return (int) ( j + 1 );
        d = sqrt ( d );
        A[j + j * lda] = d;
        for ( Long i = j + 1; i < n; i++ )
        {
            double s = A[i + j * lda];
            for ( Long p = 0; p < j; p++ ) s -= A[i + p * lda] * A[j + p * lda];
            A[i + j * lda] = s / d;
        }
    }
    return 0;
}

/* dtrsm_ "R","L","C"/"T",diag (CH.c:2142; LU.c:2660 with diag = "U"): B(m x n) := B * L^{-T} */
static void trsm_rlt ( Long m, Long n, int unit, const double *L, Long ldl, double *B, Long ldb )
{
    for ( Long j = 0; j < n; j++ )
        for ( Long i = 0; i < m; i++ )
        {
            double s = B[i + j * ldb];
            for ( Long p = 0; p < j; p++ ) s -= B[i + p * ldb] * L[j + p * ldl];
            B[i + j * ldb] = unit ? s : s / L[j + j * ldl];
        }
}

/* magma_dgetrf_nopiv (LU.c:2653): no-pivot LU of an m x n panel */
static int getrf_nopiv ( Long m, Long n, double *A, Long lda )
{
    int info = 0;
    Long mn = m < n ? m : n;
    for ( Long k = 0; k < mn; k++ )
    {
        double piv = A[k + k * lda];
        if ( piv == 0 ) { if ( !info ) info = (int) ( k + 1 ); continue; }
        for ( Long i = k + 1; i < m; i++ ) A[i + k * lda] /= piv;
        for ( Long j = k + 1; j < n; j++ )
        {
            double u = A[k + j * lda];
            for ( Long i = k + 1; i < m; i++ ) A[i + j * lda] -= A[i + k * lda] * u;
        }
    }
    return info;
}

/* ---- assembly (loadA) ---------------------------------------------------------------------------------------- */
static void load_front ( int lu, const Long *Lp, const Long *Li, const double *Lx, const Long *Up, const Long *Ui,
        const double *Ux, const Long *Super, const Long *Map, Long s, Long nscol, Long nsrow, Long slda, double *A )
{
    memset ( A, 0, (size_t) ( nscol * slda ) * sizeof(double) );
    for ( Long j = Super[s]; j < Super[s+1]; j++ )
    {
        Long sj = j - Super[s];
        for ( Long p = Lp[j]; p < Lp[j+1]; p++ )
        {
            Long i = Li[p];
            if ( !lu || i > j ) A[sj * slda + Map[i]] = Lx[p];
        }
        if ( lu )
            for ( Long p = Up[j]; p < Up[j+1]; p++ )
            {
                Long si = Map[Ui[p]];
                if ( si < nscol ) A[si * slda + sj] = Ux[p];
                else A[( nsrow - nscol ) + sj * slda + si] = Ux[p];
            }
    }
}

/* ---- one descendant update (cpuApply) ------------------------------------------------------------------------ */
static void apply_descendant ( int lu, const Long *SuperMap, const Long *Super, const Long *Lsip, const Long *Lsi,
        const Long *Lsxp, double *Lsx, Long *Head, Long *Next, Long *Lpos, const Long *Map, Long *Rel,
        Long s, Long nscol, Long nsrow, double *A, Long d, double *C )
{
    Long slda = lu ? 2 * nsrow - nscol : nsrow;
    Long ndcol = Super[d+1] - Super[d], ndrow = Lsip[d+1] - Lsip[d];
    Long lpos = Lpos[d], lpos_next;
    for ( lpos_next = lpos; lpos_next < ndrow && Lsi[Lsip[d] + lpos_next] < Super[s+1]; lpos_next++ );
    Long dn = lpos_next - lpos, dm = ndrow - lpos - dn, dk = ndcol;
    Long dlda = lu ? 2 * ndrow - ndcol : ndrow;
    Long dldc = lu ? dn + 2 * dm : ndrow - lpos;
    double *D = Lsx + Lsxp[d];

    for ( Long di = 0; di < ndrow - lpos; di++ ) Rel[di] = Map[Lsi[Lsip[d] + lpos + di]];

    if ( !lu )
    {
        syrk_lower ( dn, dk, D + lpos, dlda, C, dldc );
        if ( dm > 0 ) gemm_nt ( dm, dn, dk, D + lpos_next, dlda, D + lpos, dlda, C + dn, dldc );
        for ( Long cj = 0; cj < dn; cj++ )
            for ( Long ci = cj; ci < dn + dm; ci++ )
                A[Rel[cj] * slda + Rel[ci]] -= C[cj * dldc + ci];
    }
    else
    {
        double *UT = D + ( ndrow - ndcol );              /* row r of U12^T lives at panel row (ndrow - ndcol) + r */
        gemm_nt ( dn + dm, dn, dk, D + lpos, dlda, UT + lpos, dlda, C, dldc );
        if ( dm > 0 ) gemm_nt ( dm, dn, dk, UT + lpos_next, dlda, D + lpos, dlda, C + ( dn + dm ), dldc );
        for ( Long cj = 0; cj < dn; cj++ )
            for ( Long ci = 0; ci < dn + dm; ci++ )
            {
                A[Rel[cj] * slda + Rel[ci]] -= C[cj * dldc + ci];
                if ( ci >= dn )
                    A[( nsrow - nscol ) + Rel[cj] * slda + Rel[ci]] -= C[dm + cj * dldc + ci];
            }
    }

    if ( lpos_next < ndrow )
    {
        Long anc = SuperMap[Lsi[Lsip[d] + lpos_next]];
        Next[d] = Head[anc];
        Head[anc] = d;
    }
    Lpos[d] = lpos_next;
}

/*
 * Numeric factorization, one worker.  kind: 0 Cholesky, 1 LU.  Up/Ui/Ux may be NULL for a symmetric LU input
 * (the reference aliases them to Lp/Li/Lx, LU.c:2718-2729).  LeafQueue (nsuper entries, first nsleaf valid) is the
 * initial ready queue from the analysis; it is copied, not modified.  Returns the number of failed pivots.
 */
long sforacle_factorize ( int kind, Long nrow, Long nsuper, const Long *Super, const Long *SuperMap,
        const Long *Lsip, const Long *Lsxp, const Long *Lsi, Long csize, Long nsleaf, const Long *LeafQueue,
        const Long *Lp, const Long *Li, const double *Lx, const Long *Up, const Long *Ui, const double *Ux,
        double *Lsx )
{
    int lu = kind != 0;
    long bad = 0;
    if ( lu && !Up ) { Up = Lp; Ui = Li; Ux = Lx; }

    Long *Lpos = calloc ( (size_t) nsuper + 1, sizeof(Long) );
    Long *Head = malloc ( ( (size_t) nsuper + 1 ) * sizeof(Long) );
    Long *Next = malloc ( ( (size_t) nsuper + 1 ) * sizeof(Long) );
    Long *Nschild = calloc ( (size_t) nsuper + 1, sizeof(Long) );
    Long *Queue = malloc ( ( (size_t) nsuper + 1 ) * sizeof(Long) );
    Long *Map = malloc ( ( (size_t) nrow + 1 ) * sizeof(Long) );
    Long *Rel = malloc ( ( (size_t) nrow + 1 ) * sizeof(Long) );
    double *C = malloc ( ( (size_t) csize + 1 ) * sizeof(double) );

    for ( Long s = 0; s < nsuper; s++ ) { Head[s] = -1; Next[s] = -1; }
    for ( Long s = 0; s < nsuper; s++ )
    {
        Long nscol = Super[s+1] - Super[s], nsrow = Lsip[s+1] - Lsip[s];
        if ( nscol < nsrow ) Nschild[SuperMap[Lsi[Lsip[s] + nscol]]]++;
    }
    memcpy ( Queue, LeafQueue, (size_t) nsuper * sizeof(Long) );
    Long qhead = 0, qtail = nsleaf;

    while ( qhead < qtail )
    {
        Long s = Queue[qhead++];
        Long nscol = Super[s+1] - Super[s], nsrow = Lsip[s+1] - Lsip[s];
        Long sn = nscol, sm = nsrow - nscol, slda = lu ? 2 * nsrow - nscol : nsrow;
        double *A = Lsx + Lsxp[s];

        for ( Long si = 0; si < nsrow; si++ ) Map[Lsi[Lsip[s] + si]] = si;
        load_front ( lu, Lp, Li, Lx, Up, Ui, Ux, Super, Map, s, nscol, nsrow, slda, A );
        while ( Head[s] >= 0 )
        {
            Long d = Head[s];
            Head[s] = Next[d];
            apply_descendant ( lu, SuperMap, Super, Lsip, Lsi, Lsxp, Lsx, Head, Next, Lpos, Map, Rel, s, nscol, nsrow, A, d, C );
        }
        if ( !lu )
        {
            if ( potrf_lower ( sn, A, slda ) ) bad++;
            if ( sm > 0 ) trsm_rlt ( sm, sn, 0, A, slda, A + sn, slda );
        }
        else
        {
            if ( getrf_nopiv ( sn + sm, sn, A, slda ) ) bad++;
            if ( sm > 0 ) trsm_rlt ( sm, sn, 1, A, slda, A + ( sn + sm ), slda );
        }

        Lpos[s] = nscol;
        if ( nscol < nsrow )
        {
            Long sparent = SuperMap[Lsi[Lsip[s] + nscol]];
            Next[s] = Head[sparent];
            Head[sparent] = s;
            if ( --Nschild[sparent] <= 0 ) Queue[qtail++] = sparent;
        }
    }

    free ( Lpos ); free ( Head ); free ( Next ); free ( Nschild ); free ( Queue ); free ( Map ); free ( Rel ); free ( C );
    return bad;
}

/* ---- solve: x := A^{-1} b through the supernodal factor (scalar loops, as the reference) ------------------------ */
void sforacle_solve ( int kind, Long nrow, Long nsuper, const Long *Super, const Long *Lsip, const Long *Lsxp,
        const Long *Lsi, const double *Lsx, const double *b, double *x )
{
    int lu = kind != 0;
    memcpy ( x, b, (size_t) nrow * sizeof(double) );
    for ( Long s = 0; s < nsuper; s++ )
    {
        Long nscol = Super[s+1] - Super[s], nsrow = Lsip[s+1] - Lsip[s];
        Long slda = lu ? 2 * nsrow - nscol : nsrow;
        for ( Long sj = 0; sj < nscol; sj++ )
        {
            Long j = Lsi[Lsip[s] + sj];
            if ( !lu ) x[j] /= Lsx[Lsxp[s] + sj * slda + sj];
            for ( Long si = sj + 1; si < nsrow; si++ )
                x[Lsi[Lsip[s] + si]] -= Lsx[Lsxp[s] + sj * slda + si] * x[j];
        }
    }
    for ( Long s = nsuper - 1; s >= 0; s-- )
    {
        Long nscol = Super[s+1] - Super[s], nsrow = Lsip[s+1] - Lsip[s];
        Long slda = lu ? 2 * nsrow - nscol : nsrow;
        for ( Long sj = nscol - 1; sj >= 0; sj-- )
        {
            Long j = Lsi[Lsip[s] + sj];
            if ( !lu )
            {
                for ( Long si = sj + 1; si < nsrow; si++ )
                    x[j] -= Lsx[Lsxp[s] + sj * slda + si] * x[Lsi[Lsip[s] + si]];
            }
            else
            {
                /* row sj of U: inside the diagonal block it is stored along the row, beyond it in U12^T */
                for ( Long si = sj + 1; si < nscol; si++ )
                    x[j] -= Lsx[Lsxp[s] + si * slda + sj] * x[Lsi[Lsip[s] + si]];
                for ( Long si = nscol; si < nsrow; si++ )
                    x[j] -= Lsx[Lsxp[s] + sj * slda + ( nsrow - nscol ) + si] * x[Lsi[Lsip[s] + si]];
            }
            x[j] /= Lsx[Lsxp[s] + sj * slda + sj];
        }
    }
}

/* ---- residual of the permuted system, the reference's self-check (CH.c:3197-3263) ------------------------------- */
/* out[0] = ||A||_1, out[1] = ||b||_inf, out[2] = ||x||_inf, out[3] = ||Ax-b||_inf; returns rnorm/(anorm*xnorm+bnorm) */
double sforacle_residual ( int kind, Long nrow, const Long *Lp, const Long *Li, const double *Lx, const Long *Up,
        const Long *Ui, const double *Ux, const double *b, const double *x, double *out )
{
    int split = ( kind != 0 ) && Up != NULL;
    double *r = malloc ( (size_t) ( nrow > 0 ? nrow : 1 ) * sizeof(double) );
    double *w = calloc ( (size_t) ( nrow > 0 ? nrow : 1 ), sizeof(double) );
    for ( Long i = 0; i < nrow; i++ ) r[i] = -b[i];
    for ( Long j = 0; j < nrow; j++ )
        for ( Long p = Lp[j]; p < Lp[j+1]; p++ )
        {
            Long i = Li[p];
            if ( !split )
            {
                r[i] += Lx[p] * x[j];
                if ( i != j ) r[j] += Lx[p] * x[i];
                w[j] += fabs ( Lx[p] );
                if ( i != j ) w[i] += fabs ( Lx[p] );
            }
            else if ( i > j )
            {
                r[i] += Lx[p] * x[j];
                w[j] += fabs ( Lx[p] );
            }
        }
    if ( split )
        for ( Long i = 0; i < nrow; i++ )
            for ( Long p = Up[i]; p < Up[i+1]; p++ )
            {
                Long j = Ui[p];
                r[i] += Ux[p] * x[j];
                w[j] += fabs ( Ux[p] );
            }
    double anorm = 0, bnorm = 0, xnorm = 0, rnorm = 0;
    for ( Long i = 0; i < nrow; i++ )
    {
        if ( w[i] > anorm ) anorm = w[i];
        if ( fabs ( b[i] ) > bnorm ) bnorm = fabs ( b[i] );
        if ( fabs ( x[i] ) > xnorm ) xnorm = fabs ( x[i] );
        if ( fabs ( r[i] ) > rnorm ) rnorm = fabs ( r[i] );
    }
    if ( out ) { out[0] = anorm; out[1] = bnorm; out[2] = xnorm; out[3] = rnorm; }
    free ( r ); free ( w );
    return rnorm / ( anorm * xnorm + bnorm );
}
