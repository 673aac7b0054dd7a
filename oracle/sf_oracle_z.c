/*
 * oracle/sf_oracle_z.c -- TEST INFRASTRUCTURE (the CPU oracle for COMPLEX matrices), never linked into or called by the product.
 *
 * Plain-C restatement of the isComplex branches of the reference's numeric factorization path for one CPU worker; the real
 * branches are restated in sf_oracle.c, and the control flow (scheduler, descendant lists, relative maps) is the same:
 *     SparseFrame_loadA              CH.c:1998-2028 (complex assignment 2019-2023)      LU.c:2478-2536
 *     SparseFrame_cpuApply           CH.c:2030-2102: zherk_ "L","N" 2063, zgemm_ "N","C" 2070, complex scatter 2080-2083
 *                                    LU.c:2538-2620: zgemm_ "N","T" 2572 and 2579
 *     SparseFrame_cpuApplyFactorize  CH.c:2104-2148: zpotrf_ "L" 2137, ztrsm_ "R","L","C","N" 2144
 *                                    LU.c:2622-2666: magma_zgetrf_nopiv 2655, ztrsm_ "R","L","T","U" 2662
 * A complex "symmetric" matrix means two different things there: the Cholesky path reads the permuted lower triangle as the
 * lower triangle of a HERMITIAN matrix (zherk / zpotrf never look at anything else, the imaginary part of the diagonal
 * included), the LU path reads it as A = A^T (U aliased to L, LU.c:2718-2729).
 * The reference's complex solve and validate are `// TODO` (CH.c:3088-3131, 3229-3256; LU.c:3644-3690): sforacle_zsolve /
 * sforacle_zresidual are what those loops compute once the TODOs are filled in (complex arithmetic, moduli in the norms).
 * BLAS / LAPACK / MAGMA calls are restated by their textbook definitions: agreement with the compiled reference
 * (oracle/_ref/libsfref_*.so) is to rounding, not bitwise.
 *
 * PINNING: tests/test_complex.py checks this file against the golden vectors in tests/golden/complex/ (dumped from the
 * compiled reference with isComplex set, tests/golden/make_golden.py) and against the compiled reference itself.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long Long;
typedef double _Complex Cplx;

/* ---- dense kernels, column-major ----------------------------------------------------------------------------- */

/* zherk_ "L","N", beta = 0 (CH.c:2063): C(0:n,0:n) lower = A A^H, real diagonal */
static void herk_lower ( Long n, Long k, const Cplx *A, Long lda, Cplx *C, Long ldc )
{
    for ( Long j = 0; j < n; j++ )
        for ( Long i = j; i < n; i++ )
        {
            Cplx s = 0;
            for ( Long p = 0; p < k; p++ ) s += A[i + p * lda] * conj ( A[j + p * lda] );
            C[i + j * ldc] = ( i == j ) ? creal ( s ) : s;
        }
}

/* zgemm_ "N","C" (CH.c:2070: conj = 1) or "N","T" (LU.c:2572, 2579: conj = 0), beta = 0: C(m x n) = A op(B), B is n x k */
static void gemm_n_op ( int conj_b, Long m, Long n, Long k, const Cplx *A, Long lda, const Cplx *B, Long ldb, Cplx *C, Long ldc )
{
    for ( Long j = 0; j < n; j++ )
        for ( Long i = 0; i < m; i++ )
        {
            Cplx s = 0;
            for ( Long p = 0; p < k; p++ ) s += A[i + p * lda] * ( conj_b ? conj ( B[j + p * ldb] ) : B[j + p * ldb] );
            C[i + j * ldc] = s;
        }
}

/* zpotrf_ "L" (CH.c:2137): only the real part of the diagonal is read, the factor's diagonal is real */
static int potrf_lower ( Long n, Cplx *A, Long lda )
{
    for ( Long j = 0; j < n; j++ )
    {
        double d = creal ( A[j + j * lda] );
        for ( Long p = 0; p < j; p++ ) d -= creal ( A[j + p * lda] * conj ( A[j + p * lda] ) );
        if ( !( d > 0 ) ) return (int) ( j + 1 );
        d = sqrt ( d );
        A[j + j * lda] = d;
        for ( Long i = j + 1; i < n; i++ )
        {
            Cplx s = A[i + j * lda];
            for ( Long p = 0; p < j; p++ ) s -= A[i + p * lda] * conj ( A[j + p * lda] );
            A[i + j * lda] = s / d;
        }
    }
    return 0;
}

/* ztrsm_ "R","L","C","N" (CH.c:2144: B := B L^{-H}, herm = 1, unit = 0) or "R","L","T","U" (LU.c:2662: B := B L^{-T}, unit) */
static void trsm_right_lower ( int herm, int unit, Long m, Long n, const Cplx *L, Long ldl, Cplx *B, Long ldb )
{
    for ( Long j = 0; j < n; j++ )
        for ( Long i = 0; i < m; i++ )
        {
            Cplx s = B[i + j * ldb];
            for ( Long p = 0; p < j; p++ ) s -= B[i + p * ldb] * ( herm ? conj ( L[j + p * ldl] ) : L[j + p * ldl] );
            B[i + j * ldb] = unit ? s : s / ( herm ? conj ( L[j + j * ldl] ) : L[j + j * ldl] );
        }
}

/* magma_zgetrf_nopiv (LU.c:2655): LU without pivoting of an m x n panel, unit lower L and upper U in place */
static int getrf_nopiv ( Long m, Long n, Cplx *A, Long lda )
{
    int info = 0;
    Long mn = m < n ? m : n;
    for ( Long k = 0; k < mn; k++ )
    {
        Cplx piv = A[k + k * lda];
        if ( piv == 0 ) { if ( !info ) info = (int) ( k + 1 ); continue; }
        for ( Long i = k + 1; i < m; i++ ) A[i + k * lda] /= piv;
        for ( Long j = k + 1; j < n; j++ )
        {
            Cplx u = A[k + j * lda];
            for ( Long i = k + 1; i < m; i++ ) A[i + j * lda] -= A[i + k * lda] * u;
        }
    }
    return info;
}

/* ---- assembly (loadA, CH.c:2004-2025 / LU.c:2484-2533): assignment, so the last copy of a duplicate entry stays ------ */
static void load_front ( int lu, const Long *Lp, const Long *Li, const Cplx *Lx, const Long *Up, const Long *Ui,
        const Cplx *Ux, const Long *Super, const Long *Map, Long s, Long nscol, Long nsrow, Long slda, Cplx *A )
{
    memset ( A, 0, (size_t) ( nscol * slda ) * sizeof(Cplx) );
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
        const Long *Lsxp, Cplx *Lsx, Long *Head, Long *Next, Long *Lpos, const Long *Map, Long *Rel,
        Long s, Long nscol, Long nsrow, Cplx *A, Long d, Cplx *C )
{
    Long slda = lu ? 2 * nsrow - nscol : nsrow;
    Long ndcol = Super[d+1] - Super[d], ndrow = Lsip[d+1] - Lsip[d];
    Long lpos = Lpos[d], lpos_next;
    for ( lpos_next = lpos; lpos_next < ndrow && Lsi[Lsip[d] + lpos_next] < Super[s+1]; lpos_next++ );
    Long dn = lpos_next - lpos, dm = ndrow - lpos - dn, dk = ndcol;
    Long dlda = lu ? 2 * ndrow - ndcol : ndrow;
    Long dldc = lu ? dn + 2 * dm : ndrow - lpos;
    Cplx *D = Lsx + Lsxp[d];

    for ( Long di = 0; di < ndrow - lpos; di++ ) Rel[di] = Map[Lsi[Lsip[d] + lpos + di]];

    if ( !lu )
    {
        herk_lower ( dn, dk, D + lpos, dlda, C, dldc );
        if ( dm > 0 ) gemm_n_op ( 1, dm, dn, dk, D + lpos_next, dlda, D + lpos, dlda, C + dn, dldc );
        for ( Long cj = 0; cj < dn; cj++ )
            for ( Long ci = cj; ci < dn + dm; ci++ )
                A[Rel[cj] * slda + Rel[ci]] -= C[cj * dldc + ci];
    }
    else
    {
        Cplx *UT = D + ( ndrow - ndcol );                /* row r of U12^T lives at panel row (ndrow - ndcol) + r */
        gemm_n_op ( 0, dn + dm, dn, dk, D + lpos, dlda, UT + lpos, dlda, C, dldc );
        if ( dm > 0 ) gemm_n_op ( 0, dm, dn, dk, UT + lpos_next, dlda, D + lpos, dlda, C + ( dn + dm ), dldc );
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
 * Numeric factorization, one worker (scheduler: CH.c:2221-2265, 2300-2331, 2960-2986).  kind: 0 Cholesky, 1 LU.  All value
 * arrays are interleaved (re, im) pairs = C99 double _Complex.  Up/Ui/Ux NULL: symmetric LU input (U aliased to L).
 * Returns the number of failed pivots.
 */
long sforacle_zfactorize ( int kind, Long nrow, Long nsuper, const Long *Super, const Long *SuperMap,
        const Long *Lsip, const Long *Lsxp, const Long *Lsi, Long csize, Long nsleaf, const Long *LeafQueue,
        const Long *Lp, const Long *Li, const Cplx *Lx, const Long *Up, const Long *Ui, const Cplx *Ux,
        Cplx *Lsx )
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
    Cplx *C = malloc ( ( (size_t) csize + 1 ) * sizeof(Cplx) );

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
        Cplx *A = Lsx + Lsxp[s];

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
            if ( sm > 0 ) trsm_right_lower ( 1, 0, sm, sn, A, slda, A + sn, slda );
        }
        else
        {
            if ( getrf_nopiv ( sn + sm, sn, A, slda ) ) bad++;
            if ( sm > 0 ) trsm_right_lower ( 0, 1, sm, sn, A, slda, A + ( sn + sm ), slda );
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

/* ---- solve: the loops of SparseFrame_solve_supernodal with the `// TODO` statements filled in (CH.c:3074-3134) ------ */
void sforacle_zsolve ( int kind, Long nrow, Long nsuper, const Long *Super, const Long *Lsip, const Long *Lsxp,
        const Long *Lsi, const Cplx *Lsx, const Cplx *b, Cplx *x )
{
    int lu = kind != 0;
    memcpy ( x, b, (size_t) nrow * sizeof(Cplx) );
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
                for ( Long si = sj + 1; si < nsrow; si++ )           /* L^H x = y */
                    x[j] -= conj ( Lsx[Lsxp[s] + sj * slda + si] ) * x[Lsi[Lsip[s] + si]];
            }
            else
            {
                for ( Long si = sj + 1; si < nscol; si++ )
                    x[j] -= Lsx[Lsxp[s] + si * slda + sj] * x[Lsi[Lsip[s] + si]];
                for ( Long si = nscol; si < nsrow; si++ )
                    x[j] -= Lsx[Lsxp[s] + sj * slda + ( nsrow - nscol ) + si] * x[Lsi[Lsip[s] + si]];
            }
            x[j] /= Lsx[Lsxp[s] + sj * slda + sj];
        }
    }
}

/* ---- residual of the permuted system (CH.c:3197-3263 with the TODO norms as complex moduli) ----------------------- */
/* out[0] = ||A||_1, out[1] = max|b_i|, out[2] = max|x_i|, out[3] = max|(Ax-b)_i|; returns rnorm/(anorm*xnorm+bnorm) */
double sforacle_zresidual ( int kind, Long nrow, const Long *Lp, const Long *Li, const Cplx *Lx, const Long *Up,
        const Long *Ui, const Cplx *Ux, const Cplx *b, const Cplx *x, double *out )
{
    int lu = kind != 0;
    int split = lu && Up != NULL;
    Cplx *r = malloc ( (size_t) ( nrow > 0 ? nrow : 1 ) * sizeof(Cplx) );
    double *w = calloc ( (size_t) ( nrow > 0 ? nrow : 1 ), sizeof(double) );
    for ( Long i = 0; i < nrow; i++ ) r[i] = -b[i];
    for ( Long j = 0; j < nrow; j++ )
        for ( Long p = Lp[j]; p < Lp[j+1]; p++ )
        {
            Long i = Li[p];
            Cplx v = ( !lu && i == j ) ? creal ( Lx[p] ) : Lx[p];        /* Hermitian: a real diagonal */
            if ( !split )
            {
                r[i] += v * x[j];
                w[j] += cabs ( v );
                if ( i != j ) { r[j] += ( lu ? v : conj ( v ) ) * x[i]; w[i] += cabs ( v ); }
            }
            else if ( i > j )
            {
                r[i] += v * x[j];
                w[j] += cabs ( v );
            }
        }
    if ( split )
        for ( Long i = 0; i < nrow; i++ )
            for ( Long p = Up[i]; p < Up[i+1]; p++ )
            {
                Long j = Ui[p];
                r[i] += Ux[p] * x[j];
                w[j] += cabs ( Ux[p] );
            }
    double anorm = 0, bnorm = 0, xnorm = 0, rnorm = 0;
    for ( Long i = 0; i < nrow; i++ )
    {
        if ( w[i] > anorm ) anorm = w[i];
        if ( cabs ( b[i] ) > bnorm ) bnorm = cabs ( b[i] );
        if ( cabs ( x[i] ) > xnorm ) xnorm = cabs ( x[i] );
        if ( cabs ( r[i] ) > rnorm ) rnorm = cabs ( r[i] );
    }
    if ( out ) { out[0] = anorm; out[1] = bnorm; out[2] = xnorm; out[3] = rnorm; }
    free ( r ); free ( w );
    return rnorm / ( anorm * xnorm + bnorm );
}
