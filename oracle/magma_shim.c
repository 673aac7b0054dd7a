/*
 * oracle/magma_shim.c -- TEST INFRASTRUCTURE.
 * MAGMA is an un-vendored, unpinned dependency of the LU reference (Config/SparseFrame.mk:37); it is absent
 * here.  The reference calls its CPU-interface routine magma_dgetrf_nopiv(m, n, A, lda, &info) on the host panel
 * (LU.c:2653): LU factorisation WITHOUT pivoting, unit-lower L and upper U overwriting A.  This restates the
 * published algorithm (right-looking, column by column).  MAGMA's own blocking differs, so results agree with a
 * real MAGMA build only to rounding: parity is UNPINNED at this one call.
 */
#include <cuComplex.h>

int magma_dgetrf_nopiv ( int m, int n, double *A, int lda, int *info )
{
    int mn = m < n ? m : n;
    *info = 0;
    for ( int k = 0; k < mn; k++ )
    {
        double *colk = A + (long) k * lda;
        double pivot = colk[k];
        if ( pivot == 0.0 )
        {
            if ( *info == 0 ) *info = k + 1;
            continue;
        }
        for ( int i = k + 1; i < m; i++ )
            colk[i] /= pivot;
        for ( int j = k + 1; j < n; j++ )
        {
            double *colj = A + (long) j * lda;
            double u = colj[k];
            if ( u != 0.0 )
                for ( int i = k + 1; i < m; i++ )
                    colj[i] -= colk[i] * u;
        }
    }
    return 0;
}

/* Complex twin (LU.c:2655): same column-by-column elimination in double-complex arithmetic (division as a multiplication
 * by the reciprocal pivot, formed with a scaled modulus; products as four real multiplies).  Parity UNPINNED at this call
 * for the same reason as above. */
int magma_zgetrf_nopiv ( int m, int n, cuDoubleComplex *A, int lda, int *info )
{
    int mn = m < n ? m : n;
    *info = 0;
    for ( int k = 0; k < mn; k++ )
    {
        cuDoubleComplex *colk = A + (long) k * lda;
        cuDoubleComplex pivot = colk[k];
        if ( pivot.x == 0.0 && pivot.y == 0.0 )
        {
            if ( *info == 0 ) *info = k + 1;
            continue;
        }
        for ( int i = k + 1; i < m; i++ )
            colk[i] = cuCdiv ( colk[i], pivot );
        for ( int j = k + 1; j < n; j++ )
        {
            cuDoubleComplex *colj = A + (long) j * lda;
            cuDoubleComplex u = colj[k];
            if ( u.x != 0.0 || u.y != 0.0 )
                for ( int i = k + 1; i < m; i++ )
                {
                    colj[i].x -= colk[i].x * u.x - colk[i].y * u.y;
                    colj[i].y -= colk[i].x * u.y + colk[i].y * u.x;
                }
        }
    }
    return 0;
}
