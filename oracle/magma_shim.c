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

int magma_zgetrf_nopiv ( int m, int n, cuDoubleComplex *A, int lda, int *info )
{
    (void) m; (void) n; (void) A; (void) lda;
    *info = -1;     /* complex LU is outside the scope of this oracle */
    return -1;
}
