/* Shim for MAGMA (LU/Include/SparseFrame.h:15).  The LU reference calls the
 * CPU-interface routine magma_dgetrf_nopiv (LU.c:2653); oracle/magma_shim.c
 * restates it.  Cholesky uses nothing from MAGMA. */
#pragma once
#include <cuComplex.h>
int magma_dgetrf_nopiv(int m, int n, double *A, int lda, int *info);
int magma_zgetrf_nopiv(int m, int n, cuDoubleComplex *A, int lda, int *info);
