/* Cholesky/Include/extern.h -- Fortran BLAS / LAPACK prototypes the reference's CPU path calls (reference:
 * Include/extern.h:6-13; LU adds getrf).  The B200 engine calls none of them -- its dense kernels are its own (csrc/kernels.cu);
 * the prototypes are kept so that code written against the reference's headers still compiles. */
#ifndef INCLUDE_EXTERN_H
#define INCLUDE_EXTERN_H

#include "arch.h"

#ifdef __cplusplus
extern "C" {
#endif
extern int dsyrk_(char *uplo, char *trans, Long *n, Long *k, const Float *alpha, Float *a, Long *lda, const Float *beta, Float *c, Long *ldc);
extern int dgemm_(char *ta, char *tb, Long *m, Long *n, Long *k, const Float *alpha, Float *a, Long *lda, Float *b, Long *ldb, const Float *beta, Float *c, Long *ldc);
extern int dpotrf_(char *uplo, Long *n, Float *a, Long *lda, int *info);
extern int dtrsm_(char *side, char *uplo, char *ta, char *diag, Long *m, Long *n, const Float *alpha, Float *a, Long *lda, Float *b, Long *ldb);
extern int zherk_(char *uplo, char *trans, Long *n, Long *k, const Complex *alpha, Complex *a, Long *lda, const Complex *beta, Complex *c, Long *ldc);
extern int zgemm_(char *ta, char *tb, Long *m, Long *n, Long *k, const Complex *alpha, Complex *a, Long *lda, Complex *b, Long *ldb, const Complex *beta, Complex *c, Long *ldc);
extern int zpotrf_(char *uplo, Long *n, Complex *a, Long *lda, int *info);
extern int ztrsm_(char *side, char *uplo, char *ta, char *diag, Long *m, Long *n, const Complex *alpha, Complex *a, Long *lda, Complex *b, Long *ldb);
#ifdef SPARSEFRAME_LU
extern int dgetrf_(Long *m, Long *n, Float *a, Long *lda, Long *ipiv, int *info);
extern int zgetrf_(Long *m, Long *n, Complex *a, Long *lda, Long *ipiv, int *info);
#endif
#ifdef __cplusplus
}
#endif

#endif
