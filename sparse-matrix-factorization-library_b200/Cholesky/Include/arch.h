/* Cholesky/Include/arch.h -- element types of the library (reference: Include/arch.h:6-16).
 * Shared by the Cholesky and LU trees (LU/Include/arch.h selects the LU flavour and includes this file). */
#ifndef INCLUDE_ARCH_H
#define INCLUDE_ARCH_H

#include <stddef.h>

#define Int int
#define uInt unsigned int
#define Long long
#define uLong unsigned long
#define Float double

/* complex values are interleaved (re, im) pairs of Float, the layout of cuDoubleComplex / LAPACK COMPLEX*16 */
typedef struct { Float x, y; } SparseFrameComplex;
#define Complex SparseFrameComplex

#endif
