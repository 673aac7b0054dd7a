/* Cholesky/Include/macro.h -- helper macros (reference: Include/macro.h). */
#ifndef INCLUDE_MACRO_H
#define INCLUDE_MACRO_H

#ifndef TRUE
#define FALSE 0
#define TRUE 1
#endif

#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif

#endif
