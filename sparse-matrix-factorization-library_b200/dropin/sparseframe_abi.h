/* dropin/sparseframe_abi.h -- the data contract of the SparseFrame library as the drop-in sees it: the header tree of
 * Cholesky/Include (LU flavour when SPARSEFRAME_LU is defined before inclusion). */
#ifndef SPARSEFRAME_ABI_H
#define SPARSEFRAME_ABI_H
#include "../Cholesky/Include/SparseFrame.h"
#endif
