/* LU/Include/SparseFrame.h -- public header of the drop-in LU library (reference: LU/Include/SparseFrame.h:43 declares
 * only SparseFrame(argc, argv); the analyze / factorize / solve entry points the reference exports without declaring
 * are declared here too).  Types and prototypes live in dropin/sparseframe_abi.h. */
#ifndef INCLUDE_SPARSEFRAME_H
#define INCLUDE_SPARSEFRAME_H
#define SPARSEFRAME_LU 1
#include "../../dropin/sparseframe_abi.h"
#endif
