/* LU/Source/SparseFrame.cpp -- translation unit of LU/Lib/libSparseFrame.so; the implementation is shared between the
 * Cholesky and LU trees (the reference keeps two edited copies of one file). */
#define SPARSEFRAME_LU 1
#include "../../dropin/sparseframe_dropin.cpp"
