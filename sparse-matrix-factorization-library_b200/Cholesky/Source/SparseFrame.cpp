/* Cholesky/Source/SparseFrame.cpp -- translation unit of Cholesky/Lib/libSparseFrame.so; the implementation is shared between the
 * Cholesky and LU trees (the reference keeps two edited copies of one file). */

#include "../../dropin/sparseframe_dropin.cpp"
