/* Shim for the un-vendored OpenBLAS header the reference includes
 * (Cholesky/Include/SparseFrame.h:14).  Only the one call it makes. */
#pragma once
void openblas_set_num_threads(int);
