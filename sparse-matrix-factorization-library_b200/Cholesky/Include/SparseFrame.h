/* Cholesky/Include/SparseFrame.h -- public header of the drop-in library (reference: Include/SparseFrame.h:1-49).  The reference
 * declares only SparseFrame(argc, argv) and exports its analyze / factorize / solve entry points undeclared; they are
 * declared here.  No BLAS / MAGMA / SuiteSparse / METIS / cuBLAS header is needed to use the library. */
#ifndef INCLUDE_SPARSEFRAME_H
#define INCLUDE_SPARSEFRAME_H

#include <stddef.h>
#include <stdio.h>
#include <omp.h>

#include "extern.h"
#include "arch.h"
#include "macro.h"
#include "type.h"
#include "constant.h"
#include "parameter.h"
#include "info.h"
#include "cuda_kernel.cuh"

#ifdef __cplusplus
extern "C" {
#endif

/* the one declared entry point (reference Include/SparseFrame.h:43) */
int SparseFrame(int argc, char **argv);

/* exported-but-undeclared entry points of the reference library, same names and signatures (CH.c:3, 16, 287, 368, 385, 526,
 * 589, 652, 1916, 3019, 3036, 3141, 3268) */
double SparseFrame_time(void);
int SparseFrame_allocate_gpu(struct common_info_struct *common_info, struct gpu_info_struct **gpu_info_list_ptr);
int SparseFrame_free_gpu(struct common_info_struct *common_info, struct gpu_info_struct **gpu_info_list_ptr);
int SparseFrame_allocate_matrix(struct common_info_struct *common_info, struct matrix_info_struct **matrix_info_list_ptr);
int SparseFrame_free_matrix(struct common_info_struct *common_info, struct matrix_info_struct **matrix_info_list_ptr);
int SparseFrame_initialize_matrix(struct matrix_info_struct *matrix_info);
int SparseFrame_read_matrix(struct matrix_info_struct *matrix_info);
int SparseFrame_compress(struct matrix_info_struct *matrix_info);
#ifdef SPARSEFRAME_LU
int SparseFrame_pivot(struct matrix_info_struct *matrix_info);      /* LU.c:589-673; called by the reader when SF_STATIC_PIVOT=1 */
#endif
int SparseFrame_analyze(struct common_info_struct *common_info, struct matrix_info_struct *matrix_info);
int SparseFrame_factorize(struct common_info_struct *common_info, struct gpu_info_struct *gpu_info_list,
                          struct matrix_info_struct *matrix_info);
int SparseFrame_solve_supernodal(struct matrix_info_struct *matrix_info);
int SparseFrame_validate(struct matrix_info_struct *matrix_info);
int SparseFrame_cleanup_matrix(struct matrix_info_struct *matrix_info);

/* addition of the B200 drop-in (not in the reference): binary triplet container, see dropin/matrix_reader.hpp */
int SparseFrame_write_binary_matrix(const char *path, int isSymmetric, Long nrow, Long ncol, Long nz, const Long *Ti,
                                    const Long *Tj, const Float *Tx);
int SparseFrame_write_binary_matrix_complex(const char *path, int isSymmetric, Long nrow, Long ncol, Long nz, const Long *Ti,
                                            const Long *Tj, const Complex *Tz);

#ifdef __cplusplus
}
#endif

#endif
