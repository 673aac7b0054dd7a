/* Cholesky/Include/cuda_kernel.cuh -- the reference's four helper-kernel launchers (reference: Include/cuda_kernel.cuh:9-23,
 * Source/cuda_kernel.cu:22-158) as exported symbols of the drop-in library.
 *
 * The B200 engine has no use for them: relative maps are built once on the host (csrc/plan.cpp), the mapped subtract is the
 * fused scatter epilogue of the update kernels (csrc/kernels.cu) and the factor never leaves HBM in pieces.  They are kept
 * as symbols so that a program linked against the reference library still links; calling one launches nothing, says so
 * on stderr and returns. */
#ifndef INCLUDE_CUDA_KERNEL_CUH
#define INCLUDE_CUDA_KERNEL_CUH

#include "arch.h"

#if !defined(__CUDACC__) && !defined(__DRIVER_TYPES_H__)
typedef struct CUstream_st *cudaStream_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

void createMap(Long *d_Map, Long *d_Lsi, Long sip_offset, Long lds, cudaStream_t stream);
void createRelativeMap(Long *d_RelativeMap, Long *d_Map, Long *d_Lsi, Long dip_offset, Long ldd, cudaStream_t stream);
#ifdef SPARSEFRAME_LU
void mappedSubtract(int isAtomic, int isComplex, void *d_A, Long nscol, Long nsrow, Long lda, void *d_C, Long cj_offset,
                    Long ci_offset, Long nccol, Long ncrow, Long ldc, Long *d_RelativeMap, cudaStream_t stream);
#else
void mappedSubtract(int isAtomic, int isComplex, void *d_A, Long lda, void *d_C, Long cj_offset, Long ci_offset, Long nccol,
                    Long ncrow, Long ldc, Long *d_RelativeMap, cudaStream_t stream);
#endif
void deviceSum(int isComplex, void *d_A, void *d_B, void *d_C, Long nscol, Long nsrow, Long lda, cudaStream_t stream);

#ifdef __cplusplus
}
#endif

#endif
