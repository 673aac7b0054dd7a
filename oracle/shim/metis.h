/* Shim for METIS: the CUDA-12.9 bundled libmetis_static.a is built with
 * 64-bit idx_t, matching the reference's sizeof(idx_t)==sizeof(Long) branch
 * (CH.c:888-891). */
#pragma once
#include <stdint.h>
typedef int64_t idx_t;
#ifdef __cplusplus
extern "C"
#endif
int METIS_NodeND(idx_t *nvtxs, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *options, idx_t *perm, idx_t *iperm);
