/* Cholesky/Include/info.h -- the three structs the library's entry points exchange (reference: Include/info.h:12-150; the LU
 * tree's variant has ten more members, selected here by SPARSEFRAME_LU).  A drop-in library must keep member ORDER and
 * TYPES -- that is the ABI; tests/test_dropin.py checks sizeof / offsetof against numbers measured on the reference
 * headers.  Members are grouped and commented by what the B200 engine does with them. */
#ifndef INCLUDE_INFO_H
#define INCLUDE_INFO_H

#include <stddef.h>
#include <stdio.h>
#include <omp.h>

#include "arch.h"
#include "type.h"
#include "parameter.h"

/* slot multiplicities of the reference's device arena (parameter.h): the arrays below are sized by them */
#define SF_MAX_D_STREAM MAX_D_STREAM
#define SF_A_MULTIPLE A_MULTIPLE
#define SF_B_MULTIPLE B_MULTIPLE
#define SF_C_MULTIPLE C_MULTIPLE

/* CUDA / cuBLAS / cuSOLVER object handles are all pointers; the drop-in never dereferences them */
typedef void *sf_handle_t;

struct common_info_struct
{
    int numCPU;                 /* 0 here: the engine has no CPU worker */
    int numGPU;                 /* devices the engine will use (no "GPU splitting") */
    int numGPU_physical;
    size_t minDevMemSize;
    size_t minHostMemSize;
    int matrixThreadNum;
    int numSparseMatrix;
    size_t devSlotSize;         /* READ by analysis: caps supernodes / drives relaxed merging and stages */
    double allocateTime;
    double computeTime;
    double freeTime;
};

struct gpu_info_struct
{
    int gpuIndex_physical;
    omp_lock_t gpuLock;
    /* the reference's arenas, slots, events, streams and library handles: left NULL by the drop-in, which keeps its
     * device state inside the engine (one sfb200_plan per analysed matrix) */
    void *devMem;
    size_t devMemSize;
    void *hostMem;
    size_t hostMemSize;
    size_t sharedMemSize;
    void *h_A[SF_A_MULTIPLE];
    void *h_B[SF_B_MULTIPLE];
    void *h_Lsi;
    void *d_A[SF_A_MULTIPLE + 1];
    void *d_B[SF_B_MULTIPLE];
    void *d_C[SF_C_MULTIPLE];
    void *d_Lsi;
    sf_handle_t s_cudaEvent_factorized;
    sf_handle_t d_cudaEvent_onDevice[SF_MAX_D_STREAM];
    sf_handle_t d_cudaEvent_applied[SF_MAX_D_STREAM];
    sf_handle_t s_cudaStream;
    sf_handle_t s_cudaStream_copyback;
    sf_handle_t d_cudaStream[SF_MAX_D_STREAM];
    sf_handle_t d_cudaStream_copy[SF_MAX_D_STREAM];
    sf_handle_t s_cublasHandle;
    sf_handle_t d_cublasHandle[SF_MAX_D_STREAM];
    sf_handle_t s_cusolverDnHandle;
    int lastMatrix;
};

struct matrix_info_struct
{
    int serial;
    const char *path;
    FILE *file;
    enum FactorizeType factorizeType;
    int isSymmetric;
    int isComplex;              /* only real matrices are supported by the drop-in */

    Long ncol, nrow, nzmax;

    Long *Tj, *Ti;              /* triplets (freed by compress) */
    Float *Tx;

    Long *Cp, *Ci;              /* compressed column input */
    Float *Cx;

#ifdef SPARSEFRAME_LU
    Long nzCPCT;                /* pattern of lower(C + C^T), the graph handed to METIS */
    Long *CPCTp, *CPCTi;
#endif

    Long *Lp, *Li;              /* lower(P A P^T) by column: INPUT of factorize */
    Float *Lx;
    Long *LTp, *LTi;
    Float *LTx;

#ifdef SPARSEFRAME_LU
    Long *Up, *Ui;              /* upper part by row: INPUT of factorize (aliases L for symmetric input) */
    Float *Ux;
    Long *UTp, *UTi;
    Float *UTx;
#endif

    enum PermMethod permMethod;
#ifdef SPARSEFRAME_LU
    Long *PivInv;               /* static pre-pivoting is disabled in the reference (LU.c:784-787): NULL unless SF_STATIC_PIVOT=1 */
#endif
    Long *Perm;
    Long *Parent, *Post, *ColCount;

    Long nsuper;                /* INPUT of factorize: the supernodal structure */
    Long *Super, *SuperMap, *Sparent;

    Long nsleaf;
    Long *LeafQueue;

    Long isize, xsize;
    Long *Lsip, *Lsxp, *Lsi;
    Float *Lsx;                 /* OUTPUT of factorize: dense panel per supernode */

    Long csize;

    Long nstage;                /* the reference's device-cache stages: reproduced, unused by the engine */
    Long *ST_Map, *ST_Pointer, *ST_Index, *ST_Parent;
    size_t *Aoffset, *Moffset;

    void *workspace;
    size_t workSize;

    Float *Bx, *Xx, *Rx;
    Float residual;

    double readTime, analyzeTime, factorizeTime, solveTime;
};

#endif
