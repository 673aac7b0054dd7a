/*
 * sparseframe_abi.h -- the data contract of the SparseFrame library, restated for the B200 drop-in.
 *
 * The reference never declares its analyze / factorize / solve entry points in a header, but it exports them from
 * libSparseFrame.so and they exchange three structs (reference: Include/info.h:12-150, LU variant Include/info.h of the
 * LU tree with ten more fields).  A drop-in library must keep the member ORDER and TYPES of those structs -- that is
 * the ABI -- so they are restated here, grouped and commented by what the B200 engine does with each member.
 * Element types follow Include/arch.h:6-16 (Int=int, Long=long, Float=double).
 *
 * Define SPARSEFRAME_LU before including to get the LU flavour of matrix_info_struct.
 */
#ifndef SPARSEFRAME_ABI_H
#define SPARSEFRAME_ABI_H

#include <stddef.h>
#include <stdio.h>
#include <omp.h>

#ifndef Long
#define Int int
#define uInt unsigned int
#define Long long
#define uLong unsigned long
#define Float double
#endif

#ifndef TRUE
#define FALSE 0
#define TRUE 1
#endif

/* slot multiplicities of the reference's device arena (Include/parameter.h:105-109); kept because the arrays below are
 * sized by them and devSlotSize is derived from them */
#define SF_MAX_D_STREAM 2
#define SF_A_MULTIPLE 2
#define SF_B_MULTIPLE SF_MAX_D_STREAM
#define SF_C_MULTIPLE SF_MAX_D_STREAM

enum FactorizeType { TYPE_CHOLESKY, TYPE_QR, TYPE_LU };
enum PermMethod { PERM_IDENTITY, PERM_AMD, PERM_METIS };

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
    Long *PivInv;               /* static pre-pivoting is disabled in the reference (LU.c:784-787): always NULL */
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

#ifdef __cplusplus
extern "C" {
#endif

/* the one declared entry point (reference Include/SparseFrame.h:43) */
int SparseFrame(int argc, char **argv);

/* exported-but-undeclared entry points of the reference library, same names and signatures */
double SparseFrame_time(void);
int SparseFrame_allocate_gpu(struct common_info_struct *common_info, struct gpu_info_struct **gpu_info_list_ptr);
int SparseFrame_free_gpu(struct common_info_struct *common_info, struct gpu_info_struct **gpu_info_list_ptr);
int SparseFrame_allocate_matrix(struct common_info_struct *common_info, struct matrix_info_struct **matrix_info_list_ptr);
int SparseFrame_free_matrix(struct common_info_struct *common_info, struct matrix_info_struct **matrix_info_list_ptr);
int SparseFrame_initialize_matrix(struct matrix_info_struct *matrix_info);
int SparseFrame_read_matrix(struct matrix_info_struct *matrix_info);
int SparseFrame_compress(struct matrix_info_struct *matrix_info);
int SparseFrame_analyze(struct common_info_struct *common_info, struct matrix_info_struct *matrix_info);
int SparseFrame_factorize(struct common_info_struct *common_info, struct gpu_info_struct *gpu_info_list,
                          struct matrix_info_struct *matrix_info);
int SparseFrame_solve_supernodal(struct matrix_info_struct *matrix_info);
int SparseFrame_validate(struct matrix_info_struct *matrix_info);
int SparseFrame_cleanup_matrix(struct matrix_info_struct *matrix_info);

#ifdef __cplusplus
}
#endif

#endif /* SPARSEFRAME_ABI_H */
