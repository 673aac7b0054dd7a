/* Cholesky/Include/parameter.h -- tunables and policy functions (reference: Include/parameter.h:1-114).
 *
 * What the B200 engine does with each of them:
 *   MATRIX_THREAD_NUM          matrices in flight in the SparseFrame() driver (CH.c:3371-3444): same here
 *   OMP_THREAD_NUM ...         BLAS threading of the reference's CPU workers: no CPU worker exists here
 *   GPU_SPLIT_LIMIT, *_MULTIPLE  only enter the devSlotSize formula (CH.c:36-41, 81-87, 199), which the symbolic structure
 *                              depends on; the engine itself keeps the whole factor resident and has no slots
 *   should_relax               relaxed amalgamation rule, applied by the host analysis (csrc/symbolic.cpp) bit for bit
 *   get_node_score_*, set_factorize_location, potrf_split_*   CPU-vs-GPU routing of the reference's dynamic scheduler:
 *                              replaced by the static plan's small-front / large-front classes (csrc/plan.hpp PlanParams)
 * Functions are static inline (the reference defines them non-static in the header). */
#ifndef INCLUDE_PARAMETER_H
#define INCLUDE_PARAMETER_H

#include "arch.h"
#include "macro.h"
#include "type.h"

#define PRINT_INFO

#define MATRIX_THREAD_NUM (2)

#define OMP_THREAD_NUM (16)
#define CP_NUM_THREAD (4)
#define CP_THREAD_THRESHOLD (256)

#define MAX_NUM_CPU (0)
#define MAX_NUM_GPU (-1)
#define GPU_SPLIT_LIMIT (4)

#define MAXMMLINE (1024)

static const enum PermMethod perm_method = PERM_METIS;
static const double prune_dense = 10.0;
static const double aggressive = 1;

#define RELAXED_SUPERNODE

/* merge a child into its parent?  `col` = columns of the merged supernode, `rate` = fraction of explicit zeros it would hold */
static inline int should_relax(Long col, double rate)
{
    static const Long col_limit[3] = { 16, 64, 256 };
    static const double rate_limit[3] = { 0.8, 0.1, 0.05 };
    for (int k = 2; k >= 0; --k)
        if (col > col_limit[k] && rate > rate_limit[k]) return FALSE;
    return TRUE;
}

struct node_size_struct
{
    Long node;
    Long n;
    Long m;
    Long k;
    Long score;
};

/* size class of one descendant update (n x k times k x (n+m)): 0 = tiny, > 0 = worth a GPU, < 0 = CPU */
static inline Long get_node_score_from_dimension(Long n, Long m, Long k)
{
    if (n <= 16 && n + m <= 16 && k <= 16) return 0;
    if (n >= 32 && n + m >= 64 && k >= 32) return (n + m) * k;
    return -(n + m) * k;
}
static inline Long set_node_score(struct node_size_struct *node)
{
    return node->score = get_node_score_from_dimension(node->n, node->m, node->k);
}
static inline Long get_node_score(const struct node_size_struct *node) { return node->score; }

#define MAX_BLAS_BATCH_SIZE (128)

/* TRUE: factorize the supernode on the CPU (reference); the engine's equivalent is the small-front class of the plan */
static inline int set_factorize_location(Long nscol, Long nsrow)
{
    return nscol < 256 && (nsrow - nscol) < 256;
}

static const Long potrf_split_threshold = 512;
static const Long potrf_split_block = 256;

#define MAX_D_STREAM (2)
#define A_MULTIPLE (2)
#define B_MULTIPLE MAX_D_STREAM
#define C_MULTIPLE MAX_D_STREAM

#define CUDA_BLOCKDIM_X (16)
#define CUDA_BLOCKDIM_Y (24)

#endif
