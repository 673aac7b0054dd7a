/*
 * sfb200.h -- C ABI of the B200-native numeric factorization engine for SparseFrame's supernodal
 * Cholesky and LU.  Plain pointers and sizes only; every array uses the reference's own element
 * types (Long = int64_t, Float = double; Include/arch.h:6-16).
 *
 * The drop-in boundary is the reference's factorize step:
 *     int SparseFrame_factorize(struct common_info_struct*, struct gpu_info_struct*, struct matrix_info_struct*)
 *         Cholesky/Source/SparseFrame.c:3019-3034, LU/Source/SparseFrame.c:3575-3590
 * which reads the symbolic structure left in matrix_info by SparseFrame_analyze and fills matrix_info->Lsx.
 * The entry points below are what that function binds to (see INTEGRATION.md for the stub).
 *
 * Three opaque objects:
 *   sfb200_symbolic : result of symbolic analysis (host).       replaces SparseFrame_analyze, CH.c:1916-1978
 *   sfb200_plan     : device-resident schedule for one symbolic structure: etree levels, batched task lists,
 *                     relative-index maps, assembly map.          replaces the per-call setup of
 *                     SparseFrame_factorize_supernodal, CH.c:2150-2265, and RelativeMap, CH.c:2055-2058
 *   (numeric calls) : assemble + factorize on the GPU.            replaces CH.c:2104-2148 / LU.c:2622-2666
 *
 * All functions return 0 on success.  There is no CPU fallback: numeric entry points fail with
 * SFB200_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef SFB200_H
#define SFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB200_CHOLESKY 0
#define SFB200_LU       1

#define SFB200_OK               0
#define SFB200_ERR_ARG          1
#define SFB200_ERR_ORDERING     2
#define SFB200_ERR_NO_DEVICE    3
#define SFB200_ERR_CUDA         4
#define SFB200_ERR_NOT_POSDEF   5   /* a non-positive (Cholesky) or zero (LU) pivot was met; factor is unusable */
#define SFB200_ERR_STATE        6

typedef struct sfb200_symbolic sfb200_symbolic;
typedef struct sfb200_plan sfb200_plan;

const char *sfb200_version(void);
const char *sfb200_last_error(void);

/* ---- symbolic analysis (host only) ------------------------------------------------------------------ */

/* Full analysis of a compressed-column matrix as SparseFrame_compress leaves it (CH.c:526-587):
 * METIS ordering, etree, postorder, column counts, relaxed supernodes, row lists, stages.
 * kind = SFB200_CHOLESKY | SFB200_LU; dev_slot_size = common_info->devSlotSize (the structure depends on it). */
int sfb200_analyze(int kind, int is_symmetric, int64_t n, const int64_t *Cp, const int64_t *Ci, const double *Cx,
                   uint64_t dev_slot_size, sfb200_symbolic **out);

/* Same analysis with the fill-reducing ordering supplied by the caller instead of METIS_NodeND (which is 96 % of the
 * analysis time on the 3-D configs): order[k] = original index placed at position k, i.e. the array "Order"
 * of an earlier analysis of the same pattern -- which this call then reproduces bit for bit -- or any other permutation
 * (geometric nested dissection, AMD, identity).  SFB200_ERR_ORDERING if `order` is not a permutation of 0..n-1. */
int sfb200_analyze_ordered(int kind, int is_symmetric, int64_t n, const int64_t *Cp, const int64_t *Ci, const double *Cx,
                           uint64_t dev_slot_size, const int64_t *order, sfb200_symbolic **out);

/* Opt-in parallel ordering (METIS_NodeND is 93-96 % of the analysis time and single-threaded).  procs > 1:
 * nested dissection whose top log2(procs) separator levels are computed with METIS_ComputeVertexSeparator and whose subgraphs
 * are ordered by METIS_NodeND in forked processes (csrc/parallel_order.cpp).  The ordering is of the same kind and quality but
 * NOT the permutation a single METIS_NodeND call returns, so the structure that follows is self-consistent, not bit-identical
 * to the reference's: the default (1) stays the reference's ordering.  Process-wide; returns the value in force.  Also set by the
 * environment variable SFB200_PARALLEL_ORDER (when this function was never called). */
int sfb200_set_ordering_processes(int procs);

/* Wrap a structure computed elsewhere (e.g. by the reference's own SparseFrame_analyze): only what the
 * factorize step reads.  Arrays are copied.  Up/Ui may be NULL (Cholesky, or LU on a symmetric matrix). */
int sfb200_symbolic_wrap(int kind, int64_t n, int64_t nsuper, const int64_t *Super, const int64_t *Lsip,
                         const int64_t *Lsxp, const int64_t *Lsi, const int64_t *Lp, const int64_t *Li,
                         const int64_t *Up, const int64_t *Ui, sfb200_symbolic **out);

/* Complex matrices (matrix_info->isComplex; the reference's Complex = interleaved (re, im) doubles, Include/arch.h:9-16).
 * Same analysis, with the element size of a Complex wherever the structure depends on it (supernode caps CH.c:1480-1496,
 * relaxed merging 1558-1574, stages 1743-1790, Aoffset 1893-1896); values are carried as pairs ("Lx" etc. then hold 2 doubles
 * per entry).  A complex "symmetric" matrix keeps the reference's meaning: Cholesky reads the permuted lower triangle as the
 * lower triangle of a Hermitian matrix (zherk / zpotrf), LU as A = A^T.  `order` may be NULL (METIS). */
int sfb200_analyze_complex(int kind, int is_symmetric, int64_t n, const int64_t *Cp, const int64_t *Ci, const double *Cz,
                           uint64_t dev_slot_size, const int64_t *order, sfb200_symbolic **out);
int sfb200_symbolic_wrap_complex(int kind, int64_t n, int64_t nsuper, const int64_t *Super, const int64_t *Lsip,
                                 const int64_t *Lsxp, const int64_t *Lsi, const int64_t *Lp, const int64_t *Li,
                                 const int64_t *Up, const int64_t *Ui, sfb200_symbolic **out);

/* Scalars: "n","nnz","nsuper","nfsuper","isize","xsize","csize","nstage","nsleaf","unnz","complex".  -1 if unknown. */
int64_t sfb200_symbolic_int(const sfb200_symbolic *s, const char *name);

/* Wall time of the analysis stages in seconds: "t_order" (adjacency + METIS), "t_tree" (permute, etree, postorder,
 * column counts), "t_structure" (supernodes, row lists, stages); "flops".  -1 if unknown. */
double sfb200_symbolic_double(const sfb200_symbolic *s, const char *name);

/* Arrays by the reference's field names: "Order" (the ordering METIS returned, input of sfb200_analyze_ordered), "Perm","Parent","Post","ColCount","Lp","Li","LTp","LTi","Up","Ui",
 * "UTp","UTi","Super","SuperMap","Sparent","Lsip","Lsxp","Lsi","ST_Map","ST_Pointer","ST_Index","LeafQueue",
 * "Aoffset","Moffset" (int64) and "Lx","LTx","Ux","UTx" (double).  Returns NULL if absent; *count = length. */
const void *sfb200_symbolic_array(const sfb200_symbolic *s, const char *name, int64_t *count);

/* Algorithmic flop count of the numeric factorization (SURVEY.md section 8d formulas). */
double sfb200_symbolic_flops(const sfb200_symbolic *s);

void sfb200_symbolic_free(sfb200_symbolic *s);

/* ---- numeric plan (host build + one upload) ------------------------------------------------------------ */

int sfb200_device_count(void);

int sfb200_plan_create(const sfb200_symbolic *s, int device, sfb200_plan **out);

/* One rank of an etree-subtree partition over `world` GPUs of one box.  Whole subtrees are local to a rank; the
 * supernodes above them ("top") are shared by a contiguous group of ranks [lo,hi).  The exchange medium is peer memory over
 * NVLink -- the reference's only medium is host RAM (CH.c:2596-2685):
 *   - a rank's image holds the panels it stores (the ones it owns plus the shared fronts its group works on together);
 *   - an update aimed at a panel stored elsewhere is added straight into the destination rank's image by the scatter
 *     epilogue of the update kernels (extend-add fused into the producing GEMM);
 *   - wide shared Cholesky fronts are factorized by their group (block columns owned round-robin), the Schur update of a
 *     shared front is split over its group; factored blocks are pushed to the group with peer copies;
 *   - the remaining synchronisation is flags in peer memory (barrier per group and level, one flag per pushed block):
 *     no host round trip, the whole factorization of a rank is one CUDA graph.
 * The ranks are connected either across processes (one process per GPU: sfb200_plan_exchange_handle on every rank, exchange
 * the 128-byte handles by any means -- torch.distributed in sparseframe_b200/parallel.py --, sfb200_plan_connect) or inside
 * one process (sfb200_group_create, one host thread per GPU like the reference, CH.c:2267-2290).  Once connected,
 * sfb200_factorize_device / sfb200_factorize run this rank's part; all ranks must call them (they wait for each other on the
 * device; a rank that never arrives is reported after SFB200_PEER_TIMEOUT_S seconds, default 30).
 *   sfb200_plan_int: "nsegments", "ncolls", "rank", "world", "connected", "owned_elems"
 *   sfb200_plan_double: "flops", "local_flops", "remote_update_bytes", "push_bytes"
 *   segment k = {coll_first, coll_count, launch_first, launch_count, level}
 *   exchange step i = {type (0 barrier, 1 push), root, lo, hi, offset in this rank's image, rows, cols, ld} */
int sfb200_plan_create_part(const sfb200_symbolic *s, int device, int rank, int world, sfb200_plan **out);
double sfb200_plan_double(const sfb200_plan *p, const char *name);
int sfb200_plan_segment_info(const sfb200_plan *p, int k, int64_t out[5]);
int sfb200_plan_coll_info(const sfb200_plan *p, int i, int64_t out[8]);
int sfb200_plan_owner(const sfb200_plan *p, int32_t *owner /* nsuper */);
void *sfb200_plan_factor_ptr(sfb200_plan *p);                 /* device pointer of this rank's image of the padded factor */
/* offsets of the panels in that image (-1: not stored by this rank); xoff[nsuper] = elements of the image */
int sfb200_plan_panel_offsets(const sfb200_plan *p, int64_t *xoff /* nsuper + 1 */);
int sfb200_plan_exchange_handle(sfb200_plan *p, void *out /* 128 bytes */);
int sfb200_plan_connect(sfb200_plan *p, const void *handles /* world x 128 bytes, in rank order */);
void sfb200_plan_free(sfb200_plan *p);

/* The whole partition inside one process: rank r on devices[r] (entries may repeat). */
typedef struct sfb200_group sfb200_group;
int sfb200_group_create(const sfb200_symbolic *s, int world, const int *devices, sfb200_group **out);
int sfb200_group_size(const sfb200_group *g);
sfb200_plan *sfb200_group_plan(sfb200_group *g, int rank);
/* host buffers, one host thread per rank; every rank writes the panels it owns into Lsx (may be NULL) */
int sfb200_group_factorize(sfb200_group *g, const double *Lx, const double *Ux, double *Lsx);
double sfb200_group_last_ms(const sfb200_group *g, int which);
void sfb200_group_free(sfb200_group *g);

/* The caller has written a complete factor into the image of this single-GPU plan (sfb200_plan_factor_ptr +
 * sfb200_plan_panel_offsets), e.g. gathered from the ranks of a partition; d_Lx / d_Ux = the matrix values it belongs to.
 * Afterwards sfb200_solve / sfb200_residual / sfb200_download_factor work on it. */
int sfb200_plan_adopt_factor(sfb200_plan *p, const double *d_Lx, const double *d_Ux);

/* Scalars: "nlevels","nlaunches","nsmall","nlarge","relmap_entries","xsize_padded","device_bytes", ... */
int64_t sfb200_plan_int(const sfb200_plan *p, const char *name);

/* ---- numeric factorization ---------------------------------------------------------------------------- */

/* Host-buffer call, the one SparseFrame_factorize maps to: uploads the values of the permuted matrix
 * (Lx, and Ux for unsymmetric LU; same order as Lp/Li, Up/Ui), assembles, factorizes, and writes the
 * factor into Lsx (xsize doubles, the reference's panel layout).  Lsx may be NULL to leave the factor on
 * the device.  Blocks until done. */
int sfb200_factorize(sfb200_plan *p, const double *Lx, const double *Ux, double *Lsx);

/* One process of a multi-process partition: like sfb200_factorize, but the output is only the panels this rank owns, packed back
 * to back in supernode order ("owned_elems" doubles; sfb200_plan_owned_runs maps the pieces into Lsx) -- a full-size Lsx per
 * process would cost world times the factor in host memory. */
int sfb200_factorize_owned(sfb200_plan *p, const double *Lx, const double *Ux, double *packed);
int sfb200_download_owned(sfb200_plan *p, double *packed);
/* runs[3k..3k+2] = {offset in the packed buffer, offset in Lsx, elements}; returns the number of runs (runs = NULL: just count) */
int64_t sfb200_plan_owned_runs(const sfb200_plan *p, int64_t *runs);

/* Device-resident call: values already in HBM (device pointers), factor stays in HBM. Asynchronous on the
 * plan's stream; sfb200_plan_sync() waits. */
int sfb200_factorize_device(sfb200_plan *p, const double *d_Lx, const double *d_Ux);
int sfb200_plan_sync(sfb200_plan *p);

/* Run the plan on a caller-owned CUDA stream (cudaStream_t passed as void*) instead of the plan's own one.  The stream is
 * BORROWED: it must outlive the plan or be replaced by another call before it is destroyed. */
int sfb200_plan_set_stream(sfb200_plan *p, void *cuda_stream);

/* Instrumented replay (CUDA events around every launch) of the last factorization's values; per-kernel totals are
 * read back with sfb200_plan_profile_get(kind, what): kind 0/1 small fronts, 2 potrf, 3 getrf, 4 trsm, 5 gemm128,
 * 6 gemm64; what 0 = ms, 1 = launches, 2 = algorithmic flops, 3 = algorithmic bytes. */
int sfb200_plan_profile(sfb200_plan *p);
double sfb200_plan_profile_get(const sfb200_plan *p, int kind, int what);
/* per launch i of the schedule: out[8] = {kind, level, grid, ops, ms, algorithmic flops, max K, min K} */
int sfb200_plan_profile_launch(const sfb200_plan *p, int i, double *out);

/* Copy the device-resident factor into the reference's host layout (Lsx = the whole array of xsize doubles; a rank of a
 * partitioned plan writes only the panels it owns). */
int sfb200_download_factor(sfb200_plan *p, double *Lsx);

/* ---- consumers of the device-resident factor ---------------------------------------------------------
 * Forward/backward substitution on the GPU, replacing SparseFrame_solve_supernodal (CH.c:3036-3139, LU.c:3592-3700).
 * B: n x nrhs, column-major, leading dimension n, in the PERMUTED ordering (like matrix_info->Bx); overwritten by X. */
int sfb200_solve(sfb200_plan *p, double *B, int nrhs);                /* host buffer */
int sfb200_solve_device(sfb200_plan *p, double *d_B, int nrhs);       /* device buffer */
/* Residual of the permuted system with the matrix values of the last factorize (SparseFrame_validate, CH.c:3197-3263):
 * out[5] = {||A||_1, ||b||_inf, ||x||_inf, ||Ax-b||_inf, ||Ax-b||_inf / (||A||_1 ||x||_inf + ||b||_inf)}. */
int sfb200_residual(sfb200_plan *p, const double *x, const double *b, double out[5]);

/* ---- complex factorization: zherk/zgemm/zpotrf/ztrsm (CH.c:2060-2070, 2134-2145) and
 * zgemm/magma_zgetrf_nopiv/ztrsm (LU.c:2569-2581, 2652-2663) ---------------------------------------------------------------
 * A zplan runs the FP64 engine on the real embedding of the complex matrix (every entry a+ib becomes the 2x2 block
 * [a -b; b a]; csrc/complex_embed.hpp): chol(phi(A)) = phi(L), and the no-pivot LU factors are read off the embedded ones
 * without arithmetic.  It needs a single GPU, twice the factor's bytes and executes twice the real flops of a native complex
 * kernel; every kernel on its path is the FP64 engine's.  LU: rows are first rotated by the phase of their diagonal entry (the
 * embedded elimination divides by Re u_kk; sfb200_zplan_set_row_phase(zp, 0) or SFB200_ZPHASE=0 switches that off), and the
 * factors handed back are those of the ORIGINAL matrix either way.
 * All complex arrays are interleaved (re, im) doubles: Lz / Uz like Lx / Ux (same order as Lp/Li, Up/Ui), Lsz = xsize
 * Complex in the reference's panel layout, Bz = n x nrhs Complex. */
typedef struct sfb200_zplan sfb200_zplan;
int sfb200_zplan_create(const sfb200_symbolic *complex_structure, int device, sfb200_zplan **out);
void sfb200_zplan_free(sfb200_zplan *zp);
sfb200_plan *sfb200_zplan_real(sfb200_zplan *zp);              /* the plan of the embedded structure (owned by the zplan) */
int64_t sfb200_zplan_int(const sfb200_zplan *zp, const char *name);   /* "n","nnz","unnz","xsize","row_phase", else as sfb200_plan_int of the embedded plan */
int sfb200_zplan_set_row_phase(sfb200_zplan *zp, int on);
int sfb200_zfactorize(sfb200_zplan *zp, const double *Lz, const double *Uz, double *Lsz /* may be NULL */);
int sfb200_zfactorize_device(sfb200_zplan *zp, const double *d_Lz, const double *d_Uz);    /* asynchronous (with row phases the call returns once the n phases have reached the host) */
int sfb200_zplan_sync(sfb200_zplan *zp);
int sfb200_zdownload_factor(sfb200_zplan *zp, double *Lsz);
/* The reference's complex solve / validate are `// TODO` (CH.c:3088-3131, 3229-3256); these are what they would be. */
int sfb200_zsolve(sfb200_zplan *zp, double *Bz, int nrhs);
int sfb200_zresidual(sfb200_zplan *zp, const double *xz, const double *bz, double out[5]);
double sfb200_zlast_ms(const sfb200_zplan *zp);

/* Complex matrices on several GPUs of one box: the etree partition of the embedded structure as an sfb200_group (rank r on
 * devices[r]); values expanded and complex panels read back on the host (a page-locked copy of the embedded factor, 4 x xsize
 * doubles, is staged there).  A composition of the group and of the host halves of the complex path: no device code of its own. */
typedef struct sfb200_zgroup sfb200_zgroup;
int sfb200_zgroup_create(const sfb200_symbolic *complex_structure, int world, const int *devices, sfb200_zgroup **out);
sfb200_group *sfb200_zgroup_real(sfb200_zgroup *g);
int sfb200_zgroup_factorize(sfb200_zgroup *g, const double *Lz, const double *Uz, double *Lsz /* may be NULL */);
double sfb200_zgroup_last_ms(const sfb200_zgroup *g, int which);
void sfb200_zgroup_free(sfb200_zgroup *g);

/* Device time of the last factorize call in milliseconds (which = 2: the last solve) (CUDA events on the plan's stream):
 * which = 0 numeric only (assemble + factor), 1 = including H2D/D2H copies of the host-buffer call. */
double sfb200_last_ms(const sfb200_plan *p, int which);

/* Number of kernel launches the last factorize issued. */
int64_t sfb200_last_launches(const sfb200_plan *p);

/* Host pinned allocation helpers (so that callers can hand page-locked Lx/Lsx buffers). */
void *sfb200_host_alloc(size_t bytes);
void sfb200_host_free(void *ptr);

/* FP64 pipe probe for the roofline denominators (TFLOP/s): which = 0 DFMA vector pipe, 1 DMMA tensor pipe. */
double sfb200_probe_fp64_peak(int device, int which, int iters);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H */
