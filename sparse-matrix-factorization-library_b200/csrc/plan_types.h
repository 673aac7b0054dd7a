// POD records shared by the host plan builder (plan.cpp), the device engine (engine.cu / kernels.cu) and the
// CPU plan interpreter used only by the tests (oracle/plan_host_exec.cpp).
//
// Data layout in HBM ("padded Lsx"): one dense column-major panel per supernode, exactly the reference's Lsx
// panel (CH.c:1632-1658; LU: LU.c:1946) except that the leading dimension is rounded up to a multiple of
// ld_align doubles and every panel starts on a multiple of ld_align, so that 16-byte vector / TMA access is legal.
//   Cholesky panel  : ld x nscol, rows [0,nscol) diagonal block (lower), rows [nscol,nsrow) L21
//   LU panel        : ld x nscol, rows [0,nscol) L11\U11, [nscol,nsrow) L21, [nsrow, 2 nsrow - nscol) U12^T
#pragma once
#include <stdint.h>

namespace sfb200 {

struct SnodeDev {
    int64_t xoff;    // element offset of the panel in the padded factor
    int64_t lsip;    // offset of the row list in Lsi (and in the per-row update tables)
    int32_t nscol;
    int32_t nsrow;
    int32_t ld;      // padded leading dimension
    int32_t level;   // height above the deepest leaf below it (leaves = 0)
};

// ---- GEMM-shaped work: C (-)= A * B^T on DMMA tiles -------------------------------------------------------
enum GemmFlags : int32_t {
    G_LOWER     = 1 << 0,   // C is diagonal-aligned: only tiles/elements with row >= col are touched
    G_SCATTER   = 1 << 1,   // C is not a dense block: elements are subtracted from ancestor panels through the
                            // relative-index tables (replaces mappedSubtract, CH.cu:62-124 / cpuApply CH.c:2073-2086)
    G_B_KMAJOR  = 1 << 2,   // B is stored K x N (k contiguous) instead of N x K        (LU: U11 row blocks)
    G_LIM_SEGB  = 1 << 3,   // scatter mask: row >= first row of the column's ancestor segment   (LU, L side)
    G_LIM_SEGE  = 1 << 4,   // scatter mask: row >= end of the column's ancestor segment, dest += u-offset (LU, U side)
    G_SKIP_HEAD = 1 << 5,   // the leading meta x meta block of C is left alone: the next diagonal block applies this
                            // update itself (DiagOp::upd_k) so that it can be factorized while this GEMM still runs
};

struct GemmOp {
    int64_t a_off, b_off, c_off;   // element offsets into the padded factor (c_off unused for scatter)
    int32_t lda, ldb, ldc;
    int32_t M, N, K;
    int32_t flags;
    int32_t tile0;                 // index of this op's first CTA inside its launch
    int64_t meta;                  // scatter: index of off-diagonal row 0 of the source supernode in the per-row tables
    int32_t tile_mod, tile_rem;    // multi-GPU: this rank computes the tiles t with t % tile_mod == tile_rem (1, 0 = all)
};

// ---- triangular solves with one nb x nb diagonal block ------------------------------------------------------
// Every variant is the recurrence  x_j = (b_j - sum_{p<j} x_p T[j,p]) / T[j,j]  over independent vectors b.
enum TrsmFlags : int32_t {
    T_TRANS = 1 << 0,   // T[j,p] is read from the upper triangle: T[j,p] = D[p,j]      (LU: X U11 = B)
    T_UNIT  = 1 << 1,   // unit diagonal                                                (LU: X L11^T = B, L11 X = B)
    T_COPY_T = 1 << 2,  // also store the solution transposed: element j of vector v at w_off + v + j * ldw
                        // (LU: the U11 row blocks are kept a second time as U11^T, so that the in-panel updates
                        // C -= L21 * U12 read their B operand row-major like every other GEMM of the schedule)
};

struct TrsmOp {
    int64_t t_off;      // diagonal block D (ld = ldt)
    int64_t b_off;      // first element of vector 0
    int32_t ldt;
    int32_t nb;         // block order (<= 64)
    int32_t nvec;       // number of vectors
    int32_t estride;    // stride between consecutive elements of one vector
    int32_t vstride;    // stride between vectors
    int32_t flags;
    int32_t tile0;      // first CTA of this op inside its launch (one CTA per 64 vectors)
    int32_t ldw;        // T_COPY_T: leading dimension of the transposed copy
    int64_t w_off;      // T_COPY_T: element offset of (vector 0, element 0) of the transposed copy
};

// ---- diagonal block factorization --------------------------------------------------------------------------
struct DiagOp {
    int64_t a_off;      // the nb x nb diagonal block (ld)
    int64_t l_off;      // leaf look-ahead: before factorizing, A -= L * U with L = nb x upd_k at l_off (ld; the panel
    int64_t u_off;      // rows of this block in the previous leaf's columns) and U = upd_k x nb at u_off (ld; LU only --
    int32_t ld;         // Cholesky uses L^T and touches the lower triangle only).  upd_k = 0: nothing pending.
    int32_t nb;
    int32_t upd_k;
    int32_t pad;
};

// ---- one kernel launch of the static schedule ---------------------------------------------------------------
enum LaunchKind : int32_t {
    L_SMALL_A = 0,   // whole-front-in-shared-memory kernel, tiny fronts
    L_SMALL_B = 1,   // same kernel, larger CTA / more shared memory
    L_POTRF   = 2,   // Cholesky of nb x nb diagonal blocks
    L_GETRF   = 3,   // no-pivot LU of nb x nb diagonal blocks
    L_TRSM    = 4,
    L_GEMM128 = 5,   // 128x128 CTA tiles
    L_GEMM64  = 6,   // 64x64 CTA tiles
    L_GEMM128WS = 7, // 128x128 tiles, persistent warp-specialized kernel: bulk-copy (TMA) producer warp + 8 DMMA consumer warps
    L_TRSM_MMA = 8,  // Cholesky row solves X L^T = B on the tensor pipe: 128 rows per CTA, blocked substitution with DMMA
    L_KINDS = 9,
};

struct Launch {
    int32_t kind;
    int32_t level;
    int32_t first;   // first op / task in the kind's array
    int32_t count;   // ops / tasks
    int32_t grid;    // CTAs
    int32_t smem;    // dynamic shared memory bytes
    int32_t top_block; // large-front launches: index of the widest-level block column this launch works on (-1: small fronts
                     // / scatter).  When it moves from i to i+1 inside a level, block column i of every large front of the
                     // level is final -> can be copied out while the rest is still being factorized
    int32_t flags;   // LaunchFlags
};

// Look-ahead: the top-level trailing update of block column b is split into NEXT (the columns of block b+1, main stream)
// and REST (everything right of it, side stream, on all but a few SMs).  Block column b+1 is then factorized on the main
// stream -- a chain of small latency-bound kernels -- while REST(b) keeps the machine busy.  NEXT(b+1) touches columns
// REST(b) also updates, so it waits for it.
enum LaunchFlags : int32_t {
    LF_SIDE      = 1 << 0,   // issue on the side stream (after everything issued so far on the main stream)
    LF_WAIT_SIDE = 1 << 1,   // main-stream launch that must wait for the side stream's work issued so far
    LF_DIAG_AHEAD = 1 << 2,  // diagonal blocks factorized on their own stream while the previous leaf's update still runs
    LF_WAIT_DIAG = 1 << 3,   // main-stream launch that must wait for those
    LF_LEAVE_SMS = 1 << 4,   // persistent GEMM that runs next to an LF_DIAG_AHEAD launch: do not occupy every SM
};

// ---- multi-GPU: the schedule of one rank is a list of SEGMENTS: first the segment's exchange steps, then its launches.
// The step lists are identical on every rank (steps are matched by order); a rank outside [lo,hi) skips the entry.
//
// Exchange medium = peer memory over NVLink (CUDA peer access inside one process, CUDA IPC between processes):
//   * every rank keeps an image of the panels it STORES -- the ones it owns plus the shared panels its group works on
//     together -- at its own offsets (compact: a rank's image is about 1/world of the factor plus the shared tops);
//   * an update aimed at a panel stored elsewhere is added straight into the destination rank's image by the scatter
//     epilogue of the update kernels (red.global.add.f64 through the peer mapping): upd_coloff carries the destination
//     rank in its top byte and the element offset inside THAT rank's image below it (extend-add fused into the producing
//     GEMM);
//   * what is left of the collectives are two kinds of steps, both device-side flags in peer memory (no host involvement,
//     capturable into the CUDA graph of the factorization):
enum CollType : int32_t {
    C_BARRIER = 0,   // every rank of [lo,hi) signals the others and waits for them: all updates the group has aimed at the
                     // shared panels of this level (or, first step of a factorization: all images are zeroed) are in place
    C_BCAST   = 1,   // a rows x cols block (column stride ld) of a factored panel / block column: root's copy is pushed
                     // into the image of every other rank of [lo,hi) (peer-to-peer copies), then a flag per destination
};

constexpr int kOwnerShift = 56;                                   // upd_coloff = (destination rank << 56) | element offset
constexpr int64_t kOffsetMask = ((int64_t)1 << kOwnerShift) - 1;
constexpr int kMaxWorld = 16;

struct Coll {
    int32_t type, root, lo, hi;
    int64_t offset;        // C_BCAST: first element of the block in THIS rank's image (-1 if this rank is outside [lo,hi))
    int32_t rows, cols;    // C_BCAST: block shape (rows contiguous)
    int32_t ld;            // C_BCAST: column stride (the panel's leading dimension, the same on every rank)
    int32_t dst_first;     // C_BCAST on its root: HostPlan::coll_dst[dst_first + (q - lo)] = offset of the block in rank q's image
};

// Flags of the exchange steps: both ends of a (signaller, waiter) pair walk the same step list and count the signals of that
// pair per channel, so the n-th signal from rank a to rank b on a channel is "n" on both sides without any communication.
// Channel 0: barriers (signalled from the main stream), channel 1: pushed blocks (signalled from the push stream).
struct SyncPoint {
    int32_t peer;
    uint32_t seq;
};
struct CollSync {               // per exchange step: ranges into HostPlan::sync_points (empty on ranks outside [lo,hi))
    int32_t sig_first, sig_count;     // signals this rank sends (barrier: to every other rank of the group; push: root to each destination,
                                      // the owner of the next block column first)
    int32_t wait_first, wait_count;   // flags this rank waits for
    int32_t chan, pad;
};

struct Segment {
    int32_t level;
    int32_t coll_first, coll_count;
    int32_t launch_first, launch_count;
    int32_t pad;
};

// all launches of one etree level (single-GPU consumers: streamed copy-out, look-ahead joins)
struct Phase {
    int32_t level;
    int32_t launch_first, launch_count;
    int32_t seg_first, seg_count;
    int32_t pad;
};

// ---- device-resident triangular solves (SparseFrame_solve_supernodal, CH.c:3036-3139 / LU.c:3592-3700) --------------
// One task = one 64-column block of one front, optionally restricted to a slab of rows.
struct SolveTask {
    int64_t xoff;      // panel offset in the padded factor
    int64_t lsip;      // offset of the front's row list
    int32_t ld, nsrow, nscol;
    int32_t col0;      // global index of the front's first column
    int32_t j0, w;     // block columns [j0, j0 + w)
    int32_t r0, r1;    // update tasks: panel rows [r0, r1)  (LU backward, U11 part: panel COLUMNS [r0, r1))
};

enum SolveKind : int32_t {
    S_DIAG_F = 0,   // x_b = T^-1 b_b, T = lower triangle of the block (unit diagonal for LU)
    S_UPD_F  = 1,   // b[rows] -= L[rows, b] x_b                         (rows below the block, atomics)
    S_UPD_B  = 2,   // b_b -= L[rows, b]^T x[rows]   (Cholesky) / b_b -= U12[b, rows] x[rows] (LU, from the U12^T rows)
    S_UPD_BU = 3,   // LU: b_b -= U11[b, later columns] x[later columns]
    S_DIAG_B = 4,   // x_b = T^-T b_b (Cholesky) / x_b = U_bb^-1 b_b (LU)
};

struct SolveLaunch {
    int32_t kind, first, count, pad;
};

// ---- final copy into the reference's unpadded layout -------------------------------------------------------
struct PackTask {
    int64_t src;     // padded offset
    int64_t dst;     // offset in the reference's Lsx
    int32_t rows;    // contiguous run length (panel rows)
    int32_t ld;      // padded leading dimension
    int32_t cols;
    int32_t pad;
};

}  // namespace sfb200
