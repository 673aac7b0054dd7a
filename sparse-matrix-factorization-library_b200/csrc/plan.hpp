// Static numeric schedule ("plan") for one symbolic structure.  Built once on the host from the supernodal
// structure; the numeric phase replays it as a fixed sequence of batched kernel launches, one etree level after
// another (replaces the dynamic ready-queue scheduler of SparseFrame_factorize_supernodal, CH.c:2264-2331,
// 2960-2986, and the per-update RelativeMap construction of CH.c:2055-2058).
#pragma once
#include <vector>

#include "plan_types.h"
#include "symbolic.hpp"

namespace sfb200 {

struct PlanParams {
    int ld_align = 2;            // panel leading dimensions / offsets are multiples of this many doubles
    int nb = 64;                 // order of the diagonal blocks factored inside one CTA
    int outer = 256;             // width of the second blocking level (its trailing updates run with K = outer)
    int outer2 = 1024;           // third level (0 = off): trailing updates of the widest fronts with K = outer2
    int small_a_bytes = 6144;    // fronts whose whole panel fits: 64-thread CTA
    int small_b_bytes = 49152;   //                              : 256-thread CTA
    int small_max_m = 160;       // ... and only if the update (m x m) is small enough for one CTA
    int gemm128_min_m = 192;     // GEMM ops with at least this many rows use 128x128 tiles, others 64x64
    int gemm128_min_k = -1;      // ... and only if K is at least this; -1 = auto: 257 on single-GPU plans -- the K = 64 and
                                 // K = 256 in-panel updates are RMW-bound short tiles that do as well on the 64x64 kernel at
                                 // 3 CTAs per SM, and the step gets 1.5-4 % faster (cfg2 211.2 -> 208.0 ms, cfg4 416.5 -> 408.6,
                                 // cfg1 7.12 -> 6.81; profiles/r01_sweep_min_k_cfg4_cfg1_cfg2.log) -- and 0 on partitioned
                                 // plans, where it has not been measured
    int leaf_lookahead = 1;      // diagonal block j+1 applies leaf j's update to itself and is factorized next to that update
    int trsm_mma = 0;            // Cholesky: the solves below a diagonal block on the tensor pipe (k_trsm_mma).  Off: measured
                                 // no faster than the vector-pipe kernel, both are bound by the latency of their staging
                                 // (profiles/r02_ncu_k_trsm_mma_v1.txt)
    int left_looking = 1;        // Cholesky: blocks below the top level pull their updates (left-looking) instead of pushing
                                 // rank-nb / rank-outer updates to the right: half the C traffic, longer K (PanelEmitter::block)
    int lookahead = 1;           // split the top-level trailing updates (see LaunchFlags) -- needs outer2 > 0
    int use_ws = 1;              // big GEMM ops with a row-major B go to the warp-specialized bulk-copy kernel
    int lu_ut_copy = 1;          // LU: large fronts keep U11^T in a scratch area behind the factor (written by the row-block
                                 // solves), the in-panel updates read it as a row-major B operand -> warp-specialized kernel
    int ws_kmajor = 0;           // ... and so do LU's k-major B ops (U11 row blocks, staged column by column) when their
                                 // sources are 16-byte aligned and K is a multiple of 16
    double balance_tol = 1.10;   // multi-GPU: subtrees are opened (their roots become shared fronts) until the heaviest rank's
                                 // local work is within this factor of the average.  Measured on cfg5 (profiles/
                                 // r02_mgpu_sweep_n{4,8}_cfg5.log): 4 ranks 1.10 = 1.05 = 967 ms, 1.02 990, 1.01 1125; 8 ranks
                                 // 1.05 666 ms (528 exchange steps, 147 GB pushed), 1.10 645 ms (308 steps, 94 GB): every
                                 // opened subtree adds shared fronts whose chains cost more than the imbalance they remove
    int weighted_blocks = 0;     // multi-GPU: deal the block columns of group-factorized fronts in a weighted cyclic order that
                                 // evens out the ranks' total flops.  Off: measured no gain (cfg5 on 4 GPUs 965.8 -> 974.9 ms) --
                                 // the schedule is level-synchronous, a rank that is light in the subtree phase waits at the first
                                 // group barrier whatever it is given later (profiles/r02_bench_n4_cfg5_v2_weighted_blocks.json)
    int dist_scatter_min_m = 512;// multi-GPU: Schur updates of shared fronts with at least this many rows are split over the group
    int dist_panel_min_cols = 2048; // multi-GPU: shared Cholesky fronts at least this wide are factorized by the whole group
                                 // (1-D block-cyclic over `outer`-wide block columns, one broadcast per block column)
};

struct HostPlan {
    Kind kind = Kind::Cholesky;
    bool split = false;          // LU on an unsymmetric matrix: separate U values
    PlanParams prm;
    i64 n = 0, nsuper = 0, nnz = 0, unnz = 0, isize = 0, xsize = 0;
    i64 xsize_pad = 0;           // elements of THIS rank's image of the padded factor (world == 1: the whole factor)
    i64 wsize = 0;               // scratch elements behind the padded factor (offsets xsize_pad .. xsize_pad + wsize): U11^T copies
    int nlevels = 0;
    i64 nsmall = 0, nlarge = 0;

    std::vector<SnodeDev> snode;            // nsuper
    std::vector<i64> lsxp;                  // reference (unpadded) panel offsets, nsuper + 1
    // per-row tables, indexed like Lsi (only off-diagonal positions are meaningful)
    std::vector<i64> upd_coloff;            // padded offset of the ancestor column this row's global index owns
    std::vector<i64> upd_relbase;           // relmap[relbase + i] = row of off-diagonal row i inside that ancestor
    std::vector<int32_t> upd_segb, upd_sege;// first / one-past-last off-diagonal row of the row's ancestor segment
    std::vector<int32_t> upd_uoff;          // LU: nsrow - nscol of the ancestor (row offset of its U12^T block)
    std::vector<int32_t> relmap;
    std::vector<i64> asm_dst_l, asm_dst_u;  // destination of every Lx (Ux) entry in the padded factor

    std::vector<int32_t> small_tasks;
    std::vector<DiagOp> diag_ops;
    std::vector<TrsmOp> trsm_ops;
    std::vector<GemmOp> gemm_ops;
    std::vector<Launch> launches;
    std::vector<PackTask> pack_tasks;

    // triangular-solve schedule (single-GPU plans): forward levels ascending, backward descending
    std::vector<int32_t> lsi32;             // row lists as int32
    std::vector<int32_t> nz_col;            // column of every Lx entry (row of every Ux entry): device residual
    std::vector<int32_t> unz_row;
    std::vector<i64> nz_row, unz_col;       // row of every Lx entry, column of every Ux entry
    std::vector<SolveTask> solve_tasks;
    std::vector<SolveLaunch> solve_fwd, solve_bwd;

    double flops = 0.0;                     // algorithmic flops

    // ---- etree partition over `world` GPUs (proportional mapping of subtrees) -------------------------
    int rank = 0, world = 1;
    std::vector<int32_t> owner;             // owning rank of every supernode
    std::vector<int32_t> group_lo, group_hi;// ranks [lo,hi) that hold descendants of the supernode (hi-lo > 1: "shared")
    double local_flops = 0.0;               // flops of the supernodes this rank owns
    std::vector<Coll> colls;                // exchange steps, grouped by segment (the same list on every rank)
    std::vector<CollSync> coll_sync;        // flags of every exchange step, as THIS rank sees them (plan_types.h)
    std::vector<SyncPoint> sync_points;
    std::vector<i64> coll_dst;              // C_BCAST steps this rank is the root of: block offset in every destination's image
    std::vector<Segment> segments;          // exchange steps, then launches
    std::vector<i64> image_elems;           // image size of every rank (elements); image_elems[rank] == xsize_pad
    double remote_update_entries = 0.0;     // update entries this rank adds into other ranks' images per factorization
    std::vector<Phase> phases;              // one per etree level
};

// world == 1: every supernode is owned by rank 0, no reduce steps.
int build_plan(const Symbolic &S, const PlanParams &prm, HostPlan &P, int rank = 0, int world = 1);

// tile geometry shared by builder, kernels and the CPU interpreter
// tiles of an op that rank-share (mod, rem) computes
inline i64 share_of_tiles(i64 total, int mod, int rem) { return total > rem ? (total - rem + mod - 1) / mod : 0; }

// tile index -> (block row, block column), the enumeration k_gemm decodes (lower ops list only tiles on/below the diagonal)
inline void decode_tile(i64 t, int M, int N, int B, bool lower, int &ti, int &tj) {
    const i64 nj = (N + B - 1) / B;
    if (!lower) { ti = (int)(t / nj); tj = (int)(t - (i64)ti * nj); return; }
    const i64 tri = nj * (nj + 1) / 2;
    if (t < tri) {
        i64 r = 0;
        while ((r + 1) * (r + 2) / 2 <= t) ++r;
        ti = (int)r; tj = (int)(t - r * (r + 1) / 2);
    } else {
        const i64 q = t - tri;
        ti = (int)(nj + q / nj); tj = (int)(q % nj);
    }
}

inline i64 gemm_tiles(int M, int N, int B, bool lower) {
    const i64 ni = (M + B - 1) / B, nj = (N + B - 1) / B;
    if (!lower) return ni * nj;
    const i64 tri = std::min(ni, nj);
    return tri * (tri + 1) / 2 + (ni > nj ? (ni - nj) * nj : 0);
}

}  // namespace sfb200
