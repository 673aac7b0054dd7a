// Hand-written sm_100a kernels of the numeric factorization engine.
//
//   k_assemble   scatter of the permuted matrix values into the zeroed panels     (loadA, CH.c:1998-2028)
//   k_small      one CTA per small front: whole panel in shared memory, factor, write back, and subtract its
//                Schur update from the ancestor panels through the relative-index tables
//                                                         (cpuApplyFactorize + cpuApply for small fronts, CH.c:2030-2148)
//   k_diag       Cholesky / no-pivot LU of one nb x nb diagonal block per CTA      (dpotrf_, CH.c:2135; magma_dgetrf_nopiv, LU.c:2653)
//   k_trsm       triangular solves against one diagonal block, one vector per thread (dtrsm_, CH.c:2142, LU.c:2660)
//   k_gemm       FP64 tensor-core (DMMA m8n8k4) tiles  C -= A B^T  fed by a cp.async shared-memory pipeline; the epilogue
//                either updates a dense block in place or performs the fused relative-index scatter-subtract
//                                                         (dsyrk_/dgemm_ + mapped subtract, CH.c:2061-2086, CH.cu:62-124)
//   k_pack       padded device layout -> the reference's Lsx layout
#include "kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace sfb200 {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(8x8) += A(8x4) * B(4x8), FP64 tensor core.  lane = 4*g + t:  a = A[g][t], b = B[t][g], c0/c1 = D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Fire-and-forget FP64 add into global memory -- this GPU's HBM or, through a peer mapping, another GPU's over NVLink.
// Explicit PTX: the destination base comes out of a pointer table, so the compiler cannot prove the address space and
// would emit a generic ATOM (returns a value, plus a shared-memory CAS fallback) instead of RED.  System scope on
// partitioned plans: several GPUs add into the same entries of a shared panel concurrently.
__device__ __forceinline__ void red_add(double *addr, double v, int sys) {
    if (sys) asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;\n" ::"l"(__cvta_generic_to_global(addr)), "d"(v) : "memory");
    else asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;\n" ::"l"(__cvta_generic_to_global(addr)), "d"(v) : "memory");
}

// first op whose tile range contains `tile` (ops sorted by tile0)
template <class Op>
__device__ __forceinline__ int find_op(const Op *ops, int nops, int tile) {
    int lo = 0, hi = nops - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ops[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------------------------
// assembly
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_assemble(const double *__restrict__ vals, const int64_t *__restrict__ dst, int64_t count,
                           double *__restrict__ X) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < count; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = dst[p];
        if (d >= 0) X[d] = vals[p];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// in-shared-memory panel factorizations.  Panel is column-major with leading dimension PR (= number of rows).
// ---------------------------------------------------------------------------------------------------------------
// Right-looking Cholesky of the leading n columns of a PR x n trapezoid (POTRF of the n x n block + TRSM of the rows
// below in one sweep).  ONE barrier per column: the sweep is square-root free -- column j stays unscaled in shared memory
// and every thread forms 1 / d_j itself, A[r,c] -= A[r,j] A[c,j] / d_j -- and the columns are scaled by 1 / sqrt(d_j) in one
// pass at the end (the two-barrier version, scale then update, was barrier-bound: 65 us per front on cfg1's 256-thread
// class).  The diagonal is left un-rooted in shared memory (d_j); the caller stores sqrt(d_j).
template <int NT>
__device__ __forceinline__ void smem_cholesky(double *sh, int PR, int n, int *status) {
    __shared__ double dsc[256];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    constexpr int TY = NT / 32;
    for (int j = 0; j < n; ++j) {
        __syncthreads();
        const double d = sh[j * PR + j];
        if (!(d > 0.0) && tid == 0) atomicOr(status, 1);
        const double rd = 1.0 / d;
        const double *colj = sh + j * PR;
        for (int c = j + 1 + ty; c < n; c += TY) {
            const double lcj = colj[c] * rd;
            double *colc = sh + c * PR;
            for (int r = c + tx; r < PR; r += 32) colc[r] = fma(-colj[r], lcj, colc[r]);
        }
    }
    __syncthreads();
    for (int c = tid; c < n; c += NT) dsc[c] = rsqrt(sh[c * PR + c]);
    __syncthreads();
    const int total = PR * n;
    for (int idx = tid; idx < total; idx += NT) {
        const int c = idx / PR, r = idx - c * PR;
        if (r > c) sh[idx] *= dsc[c];
    }
    __syncthreads();
}

// Right-looking LU without pivoting of a panel whose first R rows are [L11\U11 ; L21] and whose rows [R, PR) are
// U12^T (updated with the L entries of the diagonal block: U12^T[:,c] -= U12^T[:,k] * L[c,k]).
template <int NT>
__device__ __forceinline__ void smem_lu(double *sh, int PR, int R, int n, int *status) {
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    constexpr int TY = NT / 32;
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        const double piv = sh[k * PR + k];
        if (piv == 0.0 && tid == 0) atomicOr(status, 1);
        const double rinv = 1.0 / piv;
        double *colk = sh + k * PR;
        for (int r = k + 1 + tid; r < R; r += NT) colk[r] *= rinv;
        __syncthreads();
        for (int c = k + 1 + ty; c < n; c += TY) {
            double *colc = sh + c * PR;
            const double u = colc[k], lck = colk[c];
            for (int r = k + 1 + tx; r < R; r += 32) colc[r] = fma(-colk[r], u, colc[r]);
            for (int r = R + tx; r < PR; r += 32) colc[r] = fma(-colk[r], lck, colc[r]);
        }
    }
    __syncthreads();
}

// Schur update of a small front from shared memory, 4x4 register tiles, subtracted from the ancestors with
// red.global.add.f64.  Sa/Sb: m x n blocks (leading dimension PR).  mode 0: rows >= column (Cholesky);
// mode 1: rows >= first row of the column's ancestor segment (LU, L side); mode 2: rows >= segment end, U side.
template <int NT>
__device__ __forceinline__ void smem_scatter(const PlanDev &P, const double *Sa, const double *Sb, int m, int n, int PR,
                                             int64_t meta, int mode) {
    const int nb4 = (m + 3) >> 2;
    const int ntile = mode == 0 ? nb4 * (nb4 + 1) / 2 : nb4 * nb4;
    for (int t = threadIdx.x; t < ntile; t += NT) {
        int bi, bj;
        if (mode == 0) {
            bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while (bi * (bi + 1) / 2 > t) --bi;
            while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
            bj = t - bi * (bi + 1) / 2;
        } else {
            bi = t / nb4;
            bj = t - bi * nb4;
        }
        const int i0 = bi << 2, j0 = bj << 2;
        if (mode != 0) {
            const int lim0 = mode == 1 ? P.upd_segb[meta + j0] : P.upd_sege[meta + j0];
            if (i0 + 3 < lim0) continue;
        }
        int ia[4], ja[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { ia[e] = min(i0 + e, m - 1); ja[e] = min(j0 + e, m - 1); }
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double *ca = Sa + k * PR, *cb = Sb + k * PR;
            double av[4], bv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { av[e] = ca[ia[e]]; bv[e] = cb[ja[e]]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + b;
            if (j >= m) break;
            const int64_t q = meta + j;
            const int64_t tgt = P.upd_coloff[q];                  // (destination rank << 56) | offset in its image
            int64_t col = tgt & kOffsetMask;
            double *img = P.peer[tgt >> kOwnerShift];
            const int64_t relb = P.upd_relbase[q];
            int lim = j;
            if (mode == 1) lim = P.upd_segb[q];
            if (mode == 2) { lim = P.upd_sege[q]; col += P.upd_uoff[q]; }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = i0 + a;
                if (i < m && i >= lim) red_add(img + col + P.relmap[relb + i], -acc[a][b], P.multi);
            }
        }
    }
}

template <int NT, bool LU>
__global__ void __launch_bounds__(NT) k_small(PlanDev P, const int32_t *__restrict__ tasks) {
    extern __shared__ double sh[];
    const SnodeDev sn = P.snode[tasks[blockIdx.x]];
    const int n = sn.nscol, R = sn.nsrow, m = R - n;
    const int PR = LU ? R + m : R;
    const int ld = sn.ld, tid = threadIdx.x;
    double *A = P.X + sn.xoff;
    const int total = PR * n;
    for (int idx = tid; idx < total; idx += NT) {
        const int c = idx / PR, r = idx - c * PR;
        sh[idx] = A[(int64_t)c * ld + r];
    }
    if (LU) smem_lu<NT>(sh, PR, R, n, P.status);
    else smem_cholesky<NT>(sh, PR, n, P.status);
    for (int idx = tid; idx < total; idx += NT) {
        const int c = idx / PR, r = idx - c * PR;
        double v = sh[idx];
        if (LU) {
            A[(int64_t)c * ld + r] = v;
        } else if (r >= c) {
            if (r == c) { v = sqrt(v); }
            A[(int64_t)c * ld + r] = v;
        }
    }
    if (m > 0) {
        const int64_t meta = sn.lsip + n;
        if (LU) {
            smem_scatter<NT>(P, sh + n, sh + R, m, n, PR, meta, 1);
            smem_scatter<NT>(P, sh + R, sh + n, m, n, PR, meta, 2);
        } else {
            smem_scatter<NT>(P, sh + n, sh + n, m, n, PR, meta, 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// diagonal blocks of the large fronts: 64 x 64 (smaller blocks are padded with the identity) held in REGISTERS,
// 256 threads as a 16 x 16 grid of 4 x 4 blocks.  One __syncthreads per column: in phase j every thread applies the
// rank-1 update of column j to its block; the threads that own column j+1 then know its final entries and publish the
// scaled column (and, for LU, the owners of row j+1 publish the row) for the next phase.  The running diagonal is
// kept in shared memory (double-buffered) so that the owners of column j+1 can form the pivot without waiting for the
// thread that owns the diagonal entry.
// ---------------------------------------------------------------------------------------------------------------
template <bool LU>
__global__ void __launch_bounds__(256) k_diag(PlanDev P, const DiagOp *__restrict__ ops) {
    __shared__ double lcol[2][64];
    __shared__ double urow[2][64];
    __shared__ double dsum[2][64];
    const DiagOp op = ops[blockIdx.x];
    const int nb = op.nb, tid = threadIdx.x;
    // a warp holds two block COLUMNS (tx) and all 16 block rows (ty): the owners of a column -- the only threads that
    // need the pivot's reciprocal square root, a long dependent sequence -- sit in one warp, the other seven skip it
    const int ty = tid & 15, tx = tid >> 4;
    const int r0 = ty << 2, c0 = tx << 2;
    double *A = P.X + op.a_off;
    const bool active = LU || ty >= tx;

    double a[4][4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = r0 + rr, c = c0 + cc;
            a[rr][cc] = (r < nb && c < nb) ? A[(int64_t)c * op.ld + r] : (r == c ? 1.0 : 0.0);
        }
    if (op.upd_k > 0) {
        // leaf look-ahead: the K = upd_k update of the previous leaf was left to this block (G_SKIP_HEAD on the GEMM side),
        // A -= L * U with L = rows of this block in the previous leaf's columns; Cholesky: U = L^T, lower triangle only
        __shared__ double Ls[16][65];
        __shared__ double Us[LU ? 16 : 1][65];
        const double *Lg = P.X + op.l_off, *Ug = P.X + op.u_off;
        for (int k0 = 0; k0 < op.upd_k; k0 += 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = tid + 256 * q;
                {
                    const int kk = idx >> 6, r = idx & 63, k = k0 + kk;          // rows contiguous
                    Ls[kk][r] = (r < nb && k < op.upd_k) ? Lg[r + (int64_t)k * op.ld] : 0.0;
                }
                if (LU) {
                    const int c = idx >> 4, kk = idx & 15, k = k0 + kk;          // k contiguous
                    Us[kk][c] = (c < nb && k < op.upd_k) ? Ug[k + (int64_t)c * op.ld] : 0.0;
                }
            }
            __syncthreads();
            if (active) {
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    double lr[4], uc[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        lr[e] = Ls[kk][r0 + e];
                        uc[e] = LU ? Us[kk][c0 + e] : Ls[kk][c0 + e];
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) a[rr][cc] = fma(-lr[rr], uc[cc], a[rr][cc]);
                }
            }
            __syncthreads();
        }
    }
    if (ty == tx) {
#pragma unroll
        for (int e = 0; e < 4; ++e) dsum[0][r0 + e] = a[e][e];
    }
    if (tid < 64) { lcol[0][tid] = 0.0; urow[0][tid] = 0.0; }
    __syncthreads();

    // phase j = -1 .. 62 prepares column j+1; phases 0..62 also apply column j.  (column 63 needs no update phase)
    int buf = 0;
#pragma unroll 1
    for (int jb = 0; jb < 16; ++jb) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int jn = jb * 4 + jj;          // column being prepared in this phase (= j + 1)
            // ---- apply column j = jn - 1 (its scaled column / row are in lcol[buf], urow[buf]) ----------------------
            double lr[4], uc[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                lr[e] = lcol[buf][r0 + e];
                uc[e] = LU ? urow[buf][c0 + e] : lcol[buf][c0 + e];
            }
            if (active) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) a[rr][cc] = fma(-lr[rr], uc[cc], a[rr][cc]);
            }
            const int nbuf = buf ^ 1;
            if (tid < 64) dsum[nbuf][tid] = fma(-lcol[buf][tid], LU ? urow[buf][tid] : lcol[buf][tid], dsum[buf][tid]);
            if (tx == jb) {                      // owners of column jn (all 16, so every row of lcol is rewritten)
                // ---- pivot of column jn from the running diagonal -----------------------------------------------------
                const double dprev = dsum[buf][jn];
                const double lj = lcol[buf][jn], uj = LU ? urow[buf][jn] : lj;
                const double d = fma(-lj, uj, dprev);
                double rinv;
                if (LU) {
                    if (d == 0.0 && ty == 0) atomicOr(P.status, 1);
                    rinv = 1.0 / d;
                } else {
                    if (!(d > 0.0) && ty == 0) atomicOr(P.status, 1);
                    rinv = rsqrt(d);
                }
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int r = r0 + rr;
                    double v = 0.0;
                    if (r > jn) { v = a[rr][jj] * rinv; a[rr][jj] = v; }
                    else if (r == jn && !LU) a[rr][jj] = d * rinv;       // sqrt(d)
                    lcol[nbuf][r] = v;
                }
            }
            if (LU && ty == jb) {                // owners of row jn: publish the entries right of the diagonal
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = c0 + cc;
                    urow[nbuf][c] = c > jn ? a[jj][cc] : 0.0;
                }
            }
            buf = nbuf;
            __syncthreads();
        }
    }

#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = r0 + rr, c = c0 + cc;
            if (r < nb && c < nb && (LU || r >= c)) A[(int64_t)c * op.ld + r] = a[rr][cc];
        }
}

// ---------------------------------------------------------------------------------------------------------------
// triangular solve: x_j = (b_j - sum_{p<j} x_p T[j,p]) / T[j,j], one vector per thread held in registers,
// T broadcast from shared memory.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_trsm(PlanDev P, const TrsmOp *__restrict__ ops, int nops) {
    __shared__ __align__(16) double Ts[64 * 64];   // Ts[j*64 + p] = T[p][j] for p > j, zero elsewhere
    __shared__ double rinv[64];
    const TrsmOp op = ops[find_op(ops, nops, (int)blockIdx.x)];
    const int tid = threadIdx.x, nb = op.nb;
    const double *D = P.X + op.t_off;
    const bool trans = (op.flags & T_TRANS) != 0, unit = (op.flags & T_UNIT) != 0;
    {
        // thread p fetches row p of T: 64 independent loads in flight (a dependent load loop here used to cost more
        // than the solve itself), then a conflict-free store of Ts[j*64 + p]
        const int p = tid;
        double t[64];
        const double *rowp = trans ? D + (int64_t)p * op.ldt : D + p;
        const int64_t step = trans ? 1 : op.ldt;
#pragma unroll
        for (int j = 0; j < 64; ++j) t[j] = (p < nb && j < p) ? rowp[(int64_t)j * step] : 0.0;
#pragma unroll
        for (int j = 0; j < 64; ++j) Ts[j * 64 + p] = t[j];
    }
    rinv[tid] = (tid < nb && !unit) ? 1.0 / D[tid + (int64_t)tid * op.ldt] : 1.0;
    __syncthreads();

    const int v = ((int)blockIdx.x - op.tile0) * 64 + tid;
    if (v >= op.nvec) return;
    double *B = P.X + op.b_off + (int64_t)v * op.vstride;
    double b[64];
    {
        const double *bp = B;       // running pointer: 64 precomputed 64-bit addresses would not fit in registers
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            b[j] = j < nb ? *bp : 0.0;
            bp += op.estride;
        }
    }
    double *bout = B;
    // T_COPY_T: second, transposed copy (consecutive vectors -> consecutive addresses: coalesced)
    const bool copy_t = (op.flags & T_COPY_T) != 0;
    double *wout = P.X + (copy_t ? op.w_off + v : 0);
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        const double x = b[j] * rinv[j];
        if (j < nb) *bout = x;      // x_j is final: store now so that its register dies here
        bout += op.estride;
        if (copy_t) {
            if (j < nb) *wout = x;
            wout += op.ldw;
        }
        const double nx = -x;
        // chunks of 16 columns with a compiler barrier in between: otherwise every shared-memory load of the
        // column (up to 63 doubles) is hoisted in front of the FMAs and the vector no longer fits in registers
#pragma unroll
        for (int p0 = (j + 1) & ~15; p0 < 64; p0 += 16) {
#pragma unroll
            for (int p = p0; p < p0 + 16; ++p)
                if (p > j) b[p] = fma(nx, Ts[j * 64 + p], b[p]);
            asm volatile("" ::: "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Triangular solve on the FP64 tensor pipe (Cholesky: X L^T = B for the panel rows below a diagonal block L, nb <= 64).
// One CTA = 128 rows: the tile B (rows x 64) and L live in shared memory, each warp owns 32 rows -- rows are independent
// in this solve, so there is no CTA-wide barrier after the staging.  Blocked forward substitution over 8-column blocks:
//     X_j = (B_j - sum_{p<j} X_p L[j,p]^T) L[j,j]^-T
// the sum is a chain of DMMA 8x8x4 (A = X columns already solved, from shared memory; B = rows of L), the 8 x 8 diagonal
// blocks are inverted once per CTA (64 threads, 8 steps each) and applied with two more DMMA.  Replaces the one-vector-
// per-thread k_trsm (vector pipe, 4 TFLOP/s, 24 us per launch on the critical path of every leaf) for Cholesky.
// ---------------------------------------------------------------------------------------------------------------
namespace tm {
constexpr int ROWS = 128, SR = ROWS + 4, ST = 64 + 4;
constexpr int SMEM_BYTES = (64 * SR + 64 * ST + 8 * 64) * 8;
}  // namespace tm

__global__ void __launch_bounds__(128) k_trsm_mma(PlanDev P, const TrsmOp *__restrict__ ops, int nops) {
    using namespace tm;
    extern __shared__ __align__(16) double smem_tm[];
    double *Xs = smem_tm;                 // Xs[c * SR + r]   : tile element (row r, column c)
    double *Ts = Xs + 64 * SR;            // Ts[k * ST + j]   : L[j][k] for k <= j (identity beyond nb)
    double *Ti = Ts + 64 * ST;            // Ti[jb*64 + k*8 + n] : inverse of diagonal block jb, entry (n, k)
    const TrsmOp op = ops[find_op(ops, nops, (int)blockIdx.x)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = op.nb;
    const int r0 = ((int)blockIdx.x - op.tile0) * ROWS;
    const int nrows = min(ROWS, op.nvec - r0);
    const double *D = P.X + op.t_off;
    double *B = P.X + op.b_off + r0;       // element (r, c) at B[r + c * estride]
    const int64_t ldb = op.estride;
    {
        double v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int idx = tid + 128 * q, j = idx & 63, k = idx >> 6;
            v[q] = (j < nb && k <= j) ? D[j + (int64_t)k * op.ldt] : (j == k ? 1.0 : 0.0);
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int idx = tid + 128 * q, j = idx & 63, k = idx >> 6;
            Ts[k * ST + j] = v[q];
        }
    }
    {
        const bool rok = tid < nrows;
#pragma unroll 16
        for (int c = 0; c < 64; ++c) Xs[c * SR + tid] = (rok && c < nb) ? B[tid + (int64_t)c * ldb] : 0.0;
    }
    __syncthreads();
    if (tid < 64) {
        // column c of the inverse of diagonal block jb: L_jb x = e_c
        const int jb = tid >> 3, c = tid & 7;
        const double *Lb = Ts + (jb * 8) * ST + jb * 8;       // Lb[k * ST + i] = block entry (i, k)
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double acc = i == c ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < i && k >= c) acc = fma(-Lb[k * ST + i], x[k], acc);
            x[i] = i >= c ? acc / Lb[i * ST + i] : 0.0;
            Ti[jb * 64 + c * 8 + i] = x[i];
        }
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int wr = warp * 32;
#pragma unroll 1
    for (int jb = 0; jb < 8; ++jb) {
        if (jb * 8 >= nb) break;
        double acc[4][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
        for (int p = 0; p < 8 * jb; p += 4) {
            const double b = Ts[(p + t) * ST + jb * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) dmma(acc[mt][0], acc[mt][1], Xs[(p + t) * SR + wr + 8 * mt + g], b);
        }
        double *y0 = Xs + (jb * 8 + 2 * t) * SR + wr + g, *y1 = y0 + SR;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            y0[8 * mt] -= acc[mt][0];
            y1[8 * mt] -= acc[mt][1];
            acc[mt][0] = acc[mt][1] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 8; ks += 4) {
            const double b = Ti[jb * 64 + (ks + t) * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) dmma(acc[mt][0], acc[mt][1], Xs[(jb * 8 + ks + t) * SR + wr + 8 * mt + g], b);
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            y0[8 * mt] = acc[mt][0];
            y1[8 * mt] = acc[mt][1];
        }
        __syncwarp();
    }
    __syncthreads();
    if (tid < nrows) {
#pragma unroll 16
        for (int c = 0; c < 64; ++c)
            if (c < nb) B[tid + (int64_t)c * ldb] = Xs[c * SR + tid];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// DMMA GEMM tiles.  C(BM x BN) -= A(BM x K) * B(BN x K)^T, operands column-major (rows contiguous).
//   shared tile layout: [k][row] with a row stride of (BM + 4) doubles -> the 8-byte fragment loads of one half-warp
//   (4 k x 4 rows) fall in 16 distinct 8-byte banks.
//   The tensor-core product is issued transposed (D' = Bfrag x Afrag) so that each thread ends up holding two
//   CONSECUTIVE ROWS of one column of C: 16 contiguous bytes of the column-major destination.
// ---------------------------------------------------------------------------------------------------------------

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int BK = 16>
struct GemmCfg {
    static constexpr int GEMM_BK = BK;
    static constexpr int NT = WARPS_M * WARPS_N * 32;
    static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    static constexpr int FM = WM / 8, FN = WN / 8;
    static constexpr int SA = BM + 4, SB = BN + 4;
    static constexpr int STAGE_DOUBLES = GEMM_BK * (SA + SB);
    static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
};

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int BK>
__global__ void __launch_bounds__(WARPS_M *WARPS_N * 32)
    k_gemm(PlanDev P, const GemmOp *__restrict__ ops, int nops) {
    using Cfg = GemmCfg<BM, BN, WARPS_M, WARPS_N, STAGES, BK>;
    constexpr int GEMM_BK = BK;
    constexpr int NT = Cfg::NT, FM = Cfg::FM, FN = Cfg::FN, SA = Cfg::SA, SB = Cfg::SB;
    static_assert(BM == BN, "triangular tile decode assumes square tiles");
    extern __shared__ __align__(16) double smem[];

    const GemmOp op = ops[find_op(ops, nops, (int)blockIdx.x)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = op.M, N = op.N, K = op.K;

    // ---- which tile -------------------------------------------------------------------------------------------
    int ti, tj;
    {
        const int t = ((int)blockIdx.x - op.tile0) * op.tile_mod + op.tile_rem;      // multi-GPU: this rank's share of the tiles
        const int nj = (N + BN - 1) / BN;
        if (op.flags & G_LOWER) {
            const int tri = nj * (nj + 1) / 2;
            if (t < tri) {
                ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
                while (ti * (ti + 1) / 2 > t) --ti;
                while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
                tj = t - ti * (ti + 1) / 2;
            } else {
                const int r = t - tri;
                ti = nj + r / nj;
                tj = r - (r / nj) * nj;
            }
        } else {
            ti = t / nj;
            tj = t - ti * nj;
        }
    }
    const int row0 = ti * BM, col0 = tj * BN;

    // LU scatter ops are enumerated as full rectangles: leave early when the whole tile is masked
    if (op.flags & (G_LIM_SEGB | G_LIM_SEGE)) {
        const int lim0 = (op.flags & G_LIM_SEGB) ? P.upd_segb[op.meta + col0] : P.upd_sege[op.meta + col0];
        if (min(row0 + BM, M) - 1 < lim0) return;
    }

    const double *Ag = P.X + op.a_off + row0;
    const double *Bg = P.X + op.b_off + ((op.flags & G_B_KMAJOR) ? (int64_t)col0 * op.ldb : (int64_t)col0);
    const int mrem = M - row0, nrem = N - col0;
    const bool b_kmajor = (op.flags & G_B_KMAJOR) != 0;

    // ---- operand staging: per-thread copy descriptors computed ONCE ----------------------------------------------
    // Every thread copies the same row offset of NP consecutive passes of k-rows, so a stage is NP cp.async per operand
    // whose source advances by a fixed stride; per k-iteration only one pointer add and one k-bound compare per copy
    // remain (a generic index -> (k,row) decode per copy costs more issue slots than the DMMAs of the stage).
    // 16-byte copies whenever the operand's columns start on even element offsets (always true for the in-place
    // trailing updates; true for a scatter op when the front has an even number of columns).
    const bool a_vec = (((op.a_off | (int64_t)op.lda) & 1) == 0);
    const bool b_vec = !b_kmajor && (((op.b_off | (int64_t)op.ldb) & 1) == 0);

    struct Stager {
        const double *src;      // source of copy 0 for the current k-iteration
        int64_t pass_step;      // elements between consecutive copies of one stage
        int64_t iter_step;      // elements between consecutive k-iterations
        int k_first;            // k offset (inside the stage) of copy 0   [row-major operands]
        int k_pass;             // k-rows per pass                           [row-major operands]
        int dst;                // shared-memory offset (doubles) of copy 0 inside the operand's stage buffer
        int dst_step;           // ... and between consecutive copies
        int bytes;              // 16 / 8 / 0: valid bytes of this thread's row position (row-major operands)
        int r_first, r_pass;    // k-major B only: row of copy 0 and rows per pass
    };
    auto make_rowmajor = [&](const double *base, int ld, int rows_rem, int stride_s, bool vec) {
        Stager st{};
        const int tpr = vec ? (BM / 2) : BM;                 // threads per k-row (BM == BN)
        const int r = (tid % tpr) * (vec ? 2 : 1);
        st.k_first = tid / tpr;
        st.k_pass = NT / tpr;
        const int rem = rows_rem - r;
        st.bytes = vec ? (rem >= 2 ? 16 : (rem == 1 ? 8 : 0)) : (rem >= 1 ? 8 : 0);
        st.src = base + r + (int64_t)st.k_first * ld;
        st.pass_step = (int64_t)st.k_pass * ld;
        st.iter_step = (int64_t)GEMM_BK * ld;
        st.dst = st.k_first * stride_s + r;
        st.dst_step = st.k_pass * stride_s;
        return st;
    };
    Stager sa = make_rowmajor(Ag, op.lda, mrem, SA, a_vec);
    Stager sb;
    if (!b_kmajor) {
        sb = make_rowmajor(Bg, op.ldb, nrem, SB, b_vec);
    } else {                                                  // element (k, r) of the tile lives at Bg[k + r * ldb]
        sb = Stager{};
        const int k = tid % GEMM_BK;
        sb.r_first = tid / GEMM_BK;
        sb.r_pass = NT / GEMM_BK;
        sb.k_first = k;
        sb.src = Bg + k + (int64_t)sb.r_first * op.ldb;
        sb.pass_step = (int64_t)sb.r_pass * op.ldb;
        sb.iter_step = GEMM_BK;
        sb.dst = k * SB + sb.r_first;
        sb.dst_step = sb.r_pass;
    }

    // issues the copies of k-iteration `kt` into `stage`; must be called with kt increasing by one each time
    auto load_stage = [&](int stage, int kt) {
        double *As = smem + stage * Cfg::STAGE_DOUBLES;
        double *Bs = As + GEMM_BK * SA;
        const int krem = K - kt * GEMM_BK;                    // valid k-rows of this stage
        {
            const double *p = sa.src;
            if (a_vec) {
                constexpr int NP = GEMM_BK / (NT / (BM / 2));
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int bytes = (sa.k_first + i * (NT / (BM / 2)) < krem) ? sa.bytes : 0;
                    cp_async16(As + sa.dst + i * sa.dst_step, bytes ? p : Ag, bytes);
                    p += sa.pass_step;
                }
            } else {
                constexpr int NP = GEMM_BK / (NT / BM);
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int bytes = (sa.k_first + i * (NT / BM) < krem) ? sa.bytes : 0;
                    cp_async8(As + sa.dst + i * sa.dst_step, bytes ? p : Ag, bytes);
                    p += sa.pass_step;
                }
            }
            sa.src += sa.iter_step;
        }
        {
            const double *p = sb.src;
            if (b_vec) {
                constexpr int NP = GEMM_BK / (NT / (BN / 2));
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int bytes = (sb.k_first + i * (NT / (BN / 2)) < krem) ? sb.bytes : 0;
                    cp_async16(Bs + sb.dst + i * sb.dst_step, bytes ? p : Bg, bytes);
                    p += sb.pass_step;
                }
            } else if (!b_kmajor) {
                constexpr int NP = GEMM_BK / (NT / BN);
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int bytes = (sb.k_first + i * (NT / BN) < krem) ? sb.bytes : 0;
                    cp_async8(Bs + sb.dst + i * sb.dst_step, bytes ? p : Bg, bytes);
                    p += sb.pass_step;
                }
            } else {
                constexpr int NP = BN / (NT / GEMM_BK);
                const bool kok = sb.k_first < krem;
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int bytes = (kok && sb.r_first + i * sb.r_pass < nrem) ? 8 : 0;
                    cp_async8(Bs + sb.dst + i * sb.dst_step, bytes ? p : Bg, bytes);
                    p += sb.pass_step;
                }
            }
            sb.src += sb.iter_step;
        }
    };

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int wm0 = (warp % WARPS_M) * Cfg::WM, wn0 = (warp / WARPS_M) * Cfg::WN;
    const int g = lane >> 2, t4 = lane & 3;
    const int KT = (K + GEMM_BK - 1) / GEMM_BK;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    const bool dbg_nocopy = (P.dbg & 1) != 0, dbg_nobar = (P.dbg & 2) != 0;
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        if (!dbg_nobar) __syncthreads();
        if (kt + STAGES - 1 < KT && !dbg_nocopy) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
        cp_async_commit();
        const double *As = smem + (kt % STAGES) * Cfg::STAGE_DOUBLES;
        const double *Bs = As + GEMM_BK * SA;
#pragma unroll
        for (int ks = 0; ks < GEMM_BK; ks += 4) {
            double af[FM], bf[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) af[i] = As[(ks + t4) * SA + wm0 + 8 * i + g];
#pragma unroll
            for (int j = 0; j < FN; ++j) bf[j] = Bs[(ks + t4) * SB + wn0 + 8 * j + g];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: thread holds C[row0 + wm0 + 8i + 2*t4 + {0,1}][col0 + wn0 + 8j + g] -------------------------
    if (op.flags & G_SCATTER) {
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int c = col0 + wn0 + 8 * j + g;
            if (c >= N) continue;
            const int64_t q = op.meta + c;
            const int64_t tgt = P.upd_coloff[q];                  // (destination rank << 56) | offset in its image
            int64_t col = tgt & kOffsetMask;
            const int64_t relb = P.upd_relbase[q];
            int lim = c;
            if (op.flags & G_LIM_SEGB) lim = P.upd_segb[q];
            if (op.flags & G_LIM_SEGE) { lim = P.upd_sege[q]; col += P.upd_uoff[q]; }
            double *dst = P.peer[tgt >> kOwnerShift] + col;       // possibly another GPU's memory: the add travels over NVLink
            const int32_t *rel = P.relmap + relb;
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                const int r = row0 + wm0 + 8 * i + 2 * t4;
                if (r < M && r >= lim) red_add(dst + rel[r], -acc[i][j][0], P.multi);
                if (r + 1 < M && r + 1 >= lim) red_add(dst + rel[r + 1], -acc[i][j][1], P.multi);
            }
        }
    } else {
        double *Cg = P.X + op.c_off;
        const bool lower = (op.flags & G_LOWER) != 0;
        const int head = (op.flags & G_SKIP_HEAD) ? (int)op.meta : 0;
        // rows r are even; with even c_off and ldc every (r, r+1) pair is one aligned 16-byte word of the column
        const bool vec = (((op.c_off | (int64_t)op.ldc) & 1) == 0);
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int c = col0 + wn0 + 8 * j + g;
            if (c >= N) continue;
            double *colp = Cg + (int64_t)c * op.ldc;
            // G_SKIP_HEAD: rows < head of the columns < head belong to the next diagonal block's own update
            const int lim = max(lower ? c : 0, c < head ? head : 0);
            // all loads of the column first (independent, all in flight), then the stores: a load/modify/store chain
            // per element would expose one global-memory round trip per element
            double2 cv[FM];
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                const int r = row0 + wm0 + 8 * i + 2 * t4;
                cv[i] = make_double2(0.0, 0.0);
                if (r + 1 < M && r + 1 >= lim) {
                    if (vec) cv[i] = *reinterpret_cast<const double2 *>(colp + r);
                    else { cv[i].x = colp[r]; cv[i].y = colp[r + 1]; }
                } else if (r < M && r >= lim) {
                    cv[i].x = colp[r];
                }
            }
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                const int r = row0 + wm0 + 8 * i + 2 * t4;
                const bool w0 = r < M && r >= lim, w1 = r + 1 < M && r + 1 >= lim;
                const double v0 = cv[i].x - acc[i][j][0], v1 = cv[i].y - acc[i][j][1];
                if (w0 && w1 && vec) *reinterpret_cast<double2 *>(colp + r) = make_double2(v0, v1);
                else {
                    if (w0) colp[r] = v0;
                    if (w1) colp[r + 1] = v1;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialized persistent DMMA GEMM (128 x 128 tiles).
//
//   warp 0 of the producer warpgroup : stages operands with cp.async.bulk (the TMA engine: SASS UBLKCP), one 1-KB bulk
//       copy per k-row and operand, completion counted in bytes on a "full" mbarrier.  No per-thread address math,
//       no LDGSTS, no CTA-wide barrier anywhere in the main loop.
//   8 consumer warps (2 x 4, 64 x 32 register tiles) : wait on "full", DMMA from shared memory, release the stage on an
//       "empty" mbarrier; epilogue (in-place update or fused relative-index scatter) straight from registers.
//   setmaxnreg moves the producer group's registers to the consumers.
//   Persistent: a CTA walks the launch's tile list with a stride of gridDim.x, and the stage ring keeps running across
//   tiles, so the producer is already filling the next tile's stages while the consumers are in the epilogue.
//
// Bulk copies need 16-byte aligned global addresses: when a tile's first row sits on an odd element offset the copy
// starts one element earlier and the consumers read the tile with a one-row shift (the shared tile has 4 spare rows).
// Rows read beyond the op's extent are finite garbage or zeros and only feed outputs the epilogue masks; k-rows beyond K
// are zero-filled by the producer warp.
// ---------------------------------------------------------------------------------------------------------------
namespace ws {
#ifndef SFB200_WS_STAGES
#define SFB200_WS_STAGES 4
#endif
constexpr int BM = 128, BN = 128, BK = 16, STAGES = SFB200_WS_STAGES, SA = 132;
// k-major B operands (LU: the U11 row blocks, element (k, n) at B[k + n * ldb]) are staged column by column: BK
// contiguous doubles (one 128-byte bulk copy) per column n, at a column stride of SBK doubles -- the fragment loads of a
// half-warp (4 columns x 4 k) then fall in 16 distinct 8-byte banks (20 * 8 B = 160 B: column c starts at bank 4c mod 16).
constexpr int SBK = 20;
constexpr int B_STAGE_DOUBLES = (BN * SBK > BK * SA) ? BN * SBK : BK * SA;
constexpr int STAGE_DOUBLES = BK * SA + B_STAGE_DOUBLES;
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8 + 2 * STAGES * 8 + 64;
constexpr int NT = 384;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct TileInfo {
    GemmOp op;
    int row0, col0;
    bool skip;
};

__device__ __forceinline__ TileInfo locate_tile(const PlanDev &P, const GemmOp *ops, int nops, int cta_tile) {
    TileInfo ti;
    ti.op = ops[find_op(ops, nops, cta_tile)];
    const GemmOp &op = ti.op;
    const int t = (cta_tile - op.tile0) * op.tile_mod + op.tile_rem;
    const int nj = (op.N + BN - 1) / BN;
    int bi, bj;
    if (op.flags & G_LOWER) {
        const int tri = nj * (nj + 1) / 2;
        if (t < tri) {
            bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
            while (bi * (bi + 1) / 2 > t) --bi;
            while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
            bj = t - bi * (bi + 1) / 2;
        } else {
            const int r = t - tri;
            bi = nj + r / nj;
            bj = r - (r / nj) * nj;
        }
    } else {
        bi = t / nj;
        bj = t - bi * nj;
    }
    ti.row0 = bi * BM;
    ti.col0 = bj * BN;
    ti.skip = false;
    if (op.flags & (G_LIM_SEGB | G_LIM_SEGE)) {       // LU scatter ops are enumerated as full rectangles
        const int lim0 = (op.flags & G_LIM_SEGB) ? P.upd_segb[op.meta + ti.col0] : P.upd_sege[op.meta + ti.col0];
        ti.skip = min(ti.row0 + BM, op.M) - 1 < lim0;
    }
    return ti;
}
}  // namespace ws

__global__ void __launch_bounds__(ws::NT, 1) k_gemm_ws(PlanDev P, const GemmOp *__restrict__ ops, int nops, int total_tiles) {
    using namespace ws;
    extern __shared__ __align__(128) double smem_ws[];
    double *smem = smem_ws;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_DOUBLES);
    uint64_t *empty = full + STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp < 4) {
        // =========================================== producer warpgroup ===========================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp != 0) return;
        int stage = 0;
        unsigned phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const TileInfo ti = locate_tile(P, ops, nops, tile);
            if (ti.skip) continue;
            const GemmOp &op = ti.op;
            const int mrem = min(BM, op.M - ti.row0), nrem = min(BN, op.N - ti.col0);
            const int KT = (op.K + BK - 1) / BK;
            if (op.flags & G_B_KMAJOR) {
                // A: lanes 0-15 stage one k-row each (as below); B: every lane stages the columns lane, lane + 32, ... of the
                // tile, BK contiguous doubles each.  The plan routes a k-major op here only if b_off and ldb are even (16-byte
                // aligned sources) and K is a multiple of BK (plan.cpp: emit_gemm).
                const int sha = (int)((op.a_off + ti.row0) & 1);
                const unsigned bytes_a = (unsigned)(((mrem + sha + 1) & ~1) * 8);
                const double *srcA = P.X + op.a_off + ti.row0 - sha + (int64_t)(lane & 15) * op.lda;
                const double *srcB = P.X + op.b_off + (int64_t)(ti.col0 + lane) * op.ldb;
                const int64_t stepA = (int64_t)BK * op.lda, colB = (int64_t)32 * op.ldb;
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(empty + stage, phase ^ 1);
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // see below
                    double *As = smem + stage * STAGE_DOUBLES, *Bs = As + BK * SA;
                    if (lane < 16) bulk_g2s(As + lane * SA, srcA, bytes_a, full + stage);
                    const double *sb = srcB;
                    for (int c = lane; c < nrem; c += 32, sb += colB) bulk_g2s(Bs + c * SBK, sb, BK * 8, full + stage);
                    srcA += stepA;
                    srcB += BK;
                    __syncwarp();
                    if (lane == 0) mbar_arrive_expect_tx(full + stage, (unsigned)BK * bytes_a + (unsigned)nrem * BK * 8);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                continue;
            }
            const double *Ag = P.X + op.a_off + ti.row0, *Bg = P.X + op.b_off + ti.col0;
            const int sha = (int)((op.a_off + ti.row0) & 1), shb = (int)((op.b_off + ti.col0) & 1);
            const unsigned bytes_a = (unsigned)(((mrem + sha + 1) & ~1) * 8), bytes_b = (unsigned)(((nrem + shb + 1) & ~1) * 8);
            // lanes 0-15 stage the 16 k-rows of A, lanes 16-31 those of B: one bulk copy per lane and stage
            const int k = lane & 15;
            const bool is_b = lane >= 16;
            const double *src = (is_b ? Bg - shb + (int64_t)k * op.ldb : Ag - sha + (int64_t)k * op.lda);
            const int64_t step = (int64_t)BK * (is_b ? op.ldb : op.lda);
            const unsigned bytes = is_b ? bytes_b : bytes_a;
            for (int kt = 0; kt < KT; ++kt) {
                mbar_wait(empty + stage, phase ^ 1);
                // The stage was last read through the generic proxy (the consumers' shared-memory loads, this warp's zero
                // fill); the bulk copies below write it through the async proxy.  The mbarrier orders the warps, the proxy
                // fence orders the two proxies: without it a 6-stage build of this kernel produced wrong factors (backward
                // error 1e-5 on lap3d 64^3, run-to-run different), with it 6/6 runs are clean.
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                double *dst = smem + stage * STAGE_DOUBLES + (is_b ? BK * SA : 0) + k * SA;
                const int krem = op.K - kt * BK;
                if (k < krem) {
                    bulk_g2s(dst, src, bytes, full + stage);
                } else {
                    for (int e = 0; e < SA; ++e) dst[e] = 0.0;      // k-rows beyond K must contribute nothing
                }
                src += step;
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(full + stage, (unsigned)min(krem, BK) * (bytes_a + bytes_b));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ================================================ consumer warps ================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int cw = warp - 4;
    const int wm0 = (cw % 2) * 64, wn0 = (cw / 2) * 32;
    const int g = lane >> 2, t4 = lane & 3;
    constexpr int FM = 8, FN = 4;
    int stage = 0;
    unsigned phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileInfo ti = locate_tile(P, ops, nops, tile);
        if (ti.skip) continue;
        const GemmOp &op = ti.op;
        const int M = op.M, N = op.N, row0 = ti.row0, col0 = ti.col0;
        const int sha = (int)((op.a_off + row0) & 1), shb = (int)((op.b_off + col0) & 1);
        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        const int KT = (op.K + BK - 1) / BK;
        if (op.flags & G_B_KMAJOR) {
            // B staged [column][k] with a column stride of SBK (see the producer): no alignment shift on this path
            for (int kt = 0; kt < KT; ++kt) {
                mbar_wait(full + stage, phase);
                const double *As = smem + stage * STAGE_DOUBLES + sha + wm0 + g;
                const double *Bs = smem + stage * STAGE_DOUBLES + BK * SA + (wn0 + g) * SBK + t4;
#pragma unroll
                for (int ks = 0; ks < BK; ks += 4) {
                    double af[FM], bf[FN];
#pragma unroll
                    for (int i = 0; i < FM; ++i) af[i] = As[(ks + t4) * SA + 8 * i];
#pragma unroll
                    for (int j = 0; j < FN; ++j) bf[j] = Bs[8 * j * SBK + ks];
#pragma unroll
                    for (int i = 0; i < FM; ++i)
#pragma unroll
                        for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        } else {
            for (int kt = 0; kt < KT; ++kt) {
                mbar_wait(full + stage, phase);
                const double *As = smem + stage * STAGE_DOUBLES + sha + wm0 + g;
                const double *Bs = smem + stage * STAGE_DOUBLES + BK * SA + shb + wn0 + g;
#pragma unroll
                for (int ks = 0; ks < BK; ks += 4) {
                    double af[FM], bf[FN];
#pragma unroll
                    for (int i = 0; i < FM; ++i) af[i] = As[(ks + t4) * SA + 8 * i];
#pragma unroll
                    for (int j = 0; j < FN; ++j) bf[j] = Bs[(ks + t4) * SA + 8 * j];
#pragma unroll
                    for (int i = 0; i < FM; ++i)
#pragma unroll
                        for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }

        // ---- epilogue: thread holds C[row0 + wm0 + 8i + 2*t4 + {0,1}][col0 + wn0 + 8j + g] ---------------------
        if (op.flags & G_SCATTER) {
#pragma unroll
            for (int j = 0; j < FN; ++j) {
                const int c = col0 + wn0 + 8 * j + g;
                if (c >= N) continue;
                const int64_t q = op.meta + c;
                const int64_t tgt = P.upd_coloff[q];                  // (destination rank << 56) | offset in its image
                int64_t col = tgt & kOffsetMask;
                const int64_t relb = P.upd_relbase[q];
                int lim = c;
                if (op.flags & G_LIM_SEGB) lim = P.upd_segb[q];
                if (op.flags & G_LIM_SEGE) { lim = P.upd_sege[q]; col += P.upd_uoff[q]; }
                double *dst = P.peer[tgt >> kOwnerShift] + col;       // possibly another GPU's memory: the add travels over NVLink
                const int32_t *rel = P.relmap + relb;
#pragma unroll
                for (int i = 0; i < FM; ++i) {
                    const int r = row0 + wm0 + 8 * i + 2 * t4;
                    if (r < M && r >= lim) red_add(dst + rel[r], -acc[i][j][0], P.multi);
                    if (r + 1 < M && r + 1 >= lim) red_add(dst + rel[r + 1], -acc[i][j][1], P.multi);
                }
            }
        } else {
            double *Cg = P.X + op.c_off;
            const bool lower = (op.flags & G_LOWER) != 0;
        const int head = (op.flags & G_SKIP_HEAD) ? (int)op.meta : 0;
            const bool vec = (((op.c_off | (int64_t)op.ldc) & 1) == 0);
#pragma unroll
            for (int j = 0; j < FN; ++j) {
                const int c = col0 + wn0 + 8 * j + g;
                if (c >= N) continue;
                double *colp = Cg + (int64_t)c * op.ldc;
                // G_SKIP_HEAD: rows < head of the columns < head belong to the next diagonal block's own update
            const int lim = max(lower ? c : 0, c < head ? head : 0);
                double2 cv[FM];
#pragma unroll
                for (int i = 0; i < FM; ++i) {
                    const int r = row0 + wm0 + 8 * i + 2 * t4;
                    cv[i] = make_double2(0.0, 0.0);
                    if (r + 1 < M && r + 1 >= lim) {
                        if (vec) cv[i] = *reinterpret_cast<const double2 *>(colp + r);
                        else { cv[i].x = colp[r]; cv[i].y = colp[r + 1]; }
                    } else if (r < M && r >= lim) {
                        cv[i].x = colp[r];
                    }
                }
#pragma unroll
                for (int i = 0; i < FM; ++i) {
                    const int r = row0 + wm0 + 8 * i + 2 * t4;
                    const bool w0 = r < M && r >= lim, w1 = r + 1 < M && r + 1 >= lim;
                    const double v0 = cv[i].x - acc[i][j][0], v1 = cv[i].y - acc[i][j][1];
                    if (w0 && w1 && vec) *reinterpret_cast<double2 *>(colp + r) = make_double2(v0, v1);
                    else {
                        if (w0) colp[r] = v0;
                        if (w1) colp[r + 1] = v1;
                    }
                }
            }
        }
    }
}

using Gemm128 = GemmCfg<128, 128, 2, 4, 3, 16>;
using Gemm128w16 = GemmCfg<128, 128, 4, 4, 3, 16>;
using Gemm128k32 = GemmCfg<128, 128, 2, 4, 3, 32>;
using Gemm128w16k32 = GemmCfg<128, 128, 4, 4, 3, 32>;
using Gemm64 = GemmCfg<64, 64, 2, 2, 4, 16>;

// ---------------------------------------------------------------------------------------------------------------
// copy-out
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const double *__restrict__ X, double *__restrict__ out,
                                               const PackTask *__restrict__ tasks) {
    const PackTask t = tasks[blockIdx.x];
    const int total = t.rows * t.cols;
    for (int idx = threadIdx.x; idx < total; idx += 256) {
        const int c = idx / t.rows, r = idx - c * t.rows;
        out[t.dst + idx] = X[t.src + (int64_t)c * t.ld + r];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Device-resident triangular solves over the padded factor (replaces the scalar host loops of
// SparseFrame_solve_supernodal, CH.c:3074-3134 / LU.c:3630-3694).  HBM-bound: every factor entry is read once per sweep.
// Right-hand sides: n x nrhs, column-major, leading dimension n, PERMUTED ordering (as matrix_info->Bx / Xx).
// ---------------------------------------------------------------------------------------------------------------
struct SolveArgs {
    const double *X;            // padded factor
    const int32_t *lsi;         // row lists
    double *rhs;                // n x nrhs
    int64_t n;
    int nrhs;
    int is_lu;
};

__device__ __forceinline__ int solve_gidx(const SolveTask &t, const int32_t *lsi, int i) {
    return i < t.nscol ? t.col0 + i : lsi[t.lsip + i];
}

// x_b = T^-1 b_b on one 64 x 64 diagonal block.
//   mode 0: forward, T lower (unit diagonal for LU);  mode 1: backward Cholesky, T = L^T;  mode 2: backward LU, T = U
// 64 threads stage the block (thread r fetches row r: 64 independent loads in flight, coalesced across threads) and the
// reciprocals of its diagonal; then ONE warp runs the substitution with register-resident unknowns (lane l owns
// equations l and l + 32) and a shuffle broadcast per unknown -- no block-wide barrier inside the 64 dependent steps.
// x_j = v_j * (1 / t_jj): the reciprocal is formed once per block outside the dependent chain (<= 1 ulp from v_j / t_jj).
__global__ void __launch_bounds__(64) k_solve_diag(SolveArgs A, const SolveTask *__restrict__ tasks, int mode) {
    __shared__ double T[64 * 65];
    __shared__ double rd[64];
    const SolveTask t = tasks[blockIdx.x];
    const int tid = threadIdx.x, w = t.w;
    const double *D = A.X + t.xoff + t.j0 + (int64_t)t.j0 * t.ld;
    const bool unit = A.is_lu && mode == 0;
    {
        double row[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) row[c] = (tid < w && c < w) ? D[(int64_t)c * t.ld + tid] : (c == tid ? 1.0 : 0.0);
#pragma unroll
        for (int c = 0; c < 64; ++c) T[tid * 65 + c] = row[c];         // T[r][c] = block(r, c); identity beyond w
        rd[tid] = unit ? 1.0 : 1.0 / T[tid * 65 + tid];              // own row: written by this thread just above
    }
    __syncthreads();
    // one warp per right-hand side: warp 0 takes the even ones, warp 1 the odd ones (a single right-hand side leaves
    // warp 1 idle)
    const int lane = tid & 31;
    const int64_t g0 = t.col0 + t.j0;
    for (int q = tid >> 5; q < A.nrhs; q += 2) {
        double *b = A.rhs + (int64_t)q * A.n + g0;
        double v0 = lane < w ? b[lane] : 0.0, v1 = lane + 32 < w ? b[lane + 32] : 0.0;
#pragma unroll 8
        for (int s = 0; s < 64; ++s) {
            const int j = mode == 0 ? s : 63 - s;
            // coefficient of unknown j in equation i:  mode 0: block(i, j), i > j;  mode 1: block(j, i) (L^T), i < j;
            //                                           mode 2: block(i, j) (U), i < j
            const double mine = (j & 32) ? v1 : v0;
            const double xj = __shfl_sync(0xffffffffu, mine, j & 31) * rd[j];
            if (lane == (j & 31)) { if (j & 32) v1 = xj; else v0 = xj; }
            const double c0 = mode == 1 ? T[j * 65 + lane] : T[lane * 65 + j];
            const double c1 = mode == 1 ? T[j * 65 + lane + 32] : T[(lane + 32) * 65 + j];
            const bool l0 = mode == 0 ? lane > j : lane < j, l1 = mode == 0 ? lane + 32 > j : lane + 32 < j;
            if (l0) v0 = fma(-c0, xj, v0);
            if (l1) v1 = fma(-c1, xj, v1);
        }
        if (lane < w) b[lane] = v0;
        if (lane + 32 < w) b[lane + 32] = v1;
    }
}

// The update kernels take QB right-hand sides per pass over the factor (QB = 1 for a single one, 4 otherwise): every
// factor entry is loaded once per pass and used QB times from registers.

// forward: b[rows] -= L[rows, block] x_block, one thread per panel row of the slab
template <int QB>
__global__ void __launch_bounds__(256) k_solve_upd_fwd(SolveArgs A, const SolveTask *__restrict__ tasks) {
    __shared__ double xb[QB][64];
    const SolveTask t = tasks[blockIdx.x];
    const double *P = A.X + t.xoff + (int64_t)t.j0 * t.ld;
    for (int q0 = 0; q0 < A.nrhs; q0 += QB) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < QB * 64; idx += 256) {
            const int qq = idx >> 6, jj = idx & 63;
            xb[qq][jj] = (q0 + qq < A.nrhs && jj < t.w) ? A.rhs[(int64_t)(q0 + qq) * A.n + t.col0 + t.j0 + jj] : 0.0;
        }
        __syncthreads();
        for (int i = t.r0 + threadIdx.x; i < t.r1; i += 256) {
            double acc[QB];
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) acc[qq] = 0.0;
#pragma unroll 16
            for (int jj = 0; jj < t.w; ++jj) {
                const double l = P[(int64_t)jj * t.ld + i];
#pragma unroll
                for (int qq = 0; qq < QB; ++qq) acc[qq] = fma(l, xb[qq][jj], acc[qq]);
            }
            const int gi = solve_gidx(t, A.lsi, i);
#pragma unroll
            for (int qq = 0; qq < QB; ++qq)
                if (q0 + qq < A.nrhs) atomicAdd(A.rhs + (int64_t)(q0 + qq) * A.n + gi, -acc[qq]);
        }
    }
}

// backward: b_block -= M[rows, block]^T x[rows] over a slab of rows; M = L (Cholesky) or the U12^T rows (LU).
// Warp w handles block columns w, w+8, ...; lanes stride the rows (coalesced), shuffle reduction, one atomic per column.
template <int QB>
__global__ void __launch_bounds__(256) k_solve_upd_bwd(SolveArgs A, const SolveTask *__restrict__ tasks) {
    __shared__ double xs[QB][1024];
    const SolveTask t = tasks[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rowshift = A.is_lu ? t.nsrow - t.nscol : 0;         // LU: off-diagonal row i lives at panel row i + (R - n)
    const double *P = A.X + t.xoff + (int64_t)t.j0 * t.ld + rowshift;
    const int cnt = t.r1 - t.r0;
    for (int q0 = 0; q0 < A.nrhs; q0 += QB) {
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += 256) {
            const int gi = solve_gidx(t, A.lsi, t.r0 + k);
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) xs[qq][k] = q0 + qq < A.nrhs ? A.rhs[(int64_t)(q0 + qq) * A.n + gi] : 0.0;
        }
        __syncthreads();
        for (int jj = warp; jj < t.w; jj += 8) {
            const double *col = P + (int64_t)jj * t.ld + t.r0;
            double acc[QB];
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) acc[qq] = 0.0;
#pragma unroll 8
            for (int k = lane; k < cnt; k += 32) {
                const double l = col[k];
#pragma unroll
                for (int qq = 0; qq < QB; ++qq) acc[qq] = fma(l, xs[qq][k], acc[qq]);
            }
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[qq] += __shfl_xor_sync(0xffffffffu, acc[qq], o);
                if (lane == 0 && q0 + qq < A.nrhs) atomicAdd(A.rhs + (int64_t)(q0 + qq) * A.n + t.col0 + t.j0 + jj, -acc[qq]);
            }
        }
    }
}

// backward, LU only: b_block -= U11[block rows, columns c in slab] x[c]; thread jj owns block row jj
template <int QB>
__global__ void __launch_bounds__(64) k_solve_upd_bwd_u(SolveArgs A, const SolveTask *__restrict__ tasks) {
    const SolveTask t = tasks[blockIdx.x];
    const int jj = threadIdx.x;
    if (jj >= t.w) return;
    const double *P = A.X + t.xoff + t.j0 + jj;
    for (int q0 = 0; q0 < A.nrhs; q0 += QB) {
        double acc[QB];
#pragma unroll
        for (int qq = 0; qq < QB; ++qq) acc[qq] = 0.0;
#pragma unroll 8
        for (int c = t.r0; c < t.r1; ++c) {
            const double u = P[(int64_t)c * t.ld];
#pragma unroll
            for (int qq = 0; qq < QB; ++qq)
                if (q0 + qq < A.nrhs) acc[qq] = fma(u, A.rhs[(int64_t)(q0 + qq) * A.n + t.col0 + c], acc[qq]);
        }
#pragma unroll
        for (int qq = 0; qq < QB; ++qq)
            if (q0 + qq < A.nrhs) atomicAdd(A.rhs + (int64_t)(q0 + qq) * A.n + t.col0 + t.j0 + jj, -acc[qq]);
    }
}

// residual of the permuted system (SparseFrame_validate, CH.c:3197-3259): r = A x - b, column sums of |A|
//   sym: entries of lower(P A P^T) by column (row >= col, mirrored);  !sym: L part takes strictly-lower entries only
__global__ void k_residual_lower(const double *__restrict__ vals, const int64_t *__restrict__ rows, const int32_t *__restrict__ cols,
                                 int64_t nnz, int mirror, int strict, const double *__restrict__ x, double *r, double *colsum) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = rows[p];
        const int j = cols[p];
        const double v = vals[p];
        if (strict && i <= j) continue;
        atomicAdd(r + i, v * x[j]);
        atomicAdd(colsum + j, fabs(v));
        if (mirror && i != j) { atomicAdd(r + j, v * x[i]); atomicAdd(colsum + i, fabs(v)); }
    }
}
// U part of an unsymmetric matrix: entry p is (row urow[p], column ucol[p])
__global__ void k_residual_upper(const double *__restrict__ vals, const int32_t *__restrict__ urow, const int64_t *__restrict__ ucol,
                                 int64_t nnz, const double *__restrict__ x, double *r, double *colsum) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        atomicAdd(r + urow[p], vals[p] * x[ucol[p]]);
        atomicAdd(colsum + ucol[p], fabs(vals[p]));
    }
}
// out[k] = max_i |v_k[i]| for the four vectors (non-negative doubles order like their bit patterns)
__global__ void k_absmax4(const double *a, const double *b, const double *c, const double *d, int64_t n, unsigned long long *out) {
    double m[4] = {0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        m[0] = fmax(m[0], fabs(a[i])); m[1] = fmax(m[1], fabs(b[i])); m[2] = fmax(m[2], fabs(c[i])); m[3] = fmax(m[3], fabs(d[i]));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        for (int o = 16; o > 0; o >>= 1) m[k] = fmax(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
        if ((threadIdx.x & 31) == 0) atomicMax(out + k, (unsigned long long)__double_as_longlong(m[k]));
    }
}
__global__ void k_axpby(double *r, const double *b, double alpha, int64_t n) {       // r = alpha * b
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) r[i] = alpha * b[i];
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 pipe peak probes
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_peak_dfma(double *out, int iters) {
    double a[8];
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_peak_dmma(double *out, int iters) {
    double c[8][2];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory exactly like the k_gemm<128,128> inner loop (8 warps, 64x32 warp tiles, 12 fragment
// loads per 32 DMMA), but with no global traffic and no barriers: the ceiling of the inner loop itself.
// mode 0: fragments re-read from shared memory every k-step; mode 1: fragments loaded once (register operands).
__global__ void __launch_bounds__(256) k_peak_dmma_tile(double *out, int iters, int mode) {
    constexpr int SA = 132, BK = 16;
    __shared__ double As[BK * SA], Bs[BK * SA];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < BK * SA; i += 256) { As[i] = 1.0 + 1e-6 * i; Bs[i] = 1.0 - 1e-6 * i; }
    __syncthreads();
    const int wm0 = (warp % 2) * 64, wn0 = (warp / 2) * 32, g = lane >> 2, t4 = lane & 3;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double af[8], bf[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) af[i] = As[t4 * SA + wm0 + 8 * i + g];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = Bs[t4 * SA + wn0 + 8 * j + g];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            if (mode == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = As[(ks + t4) * SA + wm0 + 8 * i + g];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = Bs[(ks + t4) * SA + wn0 + 8 * j + g];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
        }
        if (mode == 0) asm volatile("" ::: "memory");
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + tid] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// Exchange steps of a partitioned factorization (plan_types.h: Coll): flags in peer memory, one warp each.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_epoch_inc(unsigned long long *epoch) { *epoch += 1; }

// Zero fill of a rank's image.  A kernel, not cudaMemsetAsync: a device memory set is one of the operations that make the
// driver serialize commands of DIFFERENT streams around it (CUDA programming guide, "implicit synchronization"); with
// several ranks on one device the memset of one rank then waits for the spinning flag wait of another, which waits for
// that very rank -- seen in round 2 as flag timeouts whenever ranks shared a GPU.
__global__ void __launch_bounds__(256) k_zero(double *__restrict__ x, int64_t n, int *status) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *status = 0;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x, tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t head = (reinterpret_cast<uintptr_t>(x) & 8) ? 1 : 0;       // to a 16-byte boundary
    if (tid == 0 && head && n > 0) x[0] = 0.0;
    double2 *v = reinterpret_cast<double2 *>(x + head);
    const int64_t nv = (n - head) >> 1;
    for (int64_t i = tid; i < nv; i += nthr) v[i] = make_double2(0.0, 0.0);
    if (tid == 0 && ((n - head) & 1)) x[n - 1] = 0.0;
}

// Everything this rank has issued before (kernels whose epilogues added into peer images, peer copies) is complete
// when this kernel starts (stream order); the fence orders those writes before the flag as seen from the peer.
__global__ void k_signal(SyncArgs a) {
    const int t = threadIdx.x;
    if (t >= a.n) return;
    __threadfence_system();
    const unsigned long long v = *a.epoch * kEpochStride + a.seq[t];
    unsigned long long *dst = a.peer_flags[a.peer[t]] + a.chan * kMaxWorld + a.me;
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(dst), "l"(v) : "memory");
}

__global__ void k_wait(SyncArgs a, int *status, unsigned long long timeout_ns) {
    const int t = threadIdx.x;
    if (t >= a.n) return;
    const unsigned long long want = *a.epoch * kEpochStride + a.seq[t];
    const unsigned long long *src = a.my_flags + a.chan * kMaxWorld + a.peer[t];
    unsigned long long t0, now, v;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
    while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(src) : "memory");
        if (v >= want) break;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(now));
        if (now - t0 > timeout_ns) { atomicOr(status, 4); break; }      // a peer never arrived: report, do not hang the device
        __nanosleep(200);
    }
    __threadfence_system();
}

// rows x cols block, column stride ld, local image -> peer image.  16-byte accesses when both sides allow it.
__global__ void __launch_bounds__(256) k_push(const double *__restrict__ src, double *__restrict__ dst, int rows, int cols, int ld) {
    const bool vec = (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) && (ld & 1) == 0;
    const int rv = vec ? rows >> 1 : 0;                      // double2 per column
    const int64_t nvec = (int64_t)rv * cols, nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += nthr) {
        const int c = (int)(i / rv), r = (int)(i - (int64_t)c * rv) << 1;
        *reinterpret_cast<double2 *>(dst + (int64_t)c * ld + r) = *reinterpret_cast<const double2 *>(src + (int64_t)c * ld + r);
    }
    const int tail = rows - 2 * rv;                          // scalar rest of every column
    const int64_t nrest = (int64_t)tail * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrest; i += nthr) {
        const int c = (int)(i / tail), r = 2 * rv + (int)(i - (int64_t)c * tail);
        dst[(int64_t)c * ld + r] = src[(int64_t)c * ld + r];
    }
}

}  // namespace

// =================================================================================================================
// launchers
// =================================================================================================================
template <class Cfg, int BM, int BN, int WM, int WN, int ST, int BK>
cudaError_t gemm_optin() {
    return cudaFuncSetAttribute(k_gemm<BM, BN, WM, WN, ST, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
}

cudaError_t kernels_init() {
    cudaError_t e;
    if ((e = gemm_optin<Gemm128, 128, 128, 2, 4, 3, 16>()) != cudaSuccess) return e;
    if ((e = gemm_optin<Gemm128w16, 128, 128, 4, 4, 3, 16>()) != cudaSuccess) return e;
    if ((e = gemm_optin<Gemm128k32, 128, 128, 2, 4, 3, 32>()) != cudaSuccess) return e;
    if ((e = gemm_optin<Gemm128w16k32, 128, 128, 4, 4, 3, 32>()) != cudaSuccess) return e;
    if ((e = gemm_optin<Gemm64, 64, 64, 2, 2, 4, 16>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_gemm_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, ws::SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_trsm_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, tm::SMEM_BYTES)) != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_small<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_small<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    return e;
}

void launch_assemble(const double *vals, const int64_t *dst, int64_t count, double *X, int sms, cudaStream_t st) {
    if (count <= 0) return;
    const int block = 256;
    const int grid = (int)std::min<int64_t>((count + block - 1) / block, (int64_t)sms * 16);
    k_assemble<<<grid, block, 0, st>>>(vals, dst, count, X);
}

void launch_small(const PlanDev &P, const int32_t *tasks, int count, int smem, bool big_cta, cudaStream_t st) {
    if (count <= 0) return;
    if (P.is_lu) {
        if (big_cta) k_small<256, true><<<count, 256, smem, st>>>(P, tasks);
        else k_small<64, true><<<count, 64, smem, st>>>(P, tasks);
    } else {
        if (big_cta) k_small<256, false><<<count, 256, smem, st>>>(P, tasks);
        else k_small<64, false><<<count, 64, smem, st>>>(P, tasks);
    }
}

void launch_diag(const PlanDev &P, const DiagOp *ops, int count, cudaStream_t st) {
    if (count <= 0) return;
    if (P.is_lu) k_diag<true><<<count, 256, 0, st>>>(P, ops);
    else k_diag<false><<<count, 256, 0, st>>>(P, ops);
}

void launch_trsm(const PlanDev &P, const TrsmOp *ops, int count, int grid, cudaStream_t st) {
    if (count <= 0 || grid <= 0) return;
    k_trsm<<<grid, 64, 0, st>>>(P, ops, count);
}

void launch_trsm_mma(const PlanDev &P, const TrsmOp *ops, int count, int grid, cudaStream_t st) {
    if (count <= 0 || grid <= 0) return;
    k_trsm_mma<<<grid, 128, tm::SMEM_BYTES, st>>>(P, ops, count);
}

void launch_gemm(const PlanDev &P, const GemmOp *ops, int count, int grid, bool tile128, cudaStream_t st) {
    if (count <= 0 || grid <= 0) return;
    // SFB200_GEMM_VARIANT: 0 = 8 warps, BK 16 (default); 1 = 16 warps, BK 16; 2 = 8 warps, BK 32; 3 = 16 warps, BK 32
    static const int variant = getenv("SFB200_GEMM_VARIANT") ? atoi(getenv("SFB200_GEMM_VARIANT")) : 0;
    if (!tile128) {
        k_gemm<64, 64, 2, 2, 4, 16><<<grid, Gemm64::NT, Gemm64::SMEM_BYTES, st>>>(P, ops, count);
        return;
    }
    switch (variant) {
        case 1: k_gemm<128, 128, 4, 4, 3, 16><<<grid, Gemm128w16::NT, Gemm128w16::SMEM_BYTES, st>>>(P, ops, count); break;
        case 2: k_gemm<128, 128, 2, 4, 3, 32><<<grid, Gemm128k32::NT, Gemm128k32::SMEM_BYTES, st>>>(P, ops, count); break;
        case 3: k_gemm<128, 128, 4, 4, 3, 32><<<grid, Gemm128w16k32::NT, Gemm128w16k32::SMEM_BYTES, st>>>(P, ops, count); break;
        default: k_gemm<128, 128, 2, 4, 3, 16><<<grid, Gemm128::NT, Gemm128::SMEM_BYTES, st>>>(P, ops, count); break;
    }
}

void launch_gemm_ws(const PlanDev &P, const GemmOp *ops, int count, int tiles, int use_sms, cudaStream_t st) {
    if (count <= 0 || tiles <= 0) return;
    const int avail = use_sms > 8 ? use_sms : 8;      // look-ahead launches leave a few SMs to the critical path
    const int grid = tiles < avail ? tiles : avail;      // persistent: one CTA per SM walks the tile list
    k_gemm_ws<<<grid, ws::NT, ws::SMEM_BYTES, st>>>(P, ops, count, tiles);
}

void launch_epoch_inc(unsigned long long *epoch, cudaStream_t st) { k_epoch_inc<<<1, 1, 0, st>>>(epoch); }
void launch_zero(double *x, int64_t n, int *status, int sms, cudaStream_t st) { k_zero<<<sms * 8, 256, 0, st>>>(x, n, status); }
void launch_signal(const SyncArgs &a, cudaStream_t st) {
    if (a.n > 0) k_signal<<<1, 32, 0, st>>>(a);
}
void launch_wait(const SyncArgs &a, int *status, unsigned long long timeout_ns, cudaStream_t st) {
    if (a.n > 0) k_wait<<<1, 32, 0, st>>>(a, status, timeout_ns);
}
void launch_push(const double *src, double *dst, int rows, int cols, int ld, int ctas, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return;
    k_push<<<ctas > 0 ? ctas : 1, 256, 0, st>>>(src, dst, rows, cols, ld);
}

void launch_pack(const double *X, double *out, const PackTask *tasks, int count, cudaStream_t st) {
    if (count <= 0) return;
    k_pack<<<count, 256, 0, st>>>(X, out, tasks);
}

void launch_solve(const double *X, const int32_t *lsi, double *rhs, int64_t n, int nrhs, bool lu, const SolveTask *tasks,
                  int kind, int count, cudaStream_t st) {
    if (count <= 0) return;
    SolveArgs A{X, lsi, rhs, n, nrhs, lu ? 1 : 0};
    switch (kind) {
        case S_DIAG_F: k_solve_diag<<<count, 64, 0, st>>>(A, tasks, 0); break;
        case S_DIAG_B: k_solve_diag<<<count, 64, 0, st>>>(A, tasks, lu ? 2 : 1); break;
        case S_UPD_F:
            if (nrhs == 1) k_solve_upd_fwd<1><<<count, 256, 0, st>>>(A, tasks);
            else k_solve_upd_fwd<4><<<count, 256, 0, st>>>(A, tasks);
            break;
        case S_UPD_B:
            if (nrhs == 1) k_solve_upd_bwd<1><<<count, 256, 0, st>>>(A, tasks);
            else k_solve_upd_bwd<4><<<count, 256, 0, st>>>(A, tasks);
            break;
        case S_UPD_BU:
            if (nrhs == 1) k_solve_upd_bwd_u<1><<<count, 64, 0, st>>>(A, tasks);
            else k_solve_upd_bwd_u<4><<<count, 64, 0, st>>>(A, tasks);
            break;
        default: break;
    }
}

void launch_residual(const double *lx, const int64_t *li, const int32_t *lcol, int64_t nnz, const double *ux, const int32_t *urow,
                     const int64_t *ui, int64_t unnz, bool split, const double *x, const double *b, double *r, double *colsum,
                     int64_t n, unsigned long long *out4, int sms, cudaStream_t st) {
    const int block = 256, grid = sms * 8;
    k_axpby<<<grid, block, 0, st>>>(r, b, -1.0, n);
    cudaMemsetAsync(colsum, 0, sizeof(double) * n, st);
    cudaMemsetAsync(out4, 0, sizeof(unsigned long long) * 4, st);
    k_residual_lower<<<grid, block, 0, st>>>(lx, li, lcol, nnz, split ? 0 : 1, split ? 1 : 0, x, r, colsum);
    if (split) k_residual_upper<<<grid, block, 0, st>>>(ux, urow, ui, unnz, x, r, colsum);
    k_absmax4<<<grid, block, 0, st>>>(colsum, b, x, r, n, out4);
}

void launch_peak_dfma(double *out, int iters, int grid, int block, cudaStream_t st) {
    k_peak_dfma<<<grid, block, 0, st>>>(out, iters);
}
void launch_peak_dmma(double *out, int iters, int grid, int block, cudaStream_t st) {
    k_peak_dmma<<<grid, block, 0, st>>>(out, iters);
}
void launch_peak_dmma_tile(double *out, int iters, int grid, int mode, cudaStream_t st) {
    k_peak_dmma_tile<<<grid, 256, 0, st>>>(out, iters, mode);
}

}  // namespace sfb200
