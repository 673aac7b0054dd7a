// Device-side interface of the numeric engine: the argument block every kernel receives and the launchers
// the engine calls.  All kernels are hand-written for sm_100a; no cuBLAS / cuSOLVER anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan_types.h"

namespace sfb200 {

struct PlanDev {
    double *X;                       // padded factor (see plan_types.h)
    const SnodeDev *snode;
    const int64_t *upd_coloff;
    const int64_t *upd_relbase;
    const int32_t *upd_segb;
    const int32_t *upd_sege;
    const int32_t *upd_uoff;
    const int32_t *relmap;
    int *status;                     // bit 0: bad pivot met
    int is_lu;
    int dbg;                         // timing experiments only (SFB200_DBG): bit0 no staging copies, bit1 no barriers
};

// once per DEVICE (function attributes are per device): opt in to large dynamic shared memory for the kernels that need it
cudaError_t kernels_init();

void launch_assemble(const double *vals, const int64_t *dst, int64_t count, double *X, int sms, cudaStream_t st);
void launch_small(const PlanDev &P, const int32_t *tasks, int count, int smem, bool big_cta, cudaStream_t st);
void launch_diag(const PlanDev &P, const DiagOp *ops, int count, cudaStream_t st);
void launch_trsm(const PlanDev &P, const TrsmOp *ops, int count, int grid, cudaStream_t st);
void launch_gemm(const PlanDev &P, const GemmOp *ops, int count, int grid, bool tile128, cudaStream_t st);
void launch_gemm_ws(const PlanDev &P, const GemmOp *ops, int count, int tiles, int use_sms, cudaStream_t st);   // persistent grid = min(tiles, use_sms)
void launch_pack(const double *X, double *out, const PackTask *tasks, int count, cudaStream_t st);

// device-resident triangular solve / residual
void launch_solve(const double *X, const int32_t *lsi, double *rhs, int64_t n, int nrhs, bool lu, const SolveTask *tasks,
                  int kind, int count, cudaStream_t st);
void launch_residual(const double *lx, const int64_t *li, const int32_t *lcol, int64_t nnz, const double *ux, const int32_t *urow,
                     const int64_t *ui, int64_t unnz, bool split, const double *x, const double *b, double *r, double *colsum,
                     int64_t n, unsigned long long *out4, int sms, cudaStream_t st);

// micro-benchmarks of the two FP64 pipes (peak measurement for the roofline denominators)
void launch_peak_dfma(double *out, int iters, int grid, int block, cudaStream_t st);
void launch_peak_dmma(double *out, int iters, int grid, int block, cudaStream_t st);
void launch_peak_dmma_tile(double *out, int iters, int grid, int mode, cudaStream_t st);

}  // namespace sfb200
