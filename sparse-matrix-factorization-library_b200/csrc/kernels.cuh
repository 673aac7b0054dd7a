// Device-side interface of the numeric engine: the argument block every kernel receives and the launchers
// the engine calls.  All kernels are hand-written for sm_100a; no cuBLAS / cuSOLVER anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan_types.h"

namespace sfb200 {

struct PlanDev {
    double *X;                       // this rank's image of the padded factor (see plan_types.h)
    double *const *peer;             // peer[r] = image of rank r as THIS device sees it (peer[rank] == X; one entry on a
                                     // single-GPU plan): scattered update entries go to the image upd_coloff names
    const SnodeDev *snode;
    const int64_t *upd_coloff;
    const int64_t *upd_relbase;
    const int32_t *upd_segb;
    const int32_t *upd_sege;
    const int32_t *upd_uoff;
    const int32_t *relmap;
    int *status;                     // bit 0: bad pivot met
    int is_lu;
    int multi;                       // partitioned plan: scattered adds may cross NVLink and race with other GPUs' (system scope)
    int dbg;                         // timing experiments only (SFB200_DBG): bit0 no staging copies, bit1 no barriers
};

// once per DEVICE (function attributes are per device): opt in to large dynamic shared memory for the kernels that need it
cudaError_t kernels_init();

void launch_assemble(const double *vals, const int64_t *dst, int64_t count, double *X, int sms, cudaStream_t st);
void launch_small(const PlanDev &P, const int32_t *tasks, int count, int smem, bool big_cta, cudaStream_t st);
void launch_diag(const PlanDev &P, const DiagOp *ops, int count, cudaStream_t st);
void launch_trsm(const PlanDev &P, const TrsmOp *ops, int count, int grid, cudaStream_t st);
void launch_trsm_mma(const PlanDev &P, const TrsmOp *ops, int count, int grid, cudaStream_t st);
void launch_gemm(const PlanDev &P, const GemmOp *ops, int count, int grid, bool tile128, cudaStream_t st);
void launch_gemm_ws(const PlanDev &P, const GemmOp *ops, int count, int tiles, int use_sms, cudaStream_t st);   // persistent grid = min(tiles, use_sms)
void launch_pack(const double *X, double *out, const PackTask *tasks, int count, cudaStream_t st);

// ---- exchange steps of a partitioned factorization: flags in peer memory ---------------------------------------------------
// flags[chan][src] in a rank's memory = value of the last signal rank `src` sent it on channel `chan`:
// epoch * kEpochStride + sequence number, where the epoch (a device-resident counter, bumped once per factorization, so that a
// captured graph can be replayed) and the sequence numbers (count of the pair's signals on that channel so far, from the
// plan's step list) are known to both sides without any host traffic.  One channel per signalling STREAM -- 0: barriers (main
// stream), 1: pushed blocks (push stream) -- so that the values of a slot arrive in issue order: a barrier signal can never
// overtake the flag of a push that is still in flight and stand in for it.
constexpr unsigned long long kEpochStride = 1ull << 24;
constexpr int kFlagSlots = 2 * kMaxWorld;      // then the epoch counter
struct SyncArgs {
    unsigned long long *const *peer_flags;     // peer_flags[r] = flag array of rank r as this device sees it
    unsigned long long *my_flags;              // == peer_flags[me]
    const unsigned long long *epoch;
    int me, n, chan;
    int peer[kMaxWorld];                       // n (peer, sequence number) pairs
    unsigned seq[kMaxWorld];
};
void launch_epoch_inc(unsigned long long *epoch, cudaStream_t st);
void launch_zero(double *x, int64_t n, int *status, int sms, cudaStream_t st);     // image = 0, *status = 0 (no cudaMemset: see kernels.cu)
void launch_signal(const SyncArgs &a, cudaStream_t st);
// spins until every listed peer's flag has reached its value; gives up after timeout_ns and sets bit 2 of *status
void launch_wait(const SyncArgs &a, int *status, unsigned long long timeout_ns, cudaStream_t st);
// rows x cols block (column stride ld) from src to dst (peer memory) with the SMs instead of the copy engines
void launch_push(const double *src, double *dst, int rows, int cols, int ld, int ctas, cudaStream_t st);

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
