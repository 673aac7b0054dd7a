// Decision probe for the multi-GPU exchange: how fast can one GPU add FP64 values into a
// PEER's memory over NVLink with red.global.add.f64 -- the instruction the fused scatter epilogue would issue -- compared
// with (a) the same atomics into local HBM, (b) plain coalesced peer reads / writes, (c) cudaMemcpyPeerAsync?
// One process, two devices with peer access (gpurun --gpus 2); standalone, not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/p2p_atomics scripts/probes/p2p_atomics.cu && /tmp/p2p_atomics
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; }       \
    } while (0)

// every thread adds `reps` values; stride 1 = fully coalesced columns (what the epilogue of an in-place block looks
// like), stride > 1 = the gather-like pattern of a relative-index scatter (elements of a column `stride` apart)
__global__ void k_red(double *dst, size_t n, int stride, int reps) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        const size_t i = ((tid + (size_t)r * nthr) * (size_t)stride) % n;
        atomicAdd(dst + i, 1.0);
    }
}
__global__ void k_read(const double *src, size_t n, double *sink) {
    const size_t nthr = (size_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += nthr) acc += src[i];
    if (acc == 12345.678) *sink = acc;
}
__global__ void k_write(double *dst, size_t n) {
    const size_t nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += nthr) dst[i] = 1.0;
}
// owner-pulls reduce: local[i] += peer[i], the alternative to NCCL reduce for a block column of a shared panel
__global__ void k_pull_add(double *local, const double *peer, size_t n) {
    const size_t nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += nthr) local[i] += peer[i];
}

template <class F>
static float timed(F f, int iters = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();                                   // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / iters;
}

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 devices (gpurun --gpus 2), found %d\n", ndev); return 1; }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (!can) { printf("no peer access between device 0 and 1\n"); return 1; }
    const size_t n = (size_t)1 << 28;      // 2 GiB of doubles on each device
    double *d0 = nullptr, *d1 = nullptr, *sink = nullptr;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&d1, n * 8)); CK(cudaMemset(d1, 0, n * 8));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&d0, n * 8)); CK(cudaMemset(d0, 0, n * 8)); CK(cudaMalloc(&sink, 8));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    const int grid = 148 * 8, block = 256, reps = 256;
    const double adds = (double)grid * block * reps;
    printf("%-44s %10s %12s\n", "case (kernels run on device 0)", "ms", "GB/s payload");
    for (int stride : {1, 2, 17, 1024}) {
        float ms = timed([&] { k_red<<<grid, block>>>(d0, n, stride, reps); });
        printf("red.add.f64 local HBM, stride %-4d             %10.3f %12.1f\n", stride, ms, adds * 8 / ms / 1e6);
        ms = timed([&] { k_red<<<grid, block>>>(d1, n, stride, reps); });
        printf("red.add.f64 PEER over NVLink, stride %-4d      %10.3f %12.1f\n", stride, ms, adds * 8 / ms / 1e6);
    }
    float ms = timed([&] { k_read<<<grid, block>>>(d1, n, sink); });
    printf("%-44s %10.3f %12.1f\n", "coalesced read of PEER memory", ms, n * 8.0 / ms / 1e6);
    ms = timed([&] { k_write<<<grid, block>>>(d1, n); });
    printf("%-44s %10.3f %12.1f\n", "coalesced write to PEER memory", ms, n * 8.0 / ms / 1e6);
    ms = timed([&] { k_pull_add<<<grid, block>>>(d0, d1, n); });
    printf("%-44s %10.3f %12.1f\n", "owner pulls: local[i] += peer[i]", ms, n * 8.0 / ms / 1e6);
    ms = timed([&] { cudaMemcpyPeerAsync(d0, 0, d1, 1, n * 8, 0); });
    printf("%-44s %10.3f %12.1f\n", "cudaMemcpyPeerAsync 1 -> 0", ms, n * 8.0 / ms / 1e6);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    printf("compare with the NCCL reduce / all-reduce / broadcast rates of scripts/mgpu_check.py on the same box\n");
    return 0;
}
