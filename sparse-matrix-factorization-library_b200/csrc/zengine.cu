// Complex (isComplex) numeric factorization on the FP64 engine through the real embedding of complex_embed.hpp:
// the reference's zherk / zgemm / zpotrf / ztrsm path (CH.c:2060-2070, 2134-2145) and its zgemm / magma_zgetrf_nopiv / ztrsm
// path (LU.c:2569-2581, 2652-2663).  A zplan owns the plan of the embedded real structure; the only device code of its own is
// three element-wise kernels: row phases, value expansion, panel read-back.  Everything between them is the FP64 engine.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sfb200.h"
#include "capi_internal.hpp"
#include "complex_embed.hpp"
#include "engine_internal.hpp"

using namespace sfb200;

struct sfb200_zplan {
    // structure of the complex matrix that the read-back and the solve need
    Kind kind = Kind::Cholesky;
    i64 n = 0, nnz = 0, unnz = 0, xsize = 0;     // complex entries: Lx, Ux (0 unless LU of an unsymmetric matrix), Lsx
    bool split = false;                          // LU of an unsymmetric matrix: separate Ux
    bool u_from_l = false;
    bool row_phase = false;
    sfb200_symbolic rsym;                        // the embedded real structure (emptied once the plan is built)
    sfb200_plan *rp = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    i64 nsrc_l = 0, nsrc_u = 0;
    int ntasks = 0;
    double *d_zl = nullptr, *d_zu = nullptr;     // complex values as uploaded
    i64 *d_src_l = nullptr, *d_src_u = nullptr, *d_diagpos = nullptr, *d_lsi = nullptr;
    int32_t *d_row_l = nullptr, *d_row_u = nullptr;
    double *d_phase = nullptr, *d_out = nullptr;
    ZExtractTask *d_tasks = nullptr;
    std::vector<double> h_phase;                 // phases of the last factorization (solve rotates the right-hand side)
    bool factored = false;
    float ms_numeric = 0.f;
    size_t device_bytes = 0;
};

namespace {

#define ZCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
            return SFB200_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

// d_k = a_kk / |a_kk| from the diagonal entry at pos[k] of the interleaved values z (1 where absent or zero)
__global__ void k_zphase(const double *__restrict__ z, const i64 *__restrict__ pos, i64 n, double *__restrict__ phase) {
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < n; k += (i64)gridDim.x * blockDim.x) {
        double re = 1.0, im = 0.0;
        const i64 p = pos[k];
        if (p >= 0) {
            const double a = z[2 * p], b = z[2 * p + 1];
            const double m = hypot(a, b);
            if (m > 0.0) { re = a / m; im = b / m; }
        }
        phase[2 * k] = re;
        phase[2 * k + 1] = im;
    }
}

// values of the embedded matrix: out[q] = +-(re | im) of complex entry src[q] >> 2, rotated by conj(d_row) when phase != null
// (complex_embed.cpp: expand_values is the same loop on the host)
__global__ void k_zexpand(const double *__restrict__ z, const i64 *__restrict__ src, const int32_t *__restrict__ rowz,
                          const double *__restrict__ phase, i64 count, double *__restrict__ out) {
    for (i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x; q < count; q += (i64)gridDim.x * blockDim.x) {
        const i64 c = src[q];
        double v = 0.0;
        if (c >= 0) {
            const i64 p = c >> 2;
            double re = z[2 * p], im = z[2 * p + 1];
            if (phase) {
                const i64 r = rowz[p];
                const double dr = phase[2 * r], di = phase[2 * r + 1];
                const double tr = dr * re + di * im;
                im = dr * im - di * re;
                re = tr;
            }
            v = (c & 1) ? im : re;
            if (c & 2) v = -v;
        }
        out[q] = v;
    }
}

// complex panels out of the image of the embedded factor, one task per CTA (complex_embed.cpp: zextract_host is the same
// arithmetic on the host and states where every number comes from)
__global__ void k_zextract(const double *__restrict__ X, const ZExtractTask *__restrict__ tasks, const double *__restrict__ phase,
                           const i64 *__restrict__ lsi, double *__restrict__ out) {
    const ZExtractTask t = tasks[blockIdx.x];
    const double *F = X + t.src;
    const i64 ld = t.ld;
    const int slda = t.mode ? 2 * t.nsrow - t.nscol : t.nsrow;
    const i64 total = (i64)t.ncols * slda;
    const i64 *rows = lsi + t.lsip;
    for (i64 e = threadIdx.x; e < total; e += blockDim.x) {
        const int k = t.c0 + (int)(e / slda);
        const int r = (int)(e % slda);
        double re, im;
        i64 prow = -1, pcol = -1;
        if (!t.mode) {
            re = r < k ? 0.0 : F[2 * (i64)r + (2 * (i64)k) * ld];
            im = r <= k ? 0.0 : F[2 * (i64)r + 1 + (2 * (i64)k) * ld];
        } else if (r < t.nsrow && r > k) {
            re = F[2 * (i64)r + 1 + (2 * (i64)k + 1) * ld];
            im = -F[2 * (i64)r + (2 * (i64)k + 1) * ld];
            if (phase) { prow = rows[r]; pcol = rows[k]; }
        } else if (r < t.nsrow) {
            re = F[2 * (i64)r + (2 * (i64)k) * ld];
            im = -F[2 * (i64)r + (2 * (i64)k + 1) * ld];
            if (phase) prow = rows[r];
        } else {
            const i64 tt = r - t.nsrow;
            re = F[2 * (i64)t.nsrow + 2 * tt + (2 * (i64)k) * ld];
            im = -F[2 * (i64)t.nsrow + 2 * tt + 1 + (2 * (i64)k) * ld];
            if (phase) prow = rows[k];
        }
        if (prow >= 0) {
            double dr = phase[2 * prow], di = phase[2 * prow + 1];
            if (pcol >= 0) {
                const double cr = phase[2 * pcol], ci = phase[2 * pcol + 1];
                const double tr = dr * cr + di * ci, ti = di * cr - dr * ci;
                dr = tr; di = ti;
            }
            const double tr = re * dr - im * di, ti = re * di + im * dr;
            re = tr; im = ti;
        }
        double *o = out + 2 * (t.dst + (i64)k * slda + r);
        o[0] = re;
        o[1] = im;
    }
}

template <class T>
int zupload(sfb200_zplan *zp, T *&dptr, const std::vector<T> &h) {
    dptr = nullptr;
    if (h.empty()) return SFB200_OK;
    ZCK(cudaMalloc((void **)&dptr, h.size() * sizeof(T)));
    zp->device_bytes += h.size() * sizeof(T);
    ZCK(cudaMemcpyAsync(dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, zp->stream));
    return SFB200_OK;
}

int zalloc(sfb200_zplan *zp, double *&dptr, size_t doubles) {
    dptr = nullptr;
    ZCK(cudaMalloc((void **)&dptr, sizeof(double) * (doubles > 0 ? doubles : 1)));
    zp->device_bytes += sizeof(double) * doubles;
    return SFB200_OK;
}

inline int grid_for(i64 count, int sms) {
    const i64 want = (count + 255) / 256;
    const i64 cap = (i64)std::max(sms, 1) * 8;
    return (int)std::max<i64>(1, std::min(want, cap));
}

// phases, expansion into the embedded plan's own value arrays, the embedded factorization: all on zp->stream, no waiting
int zissue_numeric(sfb200_zplan *zp) {
    const PlanInternals in = plan_internals(zp->rp);
    const double *phase = nullptr;
    if (zp->row_phase && zp->kind == Kind::LU && zp->n > 0) {
        const double *zdiag = zp->split ? zp->d_zu : zp->d_zl;
        k_zphase<<<grid_for(zp->n, in.sms), 256, 0, zp->stream>>>(zdiag, zp->d_diagpos, zp->n, zp->d_phase);
        phase = zp->d_phase;
    }
    if (zp->nsrc_l > 0)
        k_zexpand<<<grid_for(zp->nsrc_l, in.sms), 256, 0, zp->stream>>>(zp->d_zl, zp->d_src_l, zp->d_row_l, phase, zp->nsrc_l, in.d_lx);
    if (zp->kind == Kind::LU && zp->nsrc_u > 0)
        k_zexpand<<<grid_for(zp->nsrc_u, in.sms), 256, 0, zp->stream>>>(zp->u_from_l ? zp->d_zl : zp->d_zu, zp->d_src_u, zp->d_row_u, phase,
                                                                       zp->nsrc_u, in.d_ux);
    ZCK(cudaGetLastError());
    const int rc = sfb200_factorize_device(zp->rp, in.d_lx, in.d_ux);
    if (rc) return rc;
    if (phase) {
        zp->h_phase.resize(2 * (size_t)zp->n);
        ZCK(cudaMemcpyAsync(zp->h_phase.data(), zp->d_phase, sizeof(double) * 2 * (size_t)zp->n, cudaMemcpyDeviceToHost, zp->stream));
    } else {
        zp->h_phase.clear();
    }
    return SFB200_OK;
}

int zissue_download(sfb200_zplan *zp, double *Lsz) {
    if (!zp->d_out) { const int rc = zalloc(zp, zp->d_out, 2 * (size_t)zp->xsize); if (rc) return rc; }
    if (zp->ntasks > 0) {
        const PlanInternals in = plan_internals(zp->rp);
        const double *phase = zp->h_phase.empty() ? nullptr : zp->d_phase;
        k_zextract<<<zp->ntasks, 256, 0, zp->stream>>>(in.d_X, zp->d_tasks, phase, zp->d_lsi, zp->d_out);
        ZCK(cudaGetLastError());
    }
    ZCK(cudaMemcpyAsync(Lsz, zp->d_out, sizeof(double) * 2 * (size_t)zp->xsize, cudaMemcpyDeviceToHost, zp->stream));
    return SFB200_OK;
}

}  // namespace

extern "C" {

void sfb200_zplan_free(sfb200_zplan *zp) {
    if (!zp) return;
    cudaSetDevice(zp->device);
    if (zp->rp) sfb200_plan_free(zp->rp);           // while the borrowed stream is still alive
    void *ptrs[] = {zp->d_zl, zp->d_zu, zp->d_src_l, zp->d_src_u, zp->d_diagpos, zp->d_lsi, zp->d_row_l, zp->d_row_u,
                    zp->d_phase, zp->d_out, zp->d_tasks};
    for (void *q : ptrs) if (q) cudaFree(q);
    for (auto &e : zp->ev) if (e) cudaEventDestroy(e);
    if (zp->stream) cudaStreamDestroy(zp->stream);
    delete zp;
}

int sfb200_zplan_create(const sfb200_symbolic *zs, int device, sfb200_zplan **out) {
    if (!zs || !out) { set_error("sfb200_zplan_create: bad argument"); return SFB200_ERR_ARG; }
    const Symbolic &Z = zs->S;
    if (!Z.is_complex) { set_error("sfb200_zplan_create: the structure is not that of a complex matrix (sfb200_analyze_complex / sfb200_symbolic_wrap_complex)"); return SFB200_ERR_ARG; }
    if (sfb200_device_count() <= device || device < 0) {
        set_error("sfb200_zplan_create: no usable CUDA device (this engine has no CPU fallback)");
        return SFB200_ERR_NO_DEVICE;
    }
    auto *zp = new sfb200_zplan();
    auto fail = [&](int code) { sfb200_zplan_free(zp); return code; };
    zp->device = device;
    zp->kind = Z.kind;
    zp->n = Z.n;
    zp->nnz = Z.L.ptr.empty() ? 0 : Z.L.ptr[Z.n];
    zp->split = Z.kind == Kind::LU && !Z.symmetric;
    zp->unnz = zp->split ? Z.U.ptr[Z.n] : 0;
    zp->xsize = Z.xsize;
    zp->row_phase = Z.kind == Kind::LU;
    if (const char *e = getenv("SFB200_ZPHASE")) zp->row_phase = zp->row_phase && atoi(e) != 0;
    ComplexEmbedding E;
    if (embed_complex(Z, E) != 0) { set_error("sfb200_zplan_create: not a valid supernodal structure of a complex matrix"); return fail(SFB200_ERR_ARG); }
    zp->u_from_l = E.u_from_l;
    zp->nsrc_l = (i64)E.src_l.size();
    zp->nsrc_u = (i64)E.src_u.size();
    std::vector<i64> diagpos;
    diagonal_positions(Z, diagpos);
    zp->rsym.S = std::move(E.R);
    int rc = sfb200_plan_create(&zp->rsym, device, &zp->rp);
    if (rc) return fail(rc);
    const PlanInternals in = plan_internals(zp->rp);
    std::vector<ZExtractTask> tasks;
    build_zextract_tasks(*in.H, Z, tasks);
    zp->ntasks = (int)tasks.size();
    zp->rsym.S = Symbolic();                        // the plan keeps what it needs
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return fail(SFB200_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&zp->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream"); return fail(SFB200_ERR_CUDA); }
    for (auto &e : zp->ev) if (cudaEventCreate(&e) != cudaSuccess) { set_error("event"); return fail(SFB200_ERR_CUDA); }
    if ((rc = sfb200_plan_set_stream(zp->rp, (void *)zp->stream))) return fail(rc);
    if ((rc = zupload(zp, zp->d_src_l, E.src_l))) return fail(rc);
    if ((rc = zupload(zp, zp->d_src_u, E.src_u))) return fail(rc);
    if ((rc = zupload(zp, zp->d_row_l, E.row_l))) return fail(rc);
    if ((rc = zupload(zp, zp->d_row_u, E.row_u))) return fail(rc);
    if ((rc = zupload(zp, zp->d_diagpos, diagpos))) return fail(rc);
    if ((rc = zupload(zp, zp->d_lsi, Z.Lsi))) return fail(rc);
    if ((rc = zupload(zp, zp->d_tasks, tasks))) return fail(rc);
    if ((rc = zalloc(zp, zp->d_zl, 2 * (size_t)zp->nnz))) return fail(rc);
    if (zp->split && (rc = zalloc(zp, zp->d_zu, 2 * (size_t)zp->unnz))) return fail(rc);
    if ((rc = zalloc(zp, zp->d_phase, 2 * (size_t)zp->n))) return fail(rc);
    if (cudaStreamSynchronize(zp->stream) != cudaSuccess) { set_error("zplan upload failed"); return fail(SFB200_ERR_CUDA); }
    *out = zp;
    return SFB200_OK;
}

sfb200_plan *sfb200_zplan_real(sfb200_zplan *zp) { return zp ? zp->rp : nullptr; }

int64_t sfb200_zplan_int(const sfb200_zplan *zp, const char *name) {
    if (!zp || !name) return -1;
    const std::string k(name);
    if (k == "n") return zp->n;
    if (k == "nnz") return zp->nnz;
    if (k == "unnz") return zp->unnz;
    if (k == "xsize") return zp->xsize;
    if (k == "split") return zp->split;
    if (k == "row_phase") return zp->row_phase;
    if (k == "extract_tasks") return zp->ntasks;
    if (k == "device_bytes") return (int64_t)zp->device_bytes + sfb200_plan_int(zp->rp, "device_bytes");
    return sfb200_plan_int(zp->rp, name);           // "nlevels", "nlaunches", "xsize_padded", ... of the embedded plan
}

int sfb200_zplan_set_row_phase(sfb200_zplan *zp, int on) {
    if (!zp) return SFB200_ERR_ARG;
    zp->row_phase = on != 0 && zp->kind == Kind::LU;
    return SFB200_OK;
}

int sfb200_zfactorize_device(sfb200_zplan *zp, const double *d_Lz, const double *d_Uz) {
    if (!zp || !d_Lz || (zp->split && !d_Uz)) { set_error("sfb200_zfactorize_device: bad argument"); return SFB200_ERR_ARG; }
    ZCK(cudaSetDevice(zp->device));
    if (d_Lz != zp->d_zl) ZCK(cudaMemcpyAsync(zp->d_zl, d_Lz, sizeof(double) * 2 * (size_t)zp->nnz, cudaMemcpyDeviceToDevice, zp->stream));
    if (zp->split && d_Uz != zp->d_zu) ZCK(cudaMemcpyAsync(zp->d_zu, d_Uz, sizeof(double) * 2 * (size_t)zp->unnz, cudaMemcpyDeviceToDevice, zp->stream));
    ZCK(cudaEventRecord(zp->ev[0], zp->stream));
    const int rc = zissue_numeric(zp);
    if (rc) return rc;
    ZCK(cudaEventRecord(zp->ev[1], zp->stream));
    zp->factored = true;
    return SFB200_OK;
}

int sfb200_zplan_sync(sfb200_zplan *zp) {
    if (!zp) return SFB200_ERR_ARG;
    ZCK(cudaSetDevice(zp->device));
    const int rc = sfb200_plan_sync(zp->rp);        // waits for the stream, reads the pivot status
    if (rc) zp->factored = false;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, zp->ev[0], zp->ev[1]) == cudaSuccess) zp->ms_numeric = ms;
    else cudaGetLastError();
    return rc;
}

int sfb200_zfactorize(sfb200_zplan *zp, const double *Lz, const double *Uz, double *Lsz) {
    if (!zp || !Lz || (zp->split && !Uz)) { set_error("sfb200_zfactorize: bad argument (an unsymmetric LU needs Uz)"); return SFB200_ERR_ARG; }
    ZCK(cudaSetDevice(zp->device));
    ZCK(cudaMemcpyAsync(zp->d_zl, Lz, sizeof(double) * 2 * (size_t)zp->nnz, cudaMemcpyHostToDevice, zp->stream));
    if (zp->split) ZCK(cudaMemcpyAsync(zp->d_zu, Uz, sizeof(double) * 2 * (size_t)zp->unnz, cudaMemcpyHostToDevice, zp->stream));
    int rc = sfb200_zfactorize_device(zp, zp->d_zl, zp->d_zu);
    if (rc) return rc;
    if (Lsz && (rc = zissue_download(zp, Lsz))) return rc;
    return sfb200_zplan_sync(zp);
}

int sfb200_zdownload_factor(sfb200_zplan *zp, double *Lsz) {
    if (!zp || !Lsz) { set_error("sfb200_zdownload_factor: bad argument"); return SFB200_ERR_ARG; }
    if (!zp->factored) { set_error("sfb200_zdownload_factor: nothing factorized yet"); return SFB200_ERR_STATE; }
    ZCK(cudaSetDevice(zp->device));
    const int rc = zissue_download(zp, Lsz);
    if (rc) return rc;
    ZCK(cudaStreamSynchronize(zp->stream));
    return SFB200_OK;
}

/* A x = b for nrhs complex right-hand sides (interleaved, leading dimension n, permuted ordering): the embedded factor on
 * the interleaved vector is the complex substitution.  Rotated rows: (D^H A) x = D^H b. */
int sfb200_zsolve(sfb200_zplan *zp, double *Bz, int nrhs) {
    if (!zp || !Bz || nrhs < 1) { set_error("sfb200_zsolve: bad argument"); return SFB200_ERR_ARG; }
    if (!zp->factored) { set_error("sfb200_zsolve: nothing factorized yet"); return SFB200_ERR_STATE; }
    if (!zp->h_phase.empty())
        for (int c = 0; c < nrhs; ++c)
            for (i64 k = 0; k < zp->n; ++k) {
                double *b = Bz + 2 * ((i64)c * zp->n + k);
                const double dr = zp->h_phase[2 * k], di = zp->h_phase[2 * k + 1], re = b[0], im = b[1];
                b[0] = dr * re + di * im;
                b[1] = dr * im - di * re;
            }
    return sfb200_solve(zp->rp, Bz, nrhs);
}

/* Residual of the embedded (row-rotated) system the factorization saw, with x and b as interleaved complex vectors:
 * out[5] as sfb200_residual, norms over the real components. */
int sfb200_zresidual(sfb200_zplan *zp, const double *xz, const double *bz, double out[5]) {
    if (!zp || !xz || !bz || !out) { set_error("sfb200_zresidual: bad argument"); return SFB200_ERR_ARG; }
    if (zp->h_phase.empty()) return sfb200_residual(zp->rp, xz, bz, out);
    std::vector<double> b(bz, bz + 2 * (size_t)zp->n);
    for (i64 k = 0; k < zp->n; ++k) {
        const double dr = zp->h_phase[2 * k], di = zp->h_phase[2 * k + 1], re = b[2 * k], im = b[2 * k + 1];
        b[2 * k] = dr * re + di * im;
        b[2 * k + 1] = dr * im - di * re;
    }
    return sfb200_residual(zp->rp, xz, b.data(), out);
}

double sfb200_zlast_ms(const sfb200_zplan *zp) { return zp ? (double)zp->ms_numeric : -1.0; }

}  // extern "C"

/* ---- complex matrices on several GPUs ------------------------------------------------------------------------------------
 * Pure composition of pieces that exist: the etree partition of the EMBEDDED structure over `world` GPUs is an ordinary
 * sfb200_group (peer-memory exchange, one host thread per GPU); the complex values are expanded on the host (the arrays are
 * the size of the matrix, not of the factor), the group writes the embedded factor into a page-locked staging buffer in the
 * reference's layout, and the complex panels are read out of it by zextract_host (complex_embed.cpp) on all host threads.
 * No device code of its own.  Costs a host copy of the embedded factor (4 x xsize doubles); buys the distributed memory of
 * the group: a complex matrix whose embedded factor does not fit one GPU. */
struct sfb200_zgroup {
    Kind kind = Kind::Cholesky;
    i64 n = 0, nnz = 0, unnz = 0, xsize = 0, rxsize = 0;
    bool split = false, u_from_l = false, row_phase = false;
    std::vector<i64> src_l, src_u, diagpos, lsi;
    std::vector<int32_t> row_l, row_u;
    std::vector<ZExtractTask> tasks;             // per supernode, offsets in the REFERENCE layout of the embedded factor
    sfb200_symbolic rsym;
    sfb200_group *rg = nullptr;
    std::vector<double> lx, ux, phase;
    double *h_lsx = nullptr;
    bool h_pinned = false;
};

extern "C" {

void sfb200_zgroup_free(sfb200_zgroup *g) {
    if (!g) return;
    if (g->rg) sfb200_group_free(g->rg);
    if (g->h_lsx) { if (g->h_pinned) sfb200_host_free(g->h_lsx); else free(g->h_lsx); }
    delete g;
}

int sfb200_zgroup_create(const sfb200_symbolic *zs, int world, const int *devices, sfb200_zgroup **out) {
    if (!zs || !out || !devices || world < 1) { set_error("sfb200_zgroup_create: bad argument"); return SFB200_ERR_ARG; }
    const Symbolic &Z = zs->S;
    if (!Z.is_complex) { set_error("sfb200_zgroup_create: the structure is not that of a complex matrix"); return SFB200_ERR_ARG; }
    auto *g = new sfb200_zgroup();
    g->kind = Z.kind;
    g->n = Z.n;
    g->nnz = Z.L.ptr.empty() ? 0 : Z.L.ptr[Z.n];
    g->split = Z.kind == Kind::LU && !Z.symmetric;
    g->unnz = g->split ? Z.U.ptr[Z.n] : 0;
    g->xsize = Z.xsize;
    g->row_phase = Z.kind == Kind::LU;
    if (const char *e = getenv("SFB200_ZPHASE")) g->row_phase = g->row_phase && atoi(e) != 0;
    ComplexEmbedding E;
    if (embed_complex(Z, E) != 0) { delete g; set_error("sfb200_zgroup_create: not a valid supernodal structure of a complex matrix"); return SFB200_ERR_ARG; }
    g->u_from_l = E.u_from_l;
    g->src_l.swap(E.src_l); g->src_u.swap(E.src_u); g->row_l.swap(E.row_l); g->row_u.swap(E.row_u);
    diagonal_positions(Z, g->diagpos);
    g->lsi = Z.Lsi;
    g->rxsize = E.R.xsize;
    const bool lu = Z.kind == Kind::LU;
    for (i64 s = 0; s < Z.nsuper; ++s) {
        ZExtractTask t{};
        t.nscol = (int32_t)(Z.Super[s + 1] - Z.Super[s]);
        t.nsrow = (int32_t)(Z.Lsip[s + 1] - Z.Lsip[s]);
        t.src = E.R.Lsxp[s];
        t.dst = Z.Lsxp[s];
        t.lsip = Z.Lsip[s];
        t.c0 = 0; t.ncols = t.nscol;
        t.ld = lu ? 4 * t.nsrow - 2 * t.nscol : 2 * t.nsrow;          // no padding in the reference layout
        t.mode = lu ? 1 : 0;
        g->tasks.push_back(t);
    }
    g->lx.resize(g->src_l.size());
    g->ux.resize(g->src_u.size());
    g->rsym.S = std::move(E.R);
    const int rc = sfb200_group_create(&g->rsym, world, devices, &g->rg);
    g->rsym.S = Symbolic();
    if (rc) { sfb200_zgroup_free(g); return rc; }
    *out = g;
    return SFB200_OK;
}

sfb200_group *sfb200_zgroup_real(sfb200_zgroup *g) { return g ? g->rg : nullptr; }

/* Host buffers: Lz / Uz interleaved complex values (as for sfb200_zfactorize), Lsz = xsize Complex or NULL (the embedded factor
 * then stays on the devices). */
int sfb200_zgroup_factorize(sfb200_zgroup *g, const double *Lz, const double *Uz, double *Lsz) {
    if (!g || !Lz || (g->split && !Uz)) { set_error("sfb200_zgroup_factorize: bad argument (an unsymmetric LU needs Uz)"); return SFB200_ERR_ARG; }
    const double *phase = nullptr;
    if (g->row_phase && g->kind == Kind::LU) {
        g->phase.resize(2 * (size_t)g->n);
        row_phases(g->diagpos, g->split ? Uz : Lz, g->phase.data());
        phase = g->phase.data();
    } else {
        g->phase.clear();
    }
    expand_values(g->src_l, Lz, g->row_l.data(), phase, g->lx.data());
    if (g->kind == Kind::LU) expand_values(g->src_u, g->u_from_l ? Lz : Uz, g->row_u.data(), phase, g->ux.data());
    if (Lsz && !g->h_lsx) {
        const size_t bytes = sizeof(double) * (size_t)std::max<i64>(g->rxsize, 1);
        g->h_lsx = (double *)sfb200_host_alloc(bytes);
        g->h_pinned = g->h_lsx != nullptr;
        if (!g->h_lsx) g->h_lsx = (double *)malloc(bytes);
        if (!g->h_lsx) { set_error("sfb200_zgroup_factorize: cannot allocate the staging buffer of the embedded factor"); return SFB200_ERR_ARG; }
    }
    const int rc = sfb200_group_factorize(g->rg, g->lx.data(), g->kind == Kind::LU ? g->ux.data() : nullptr, Lsz ? g->h_lsx : nullptr);
    if (rc || !Lsz) return rc;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned w = 0; w < hw; ++w)
        th.emplace_back([&, w]() {
            for (size_t k = w; k < g->tasks.size(); k += hw) zextract_host(g->tasks[k], g->h_lsx, Lsz, phase, g->lsi.data());
        });
    for (auto &t : th) t.join();
    return SFB200_OK;
}

double sfb200_zgroup_last_ms(const sfb200_zgroup *g, int which) { return g ? sfb200_group_last_ms(g->rg, which) : -1.0; }

}  // extern "C"
