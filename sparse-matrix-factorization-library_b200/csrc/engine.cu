// Numeric half of the C ABI (include/sfb200.h): device-resident plan, schedule replay, copies.
// Replaces the numeric phase of SparseFrame_factorize_supernodal (CH.c:2150-3017 / LU.c:2668-3573): no host-driven
// per-descendant loop, no PCIe staging of panels, no cuBLAS / cuSOLVER -- a fixed sequence of batched launches over
// a factor that never leaves HBM until the final copy-out.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/sfb200.h"
#include "capi_internal.hpp"
#include "engine_internal.hpp"
#include "kernels.cuh"
#include "plan.hpp"

using namespace sfb200;

struct sfb200_plan {
    HostPlan H;
    int device = 0;
    int sms = 0;                               // multiprocessors of `device` (grid sizing of the persistent / grid-stride kernels)
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    // per-launch-kind profile of the last sfb200_plan_profile() call: ms, launches, algorithmic flops, bytes
    double prof_ms[L_KINDS] = {0}, prof_flops[L_KINDS] = {0}, prof_bytes[L_KINDS] = {0};
    int64_t prof_launches[L_KINDS] = {0};
    std::vector<float> prof_launch_ms;     // per launch of the schedule
    std::vector<double> prof_launch_flops;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // numeric begin/end, e2e begin/end
    cudaEvent_t ev_solve[2] = {nullptr, nullptr};               // around the last solve (created on first use)
    PlanDev D{};
    // device copies of the plan
    SnodeDev *d_snode = nullptr;
    int64_t *d_coloff = nullptr, *d_relbase = nullptr, *d_asm_l = nullptr, *d_asm_u = nullptr;
    int32_t *d_segb = nullptr, *d_sege = nullptr, *d_uoff = nullptr, *d_relmap = nullptr, *d_small = nullptr;
    DiagOp *d_diag = nullptr;
    TrsmOp *d_trsm = nullptr;
    GemmOp *d_gemm = nullptr;
    PackTask *d_pack = nullptr;
    double *d_X = nullptr, *d_out = nullptr, *d_lx = nullptr, *d_ux = nullptr;
    // triangular solve / residual
    int32_t *d_lsi32 = nullptr, *d_nz_col = nullptr, *d_unz_row = nullptr;
    int64_t *d_nz_row = nullptr, *d_unz_col = nullptr;
    SolveTask *d_solve = nullptr;
    double *d_rhs = nullptr, *d_work = nullptr;     // d_work: b copy, r, column sums (3 n)
    size_t rhs_capacity = 0;
    unsigned long long *d_norms = nullptr;
    float ms_solve = 0.f;
    int *d_status = nullptr;
    // look-ahead: side stream for the REST part of the top-level trailing updates
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_to_side = nullptr, ev_from_side = nullptr;
    bool side_pending = false;
    int spare_sms = 16;
    // leaf look-ahead: diagonal blocks flagged LF_DIAG_AHEAD are factorized on their own (high-priority) stream
    cudaStream_t diag_stream = nullptr;
    cudaEvent_t ev_to_diag = nullptr, ev_from_diag = nullptr;
    bool diag_pending = false;
    int leave_sms = 2;
    // The schedule is static, so each replayed piece (the whole device-resident factorization, the whole streamed
    // factorization into one destination, one multi-GPU segment) is captured once as a CUDA graph and launched from then
    // on: kernel-to-kernel gaps and the cross-stream dependencies of the look-ahead become graph edges.
    // (The first replay of a piece is issued plainly -- one-shot callers never pay for capture + instantiation, and lazily
    // created helper objects exist by the time the second replay is captured.)
    struct CachedGraph { cudaGraphExec_t exec; int64_t launches; };
    std::map<std::tuple<int, size_t, size_t, const void *>, CachedGraph> graphs;
    bool use_graphs = true, graphs_wanted = true;
    // ranks of one process (sfb200_group): called right before a freshly instantiated graph is launched, so that every rank
    // has finished capturing / instantiating (which may allocate) before the first one starts waiting on the device
    std::function<void()> before_first_graph_launch;
    // streamed copy-out (host-buffer call with a page-locked Lsx): copy stream, per-level work lists
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    size_t copy_events_used = 0;
    cudaEvent_t copy_done = nullptr;
    PackTask *d_spack = nullptr;                       // pack tasks of the SMALL panels, grouped by level
    std::vector<std::pair<int, int>> spack_range;      // per level: [first, count) in d_spack
    std::vector<std::vector<i64>> big_by_level;        // per level: supernodes copied by the DMA engine
    // small panels are first packed on the device into the dense reference layout (d_out), then leave as contiguous runs
    // of the dense buffer: per level, [offset, count) element ranges that become final at that level.  All levels up to
    // `merge_level` are flushed together (their supernodes are neighbours in the postorder, so the runs are long).
    std::vector<std::vector<std::pair<i64, i64>>> small_runs;     // per flush level: (first supernode of the run, elements)
    int merge_level = 0;
    int top_width = 256;                               // width of the widest blocking level
    // ---- partitioned plans (world > 1): exchange over peer memory, see plan_types.h (Coll) ---------------------------
    std::vector<double *> peer_img;                    // image of every rank as this device addresses it
    std::vector<unsigned long long *> peer_flags;      // flag array of every rank
    double **d_peer = nullptr;                         // device copies of the two tables
    unsigned long long **d_peer_flags = nullptr;
    unsigned long long *d_flags = nullptr;             // this rank's flags: [2][kMaxWorld] signal values, then the epoch counter
    unsigned long long *d_epoch = nullptr;
    std::vector<void *> ipc_mapped;                    // mappings opened with cudaIpcOpenMemHandle
    double *pending_out = nullptr;                     // pageable destination of a host-buffer call, filled in its second half
    bool pending_packed = false;
    double *h_stage = nullptr;                         // page-locked staging for the matrix values of a host-buffer call
    bool connected = false;
    cudaStream_t comm_stream = nullptr;                // pushes of factored panels / block columns to the group
    cudaEvent_t ev_to_comm = nullptr, ev_from_comm = nullptr;
    bool comm_pending = false;
    int push_mode = 0;                                 // 0: copy engines (cudaMemcpy2DAsync to the peer mapping), 1: k_push on the SMs
    unsigned long long peer_timeout_ns = 30ull * 1000000000ull;
    // copy-out of the panels this rank owns: packed (reference layout, owned panels back to back) in d_out, then one
    // device-to-host copy per run of consecutive owned supernodes
    struct OutRun { i64 packed, lsx, count; };
    std::vector<OutRun> out_runs;
    std::vector<i64> packed_off;                       // offset of every owned panel in the packed layout (-1: not owned)
    i64 out_elems = 0;
    int n_pack = 0;
    size_t device_bytes = 0;
    float ms_numeric = 0.f, ms_e2e = 0.f;
    int64_t launches = 0;
    bool factored = false;
};

namespace {

// cudaFuncSetAttribute is per device: every device a plan is created on gets its own opt-in, once
std::mutex g_kernels_mutex;
std::vector<char> g_kernels_ready;

int ensure_kernels(int device) {
    std::lock_guard<std::mutex> lk(g_kernels_mutex);
    if ((int)g_kernels_ready.size() <= device) g_kernels_ready.resize(device + 1, 0);
    if (g_kernels_ready[device]) return SFB200_OK;
    if (kernels_init() != cudaSuccess) { cudaGetLastError(); return SFB200_ERR_CUDA; }
    g_kernels_ready[device] = 1;
    return SFB200_OK;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
            return SFB200_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

template <class T>
int upload(sfb200_plan *p, T *&dptr, const std::vector<T> &h) {
    dptr = nullptr;
    if (h.empty()) return SFB200_OK;
    CK(cudaMalloc((void **)&dptr, h.size() * sizeof(T)));
    p->device_bytes += h.size() * sizeof(T);
    CK(cudaMemcpyAsync(dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, p->stream));
    return SFB200_OK;
}

int ensure_out(sfb200_plan *p) {
    if (p->d_out) return SFB200_OK;
    CK(cudaMalloc((void **)&p->d_out, sizeof(double) * (size_t)std::max<i64>(p->out_elems, 1)));
    p->device_bytes += sizeof(double) * (size_t)p->out_elems;
    return SFB200_OK;
}

int upload_peer_tables(sfb200_plan *p) {
    CK(cudaMemcpyAsync(p->d_peer, p->peer_img.data(), sizeof(double *) * p->peer_img.size(), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->d_peer_flags, p->peer_flags.data(), sizeof(unsigned long long *) * p->peer_flags.size(), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return SFB200_OK;
}

SyncArgs sync_args(const sfb200_plan *p, int chan, size_t first, size_t count) {      // `count` entries of H.sync_points from `first`
    const std::vector<SyncPoint> &list = p->H.sync_points;
    SyncArgs a{};
    a.peer_flags = p->d_peer_flags;
    a.my_flags = p->d_flags;
    a.epoch = p->d_epoch;
    a.me = p->H.rank;
    a.n = (int)count;
    a.chan = chan;
    for (size_t k = 0; k < count; ++k) { a.peer[k] = list[first + k].peer; a.seq[k] = list[first + k].seq; }
    return a;
}

int issue_launch_on(sfb200_plan *p, const Launch &L, cudaStream_t st, int spare_sms) {
    switch (L.kind) {
        case L_SMALL_A: launch_small(p->D, p->d_small + L.first, L.count, L.smem, false, st); break;
        case L_SMALL_B: launch_small(p->D, p->d_small + L.first, L.count, L.smem, true, st); break;
        case L_POTRF:
        case L_GETRF: launch_diag(p->D, p->d_diag + L.first, L.count, st); break;
        case L_TRSM: launch_trsm(p->D, p->d_trsm + L.first, L.count, L.grid, st); break;
        case L_TRSM_MMA: launch_trsm_mma(p->D, p->d_trsm + L.first, L.count, L.grid, st); break;
        case L_GEMM128: launch_gemm(p->D, p->d_gemm + L.first, L.count, L.grid, true, st); break;
        case L_GEMM64: launch_gemm(p->D, p->d_gemm + L.first, L.count, L.grid, false, st); break;
        case L_GEMM128WS: launch_gemm_ws(p->D, p->d_gemm + L.first, L.count, L.grid, p->sms - spare_sms, st); break;
        default: set_error("corrupt plan: unknown launch kind"); return SFB200_ERR_STATE;
    }
    return SFB200_OK;
}

// single-stream issue (profiling replays and the streamed copy-out keep everything on the main stream)
int issue_launch(sfb200_plan *p, const Launch &L) { return issue_launch_on(p, L, p->stream, 0); }

// look-ahead aware issue: LF_SIDE launches go to the side stream behind everything issued so far on the main stream and
// leave `spare` SMs to the main stream's critical path; LF_WAIT_SIDE launches first join the side stream
int issue_launch_la(sfb200_plan *p, const Launch &L) {
    if (p->side_stream && (L.flags & LF_SIDE)) {
        CK(cudaEventRecord(p->ev_to_side, p->stream));
        CK(cudaStreamWaitEvent(p->side_stream, p->ev_to_side, 0));
        const int rc = issue_launch_on(p, L, p->side_stream, p->spare_sms);
        if (rc) return rc;
        CK(cudaEventRecord(p->ev_from_side, p->side_stream));
        p->side_pending = true;
        return SFB200_OK;
    }
    if (p->diag_stream && (L.flags & LF_DIAG_AHEAD)) {
        CK(cudaEventRecord(p->ev_to_diag, p->stream));
        CK(cudaStreamWaitEvent(p->diag_stream, p->ev_to_diag, 0));
        const int rc = issue_launch_on(p, L, p->diag_stream, 0);
        if (rc) return rc;
        CK(cudaEventRecord(p->ev_from_diag, p->diag_stream));
        p->diag_pending = true;
        return SFB200_OK;
    }
    if (p->side_pending && (L.flags & LF_WAIT_SIDE)) {
        CK(cudaStreamWaitEvent(p->stream, p->ev_from_side, 0));
        p->side_pending = false;
    }
    if (p->diag_pending && (L.flags & LF_WAIT_DIAG)) {
        CK(cudaStreamWaitEvent(p->stream, p->ev_from_diag, 0));
        p->diag_pending = false;
    }
    // a persistent GEMM that runs next to a diagonal block leaves it a couple of SMs
    return issue_launch_on(p, L, p->stream, (p->diag_stream && (L.flags & LF_LEAVE_SMS)) ? p->leave_sms : 0);
}

int join_side(sfb200_plan *p) {
    if (p->side_pending) {
        CK(cudaStreamWaitEvent(p->stream, p->ev_from_side, 0));
        p->side_pending = false;
    }
    if (p->diag_pending) {
        CK(cudaStreamWaitEvent(p->stream, p->ev_from_diag, 0));
        p->diag_pending = false;
    }
    return SFB200_OK;
}

// zero the factor, scatter the matrix values of the panels this rank owns
int begin_factor(sfb200_plan *p) {
    const HostPlan &H = p->H;
    cudaStream_t st = p->stream;
    if (H.world > 1) {
        launch_zero(p->d_X, H.xsize_pad, p->d_status, p->sms, st);      // never a memset while ranks may wait for each other
    } else {
        CK(cudaMemsetAsync(p->d_status, 0, sizeof(int), st));
        CK(cudaMemsetAsync(p->d_X, 0, sizeof(double) * (size_t)H.xsize_pad, st));
    }
    launch_assemble(p->d_lx, p->d_asm_l, H.nnz, p->d_X, p->sms, st);
    p->launches = 2;
    if (H.kind == Kind::LU) {
        launch_assemble(H.split ? p->d_ux : p->d_lx, p->d_asm_u, H.unnz, p->d_X, p->sms, st);
        ++p->launches;
    }
    return SFB200_OK;
}

int run_launches(sfb200_plan *p, size_t first, size_t count) {
    const HostPlan &H = p->H;
    for (size_t i = first; i < first + count; ++i) {
        const int rc = issue_launch_la(p, H.launches[i]);
        if (rc) return rc;
        ++p->launches;
    }
    const int rc = join_side(p);          // an etree level ends with the side stream drained
    if (rc) return rc;
    CK(cudaGetLastError());
    return SFB200_OK;
}

// One exchange step of a partitioned plan (plan_types.h: Coll), entirely on the device: no host synchronisation, so the
// whole factorization of a rank -- kernels, pushes, flags -- is one stream-ordered sequence and replays as one CUDA graph.
int issue_coll(sfb200_plan *p, size_t i) {
    const HostPlan &H = p->H;
    const Coll &c = H.colls[i];
    if (H.rank < c.lo || H.rank >= c.hi) return SFB200_OK;
    const CollSync &cs = H.coll_sync[i];
    if (c.type == C_BARRIER) {
        // everything this rank has added into its peers' images so far precedes the signal in stream order (side streams
        // are joined at every segment end)
        launch_signal(sync_args(p, 0, (size_t)cs.sig_first, (size_t)cs.sig_count), p->stream);
        launch_wait(sync_args(p, 0, (size_t)cs.wait_first, (size_t)cs.wait_count), p->d_status, p->peer_timeout_ns, p->stream);
        p->launches += 2;
    } else if (H.rank == c.root) {
        // the factored block leaves on the push stream; this rank's own next launches do not wait for it
        CK(cudaEventRecord(p->ev_to_comm, p->stream));
        CK(cudaStreamWaitEvent(p->comm_stream, p->ev_to_comm, 0));
        const double *src = p->d_X + c.offset;
        for (int k = 0; k < cs.sig_count; ++k) {
            const int q = H.sync_points[cs.sig_first + k].peer;
            double *dst = p->peer_img[q] + H.coll_dst[c.dst_first + (q - c.lo)];
            if (p->push_mode == 0) {
                CK(cudaMemcpy2DAsync(dst, sizeof(double) * c.ld, src, sizeof(double) * c.ld, sizeof(double) * c.rows, (size_t)c.cols,
                                     cudaMemcpyDefault, p->comm_stream));
            } else {
                launch_push(src, dst, c.rows, c.cols, c.ld, 32, p->comm_stream);
                ++p->launches;
            }
            launch_signal(sync_args(p, 1, (size_t)(cs.sig_first + k), 1), p->comm_stream);       // one flag per destination, as soon as its copy is done
            ++p->launches;
        }
        p->comm_pending = true;
    } else {
        launch_wait(sync_args(p, 1, (size_t)cs.wait_first, (size_t)cs.wait_count), p->d_status, p->peer_timeout_ns, p->stream);
        ++p->launches;
    }
    return SFB200_OK;
}

// Replay `issue` (stream-ordered work on p->stream and the plan's helper streams, joined at the end) through a cached
// CUDA graph: captured the first time the key is seen, launched afterwards.
template <class F>
int run_graphed(sfb200_plan *p, int kind, size_t first, size_t count, const void *tag, F issue) {
    if (!p->use_graphs) return issue();
    const auto key = std::make_tuple(kind, first, count, tag);
    auto it = p->graphs.find(key);
    if (it == p->graphs.end()) {
        // the key holds the destination pointer of the streamed copy-out: a caller that hands in a fresh buffer every time
        // must not grow the cache without bound
        // (entries without a tag -- device-resident runs, multi-GPU segments -- are bounded by the schedule itself)
        if (tag) {
            size_t tagged = 0;
            for (auto &kv : p->graphs) tagged += std::get<3>(kv.first) != nullptr;
            if (tagged >= 8)
                for (auto g = p->graphs.begin(); g != p->graphs.end();) {
                    if (std::get<3>(g->first) == nullptr) { ++g; continue; }
                    if (g->second.exec) cudaGraphExecDestroy(g->second.exec);
                    g = p->graphs.erase(g);
                }
        }
        p->graphs.emplace(key, sfb200_plan::CachedGraph{nullptr, 0});
        return issue();
    }
    if (!it->second.exec) {
        const int64_t before = p->launches;
        // the legacy default stream (a caller's sfb200_plan_set_stream(0)) cannot be captured: such plans replay plainly
        if (p->stream == nullptr || p->stream == cudaStreamLegacy || p->stream == cudaStreamPerThread ||
            cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            p->use_graphs = false;
            return issue();
        }
        const int rc = issue();
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(p->stream, &g);
        if (rc || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            if (!rc) set_error(std::string("stream capture failed: ") + cudaGetErrorString(e));
            return rc ? rc : SFB200_ERR_CUDA;
        }
        cudaGraphExec_t ex = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&ex, g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) { set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei)); return SFB200_ERR_CUDA; }
        it->second = sfb200_plan::CachedGraph{ex, p->launches - before};
        p->launches = before;
        if (p->before_first_graph_launch) p->before_first_graph_launch();
    }
    CK(cudaGraphLaunch(it->second.exec, p->stream));
    p->launches += it->second.launches;
    return SFB200_OK;
}

int run_schedule(sfb200_plan *p) {
    p->launches = 0;
    return run_graphed(p, 0, 0, p->H.launches.size(), nullptr, [&]() -> int {
        int rc = begin_factor(p);
        if (rc) return rc;
        for (const Phase &ph : p->H.phases) {        // level by level: each one ends with the side stream joined
            if ((rc = run_launches(p, (size_t)ph.launch_first, (size_t)ph.launch_count))) return rc;
        }
        return SFB200_OK;
    });
}

// One rank's part of a partitioned factorization: epoch bump, zero + assemble, then segment after segment its exchange
// steps and launches.  Every rank of the partition must run it (the barriers wait for all of them).
int run_partitioned(sfb200_plan *p) {
    p->launches = 0;
    return run_graphed(p, 4, 0, p->H.segments.size(), nullptr, [&]() -> int {
        launch_epoch_inc(p->d_epoch, p->stream);
        int rc = begin_factor(p);
        if (rc) return rc;
        ++p->launches;
        for (const Segment &sg : p->H.segments) {
            for (int i = sg.coll_first; i < sg.coll_first + sg.coll_count; ++i)
                if ((rc = issue_coll(p, (size_t)i))) return rc;
            if ((rc = run_launches(p, (size_t)sg.launch_first, (size_t)sg.launch_count))) return rc;
        }
        if (p->comm_pending) {          // the last pushes are on their way before the factorization counts as done
            CK(cudaEventRecord(p->ev_from_comm, p->comm_stream));
            CK(cudaStreamWaitEvent(p->stream, p->ev_from_comm, 0));
            p->comm_pending = false;
        }
        return SFB200_OK;
    });
}

int check_status(sfb200_plan *p) {
    int status = 0;
    CK(cudaMemcpyAsync(&status, p->d_status, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (status & 4) {
        p->factored = false;
        set_error("a rank of the partition did not reach an exchange step in time (peer flag timeout)");
        return SFB200_ERR_STATE;
    }
    if (status & 1) {
        p->factored = false;          // the image holds NaNs / garbage: solve and download refuse it
        set_error("non-positive (Cholesky) or zero (LU) pivot met");
        return SFB200_ERR_NOT_POSDEF;
    }
    return SFB200_OK;
}

}  // namespace

extern "C" {

int sfb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void sfb200_plan_free(sfb200_plan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    void *sptrs[] = {p->d_lsi32, p->d_nz_col, p->d_unz_row, p->d_nz_row, p->d_unz_col, p->d_solve, p->d_rhs, p->d_work, p->d_norms};
    for (void *q : sptrs) if (q) cudaFree(q);
    void *ptrs[] = {p->d_snode, p->d_coloff, p->d_relbase, p->d_asm_l, p->d_asm_u, p->d_segb, p->d_sege, p->d_uoff,
                    p->d_relmap, p->d_small, p->d_diag, p->d_trsm, p->d_gemm, p->d_pack, p->d_X, p->d_out, p->d_lx,
                    p->d_ux, p->d_status};
    for (void *q : ptrs) if (q) cudaFree(q);
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    for (auto &e : p->ev_solve) if (e) cudaEventDestroy(e);
    for (auto &e : p->copy_events) if (e) cudaEventDestroy(e);
    if (p->copy_done) cudaEventDestroy(p->copy_done);
    if (p->d_spack) cudaFree(p->d_spack);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->ev_to_side) cudaEventDestroy(p->ev_to_side);
    if (p->ev_from_side) cudaEventDestroy(p->ev_from_side);
    if (p->side_stream) cudaStreamDestroy(p->side_stream);
    for (void *m : p->ipc_mapped) cudaIpcCloseMemHandle(m);
    if (p->h_stage) cudaFreeHost(p->h_stage);
    if (p->d_peer) cudaFree(p->d_peer);
    if (p->d_peer_flags) cudaFree(p->d_peer_flags);
    if (p->d_flags) cudaFree(p->d_flags);
    if (p->ev_to_comm) cudaEventDestroy(p->ev_to_comm);
    if (p->ev_from_comm) cudaEventDestroy(p->ev_from_comm);
    if (p->comm_stream) cudaStreamDestroy(p->comm_stream);
    for (auto &kv : p->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (p->ev_to_diag) cudaEventDestroy(p->ev_to_diag);
    if (p->ev_from_diag) cudaEventDestroy(p->ev_from_diag);
    if (p->diag_stream) cudaStreamDestroy(p->diag_stream);
    if (p->stream && p->own_stream) cudaStreamDestroy(p->stream);
    delete p;
}

int sfb200_plan_create(const sfb200_symbolic *s, int device, sfb200_plan **out) {
    return sfb200_plan_create_part(s, device, 0, 1, out);
}

int sfb200_plan_create_part(const sfb200_symbolic *s, int device, int rank, int world, sfb200_plan **out) {
    if (!s || !out || world < 1 || rank < 0 || rank >= world) { set_error("sfb200_plan_create: bad argument"); return SFB200_ERR_ARG; }
    if (sfb200_device_count() <= device || device < 0) {
        set_error("sfb200_plan_create: no usable CUDA device (this engine has no CPU fallback)");
        return SFB200_ERR_NO_DEVICE;
    }
    auto *p = new sfb200_plan();
    p->device = device;
    PlanParams prm;
    if (const char *e = getenv("SFB200_NB")) prm.nb = atoi(e);
    if (const char *e = getenv("SFB200_OUTER")) prm.outer = atoi(e);
    if (const char *e = getenv("SFB200_OUTER2")) prm.outer2 = atoi(e);
    if (const char *e = getenv("SFB200_SMALL_A")) prm.small_a_bytes = atoi(e);
    if (const char *e = getenv("SFB200_SMALL_B")) prm.small_b_bytes = atoi(e);
    if (const char *e = getenv("SFB200_GEMM128_MIN")) prm.gemm128_min_m = atoi(e);
    if (const char *e = getenv("SFB200_GEMM128_MIN_K")) prm.gemm128_min_k = atoi(e);
    if (const char *e = getenv("SFB200_DIST_MIN_M")) prm.dist_scatter_min_m = atoi(e);
    if (const char *e = getenv("SFB200_DIST_PANEL_MIN")) prm.dist_panel_min_cols = atoi(e);
    if (const char *e = getenv("SFB200_GEMM_WS")) prm.use_ws = atoi(e);
    if (const char *e = getenv("SFB200_WS_KMAJOR")) prm.ws_kmajor = atoi(e);
    if (const char *e = getenv("SFB200_LU_UT_COPY")) prm.lu_ut_copy = atoi(e);
    if (const char *e = getenv("SFB200_LOOKAHEAD")) prm.lookahead = atoi(e);
    if (const char *e = getenv("SFB200_LEAF_LOOKAHEAD")) prm.leaf_lookahead = atoi(e);
    if (const char *e = getenv("SFB200_LEFT_LOOKING")) prm.left_looking = atoi(e);
    if (const char *e = getenv("SFB200_TRSM_MMA")) prm.trsm_mma = atoi(e);
    if (const char *e = getenv("SFB200_BALANCE_TOL")) prm.balance_tol = atof(e);
    if (const char *e = getenv("SFB200_WEIGHTED_BLOCKS")) prm.weighted_blocks = atoi(e);
    const int rc = build_plan(s->S, prm, p->H, rank, world);
    if (rc != 0) {
        delete p;
        set_error("sfb200_plan_create: structure is not a valid supernodal structure (code " + std::to_string(rc) + ")");
        return SFB200_ERR_ARG;
    }
    int err = SFB200_OK;
    auto fail = [&](int code) { sfb200_plan_free(p); return code; };
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return fail(SFB200_ERR_CUDA); }
    if (ensure_kernels(device) != SFB200_OK) { set_error("kernel attribute setup failed (not an sm_100a device?)"); return fail(SFB200_ERR_CUDA); }
    if (cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || p->sms <= 0) { set_error("cannot query the SM count"); return fail(SFB200_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream"); return fail(SFB200_ERR_CUDA); }
    for (auto &e : p->ev) if (cudaEventCreate(&e) != cudaSuccess) { set_error("event"); return fail(SFB200_ERR_CUDA); }
    const HostPlan &H = p->H;
    if ((err = upload(p, p->d_snode, H.snode))) return fail(err);
    if ((err = upload(p, p->d_coloff, H.upd_coloff))) return fail(err);
    if ((err = upload(p, p->d_relbase, H.upd_relbase))) return fail(err);
    if ((err = upload(p, p->d_segb, H.upd_segb))) return fail(err);
    if ((err = upload(p, p->d_sege, H.upd_sege))) return fail(err);
    if ((err = upload(p, p->d_uoff, H.upd_uoff))) return fail(err);
    if ((err = upload(p, p->d_relmap, H.relmap))) return fail(err);
    if ((err = upload(p, p->d_asm_l, H.asm_dst_l))) return fail(err);
    if ((err = upload(p, p->d_asm_u, H.asm_dst_u))) return fail(err);
    if ((err = upload(p, p->d_small, H.small_tasks))) return fail(err);
    if ((err = upload(p, p->d_diag, H.diag_ops))) return fail(err);
    if ((err = upload(p, p->d_trsm, H.trsm_ops))) return fail(err);
    if ((err = upload(p, p->d_gemm, H.gemm_ops))) return fail(err);
    {   // copy-out of the owned panels: packed back to back in d_out, one host copy per run of consecutive owned supernodes
        const bool lu_ = H.kind == Kind::LU;
        std::vector<PackTask> ptasks;
        i64 packed = 0;
        p->packed_off.assign(H.nsuper, -1);
        for (i64 sidx = 0; sidx < H.nsuper; ++sidx) {
            if (H.owner[sidx] != H.rank) continue;
            p->packed_off[sidx] = packed;
            const SnodeDev &d = H.snode[sidx];
            const int rows = lu_ ? 2 * d.nsrow - d.nscol : d.nsrow;
            const i64 elems = (i64)rows * d.nscol;
            if (!p->out_runs.empty() && p->out_runs.back().lsx + p->out_runs.back().count == H.lsxp[sidx]) p->out_runs.back().count += elems;
            else p->out_runs.push_back({packed, H.lsxp[sidx], elems});
            const int chunk = std::max(1, 65536 / std::max(rows, 1));
            for (int c0 = 0; c0 < d.nscol; c0 += chunk) {
                PackTask t{};
                t.src = d.xoff + (i64)c0 * d.ld;
                t.dst = packed + (i64)c0 * rows;
                t.rows = rows; t.ld = d.ld; t.cols = std::min(chunk, d.nscol - c0);
                ptasks.push_back(t);
            }
            packed += elems;
        }
        p->out_elems = packed;
        p->n_pack = (int)ptasks.size();
        if ((err = upload(p, p->d_pack, ptasks))) return fail(err);
        // A rank of a partition must not allocate (cudaMalloc may wait for the device to go idle, with the context lock held)
        // between the launch of its part and the moment every rank has launched its own: a rank whose flag wait is already
        // spinning on the device would then block the host thread that is about to launch the signal it waits for (seen in
        // round 2 with several ranks in one process).  Its copy-out buffer therefore exists from the start.
        if (H.world > 1 && (err = ensure_out(p))) return fail(err);
        // ... and its matrix values never come from pageable memory: such a copy makes the driver wait inside the call (and,
        // with several ranks in one process, behind the other ranks' work); they are staged through a page-locked buffer
        if (H.world > 1 && cudaMallocHost((void **)&p->h_stage, sizeof(double) * (size_t)std::max<i64>(H.nnz + H.unnz, 1)) != cudaSuccess) {
            set_error("cudaMallocHost of the value staging buffer failed");
            return fail(SFB200_ERR_CUDA);
        }
    }
    if ((err = upload(p, p->d_lsi32, H.lsi32))) return fail(err);
    if ((err = upload(p, p->d_nz_col, H.nz_col))) return fail(err);
    if ((err = upload(p, p->d_nz_row, H.nz_row))) return fail(err);
    if ((err = upload(p, p->d_unz_row, H.unz_row))) return fail(err);
    if ((err = upload(p, p->d_unz_col, H.unz_col))) return fail(err);
    if ((err = upload(p, p->d_solve, H.solve_tasks))) return fail(err);
    auto dmalloc = [&](void **ptr, size_t bytes) {
        if (bytes == 0) bytes = 8;
        if (cudaMalloc(ptr, bytes) != cudaSuccess) { set_error("cudaMalloc of " + std::to_string(bytes) + " bytes failed"); return SFB200_ERR_CUDA; }
        p->device_bytes += bytes;
        return SFB200_OK;
    };
    // factor image, then the U11^T scratch of the large LU fronts (never zeroed again: always written before it is read;
    // zeroed once so that the rows a tile reads beyond an operand's extent are finite), then a tile's worth of slack for
    // those over-reads at the very end
    if ((err = dmalloc((void **)&p->d_X, sizeof(double) * (size_t)(H.xsize_pad + H.wsize + 256)))) return fail(err);
    if (cudaMemsetAsync(p->d_X + H.xsize_pad, 0, sizeof(double) * (size_t)(H.wsize + 256), p->stream) != cudaSuccess) {
        set_error("cudaMemsetAsync of the scratch area failed");
        return fail(SFB200_ERR_CUDA);
    }
    if ((err = dmalloc((void **)&p->d_lx, sizeof(double) * (size_t)H.nnz))) return fail(err);
    if (H.split && (err = dmalloc((void **)&p->d_ux, sizeof(double) * (size_t)H.unnz))) return fail(err);
    if ((err = dmalloc((void **)&p->d_status, sizeof(int)))) return fail(err);
    {   // exchange state: image / flag tables (one entry on a single-GPU plan), flags + epoch counter, push stream
        p->peer_img.assign(H.world, nullptr);
        p->peer_flags.assign(H.world, nullptr);
        if ((err = dmalloc((void **)&p->d_flags, sizeof(unsigned long long) * (kFlagSlots + 1)))) return fail(err);
        if ((err = dmalloc((void **)&p->d_peer, sizeof(double *) * H.world))) return fail(err);
        if ((err = dmalloc((void **)&p->d_peer_flags, sizeof(unsigned long long *) * H.world))) return fail(err);
        if (cudaMemsetAsync(p->d_flags, 0, sizeof(unsigned long long) * (kFlagSlots + 1), p->stream) != cudaSuccess) { set_error("flags"); return fail(SFB200_ERR_CUDA); }
        p->d_epoch = p->d_flags + kFlagSlots;
        p->peer_img[H.rank] = p->d_X;
        p->peer_flags[H.rank] = p->d_flags;
        if ((err = upload_peer_tables(p))) return fail(err);
        p->connected = H.world == 1;
        if (H.world > 1) {
            if (H.world > kMaxWorld) { set_error("at most 16 ranks"); return fail(SFB200_ERR_ARG); }
            if (cudaStreamCreateWithFlags(&p->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_to_comm, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_from_comm, cudaEventDisableTiming) != cudaSuccess) { set_error("push stream"); return fail(SFB200_ERR_CUDA); }
            if (const char *e = getenv("SFB200_PUSH")) p->push_mode = (std::string(e) == "sm") ? 1 : 0;
            if (const char *e = getenv("SFB200_PEER_TIMEOUT_S")) p->peer_timeout_ns = (unsigned long long)(atof(e) * 1e9);
        }
    }
    if (prm.leaf_lookahead) {
        if (const char *e = getenv("SFB200_LEAVE_SMS")) p->leave_sms = atoi(e);
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        if (cudaStreamCreateWithPriority(&p->diag_stream, cudaStreamNonBlocking, hi_prio) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_to_diag, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_from_diag, cudaEventDisableTiming) != cudaSuccess) { set_error("diagonal-block stream"); return fail(SFB200_ERR_CUDA); }
    }
    {   // copy-out work lists of the panels this rank owns: small panels are packed on the device (k_pack, packed layout in
        // d_out) and leave as long contiguous runs; panels of at least 1 MiB go straight from the image through the DMA engine
        // (cudaMemcpy2DAsync), block column by block column for the widest fronts of a single-GPU plan
        const bool lu = H.kind == Kind::LU;
        p->top_width = prm.outer2 > 0 ? prm.outer2 : prm.outer;
        p->big_by_level.assign(H.nlevels, {});
        std::vector<std::vector<PackTask>> small(H.nlevels);
        auto owned = [&](i64 sidx) { return H.owner[sidx] == H.rank; };
        auto is_small = [&](i64 sidx) {
            const SnodeDev &d = H.snode[sidx];
            const int rows = lu ? 2 * d.nsrow - d.nscol : d.nsrow;
            return (i64)rows * d.nscol * 8 < (1 << 20);
        };
        for (i64 sidx = 0; sidx < H.nsuper; ++sidx) {
            if (!owned(sidx)) continue;
            const SnodeDev &d = H.snode[sidx];
            const int rows = lu ? 2 * d.nsrow - d.nscol : d.nsrow;
            if (!is_small(sidx)) { p->big_by_level[d.level].push_back(sidx); continue; }
            const int chunk = std::max(1, 65536 / std::max(rows, 1));
            for (int c0 = 0; c0 < d.nscol; c0 += chunk) {
                PackTask t{};
                t.src = d.xoff + (i64)c0 * d.ld;
                t.dst = p->packed_off[sidx] + (i64)c0 * rows;
                t.rows = rows; t.ld = d.ld; t.cols = std::min(chunk, d.nscol - c0);
                small[d.level].push_back(t);
            }
        }
        std::vector<PackTask> flat;
        for (int l = 0; l < H.nlevels; ++l) {
            p->spack_range.push_back({(int)flat.size(), (int)small[l].size()});
            flat.insert(flat.end(), small[l].begin(), small[l].end());
        }
        if ((err = upload(p, p->d_spack, flat))) return fail(err);
        {
            // Flush point: every level up to `merge_level` leaves in one batch when that level is final.  Each run costs
            // one copy call (a few microseconds of host time during which no kernel is enqueued), and the leaf-side levels
            // interleave in the postorder -- cfg2: 18 246 runs if every level flushes on its own, 467 when levels 0..5 go
            // together -- so the merge level is the lowest one that keeps the whole factorization under 512 runs.
            // A run = consecutive owned small supernodes: contiguous in the packed layout and in the reference's Lsx.
            auto count_runs = [&](int ml) {
                i64 runs = 0;
                int prev = -1;
                for (i64 sidx = 0; sidx < H.nsuper; ++sidx) {
                    if (!owned(sidx) || !is_small(sidx)) { prev = -1; continue; }
                    const int fl = std::max<int>(H.snode[sidx].level, ml);
                    if (fl != prev) ++runs;
                    prev = fl;
                }
                return runs;
            };
            p->merge_level = 0;
            while (p->merge_level + 1 < H.nlevels && count_runs(p->merge_level) > 512) ++p->merge_level;
            p->small_runs.assign(H.nlevels, {});
            auto flush_level = [&](i64 sidx) { return std::max<int>(H.snode[sidx].level, p->merge_level); };
            for (i64 sidx = 0; sidx < H.nsuper;) {
                if (!owned(sidx) || !is_small(sidx)) { ++sidx; continue; }
                const int fl = flush_level(sidx);
                i64 e = sidx + 1, count = 0;
                while (e < H.nsuper && owned(e) && is_small(e) && flush_level(e) == fl) ++e;
                for (i64 q = sidx; q < e; ++q) {
                    const SnodeDev &d = H.snode[q];
                    count += (i64)(lu ? 2 * d.nsrow - d.nscol : d.nsrow) * d.nscol;
                }
                p->small_runs[fl].push_back({sidx, count});       // (first supernode of the run, elements)
                sidx = e;
            }
        }
        if (prm.lookahead) {
            if (const char *e = getenv("SFB200_SPARE_SMS")) p->spare_sms = atoi(e);
            int lo_prio = 0, hi_prio = 0;
            cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
            if (cudaStreamCreateWithPriority(&p->side_stream, cudaStreamNonBlocking, lo_prio) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_to_side, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_from_side, cudaEventDisableTiming) != cudaSuccess) { set_error("side stream"); return fail(SFB200_ERR_CUDA); }
        }
        if (cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->copy_done, cudaEventDisableTiming) != cudaSuccess) { set_error("copy stream"); return fail(SFB200_ERR_CUDA); }
    }
    if (cudaStreamSynchronize(p->stream) != cudaSuccess) { set_error("plan upload failed"); return fail(SFB200_ERR_CUDA); }
    p->D.X = p->d_X;
    p->D.peer = p->d_peer;
    p->D.snode = p->d_snode;
    p->D.upd_coloff = p->d_coloff;
    p->D.upd_relbase = p->d_relbase;
    p->D.upd_segb = p->d_segb;
    p->D.upd_sege = p->d_sege;
    p->D.upd_uoff = p->d_uoff;
    p->D.relmap = p->d_relmap;
    p->D.status = p->d_status;
    p->D.is_lu = H.kind == Kind::LU;
    p->D.multi = H.world > 1;
    p->D.dbg = getenv("SFB200_DBG") ? atoi(getenv("SFB200_DBG")) : 0;
    if (const char *e = getenv("SFB200_GRAPHS")) p->use_graphs = p->graphs_wanted = atoi(e) != 0;
    *out = p;
    return SFB200_OK;
}

int64_t sfb200_plan_int(const sfb200_plan *p, const char *name) {
    if (!p || !name) return -1;
    const std::string k(name);
    const HostPlan &H = p->H;
    if (k == "nlevels") return H.nlevels;
    if (k == "nlaunches") return (int64_t)H.launches.size();
    if (k == "nsmall") return H.nsmall;
    if (k == "nlarge") return H.nlarge;
    if (k == "relmap_entries") return (int64_t)H.relmap.size();
    if (k == "xsize_padded") return H.xsize_pad;
    if (k == "scratch_elems") return H.wsize;
    if (k == "xsize") return H.xsize;
    if (k == "nnz") return H.nnz;
    if (k == "unnz") return H.unnz;
    if (k == "split") return H.split;
    if (k == "device_bytes") return (int64_t)p->device_bytes;
    if (k == "gemm_ops") return (int64_t)H.gemm_ops.size();
    if (k == "trsm_ops") return (int64_t)H.trsm_ops.size();
    if (k == "diag_ops") return (int64_t)H.diag_ops.size();
    if (k == "device") return p->device;
    if (k == "nphases") return (int64_t)H.phases.size();
    if (k == "nsegments") return (int64_t)H.segments.size();
    if (k == "ncolls") return (int64_t)H.colls.size();
    if (k == "rank") return H.rank;
    if (k == "world") return H.world;
    if (k == "connected") return p->connected;
    if (k == "owned_elems") return p->out_elems;
    if (k == "sms") return p->sms;
    if (k == "nsuper") return H.nsuper;
    return -1;
}

double sfb200_plan_double(const sfb200_plan *p, const char *name) {
    if (!p || !name) return -1.0;
    const std::string k(name);
    if (k == "flops") return p->H.flops;
    if (k == "local_flops") return p->H.local_flops;
    if (k == "remote_update_bytes") return 8.0 * p->H.remote_update_entries;      // payload this rank adds into peer images
    if (k == "push_bytes") {            // factored blocks this rank pushes to its groups, per factorization
        double b = 0.0;
        for (const Coll &c : p->H.colls) if (c.type == C_BCAST && c.root == p->H.rank) b += 8.0 * c.rows * (double)c.cols * (c.hi - c.lo - 1);
        return b;
    }
    return -1.0;
}

int sfb200_plan_segment_info(const sfb200_plan *p, int k, int64_t *out) {
    if (!p || !out || k < 0 || k >= (int)p->H.segments.size()) return SFB200_ERR_ARG;
    const Segment &sg = p->H.segments[k];
    out[0] = sg.coll_first; out[1] = sg.coll_count; out[2] = sg.launch_first; out[3] = sg.launch_count; out[4] = sg.level;
    return SFB200_OK;
}

int sfb200_plan_coll_info(const sfb200_plan *p, int i, int64_t *out) {
    if (!p || !out || i < 0 || i >= (int)p->H.colls.size()) return SFB200_ERR_ARG;
    const Coll &c = p->H.colls[i];
    out[0] = c.type; out[1] = c.root; out[2] = c.lo; out[3] = c.hi; out[4] = c.offset; out[5] = c.rows; out[6] = c.cols; out[7] = c.ld;
    return SFB200_OK;
}

int sfb200_plan_owner(const sfb200_plan *p, int32_t *owner) {
    if (!p || !owner) return SFB200_ERR_ARG;
    for (i64 s = 0; s < p->H.nsuper; ++s) owner[s] = p->H.owner[s];
    return SFB200_OK;
}

void *sfb200_plan_factor_ptr(sfb200_plan *p) { return p ? (void *)p->d_X : nullptr; }

int sfb200_plan_panel_offsets(const sfb200_plan *p, int64_t *xoff) {
    if (!p || !xoff) return SFB200_ERR_ARG;
    for (i64 s = 0; s < p->H.nsuper; ++s) xoff[s] = p->H.snode[s].xoff;
    xoff[p->H.nsuper] = p->H.xsize_pad;      /* elements of this rank's image; panels it does not store are -1 */
    return SFB200_OK;
}

int sfb200_factorize_device(sfb200_plan *p, const double *d_Lx, const double *d_Ux) {
    if (!p || !d_Lx) { set_error("sfb200_factorize_device: bad argument"); return SFB200_ERR_ARG; }
    CK(cudaSetDevice(p->device));
    const HostPlan &H = p->H;
    CK(cudaEventRecord(p->ev[0], p->stream));
    if (d_Lx != p->d_lx) CK(cudaMemcpyAsync(p->d_lx, d_Lx, sizeof(double) * (size_t)H.nnz, cudaMemcpyDeviceToDevice, p->stream));
    if (H.split) {
        if (!d_Ux) { set_error("unsymmetric LU needs Ux"); return SFB200_ERR_ARG; }
        if (d_Ux != p->d_ux) CK(cudaMemcpyAsync(p->d_ux, d_Ux, sizeof(double) * (size_t)H.unnz, cudaMemcpyDeviceToDevice, p->stream));
    }
    if (H.world > 1 && !p->connected) {
        set_error("sfb200_factorize_device: this rank of a partitioned plan is not connected to its peers (sfb200_plan_connect / sfb200_group_create)");
        return SFB200_ERR_STATE;
    }
    const int rc = H.world > 1 ? run_partitioned(p) : run_schedule(p);
    if (rc) return rc;
    CK(cudaEventRecord(p->ev[1], p->stream));
    p->factored = true;
    return SFB200_OK;
}

int sfb200_plan_sync(sfb200_plan *p) {
    if (!p) return SFB200_ERR_ARG;
    CK(cudaSetDevice(p->device));
    const int rc = check_status(p);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]) == cudaSuccess) p->ms_numeric = ms;
    else cudaGetLastError();
    return rc;
}

/* The panels this rank owns (all of them on a single-GPU plan) into the reference's layout: Lsx is the WHOLE factor array
 * (xsize doubles); a rank of a partitioned plan writes its own panels and leaves the rest alone. */
static int download(sfb200_plan *p, double *out, bool packed, bool wait = true) {
    if (!p || !out) { set_error("sfb200_download_factor: bad argument"); return SFB200_ERR_ARG; }
    if (!p->factored) { set_error("sfb200_download_factor: nothing factorized yet"); return SFB200_ERR_STATE; }
    CK(cudaSetDevice(p->device));
    int rc = ensure_out(p);
    if (rc) return rc;
    launch_pack(p->d_X, p->d_out, p->d_pack, p->n_pack, p->stream);
    CK(cudaGetLastError());
    if (packed) {
        CK(cudaMemcpyAsync(out, p->d_out, sizeof(double) * (size_t)p->out_elems, cudaMemcpyDeviceToHost, p->stream));
    } else {
        for (const auto &run : p->out_runs)
            CK(cudaMemcpyAsync(out + run.lsx, p->d_out + run.packed, sizeof(double) * (size_t)run.count, cudaMemcpyDeviceToHost, p->stream));
    }
    if (wait) CK(cudaStreamSynchronize(p->stream));
    return SFB200_OK;
}

int sfb200_download_factor(sfb200_plan *p, double *Lsx) { return download(p, Lsx, false); }

/* The panels this rank owns, back to back in supernode order ("owned_elems" doubles): what one process of a multi-process
 * partition brings home.  sfb200_plan_owned_runs lists where the pieces belong in the reference's Lsx. */
int sfb200_download_owned(sfb200_plan *p, double *packed) { return download(p, packed, true); }

/* runs[3 * k .. 3 * k + 2] = {offset in the packed buffer, offset in Lsx, elements}; returns the number of runs (call with
 * runs = NULL to size the array) */
int64_t sfb200_plan_owned_runs(const sfb200_plan *p, int64_t *runs) {
    if (!p) return -1;
    if (runs)
        for (size_t k = 0; k < p->out_runs.size(); ++k) {
            runs[3 * k] = p->out_runs[k].packed; runs[3 * k + 1] = p->out_runs[k].lsx; runs[3 * k + 2] = p->out_runs[k].count;
        }
    return (int64_t)p->out_runs.size();
}

namespace {

cudaEvent_t next_copy_event(sfb200_plan *p) {
    if (p->copy_events_used == p->copy_events.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        p->copy_events.push_back(e);
    }
    return p->copy_events[p->copy_events_used++];
}

// where an owned panel goes in the caller's buffer: the reference's Lsx layout, or the packed layout of the owned panels
inline i64 host_off(const sfb200_plan *p, i64 sidx, bool packed) { return packed ? p->packed_off[sidx] : p->H.lsxp[sidx]; }

// copy columns [c0,c1) of supernode s from this rank's image to the host (DMA engine)
int copy_columns(sfb200_plan *p, i64 sidx, int c0, int c1, double *out, bool packed) {
    const HostPlan &H = p->H;
    const SnodeDev &d = H.snode[sidx];
    const int rows = H.kind == Kind::LU ? 2 * d.nsrow - d.nscol : d.nsrow;
    CK(cudaMemcpy2DAsync(out + host_off(p, sidx, packed) + (i64)c0 * rows, sizeof(double) * rows, p->d_X + d.xoff + (i64)c0 * d.ld,
                         sizeof(double) * d.ld, sizeof(double) * rows, (size_t)(c1 - c0), cudaMemcpyDeviceToHost,
                         p->copy_stream));
    return SFB200_OK;
}

// Factorization with the copy-out streamed behind it: as soon as an etree level (or, for the widest fronts of a single-GPU
// plan, a block column of the level) is final, the panels this rank owns leave for the host on a second stream while the
// next level computes.  A partitioned plan runs its segments -- exchange steps, then launches -- level by level.
int factorize_streamed_issue(sfb200_plan *p, double *out, bool packed, bool trace) {
    const HostPlan &H = p->H;
    if (H.world > 1) launch_epoch_inc(p->d_epoch, p->stream);
    int rc = begin_factor(p);
    if (rc) return rc;
    p->copy_events_used = 0;
    // SFB200_COPY_TRACE=1: per level, when the kernels and when the copies finished (stderr; timing experiments only)
    std::vector<cudaEvent_t> tr_main, tr_copy;
    auto publish = [&]() -> cudaEvent_t {            // everything issued so far on the main stream is visible to copies
        cudaEvent_t e = next_copy_event(p);
        if (!e) return nullptr;
        if (cudaEventRecord(e, p->stream) != cudaSuccess || cudaStreamWaitEvent(p->copy_stream, e, 0) != cudaSuccess) return nullptr;
        return e;
    };
    for (const Phase &ph : H.phases) {
        const int lev = ph.level;
        const std::vector<i64> &big = p->big_by_level[lev];
        std::vector<int> sent(big.size(), 0);            // columns of each big front already on their way
        if (H.world > 1) {
            // a level of a partitioned plan: its segments in order; a panel is final once the level is (pushed block columns
            // of group-factorized fronts have been waited for inside the level)
            for (int k = ph.seg_first; k < ph.seg_first + ph.seg_count; ++k) {
                const Segment &sg = H.segments[k];
                for (int i = sg.coll_first; i < sg.coll_first + sg.coll_count; ++i)
                    if ((rc = issue_coll(p, (size_t)i))) return rc;
                if ((rc = run_launches(p, (size_t)sg.launch_first, (size_t)sg.launch_count))) return rc;
            }
        } else {
            int cur_block = -1;
            for (int li = ph.launch_first; li < ph.launch_first + ph.launch_count; ++li) {
                const Launch &L = H.launches[li];
                if (L.top_block > cur_block && cur_block >= 0) {
                    // the schedule moves on to block column L.top_block: everything left of it is final in every front
                    bool any = false;
                    for (size_t b = 0; b < big.size(); ++b)
                        if (std::min<int>(H.snode[big[b]].nscol, L.top_block * p->top_width) > sent[b]) any = true;
                    if (any) {
                        if (!publish()) { set_error("event"); return SFB200_ERR_CUDA; }
                        for (size_t b = 0; b < big.size(); ++b) {
                            const int c1 = std::min<int>(H.snode[big[b]].nscol, L.top_block * p->top_width);
                            if (c1 > sent[b]) { rc = copy_columns(p, big[b], sent[b], c1, out, packed); if (rc) return rc; sent[b] = c1; }
                        }
                    }
                }
                if (L.top_block >= 0) cur_block = std::max(cur_block, (int)L.top_block);
                rc = issue_launch_la(p, L);
                if (rc) return rc;
                ++p->launches;
            }
            if ((rc = join_side(p))) return rc;
        }
        // the level is final: remaining columns of its big fronts, then all of its small fronts
        const auto range = p->spack_range[lev];
        if (!big.empty() || range.second > 0 || !p->small_runs[lev].empty()) {
            if (!publish()) { set_error("event"); return SFB200_ERR_CUDA; }
            for (size_t b = 0; b < big.size(); ++b) {
                const int n = H.snode[big[b]].nscol;
                if (sent[b] < n) { rc = copy_columns(p, big[b], sent[b], n, out, packed); if (rc) return rc; }
            }
            if (range.second > 0) launch_pack(p->d_X, p->d_out, p->d_spack + range.first, range.second, p->copy_stream);
            for (const auto &run : p->small_runs[lev])
                CK(cudaMemcpyAsync(out + host_off(p, run.first, packed), p->d_out + p->packed_off[run.first],
                                   sizeof(double) * (size_t)run.second, cudaMemcpyDeviceToHost, p->copy_stream));
        }
        if (trace) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
            CK(cudaEventRecord(a, p->stream)); CK(cudaEventRecord(b, p->copy_stream));
            tr_main.push_back(a); tr_copy.push_back(b);
        }
    }
    CK(cudaGetLastError());
    if (p->comm_pending) {          // the last pushes of a partitioned plan are on their way before the call counts as done
        CK(cudaEventRecord(p->ev_from_comm, p->comm_stream));
        CK(cudaStreamWaitEvent(p->stream, p->ev_from_comm, 0));
        p->comm_pending = false;
    }
    if (trace) CK(cudaEventRecord(p->ev[1], p->stream));
    CK(cudaEventRecord(p->copy_done, p->copy_stream));
    CK(cudaStreamWaitEvent(p->stream, p->copy_done, 0));
    if (trace) {
        CK(cudaStreamSynchronize(p->stream));
        for (size_t i = 0; i < tr_main.size(); ++i) {
            float tm = 0, tc = 0;
            cudaEventElapsedTime(&tm, p->ev[0], tr_main[i]);
            cudaEventElapsedTime(&tc, p->ev[0], tr_copy[i]);
            fprintf(stderr, "copy trace: level %2d kernels done %8.2f ms, copies done %8.2f ms\n", H.phases[i].level, tm, tc);
            cudaEventDestroy(tr_main[i]); cudaEventDestroy(tr_copy[i]);
        }
    }
    return SFB200_OK;
}

// Streamed copy-out: kernels, exchange steps, k_pack launches and device-to-host copies of one factorization into `out`,
// replayed as one CUDA graph per destination buffer from the second call on.
int factorize_streamed(sfb200_plan *p, double *out, bool packed) {
    const HostPlan &H = p->H;
    { const int rco = ensure_out(p); if (rco) return rco; }
    const bool trace = getenv("SFB200_COPY_TRACE") != nullptr;
    p->launches = 0;
    int rc;
    if (trace) rc = factorize_streamed_issue(p, out, packed, true);
    else rc = run_graphed(p, packed ? 5 : 1, 0, H.launches.size(), out, [&]() -> int { return factorize_streamed_issue(p, out, packed, false); });
    if (rc) return rc;
    if (!trace) CK(cudaEventRecord(p->ev[1], p->stream));      // includes the copy tail (the pieces are one graph)
    return SFB200_OK;
}

}  // namespace

// The host-buffer call in two halves: `enqueue` only issues work (uploads, this rank's factorization, pack and copy-out) and
// never waits for the device; `finish` waits and reads the status.  A group runs the first half on every rank before any
// rank enters the second (sfb200_group_factorize).
static int factorize_host_enqueue(sfb200_plan *p, const double *Lx, const double *Ux, double *Lsx, bool packed) {
    if (!p || !Lx) { set_error("sfb200_factorize: bad argument"); return SFB200_ERR_ARG; }
    CK(cudaSetDevice(p->device));
    const HostPlan &H = p->H;
    if (H.split && !Ux) { set_error("unsymmetric LU needs Ux"); return SFB200_ERR_ARG; }
    if (H.world > 1 && !p->connected) {
        set_error("sfb200_factorize: this rank of a partitioned plan is not connected to its peers (sfb200_plan_connect / sfb200_group_create)");
        return SFB200_ERR_STATE;
    }
    // page-locked destination -> stream the copy-out behind the factorization; pageable -> pack + one copy at the end
    double *alias = nullptr;
    if (Lsx && !getenv("SFB200_NO_STREAMED_COPY")) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, Lsx) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            alias = (double *)attr.devicePointer;
        else cudaGetLastError();
    }
    CK(cudaEventRecord(p->ev[2], p->stream));
    if (p->h_stage) {
        auto pinned = [](const void *q) {
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
            return a.type == cudaMemoryTypeHost;
        };
        if (!pinned(Lx)) {
            CK(cudaStreamSynchronize(p->stream));          // the previous call's upload has left the staging buffer
            memcpy(p->h_stage, Lx, sizeof(double) * (size_t)H.nnz);
            Lx = p->h_stage;
        }
        if (H.split && !pinned(Ux)) {
            if (Lx != p->h_stage) CK(cudaStreamSynchronize(p->stream));
            memcpy(p->h_stage + H.nnz, Ux, sizeof(double) * (size_t)H.unnz);
            Ux = p->h_stage + H.nnz;
        }
    }
    CK(cudaMemcpyAsync(p->d_lx, Lx, sizeof(double) * (size_t)H.nnz, cudaMemcpyHostToDevice, p->stream));
    if (H.split) CK(cudaMemcpyAsync(p->d_ux, Ux, sizeof(double) * (size_t)H.unnz, cudaMemcpyHostToDevice, p->stream));
    int rc;
    if (alias) {
        CK(cudaEventRecord(p->ev[0], p->stream));
        rc = factorize_streamed(p, Lsx, packed);
        if (rc) return rc;
        p->factored = true;
    } else {
        rc = sfb200_factorize_device(p, p->d_lx, p->d_ux);
        if (rc) return rc;
        // a pageable destination is filled in the second half: a copy into pageable memory blocks inside the call until the
        // stream gets there, i.e. until the other ranks have played their part
        p->pending_out = Lsx;
        p->pending_packed = packed;
    }
    CK(cudaEventRecord(p->ev[3], p->stream));
    return SFB200_OK;
}

static int factorize_host_finish(sfb200_plan *p) {
    CK(cudaSetDevice(p->device));
    if (p->pending_out) {
        double *out = p->pending_out;
        p->pending_out = nullptr;
        const int rcd = download(p, out, p->pending_packed, true);
        if (rcd) return rcd;
        CK(cudaEventRecord(p->ev[3], p->stream));
    }
    const int rc = sfb200_plan_sync(p);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]) == cudaSuccess) p->ms_e2e = ms;
    else cudaGetLastError();
    return rc;
}

static int factorize_host(sfb200_plan *p, const double *Lx, const double *Ux, double *Lsx, bool packed) {
    const int rc = factorize_host_enqueue(p, Lx, Ux, Lsx, packed);
    return rc ? rc : factorize_host_finish(p);
}

int sfb200_factorize(sfb200_plan *p, const double *Lx, const double *Ux, double *Lsx) { return factorize_host(p, Lx, Ux, Lsx, false); }

/* Same call for one process of a multi-process partition: the output is this rank's owned panels, packed
 * (sfb200_download_owned); a full-size Lsx per process would be world times the factor in host memory. */
int sfb200_factorize_owned(sfb200_plan *p, const double *Lx, const double *Ux, double *packed) { return factorize_host(p, Lx, Ux, packed, true); }

int sfb200_plan_set_stream(sfb200_plan *p, void *cuda_stream) {
    if (!p) return SFB200_ERR_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    p->stream = (cudaStream_t)cuda_stream;
    p->own_stream = false;
    // graphs captured on the old stream replay fine on the new one, but a plan that fell back to plain replay because its
    // stream could not be captured gets another chance
    p->use_graphs = p->graphs_wanted;
    return SFB200_OK;
}

/* One instrumented replay of the schedule on the values of the last factorize call: CUDA events around every launch,
 * accumulated per launch kind.  Slower than a plain replay (event records serialise nothing but add ~2 us each). */
int sfb200_plan_profile(sfb200_plan *p) {
    if (!p) return SFB200_ERR_ARG;
    if (p->H.world > 1 && !p->connected) {       // its scatter epilogues add into the peers' images
        set_error("sfb200_plan_profile: this rank of a partitioned plan is not connected to its peers");
        return SFB200_ERR_STATE;
    }
    CK(cudaSetDevice(p->device));
    const HostPlan &H = p->H;
    cudaStream_t st = p->stream;
    const size_t nl = H.launches.size();
    struct Events {          // destroyed on every exit path
        std::vector<cudaEvent_t> v;
        ~Events() { for (auto e : v) if (e) cudaEventDestroy(e); }
    } evs;
    evs.v.assign(nl + 1, nullptr);
    std::vector<cudaEvent_t> &ev = evs.v;
    for (auto &e : ev) CK(cudaEventCreate(&e));
    for (int k = 0; k < L_KINDS; ++k) { p->prof_ms[k] = p->prof_flops[k] = p->prof_bytes[k] = 0; p->prof_launches[k] = 0; }
    p->prof_launch_ms.clear();
    p->prof_launch_flops.clear();
    CK(cudaMemsetAsync(p->d_status, 0, sizeof(int), st));
    CK(cudaMemsetAsync(p->d_X, 0, sizeof(double) * (size_t)H.xsize_pad, st));
    launch_assemble(p->d_lx, p->d_asm_l, H.nnz, p->d_X, p->sms, st);
    if (H.kind == Kind::LU) launch_assemble(H.split ? p->d_ux : p->d_lx, p->d_asm_u, H.unnz, p->d_X, p->sms, st);
    for (size_t i = 0; i < nl; ++i) {
        const Launch &L = H.launches[i];
        CK(cudaEventRecord(ev[i], st));
        if (issue_launch(p, L)) return SFB200_ERR_STATE;
    }
    CK(cudaEventRecord(ev[nl], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const bool lu = H.kind == Kind::LU;
    for (size_t i = 0; i < nl; ++i) {
        const Launch &L = H.launches[i];
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
        const int k = L.kind;
        p->prof_ms[k] += ms;
        p->prof_launches[k] += 1;
        double fl = 0, by = 0;
        for (int q = 0; q < L.count; ++q) {
            if (k == L_GEMM128 || k == L_GEMM64 || k == L_GEMM128WS) {
                const GemmOp &g = H.gemm_ops[L.first + q];
                // LU scatter ops are enumerated as full rectangles but only about half of each is wanted
                const bool half = (g.flags & G_LOWER) != 0;
                const double outputs = half ? (double)g.N * (g.N + 1) / 2 + (double)(g.M - g.N) * g.N
                                            : ((g.flags & (G_LIM_SEGB | G_LIM_SEGE)) ? 0.5 * g.M * (double)g.N : (double)g.M * g.N);
                // multi-GPU: only this rank's share of the op's tiles is computed here
                const double share = g.tile_mod > 1 ? 1.0 / g.tile_mod : 1.0;
                fl += 2.0 * g.K * outputs * share;
                by += (8.0 * ((double)g.M * g.K + (double)g.N * g.K) + 16.0 * outputs) * share;
            } else if (k == L_TRSM || k == L_TRSM_MMA) {
                const TrsmOp &t = H.trsm_ops[L.first + q];
                fl += (double)t.nvec * t.nb * t.nb;
                by += 16.0 * t.nvec * t.nb + 8.0 * t.nb * t.nb;
            } else if (k == L_POTRF || k == L_GETRF) {
                const DiagOp &d = H.diag_ops[L.first + q];
                fl += (lu ? 2.0 : 1.0) * d.nb * (double)d.nb * d.nb / 3.0;
                by += 16.0 * d.nb * d.nb;
            } else {
                const SnodeDev &sn = H.snode[H.small_tasks[L.first + q]];
                const double c = sn.nscol, m = sn.nsrow - sn.nscol;
                fl += lu ? 2.0 * c * c * c / 3.0 + 2.0 * m * c * c + 2.0 * m * m * c : c * c * c / 3.0 + m * c * c + m * (m + 1.0) * c;
                const double pe = lu ? c * (2 * sn.nsrow - c) : c * sn.nsrow;
                by += 16.0 * pe + 16.0 * (lu ? m * m : m * (m + 1) / 2);   // panel read+write, update entries read-modify-written
            }
        }
        p->prof_flops[k] += fl;
        p->prof_bytes[k] += by;
        p->prof_launch_ms.push_back(ms);
        p->prof_launch_flops.push_back(fl);
    }
    // a partitioned plan replayed alone misses the other ranks' contributions: its pivots mean nothing
    if (H.world > 1) { CK(cudaStreamSynchronize(st)); return SFB200_OK; }
    return check_status(p);
}

/* kind: 0/1 small fronts (64/256-thread CTAs), 2 potrf, 3 getrf, 4 trsm, 5 gemm 128x128, 6 gemm 64x64.
 * what: 0 = total ms, 1 = launches, 2 = algorithmic flops, 3 = algorithmic bytes */
double sfb200_plan_profile_get(const sfb200_plan *p, int kind, int what) {
    if (!p || kind < 0 || kind >= L_KINDS) return -1.0;
    switch (what) {
        case 0: return p->prof_ms[kind];
        case 1: return (double)p->prof_launches[kind];
        case 2: return p->prof_flops[kind];
        case 3: return p->prof_bytes[kind];
    }
    return -1.0;
}

/* per-launch record of the last profile: out = {kind, level, grid, count, ms, algorithmic flops, max K of its ops, min K} */
int sfb200_plan_profile_launch(const sfb200_plan *p, int i, double *out) {
    if (!p || !out || i < 0 || i >= (int)p->prof_launch_ms.size()) return SFB200_ERR_ARG;
    const Launch &L = p->H.launches[i];
    out[0] = L.kind; out[1] = L.level; out[2] = L.grid; out[3] = L.count;
    out[4] = p->prof_launch_ms[i]; out[5] = p->prof_launch_flops[i];
    double kmax = 0, kmin = 1e30;
    if (L.kind == L_GEMM128 || L.kind == L_GEMM64 || L.kind == L_GEMM128WS)
        for (int q = 0; q < L.count; ++q) { const double k = p->H.gemm_ops[L.first + q].K; kmax = std::max(kmax, k); kmin = std::min(kmin, k); }
    out[6] = kmax; out[7] = kmin < 1e29 ? kmin : 0;
    return SFB200_OK;
}

/* ---- device-resident triangular solve and residual -------------------------------------------------- */
namespace {
int ensure_rhs(sfb200_plan *p, int nrhs) {
    const size_t need = (size_t)p->H.n * (size_t)nrhs;
    if (need > p->rhs_capacity) {
        if (p->d_rhs) cudaFree(p->d_rhs);
        CK(cudaMalloc((void **)&p->d_rhs, sizeof(double) * std::max<size_t>(need, 1)));
        p->rhs_capacity = need;
    }
    if (!p->d_work) CK(cudaMalloc((void **)&p->d_work, sizeof(double) * 3 * (size_t)std::max<i64>(p->H.n, 1)));
    if (!p->d_norms) CK(cudaMalloc((void **)&p->d_norms, sizeof(unsigned long long) * 4));
    return SFB200_OK;
}
}  // namespace

int sfb200_solve_device(sfb200_plan *p, double *d_B, int nrhs) {
    if (!p || !d_B || nrhs < 1) { set_error("sfb200_solve_device: bad argument"); return SFB200_ERR_ARG; }
    if (!p->factored) { set_error("sfb200_solve: nothing factorized yet"); return SFB200_ERR_STATE; }
    if (p->H.world > 1) { set_error("sfb200_solve: not available on partitioned plans"); return SFB200_ERR_STATE; }
    CK(cudaSetDevice(p->device));
    const HostPlan &H = p->H;
    const bool lu = H.kind == Kind::LU;
    if (!p->ev_solve[0]) { CK(cudaEventCreate(&p->ev_solve[0])); CK(cudaEventCreate(&p->ev_solve[1])); }
    cudaEvent_t a = p->ev_solve[0], b = p->ev_solve[1];
    CK(cudaEventRecord(a, p->stream));
    // the sweeps are a fixed launch sequence too: one graph per (right-hand-side buffer, count)
    const int64_t keep = p->launches;
    const int rcg = run_graphed(p, 3, (size_t)nrhs, 0, d_B, [&]() -> int {
        for (const SolveLaunch &L : H.solve_fwd) launch_solve(p->d_X, p->d_lsi32, d_B, H.n, nrhs, lu, p->d_solve + L.first, L.kind, L.count, p->stream);
        for (const SolveLaunch &L : H.solve_bwd) launch_solve(p->d_X, p->d_lsi32, d_B, H.n, nrhs, lu, p->d_solve + L.first, L.kind, L.count, p->stream);
        return cudaGetLastError() == cudaSuccess ? SFB200_OK : SFB200_ERR_CUDA;
    });
    p->launches = keep;
    if (rcg) return rcg;
    CK(cudaEventRecord(b, p->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(p->stream));
    cudaEventElapsedTime(&p->ms_solve, a, b);
    return SFB200_OK;
}

int sfb200_solve(sfb200_plan *p, double *B, int nrhs) {
    if (!p || !B || nrhs < 1) { set_error("sfb200_solve: bad argument"); return SFB200_ERR_ARG; }
    CK(cudaSetDevice(p->device));
    int rc = ensure_rhs(p, nrhs);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * (size_t)p->H.n * (size_t)nrhs;
    CK(cudaMemcpyAsync(p->d_rhs, B, bytes, cudaMemcpyHostToDevice, p->stream));
    rc = sfb200_solve_device(p, p->d_rhs, nrhs);
    if (rc) return rc;
    CK(cudaMemcpyAsync(B, p->d_rhs, bytes, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return SFB200_OK;
}

/* out = {||A||_1, ||b||_inf, ||x||_inf, ||Ax-b||_inf, rnorm/(anorm*xnorm+bnorm)} with the matrix values of the last factorize */
int sfb200_residual(sfb200_plan *p, const double *x, const double *b, double *out) {
    if (!p || !x || !b || !out) { set_error("sfb200_residual: bad argument"); return SFB200_ERR_ARG; }
    if (p->H.world > 1) { set_error("sfb200_residual: not available on partitioned plans"); return SFB200_ERR_STATE; }
    CK(cudaSetDevice(p->device));
    int rc = ensure_rhs(p, 1);
    if (rc) return rc;
    const HostPlan &H = p->H;
    const size_t nb = sizeof(double) * (size_t)H.n;
    double *d_b = p->d_work, *d_r = p->d_work + H.n, *d_w = p->d_work + 2 * H.n;
    CK(cudaMemcpyAsync(p->d_rhs, x, nb, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(d_b, b, nb, cudaMemcpyHostToDevice, p->stream));
    launch_residual(p->d_lx, p->d_nz_row, p->d_nz_col, H.nnz, p->d_ux, p->d_unz_row, p->d_unz_col, H.unnz, H.split, p->d_rhs, d_b,
                    d_r, d_w, H.n, p->d_norms, p->sms, p->stream);
    CK(cudaGetLastError());
    unsigned long long bits[4];
    CK(cudaMemcpyAsync(bits, p->d_norms, sizeof(bits), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int k = 0; k < 4; ++k) memcpy(out + k, bits + k, sizeof(double));
    out[4] = out[3] / (out[0] * out[2] + out[1]);
    return SFB200_OK;
}

double sfb200_last_ms(const sfb200_plan *p, int which) {
    if (!p) return -1.0;
    return which == 0 ? p->ms_numeric : which == 1 ? p->ms_e2e : p->ms_solve;
}

int64_t sfb200_last_launches(const sfb200_plan *p) { return p ? p->launches : -1; }

void *sfb200_host_alloc(size_t bytes) {
    void *ptr = nullptr;
    if (cudaMallocHost(&ptr, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return ptr;
}
void sfb200_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

/* ---- partitioned plans: connecting the ranks (peer memory) ------------------------------------------------------------- */
namespace {
int set_peers(sfb200_plan *p, const std::vector<double *> &images, const std::vector<unsigned long long *> &flags) {
    const HostPlan &H = p->H;
    for (int q = 0; q < H.world; ++q) {
        if (q == H.rank) continue;
        if (!images[q] || !flags[q]) { set_error("sfb200_plan_connect: missing peer mapping"); return SFB200_ERR_ARG; }
        p->peer_img[q] = images[q];
        p->peer_flags[q] = flags[q];
    }
    const int rc = upload_peer_tables(p);
    if (rc) return rc;
    p->connected = true;
    return SFB200_OK;
}
}  // namespace

int sfb200_plan_exchange_handle(sfb200_plan *p, void *out) {
    if (!p || !out) { set_error("sfb200_plan_exchange_handle: bad argument"); return SFB200_ERR_ARG; }
    CK(cudaSetDevice(p->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle layout");
    cudaIpcMemHandle_t h[2];
    CK(cudaIpcGetMemHandle(&h[0], p->d_X));
    CK(cudaIpcGetMemHandle(&h[1], p->d_flags));
    memcpy(out, h, sizeof(h));
    return SFB200_OK;
}

int sfb200_plan_connect(sfb200_plan *p, const void *handles) {
    if (!p || !handles) { set_error("sfb200_plan_connect: bad argument"); return SFB200_ERR_ARG; }
    CK(cudaSetDevice(p->device));
    const HostPlan &H = p->H;
    std::vector<double *> images(H.world, nullptr);
    std::vector<unsigned long long *> flags(H.world, nullptr);
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)handles;
    for (int q = 0; q < H.world; ++q) {
        if (q == H.rank) continue;
        void *a = nullptr, *b = nullptr;
        CK(cudaIpcOpenMemHandle(&a, h[2 * q], cudaIpcMemLazyEnablePeerAccess));
        p->ipc_mapped.push_back(a);
        CK(cudaIpcOpenMemHandle(&b, h[2 * q + 1], cudaIpcMemLazyEnablePeerAccess));
        p->ipc_mapped.push_back(b);
        images[q] = (double *)a;
        flags[q] = (unsigned long long *)b;
    }
    return set_peers(p, images, flags);
}

struct sfb200_group {
    std::vector<sfb200_plan *> plans;
    long calls = 0;                      // sfb200_group_factorize calls so far (the second one captures the CUDA graphs)
};

/* All ranks of a partition inside one process, one device per rank (the reference's arrangement: one host thread per GPU
 * inside SparseFrame_factorize_supernodal, CH.c:2267-2290).  devices[r] may repeat: several ranks on one GPU exchange
 * through the same mechanism (used by the one-GPU tests). */
int sfb200_group_create(const sfb200_symbolic *s, int world, const int *devices, sfb200_group **out) {
    if (!s || !out || world < 1 || world > kMaxWorld || !devices) { set_error("sfb200_group_create: bad argument"); return SFB200_ERR_ARG; }
    auto *g = new sfb200_group();
    g->plans.assign(world, nullptr);
    std::vector<int> rcs(world, SFB200_OK);
    std::vector<std::string> errs(world);
    {
        std::vector<std::thread> th;
        for (int r = 0; r < world; ++r)
            th.emplace_back([&, r]() {
                rcs[r] = sfb200_plan_create_part(s, devices[r], r, world, &g->plans[r]);
                if (rcs[r]) errs[r] = sfb200_last_error();
            });
        for (auto &t : th) t.join();
    }
    for (int r = 0; r < world; ++r)
        if (rcs[r]) { set_error(errs[r]); const int rc = rcs[r]; sfb200_group_free(g); return rc; }
    auto fail = [&](int rc) { sfb200_group_free(g); return rc; };
    for (int a = 0; a < world; ++a)
        for (int b = 0; b < world; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) {
                set_error("sfb200_group_create: no peer access between devices " + std::to_string(devices[a]) + " and " + std::to_string(devices[b]));
                return fail(SFB200_ERR_CUDA);
            }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return fail(SFB200_ERR_CUDA); }
            cudaGetLastError();
        }
    std::vector<double *> images(world);
    std::vector<unsigned long long *> flags(world);
    for (int r = 0; r < world; ++r) { images[r] = g->plans[r]->d_X; flags[r] = g->plans[r]->d_flags; }
    for (int r = 0; r < world; ++r) {
        cudaSetDevice(devices[r]);
        const int rc = set_peers(g->plans[r], images, flags);
        if (rc) return fail(rc);
    }
    *out = g;
    return SFB200_OK;
}

int sfb200_group_size(const sfb200_group *g) { return g ? (int)g->plans.size() : 0; }
sfb200_plan *sfb200_group_plan(sfb200_group *g, int rank) { return (g && rank >= 0 && rank < (int)g->plans.size()) ? g->plans[rank] : nullptr; }

/* Host-buffer factorization by the whole group: one host thread per rank uploads the values, runs its part and writes the
 * panels it owns into Lsx (may be NULL: factor stays on the devices).  Returns the first non-zero status. */
int sfb200_group_factorize(sfb200_group *g, const double *Lx, const double *Ux, double *Lsx) {
    if (!g || !Lx) { set_error("sfb200_group_factorize: bad argument"); return SFB200_ERR_ARG; }
    const int world = (int)g->plans.size();
    std::vector<int> rcs(world, SFB200_OK);
    std::vector<std::string> errs(world);
    std::vector<std::thread> th;
    // Two halves with a host-side meeting point in between: no thread may enter a call that waits for the device (or takes a
    // driver lock while doing so) before EVERY rank has launched its part -- the ranks wait for each other on the device.
    std::mutex mtx;
    std::condition_variable cv;
    int arrived = 0;
    long generation = 0;
    auto meet = [&]() {                    // reusable barrier of the `world` host threads
        std::unique_lock<std::mutex> lk(mtx);
        const long gen = generation;
        if (++arrived == world) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    };
    // the call in which the ranks capture + instantiate their graphs is the same on every rank (the second one): they meet
    // once more between instantiation and launch
    const bool will_capture = g->calls == 1 && g->plans[0]->use_graphs;
    ++g->calls;
    for (int r = 0; r < world; ++r)
        th.emplace_back([&, r]() {
            bool met = false;
            g->plans[r]->before_first_graph_launch = [&]() { if (will_capture) { met = true; meet(); } };
            rcs[r] = factorize_host_enqueue(g->plans[r], Lx, Ux, Lsx, false);
            g->plans[r]->before_first_graph_launch = nullptr;
            if (rcs[r]) errs[r] = sfb200_last_error();
            if (will_capture && !met) meet();          // a rank that failed before its launch must not leave the others waiting
            meet();
            if (rcs[r] == SFB200_OK) {
                rcs[r] = factorize_host_finish(g->plans[r]);
                if (rcs[r]) errs[r] = sfb200_last_error();
            }
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < world; ++r)
        if (rcs[r]) { set_error("rank " + std::to_string(r) + ": " + errs[r]); return rcs[r]; }
    return SFB200_OK;
}

/* numeric time of the last factorization: the slowest rank (which = 0) / including the copies (which = 1) */
double sfb200_group_last_ms(const sfb200_group *g, int which) {
    double ms = -1.0;
    if (g) for (const sfb200_plan *p : g->plans) ms = std::max(ms, sfb200_last_ms(p, which));
    return ms;
}

void sfb200_group_free(sfb200_group *g) {
    if (!g) return;
    for (sfb200_plan *p : g->plans) sfb200_plan_free(p);
    delete g;
}

/* The caller has filled this (single-GPU) plan's image -- sfb200_plan_factor_ptr, panel offsets from
 * sfb200_plan_panel_offsets -- with a complete factor of the matrix whose values are d_Lx / d_Ux, e.g. gathered from the
 * ranks of a partitioned factorization: from here on sfb200_solve / sfb200_residual / sfb200_download_factor use it. */
int sfb200_plan_adopt_factor(sfb200_plan *p, const double *d_Lx, const double *d_Ux) {
    if (!p || !d_Lx) { set_error("sfb200_plan_adopt_factor: bad argument"); return SFB200_ERR_ARG; }
    if (p->H.world != 1) { set_error("sfb200_plan_adopt_factor: single-GPU plans only"); return SFB200_ERR_STATE; }
    CK(cudaSetDevice(p->device));
    const HostPlan &H = p->H;
    if (d_Lx != p->d_lx) CK(cudaMemcpyAsync(p->d_lx, d_Lx, sizeof(double) * (size_t)H.nnz, cudaMemcpyDeviceToDevice, p->stream));
    if (H.split) {
        if (!d_Ux) { set_error("unsymmetric LU needs Ux"); return SFB200_ERR_ARG; }
        if (d_Ux != p->d_ux) CK(cudaMemcpyAsync(p->d_ux, d_Ux, sizeof(double) * (size_t)H.unnz, cudaMemcpyDeviceToDevice, p->stream));
    }
    CK(cudaMemsetAsync(p->d_status, 0, sizeof(int), p->stream));
    CK(cudaEventRecord(p->ev[0], p->stream));
    CK(cudaEventRecord(p->ev[1], p->stream));
    p->factored = true;
    return SFB200_OK;
}

/* FP64 pipe probes for the roofline denominators: returns TFLOP/s.  which = 0 DFMA (vector pipe), 1 DMMA (tensor pipe) */
double sfb200_probe_fp64_peak(int device, int which, int iters) {
    if (sfb200_device_count() <= device) return -1.0;
    cudaSetDevice(device);
    // which: 0 DFMA, 1 DMMA (register operands, 8 accumulators/warp), 2 DMMA fed from shared memory like the GEMM inner
    // loop (1 CTA of 8 warps per SM), 3 same tile shape with register-resident fragments
    const bool tile = which >= 2;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) return -1.0;
    const int grid = tile ? sms : sms * 8, block = 256;
    double *d = nullptr;
    if (cudaMalloc((void **)&d, sizeof(double) * grid * block) != cudaSuccess) return -1.0;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a, 0);
        if (which == 0) launch_peak_dfma(d, iters, grid, block, 0);
        else if (which == 1) launch_peak_dmma(d, iters, grid, block, 0);
        else launch_peak_dmma_tile(d, iters, grid, which == 2 ? 0 : 1, 0);
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || ms <= 0.f) return -1.0;
    const double threads = (double)grid * block;
    double flops;
    if (which == 0) flops = threads * iters * 64.0 * 2.0;                        /* 64 FMA per thread-iteration */
    else if (which == 1) flops = (threads / 32.0) * iters * 32.0 * 512.0;        /* 32 DMMA x 256 FMA per warp-iteration */
    else flops = (threads / 32.0) * iters * 128.0 * 512.0;                       /* 4 k-steps x 32 DMMA per warp-iteration */
    return flops / (ms * 1e-3) / 1e12;
}

}  // extern "C"

namespace sfb200 {
PlanInternals plan_internals(sfb200_plan *p) { return PlanInternals{&p->H, p->d_X, p->d_lx, p->d_ux, p->stream, p->device, p->sms}; }
}  // namespace sfb200
