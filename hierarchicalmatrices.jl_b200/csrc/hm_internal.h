// hm_internal.h -- private to the library: error plumbing of the C ABI and the plan object,
// shared by hm_api.cu (builder, plan, mul!) and hm_dist.cu (multi-GPU exchange).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <new>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "../../include/hmb200.h"
#include "hm_kernels.cuh"
#include "hm_layout.h"
#include "hm_nest.h"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
// message of the last failure on the calling thread (fixed storage: reporting a failure must not
// allocate); defined in hm_api.cu
extern thread_local char hm_err_msg[512];
int32_t fail(hm_status st, const char *fmt, ...) noexcept;

// No C++ exception may cross the extern "C" boundary (behind ccall / ctypes it would reach
// std::terminate and kill the host process): every entry point runs its body through this.
template <class Fn> int32_t guarded(Fn &&fn) noexcept
{
    try {
        return fn();
    } catch (const std::bad_alloc &) {
        return fail(HM_ERR_NOMEM, "out of host memory");
    } catch (const std::length_error &e) {
        return fail(HM_ERR_NOMEM, "container size limit exceeded: %s", e.what());
    } catch (const std::exception &e) {
        return fail(HM_ERR_INVALID, "internal error: %s", e.what());
    } catch (...) {
        return fail(HM_ERR_INVALID, "internal error (unknown exception)");
    }
}

#define HM_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            cudaGetLastError();                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? HM_ERR_NOMEM : HM_ERR_CUDA,              \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)

// restore the caller's current device on scope exit (the host process may be
// driving other devices through its own runtime, e.g. torch)
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
        ok = err == cudaSuccess;
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (ok && prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

#define HM_DEVICE(dev)                                                                             \
    DeviceGuard guard_(dev);                                                                       \
    if (!guard_.ok) {                                                                              \
        cudaGetLastError();                                                                        \
        return fail(HM_ERR_CUDA, "cannot select CUDA device %d: %s", (int)(dev),                   \
                    cudaGetErrorString(guard_.err));                                               \
    }

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T> &v, cudaStream_t st)
    {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// multi-GPU exchange state of a plan (hm_dist.cu): NCCL communicator, peer-mapped buffers
struct HmDist;
void hm_dist_release(HmDist *d) noexcept;

struct hm_plan {
    int device = 0;
    HmLayout L; // host metadata (tables are kept for the test hooks)
    int kernel_id = 0;
    HmCheb cheb{};
    // matrix-free plans (hm_assemble_kernel_free): no streams; the tables the fill kernels use stay
    // on the device together with the point sets, and the apply evaluates the entries itself
    bool matrix_free = false;
    DevBuf<HmFreeEnt> f_ent1;
    DevBuf<HmFreeRun> f_run3;
    DevBuf<double> f_px, f_py;
    int free1_units = 1, free3_zcap = HM_SMAX;
    bool free_cheb = false; // cores hold C F C' and the apply runs the Chebyshev-series kernels
    // nested-basis form of a matrix-free plan (hm_nest.h): box trees, transfer maps, shared cores
    bool nested = false;
    DevBuf<HmNestNode> nr_nodes, nc_nodes;
    DevBuf<int32_t> nr_order, nr_grp, nr_sub, nc_order, nc_grp, nc_sub, n_rleaf_begin, nr_base, nc_base, n_item_box;
    bool n_fused_eval = false;
    bool n_overlap = true; // dense leaves on side_stream beside the tree passes (HMB200_NEST_OVERLAP=0: off)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_dense = nullptr;
    // many right-hand sides in nested form: items with the coefficient run, per-box panels
    DevBuf<int32_t> n_fin;
    DevBuf<HmItem> n_items3p;
    DevBuf<HmRun> n_runsp;
    DevBuf<HmFreeRun> n_frunp;
    DevBuf<double> n_MUp, n_LAMp, n_Sp;
    int n_ws_cs = 0;
    int64_t n_fin_rows = 0;
    DevBuf<HmNestLeaf> n_rleaf;
    DevBuf<double> n_cores, n_M, n_MU, n_LAM;
    DevBuf<HmItem> n_items3; // the stage-3 items with their dense runs only
    DevBuf<HmRun> n_runs;
    DevBuf<HmFreeRun> n_frun;
    std::vector<int64_t> n_round_begin;
    int n_zcap = 2;
    HmNestDev n_rows, n_cols;
    int64_t n_distinct_cores = 0;
    // device arrays
    DevBuf<double> vstream, ustream, core, svec, partial;
    DevBuf<HmItem> items1, items3;
    DevBuf<HmRun> runs;
    DevBuf<HmCoreBlock> cores;
    DevBuf<int32_t> plist, bigcores; // bigcores: leaves with more than HM_CORE_BIG partial sums
    int64_t nbig = 0;
    // host-pointer path: chunked item orders, copy stream and events (allocated on first use)
    DevBuf<HmItem> items1c, items3c;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_x[HM_NCHUNK] = {}, ev_y[HM_NCHUNK] = {}, ev_y0 = nullptr;
    bool chunk_ready = false;
    // adjoint apply (allocated on first use)
    DevBuf<double> pq;
    DevBuf<int32_t> qlist, core_q0, core_qn, adjbig;
    int nadjbig = 0;
    DevBuf<HmColSeg> colsegs;
    DevBuf<int64_t> colbases;
    bool adj_ready = false;
    // Optional: stage 2 fused into the tail of stage 1 (HMB200_FUSE_STAGE2=1).  Measured slower
    // on one B200 at N = 2^20 (2.233 vs 2.121 ms per matvec: the fence + arrival atomics at the
    // end of every stage-1 CTA cost more than the 0.1 ms stand-alone kernel), so it is off by default.
    DevBuf<int32_t> s1ent;
    DevBuf<int> counters;
    bool fuse = false;
    // host-pointer path
    cudaStream_t stream = nullptr;
    DevBuf<double> dx, dy;
    double *hx = nullptr, *hy = nullptr; // pinned, for strided arguments
    int64_t nrhs_cap = 0;
    DevBuf<double> dX, dY;
    // panel workspace of the multi-RHS path (row pitch ws_cs)
    int ws_cs = 0;
    int panel_zcap = 0; // largest z length of a stage-3 item
    DevBuf<double> wXt, wPp, wYt;
    double *wSp = nullptr; // inside wXt's allocation, behind the x panel
    std::mutex mu;
    // per-stage timing (bench bookkeeping)
    std::vector<cudaEvent_t> tev;
    int tcap = 0, tcount = 0;
    // test hooks
    std::unordered_multimap<int32_t, size_t> idx1, idx3;
    bool indexed = false;
    // multi-GPU exchange (hm_dist_init)
    HmDist *dist = nullptr;
    // adjoint of a matrix-free plan: a second matrix-free plan over the transposed leaves with the two
    // point sets exchanged (built on the first hm_matvec_adjoint*), and the buffer for -x of the odd kernels
    double box[4] = {0, 0, 0, 0}; // a, b, c, d of hm_assemble_kernel*
    hm_plan *adj = nullptr;
    DevBuf<double> adj_x;
    ~hm_plan()
    {
        delete adj;
        hm_dist_release(dist);
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
        for (int k = 0; k < HM_NCHUNK; k++) {
            if (ev_x[k]) cudaEventDestroy(ev_x[k]);
            if (ev_y[k]) cudaEventDestroy(ev_y[k]);
        }
        if (ev_y0) cudaEventDestroy(ev_y0);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_dense) cudaEventDestroy(ev_dense);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (hx) cudaFreeHost(hx);
        if (hy) cudaFreeHost(hy);
        if (stream) cudaStreamDestroy(stream);
    }
};

// the three stages of one matvec on `stream`; peers != nullptr: stage 3 stores every owned row into
// all ranks' y buffers (hm_api.cu)
extern "C" int32_t hm_matvec_device_peers(hm_plan *p, const double *dx, double *dy, int32_t accumulate, void *stream,
                               const HmPeers *peers);
