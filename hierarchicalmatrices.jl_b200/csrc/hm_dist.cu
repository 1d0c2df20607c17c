// hm_dist.cu -- multi-GPU mul!: one process per GPU, one block-row part of the operator per plan
// (SURVEY 8e).  The plan owns the exchange:
//
//   x   replicated with an NCCL broadcast over NVLink into a plan-owned buffer (hm_dist_bcast_x);
//   y   every rank's stage 3 stores the rows it owns straight into the y buffer of *every* rank
//       (cudaIpc peer mappings of one exchange region per rank), so the all-gather of y is part of
//       the kernel and no collective moves y;
//   barrier   a one-CTA kernel: rank r writes its epoch into slot r of every peer's flag row
//       (st.release.sys over NVLink) and spins (ld.acquire.sys) until its own row shows the epoch
//       from everybody -- afterwards each rank holds the complete y.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on
// it, and in a process that already carries an NCCL (PyTorch) the same instance is used.  Only the
// broadcast of x and the bootstrap (exchange of the 64-byte IPC handles) go through NCCL.
//
// Buffers are double-buffered ("slots"): a call that writes y slot s may run while a peer still
// reads slot 1-s; the barrier at the end of every matvec makes two slots sufficient (a rank can be
// at most one matvec ahead of any other).
#include <dlfcn.h>
#include <nccl.h> // types and prototypes only; the entry points are resolved with dlsym

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "hm_internal.h"

namespace {

struct NcclApi {
    void *h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr; // optional (NCCL >= 2.17)
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    const char *why = "";

    bool load()
    {
        if (h) return true;
        const char *names[] = {getenv("HMB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            why = "libnccl.so.2 not found (set HMB200_NCCL_LIB)";
            return false;
        }
#define HM_SYM(field, name)                                                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(h, name));                                     \
    if (!field) {                                                                                  \
        why = "NCCL symbol " name " missing";                                                      \
        h = nullptr;                                                                               \
        return false;                                                                              \
    }
        HM_SYM(GetUniqueId, "ncclGetUniqueId")
        HM_SYM(CommInitRank, "ncclCommInitRank")
        HM_SYM(CommDestroy, "ncclCommDestroy")
        HM_SYM(Broadcast, "ncclBroadcast")
        HM_SYM(AllGather, "ncclAllGather")
        HM_SYM(GetErrorString, "ncclGetErrorString")
        HM_SYM(GetVersion, "ncclGetVersion")
#undef HM_SYM
        CommInitRankConfig = reinterpret_cast<decltype(CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
        return true;
    }
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

#define HM_NCCL(call)                                                                              \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail(HM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

constexpr int FLAG_STRIDE = 32; // one 128-byte line per source rank

struct BarrierArgs {
    unsigned *flags[HM_MAX_PEERS]; // flag rows of all ranks (peer-mapped); row = HM_MAX_PEERS lines
    unsigned *epoch;               // this rank's barrier counter (device memory: CUDA-graph replays advance it)
    int *err;                      // set when a peer did not arrive in time
    int n, self;
    long long timeout_cycles;
};

// Every rank has passed this point and all its earlier stores (stage 3's rows in the peers' y
// buffers) are visible, when the kernel ends.
__global__ void __launch_bounds__(32) hm_barrier_kernel(BarrierArgs b)
{
    __shared__ unsigned es;
    // programmatic dependent launch: the next kernel (stage 1 of the following matvec) may begin its
    // prologue while this one spins; this kernel itself must not signal before stage 3 has finished
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) es = ++(*b.epoch);
    __syncwarp();
    const unsigned e = es;
    const int q = threadIdx.x;
    if (q >= b.n) return;
    __threadfence_system();
    unsigned *dst = b.flags[q] + b.self * FLAG_STRIDE;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(e) : "memory");
    const unsigned *src = b.flags[b.self] + q * FLAG_STRIDE;
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        if ((int)(v - e) >= 0) break; // the peer may already be one barrier ahead
        if (clock64() - t0 > b.timeout_cycles) {
            *b.err = 1;
            break;
        }
    }
}

} // namespace

struct HmDist {
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    char *base = nullptr; // this rank's exchange region
    char *peer[HM_MAX_PEERS] = {};
    size_t bytes = 0;
    size_t off_flags = 0, off_epoch = 0, off_err = 0, off_x[2] = {}, off_y[2] = {};
    int64_t nrows = 0, ncols = 0;
    double *hx = nullptr, *hy = nullptr; // pinned staging of hm_dist_matvec
    uint64_t calls = 0;                  // hm_dist_matvec calls so far
    // hm_dist_push_x: copy-engine replication of x (root side)
    cudaStream_t push_stream[HM_MAX_PEERS] = {};
    cudaEvent_t push_start = nullptr, push_done[HM_MAX_PEERS] = {};
    bool push_pending = false;
    double *x(int s) const { return reinterpret_cast<double *>(base + off_x[s]); }
    double *y(int q, int s) const { return reinterpret_cast<double *>(peer[q] + off_y[s]); }
};

void hm_dist_release(HmDist *d) noexcept
{
    if (!d) return;
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    for (int q = 0; q < d->nranks; q++)
        if (q != d->rank && d->peer[q]) cudaIpcCloseMemHandle(d->peer[q]);
    if (d->base) cudaFree(d->base);
    if (d->hx) cudaFreeHost(d->hx);
    if (d->hy) cudaFreeHost(d->hy);
    for (int q = 0; q < HM_MAX_PEERS; q++) {
        if (d->push_stream[q]) cudaStreamDestroy(d->push_stream[q]);
        if (d->push_done[q]) cudaEventDestroy(d->push_done[q]);
    }
    if (d->push_start) cudaEventDestroy(d->push_start);
    delete d;
}

namespace {

int32_t launch_barrier(hm_plan *p, cudaStream_t st)
{
    HmDist *d = p->dist;
    if (d->push_pending) {
        // x pushed by hm_dist_push_x must have landed on every rank before this rank signals
        for (int q = 0; q < d->nranks; q++)
            if (d->push_done[q]) HM_CUDA(cudaStreamWaitEvent(st, d->push_done[q], 0));
        d->push_pending = false;
    }
    BarrierArgs b{};
    for (int q = 0; q < d->nranks; q++) b.flags[q] = reinterpret_cast<unsigned *>(d->peer[q] + d->off_flags);
    b.epoch = reinterpret_cast<unsigned *>(d->base + d->off_epoch);
    b.err = reinterpret_cast<int *>(d->base + d->off_err);
    b.n = d->nranks;
    b.self = d->rank;
    b.timeout_cycles = 20LL * 1000 * 1000 * 1000; // ~10 s at 2 GHz: a dead peer must not hang the GPU
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const char *e = getenv("HMB200_PDL");
    cfg.numAttrs = (e && e[0] == '0') ? 0 : 1;
    HM_CUDA(cudaLaunchKernelEx(&cfg, hm_barrier_kernel, b));
    return HM_OK;
}

int32_t need_dist(hm_plan *p)
{
    if (!p) return fail(HM_ERR_NULL, "plan is NULL");
    if (!p->dist) return fail(HM_ERR_STATE, "hm_dist_init has not been called on this plan");
    return HM_OK;
}

} // namespace

extern "C" {

int32_t hm_dist_get_id(void *id_out)
{
    return guarded([&]() -> int32_t {
        if (!id_out) return fail(HM_ERR_NULL, "id_out is NULL");
        std::lock_guard<std::mutex> lock(g_nccl_mu);
        if (!g_nccl.load()) return fail(HM_ERR_UNSUPPORTED, "NCCL unavailable: %s", g_nccl.why);
        static_assert(HM_DIST_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
        ncclUniqueId id;
        HM_NCCL(g_nccl.GetUniqueId(&id));
        memcpy(id_out, &id, sizeof id);
        return HM_OK;
    });
}

int32_t hm_dist_init(hm_plan *p, const void *id, int32_t nranks, int32_t rank)
{
    return guarded([&]() -> int32_t {
        if (!p || !id) return fail(HM_ERR_NULL, "NULL argument");
        if (p->dist) return fail(HM_ERR_STATE, "hm_dist_init was already called on this plan");
        if (nranks < 1 || nranks > HM_MAX_PEERS) return fail(HM_ERR_INVALID, "nranks must be 1..%d", HM_MAX_PEERS);
        if (rank < 0 || rank >= nranks) return fail(HM_ERR_INVALID, "rank out of range");
        if (p->L.nparts != nranks || p->L.part != rank)
            return fail(HM_ERR_STATE, "the plan holds part %d of %d, not rank %d of %d", p->L.part, p->L.nparts, rank,
                        nranks);
        {
            std::lock_guard<std::mutex> lock(g_nccl_mu);
            if (!g_nccl.load()) return fail(HM_ERR_UNSUPPORTED, "NCCL unavailable: %s", g_nccl.why);
        }
        std::lock_guard<std::mutex> lock(p->mu);
        HM_DEVICE(p->device);
        HmDist *d = new HmDist;
        p->dist = d; // owned by the plan from here on (released with it, also after a failure below)
        d->nranks = nranks;
        d->rank = rank;
        d->nrows = p->L.nrows;
        d->ncols = p->L.ncols;
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        // HMB200_NCCL_MAX_CTAS > 0 limits the communicator's CTAs.  Measured on 8 B200 at N = 2^20: the
        // 8 MB broadcast takes 33 us with NCCL's default, 62 us with 8 CTAs and 190 us with 2 -- it stops
        // hiding under the 0.3 ms step, so the default is left alone.
        int max_ctas = 0;
        if (const char *e = getenv("HMB200_NCCL_MAX_CTAS")) max_ctas = atoi(e);
        if (g_nccl.CommInitRankConfig && max_ctas > 0) {
            ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
            cfg.minCTAs = 1;
            cfg.maxCTAs = max_ctas;
            HM_NCCL(g_nccl.CommInitRankConfig(&d->comm, nranks, uid, rank, &cfg));
        } else {
            HM_NCCL(g_nccl.CommInitRank(&d->comm, nranks, uid, rank));
        }
        // exchange region: [flag rows | epoch | err | x0 | x1 | y0 | y1], identical layout on all ranks
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        size_t off = 0;
        d->off_flags = off;
        off = up(off + (size_t)HM_MAX_PEERS * FLAG_STRIDE * sizeof(unsigned));
        d->off_epoch = off;
        off += 128;
        d->off_err = off;
        off = up(off + 128);
        for (int s = 0; s < 2; s++) {
            d->off_x[s] = off;
            off = up(off + (size_t)std::max<int64_t>(d->ncols, 1) * 8);
        }
        for (int s = 0; s < 2; s++) {
            d->off_y[s] = off;
            off = up(off + (size_t)std::max<int64_t>(d->nrows, 1) * 8);
        }
        d->bytes = off;
        HM_CUDA(cudaMalloc((void **)&d->base, d->bytes));
        HM_CUDA(cudaMemset(d->base, 0, d->bytes));
        d->peer[rank] = d->base;
        if (nranks > 1) {
            cudaIpcMemHandle_t mine;
            HM_CUDA(cudaIpcGetMemHandle(&mine, d->base));
            DevBuf<char> dsend, drecv;
            HM_CUDA(dsend.alloc(sizeof mine));
            HM_CUDA(drecv.alloc(sizeof mine * (size_t)nranks));
            cudaStream_t st = p->stream;
            HM_CUDA(cudaMemcpyAsync(dsend.p, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
            HM_NCCL(g_nccl.AllGather(dsend.p, drecv.p, sizeof mine, ncclChar, d->comm, st));
            std::vector<cudaIpcMemHandle_t> all((size_t)nranks);
            HM_CUDA(cudaMemcpyAsync(all.data(), drecv.p, sizeof mine * (size_t)nranks, cudaMemcpyDeviceToHost, st));
            HM_CUDA(cudaStreamSynchronize(st));
            for (int q = 0; q < nranks; q++) {
                if (q == rank) continue;
                void *ptr = nullptr;
                HM_CUDA(cudaIpcOpenMemHandle(&ptr, all[(size_t)q], cudaIpcMemLazyEnablePeerAccess));
                d->peer[q] = static_cast<char *>(ptr);
            }
            // nobody signals a flag row before every rank has zeroed its own: one more collective
            HM_NCCL(g_nccl.AllGather(dsend.p, drecv.p, 1, ncclChar, d->comm, st));
            HM_CUDA(cudaStreamSynchronize(st));
        }
        return HM_OK;
    });
}

int32_t hm_dist_buffers(hm_plan *p, double **x2, double **y2)
{
    return guarded([&]() -> int32_t {
        if (int32_t st = need_dist(p)) return st;
        for (int s = 0; s < 2; s++) {
            if (x2) x2[s] = p->dist->x(s);
            if (y2) y2[s] = p->dist->y(p->dist->rank, s);
        }
        return HM_OK;
    });
}

int32_t hm_dist_bcast_x(hm_plan *p, const double *dx_root, int32_t root, int32_t slot, void *stream)
{
    return guarded([&]() -> int32_t {
        if (int32_t st = need_dist(p)) return st;
        HmDist *d = p->dist;
        if (root < 0 || root >= d->nranks) return fail(HM_ERR_INVALID, "root out of range");
        if (slot != 0 && slot != 1) return fail(HM_ERR_INVALID, "slot must be 0 or 1");
        if (d->ncols == 0) return HM_OK;
        HM_DEVICE(p->device);
        const double *src = d->rank == root ? (dx_root ? dx_root : d->x(slot)) : d->x(slot);
        HM_NCCL(g_nccl.Broadcast(src, d->x(slot), (size_t)d->ncols, ncclDouble, root, d->comm, (cudaStream_t)stream));
        return HM_OK;
    });
}

int32_t hm_dist_push_x(hm_plan *p, const double *dx_root, int32_t root, int32_t slot, void *stream)
{
    return guarded([&]() -> int32_t {
        if (int32_t st = need_dist(p)) return st;
        HmDist *d = p->dist;
        if (root < 0 || root >= d->nranks) return fail(HM_ERR_INVALID, "root out of range");
        if (slot != 0 && slot != 1) return fail(HM_ERR_INVALID, "slot must be 0 or 1");
        if (d->rank != root || d->ncols == 0) return HM_OK; // the root's copy engines do all the work
        if (!dx_root) return fail(HM_ERR_NULL, "x is NULL on the root");
        HM_DEVICE(p->device);
        cudaStream_t st = (cudaStream_t)stream;
        if (!d->push_start) {
            HM_CUDA(cudaEventCreateWithFlags(&d->push_start, cudaEventDisableTiming));
            for (int q = 0; q < d->nranks; q++) {
                HM_CUDA(cudaStreamCreateWithFlags(&d->push_stream[q], cudaStreamNonBlocking));
                HM_CUDA(cudaEventCreateWithFlags(&d->push_done[q], cudaEventDisableTiming));
            }
        }
        // fork: one copy stream per destination, so the copies spread over the copy engines
        HM_CUDA(cudaEventRecord(d->push_start, st));
        const size_t bytes = (size_t)d->ncols * 8;
        for (int q = 0; q < d->nranks; q++) {
            double *dst = reinterpret_cast<double *>(d->peer[q] + d->off_x[slot]);
            HM_CUDA(cudaStreamWaitEvent(d->push_stream[q], d->push_start, 0));
            if (dst != dx_root) HM_CUDA(cudaMemcpyAsync(dst, dx_root, bytes, cudaMemcpyDeviceToDevice, d->push_stream[q]));
            HM_CUDA(cudaEventRecord(d->push_done[q], d->push_stream[q]));
        }
        d->push_pending = true; // joined by the next barrier enqueued on this plan
        return HM_OK;
    });
}

int32_t hm_dist_matvec_device(hm_plan *p, const double *dx, int32_t yslot, int32_t accumulate, void *stream)
{
    return guarded([&]() -> int32_t {
        if (int32_t st = need_dist(p)) return st;
        HmDist *d = p->dist;
        if (yslot != 0 && yslot != 1) return fail(HM_ERR_INVALID, "slot must be 0 or 1");
        if (!dx && d->ncols > 0) return fail(HM_ERR_NULL, "x is NULL");
        HM_DEVICE(p->device);
        HmPeers pe;
        pe.n = d->nranks;
        for (int q = 0; q < d->nranks; q++) pe.y[q] = d->y(q, yslot);
        if (int32_t st = hm_matvec_device_peers(p, dx, pe.y[d->rank], accumulate, stream, &pe)) return st;
        return launch_barrier(p, (cudaStream_t)stream);
    });
}

int32_t hm_dist_barrier(hm_plan *p, void *stream)
{
    return guarded([&]() -> int32_t {
        if (int32_t st = need_dist(p)) return st;
        HM_DEVICE(p->device);
        return launch_barrier(p, (cudaStream_t)stream);
    });
}

int32_t hm_dist_check(hm_plan *p)
{
    return guarded([&]() -> int32_t {
        if (int32_t st = need_dist(p)) return st;
        HM_DEVICE(p->device);
        int err = 0;
        HM_CUDA(cudaMemcpy(&err, p->dist->base + p->dist->off_err, sizeof err, cudaMemcpyDeviceToHost));
        if (err) return fail(HM_ERR_CUDA, "a rank did not reach the barrier in time (peer lost?)");
        return HM_OK;
    });
}

int32_t hm_dist_matvec(hm_plan *p, const double *x, int64_t incx, double *y, int64_t incy, int32_t root,
                       int32_t accumulate)
{
    return guarded([&]() -> int32_t {
        if (int32_t st = need_dist(p)) return st;
        HmDist *d = p->dist;
        if (root < 0 || root >= d->nranks) return fail(HM_ERR_INVALID, "root out of range");
        if (incx <= 0 || incy <= 0) return fail(HM_ERR_INVALID, "strides must be positive");
        const bool is_root = d->rank == root;
        if (is_root && !x && d->ncols > 0) return fail(HM_ERR_NULL, "x is NULL on the root");
        if (accumulate && !y && d->nrows > 0) return fail(HM_ERR_NULL, "accumulate needs y on every rank");
        std::lock_guard<std::mutex> lock(p->mu);
        HM_DEVICE(p->device);
        cudaStream_t st = p->stream;
        const int64_t nc = d->ncols, nr = d->nrows;
        // y slots alternate between calls (all ranks make the same sequence of calls): a rank that is
        // one call ahead stores into the slot its peers are not reading back
        const int ys = (int)(d->calls++ & 1);
        if (is_root && nc > 0) {
            if (incx == 1) {
                HM_CUDA(cudaMemcpyAsync(d->x(0), x, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
            } else {
                if (!d->hx) HM_CUDA(cudaMallocHost((void **)&d->hx, (size_t)nc * 8));
                for (int64_t j = 0; j < nc; j++) d->hx[j] = x[j * incx];
                HM_CUDA(cudaMemcpyAsync(d->x(0), d->hx, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
            }
        }
        if (nc > 0) HM_NCCL(g_nccl.Broadcast(d->x(0), d->x(0), (size_t)nc, ncclDouble, root, d->comm, st));
        if (y && incy != 1 && nr > 0 && !d->hy) HM_CUDA(cudaMallocHost((void **)&d->hy, (size_t)nr * 8));
        if (accumulate && nr > 0) {
            // every rank passes the same y: its own rows are what this rank accumulates into
            const int64_t r0 = p->L.row_begin, rows = p->L.row_end - r0;
            if (incy == 1) {
                HM_CUDA(cudaMemcpyAsync(d->y(d->rank, ys) + r0, y + r0, (size_t)rows * 8, cudaMemcpyHostToDevice, st));
            } else {
                for (int64_t i = 0; i < rows; i++) d->hy[r0 + i] = y[(r0 + i) * incy];
                HM_CUDA(cudaMemcpyAsync(d->y(d->rank, ys) + r0, d->hy + r0, (size_t)rows * 8, cudaMemcpyHostToDevice, st));
            }
        }
        HmPeers pe;
        pe.n = d->nranks;
        for (int q = 0; q < d->nranks; q++) pe.y[q] = d->y(q, ys);
        if (int32_t rc = hm_matvec_device_peers(p, d->x(0), pe.y[d->rank], accumulate, st, &pe)) return rc;
        if (int32_t rc = launch_barrier(p, st)) return rc;
        if (y && nr > 0)
            HM_CUDA(cudaMemcpyAsync(incy == 1 ? y : d->hy, d->y(d->rank, ys), (size_t)nr * 8, cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        if (y && incy != 1)
            for (int64_t i = 0; i < nr; i++) y[i * incy] = d->hy[i];
        int err = 0;
        HM_CUDA(cudaMemcpy(&err, d->base + d->off_err, sizeof err, cudaMemcpyDeviceToHost));
        if (err) return fail(HM_ERR_CUDA, "a rank did not reach the barrier in time (peer lost?)");
        return HM_OK;
    });
}

} // extern "C"
