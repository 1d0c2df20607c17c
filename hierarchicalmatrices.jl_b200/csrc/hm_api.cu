// hm_api.cu -- C ABI (include/hmb200.h): builder, plan, mul!.
//
// No CPU fallback: every compute entry point needs a CUDA device and fails
// with HM_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hmb200.h"
#include "hm_kernels.cuh"
#include "hm_layout.h"
#include "hm_tree.h"

#include "hm_internal.h"

thread_local char hm_err_msg[512] = "";

int32_t fail(hm_status st, const char *fmt, ...) noexcept
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(hm_err_msg, sizeof hm_err_msg, fmt, ap);
    va_end(ap);
    return (int32_t)st;
}

namespace {


// ---------------------------------------------------------------------------
// staging of raw leaf data onto the device (builder path)
// ---------------------------------------------------------------------------
class Uploader {
  public:
    static constexpr size_t WIN = (size_t)32 << 20;    // pinned window
    static constexpr size_t CHUNK = (size_t)256 << 20; // device arena chunk

    cudaError_t init()
    {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int i = 0; i < 2; i++) {
            if ((e = cudaMallocHost((void **)&pin_[i], WIN)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }

    ~Uploader()
    {
        for (int i = 0; i < 2; i++) {
            if (pin_[i]) cudaFreeHost(pin_[i]);
            if (ev_[i]) cudaEventDestroy(ev_[i]);
        }
        for (void *c : chunks_) cudaFree(c);
        if (st_) cudaStreamDestroy(st_);
    }

    // copy a column-major rows x cols matrix (leading dimension ld) to a tight
    // device copy; returns the device pointer in *out
    cudaError_t put(const double *src, int64_t rows, int64_t cols, int64_t ld, const double **out)
    {
        size_t bytes = (size_t)rows * (size_t)cols * sizeof(double);
        *out = nullptr;
        if (bytes == 0) return cudaSuccess;
        char *d = nullptr;
        cudaError_t e = reserve(bytes, &d);
        if (e != cudaSuccess) return e;
        *out = reinterpret_cast<const double *>(d);
        if (ld == rows) return append(d, reinterpret_cast<const char *>(src), bytes);
        for (int64_t j = 0; j < cols; j++) {
            e = append(d + (size_t)j * rows * sizeof(double), reinterpret_cast<const char *>(src + j * ld),
                       (size_t)rows * sizeof(double));
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }

    cudaError_t finish()
    {
        cudaError_t e = flush();
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(st_);
    }

    size_t bytes_staged() const { return total_; }

  private:
    cudaError_t reserve(size_t bytes, char **out)
    {
        size_t need = (bytes + 15) & ~(size_t)15;
        if (need > left_) {
            size_t sz = std::max(need, CHUNK);
            void *c = nullptr;
            cudaError_t e = cudaMalloc(&c, sz);
            if (e != cudaSuccess) return e;
            chunks_.push_back(c);
            cur_ = (char *)c;
            left_ = sz;
        }
        *out = cur_;
        cur_ += need;
        left_ -= need;
        total_ += need;
        return cudaSuccess;
    }

    cudaError_t flush()
    {
        if (win_used_ == 0) {
            win_dev_ = nullptr;
            return cudaSuccess;
        }
        cudaError_t e = cudaMemcpyAsync(win_dev_, pin_[cur_buf_], win_used_, cudaMemcpyHostToDevice, st_);
        if (e != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev_[cur_buf_], st_)) != cudaSuccess) return e;
        cur_buf_ ^= 1;
        e = cudaEventSynchronize(ev_[cur_buf_]); // the other buffer's copy is done
        win_used_ = 0;
        win_dev_ = nullptr;
        return e;
    }

    cudaError_t append(char *d, const char *src, size_t bytes)
    {
        while (bytes) {
            if (win_dev_ && d > win_dev_ + win_used_ && (size_t)(d - (win_dev_ + win_used_)) < 64 &&
                (size_t)(d - win_dev_) < WIN) {
                size_t gap = (size_t)(d - (win_dev_ + win_used_));
                memset(pin_[cur_buf_] + win_used_, 0, gap); // alignment padding between leaves
                win_used_ += gap;
            }
            if (win_dev_ && (d != win_dev_ + win_used_ || win_used_ == WIN)) {
                cudaError_t e = flush();
                if (e != cudaSuccess) return e;
            }
            if (!win_dev_) {
                win_dev_ = d;
                win_used_ = 0;
            }
            size_t n = std::min(bytes, WIN - win_used_);
            memcpy(pin_[cur_buf_] + win_used_, src, n);
            win_used_ += n;
            d += n;
            src += n;
            bytes -= n;
        }
        return cudaSuccess;
    }

    cudaStream_t st_ = nullptr;
    char *pin_[2] = {nullptr, nullptr};
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    int cur_buf_ = 0;
    char *win_dev_ = nullptr;
    size_t win_used_ = 0;
    std::vector<void *> chunks_;
    char *cur_ = nullptr;
    size_t left_ = 0, total_ = 0;
};

// Layout parameters; HMB200_RMAX / HMB200_CMAX0 / HMB200_CMAX1 / HMB200_NBIG override the
// defaults (tuning experiments: finer items shorten the tail of each launch when a GPU
// owns only a small part of the operator).
HmLayoutParams layout_params()
{
    HmLayoutParams prm;
    auto geti = [](const char *name, int def, int lo, int hi) {
        const char *e = getenv(name);
        if (!e) return def;
        int v = atoi(e);
        return (v < lo || v > hi) ? def : v;
    };
    prm.rmax = geti("HMB200_RMAX", prm.rmax, 8, 512);
    prm.cmax0 = geti("HMB200_CMAX0", prm.cmax0, 8, HM_SMAX);
    prm.cmax1 = geti("HMB200_CMAX1", prm.cmax1, 64, HM_SMAX);
    prm.nbig = geti("HMB200_NBIG", prm.nbig, 64, 1 << 30);
    return prm;
}

void fill_stats(const HmLayout &L, hm_stats *s)
{
    memset(s, 0, sizeof *s);
    s->nrows = L.nrows;
    s->ncols = L.ncols;
    s->n_dense = L.n_dense;
    s->n_lowrank = L.n_lowrank;
    s->n_bary2d = L.n_bary2d;
    s->dense_words = L.dense_words;
    s->lowrank_words = L.lowrank_words;
    s->core_words = L.core_words_all;
    s->algorithmic_bytes = 8 * (L.dense_words + L.lowrank_words) + 8 * L.ncols + 8 * L.nrows;
    s->row_begin = L.row_begin;
    s->row_end = L.row_end;
    s->part_words = L.part_words;
    s->v_stream_bytes = 8 * L.vstream_words;
    s->u_stream_bytes = 8 * L.ustream_words;
    s->stored_bytes = 8 * (L.vstream_words + L.ustream_words + L.core_words);
    s->partial_bytes = 8 * L.partial_words;
    s->n_stage1_items = (int64_t)L.items1.size();
    s->n_stage2_blocks = (int64_t)L.cores.size();
    s->n_stage3_items = (int64_t)L.items3.size();
    s->n_stage3_rounds = (int64_t)L.round_begin.size() - 1;
    s->part_algorithmic_bytes = 8 * L.part_words + 8 * L.ncols + 8 * (L.row_end - L.row_begin);
    s->part_v_words = L.part_v_words;
    s->part_core_words = L.part_core_words;
    s->part_u_words = L.part_u_words;
    s->part_dense_words = L.part_dense_words;
}

} // namespace

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------
struct hm_builder {
    int64_t nrows = 0, ncols = 0;
    int device = -1;
    std::vector<HmLeaf> leaves;
    Uploader *up = nullptr;
    ~hm_builder() { delete up; }
};


namespace {

// Builds the device side of a plan from a finished layout whose leaves carry
// their sources (device copies or kernel-evaluation parameters).
int32_t materialize(hm_plan *P, const double *dpx, const double *dpy)
{
    HmLayout &L = P->L;
    HM_CUDA(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));
    cudaStream_t st = P->stream;
    if (!P->matrix_free) {
        HM_CUDA(P->vstream.alloc((size_t)L.vstream_words));
        HM_CUDA(P->ustream.alloc((size_t)L.ustream_words));
    }
    HM_CUDA(P->core.alloc((size_t)L.core_words));
    HM_CUDA(P->svec.alloc((size_t)std::max<int64_t>(L.s_words, 1)));
    HM_CUDA(P->partial.alloc((size_t)std::max<int64_t>(L.partial_words, 1)));
    if (!P->matrix_free) {
        if (L.vstream_words) HM_CUDA(cudaMemsetAsync(P->vstream.p, 0, (size_t)L.vstream_words * 8, st));
        if (L.ustream_words) HM_CUDA(cudaMemsetAsync(P->ustream.p, 0, (size_t)L.ustream_words * 8, st));
    }
    if (L.core_words) HM_CUDA(cudaMemsetAsync(P->core.p, 0, (size_t)L.core_words * 8, st));
    HM_CUDA(cudaMemsetAsync(P->partial.p, 0, P->partial.n * 8, st));
    hm_trace_point("materialize: allocations");
    HM_CUDA(P->items1.upload(L.items1, st));
    HM_CUDA(P->items3.upload(L.items3, st));
    HM_CUDA(P->runs.upload(L.runs, st));
    HM_CUDA(P->cores.upload(L.cores, st));
    HM_CUDA(P->plist.upload(L.plist, st));
    HM_CUDA(P->s1ent.upload(L.s1ent, st));
    hm_trace_point("materialize: item / run / core tables uploaded");
    HM_CUDA(P->counters.alloc(std::max<size_t>(L.cores.size(), 1)));
    HM_CUDA(cudaMemsetAsync(P->counters.p, 0, P->counters.n * sizeof(int), st));
    {
        const char *e = getenv("HMB200_FUSE_STAGE2");
        P->fuse = (e && e[0] == '1') && (size_t)L.max_r * 8 <= HM_SMAX && !P->matrix_free;
    }
    {
        std::vector<int32_t> big;
        for (size_t c = 0; c < L.cores.size(); c++)
            if (L.cores[c].npl > HM_CORE_BIG) big.push_back((int32_t)c);
        P->nbig = (int64_t)big.size();
        HM_CUDA(P->bigcores.upload(big, st));
    }
    if (P->matrix_free) {
        // the apply kernels read the fill tables themselves; only the cores are materialised
        int units = 1;
        for (const HmItem &it : L.items1) {
            int nch, CH;
            hm_free1_split(it.S, it.nrun, nch, CH);
            units = std::max(units, it.nrun * nch);
        }
        if ((size_t)units * 20 * sizeof(double) > 160 * 1024)
            return fail(HM_ERR_UNSUPPORTED, "matrix-free: a column segment is covered by too many leaves");
        for (const HmItem &it : L.items3)
            if (it.F > HM_THREADS || it.nrun > HM_MAXRUNS || it.S > HM_SMAX)
                return fail(HM_ERR_UNSUPPORTED, "matrix-free: row segment outside the kernel limits");
        P->free1_units = units;
        int zcap = 2;
        for (const HmItem &it : L.items3) zcap = std::max(zcap, (int)it.S);
        P->free3_zcap = (zcap + 1) & ~1;
        hm_trace_point("materialize: limits");
        std::vector<HmFreeEnt> ent1(L.fill1.size());
        for (const HmItem &it : L.items1)
            for (int32_t e = it.run0; e < it.run0 + it.nrun; e++) {
                const HmFill &f = L.fill1[(size_t)e];
                const HmLeaf &l = L.leaves[(size_t)f.leaf];
                ent1[(size_t)e] = HmFreeEnt{0.5 * (l.c + l.d), 0.5 * (l.d - l.c), l.yj0 + f.off, f.dst - it.slab};
            }
        std::vector<HmFreeRun> run3(L.fill3.size());
        for (size_t i = 0; i < L.fill3.size(); i++) {
            const HmFill &f = L.fill3[i];
            const HmLeaf &l = L.leaves[(size_t)f.leaf];
            run3[i] = HmFreeRun{0.5 * (l.a + l.b), 0.5 * (l.b - l.a), l.xi0 + f.off, l.yj0 + f.k0, f.k0, f.kn};
        }
        hm_trace_point("materialize: free tables built");
        DevBuf<HmLeaf> dleaves;
        DevBuf<int32_t> dcore_leaf;
        HM_CUDA(dleaves.upload(L.leaves, st));
        HM_CUDA(P->f_ent1.upload(ent1, st));
        HM_CUDA(P->f_run3.upload(run3, st));
        HM_CUDA(dcore_leaf.upload(L.core_leaf, st));
        hm_trace_point("materialize: free tables uploaded");
        HM_CUDA(hm_launch_fillcore(P->cores.p, dcore_leaf.p, (int64_t)L.cores.size(), dleaves.p, P->core.p,
                                   P->cheb, P->kernel_id, st));
        {
            // Chebyshev form (default; HMB200_FREE_FORM=bary keeps the reference's barycentric arithmetic):
            // cores become C F C', C = values at the Chebyshev nodes -> Chebyshev coefficients
            const char *e = getenv("HMB200_FREE_FORM");
            P->free_cheb = !(e && e[0] == 'b') && P->cheb.r == 20;
            if (P->free_cheb) {
                const int R = P->cheb.r;
                std::vector<double> Cm((size_t)R * R);
                for (int k = 0; k < R; k++) {
                    long double tm2 = 1.0L, tm1 = (long double)P->cheb.node[k];
                    for (int q = 0; q < R; q++) {
                        long double tq = q == 0 ? 1.0L : q == 1 ? (long double)P->cheb.node[k] : 2.0L * P->cheb.node[k] * tm1 - tm2;
                        if (q >= 2) {
                            tm2 = tm1;
                            tm1 = tq;
                        }
                        Cm[(size_t)q + (size_t)k * R] = (double)((q == 0 ? 1.0L : 2.0L) * tq / R);
                    }
                }
                // Chebyshev differentiation matrix of the first-kind points (barycentric form):
                // D[i][j] = (w_j / w_i) / (x_i - x_j), D[i][i] = -sum_{j != i} D[i][j]
                std::vector<double> Dm((size_t)R * R);
                for (int i = 0; i < R; i++) {
                    long double diag = 0.0L;
                    for (int j = 0; j < R; j++) {
                        if (j == i) continue;
                        const long double dij = ((long double)P->cheb.lam[j] / (long double)P->cheb.lam[i]) /
                                                ((long double)P->cheb.node[i] - (long double)P->cheb.node[j]);
                        Dm[(size_t)i + (size_t)j * R] = (double)dij;
                        diag -= dij;
                    }
                    Dm[(size_t)i + (size_t)i * R] = (double)diag;
                }
                DevBuf<double> dC, dD;
                HM_CUDA(dC.upload(Cm, st));
                HM_CUDA(dD.upload(Dm, st));
                HM_CUDA(hm_launch_core_cheb(P->cores.p, dcore_leaf.p, (int64_t)L.cores.size(), dleaves.p, P->core.p, dC.p,
                                            dD.p, P->cheb, st));
                HM_CUDA(cudaStreamSynchronize(st));
            }
        }
        HM_CUDA(cudaStreamSynchronize(st));
    } else {
        // temporary tables for the fill kernels
        DevBuf<HmLeaf> dleaves;
        DevBuf<HmFill> dfill1, dfill3;
        DevBuf<int32_t> dcore_leaf;
        HM_CUDA(dleaves.upload(L.leaves, st));
        HM_CUDA(dfill1.upload(L.fill1, st));
        HM_CUDA(dfill3.upload(L.fill3, st));
        HM_CUDA(dcore_leaf.upload(L.core_leaf, st));
        HM_CUDA(hm_launch_fill1(dfill1.p, (int64_t)L.fill1.size(), dleaves.p, P->vstream.p, dpy, P->cheb, st));
        HM_CUDA(hm_launch_fill3(dfill3.p, (int64_t)L.fill3.size(), dleaves.p, P->ustream.p, dpx, dpy, P->cheb,
                                P->kernel_id, st));
        HM_CUDA(hm_launch_fillcore(P->cores.p, dcore_leaf.p, (int64_t)L.cores.size(), dleaves.p, P->core.p,
                                   P->cheb, P->kernel_id, st));
        HM_CUDA(cudaStreamSynchronize(st));
    }
    // the raw sources are gone after this point: drop the dangling pointers
    for (HmLeaf &l : L.leaves) l.dU = l.dC = l.dV = nullptr;
    return HM_OK;
}

int32_t check_block(const hm_builder *b, int64_t m, int64_t n, int64_t row0, int64_t col0)
{
    if (m < 0 || n < 0) return fail(HM_ERR_SHAPE, "negative block extent %lld x %lld", (long long)m, (long long)n);
    if (row0 < 0 || col0 < 0 || row0 + m > b->nrows || col0 + n > b->ncols)
        return fail(HM_ERR_RANGE, "block [%lld,%lld) x [%lld,%lld) outside the %lld x %lld operator",
                    (long long)row0, (long long)(row0 + m), (long long)col0, (long long)(col0 + n),
                    (long long)b->nrows, (long long)b->ncols);
    return HM_OK;
}

int32_t stage(hm_builder *b, const double *src, int64_t rows, int64_t cols, int64_t ld, const char *what,
              const double **out)
{
    *out = nullptr;
    if (rows * cols == 0 || b->device < 0) return HM_OK;
    if (!src) return fail(HM_ERR_NULL, "%s is NULL", what);
    if (ld < rows) return fail(HM_ERR_SHAPE, "%s: leading dimension %lld < rows %lld", what, (long long)ld, (long long)rows);
    HM_CUDA(b->up->put(src, rows, cols, ld, out));
    return HM_OK;
}

} // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *hm_last_error(void) { return hm_err_msg; }

int32_t hm_version(void) { return 100; }

int32_t hm_blockrank_f64(void) { return hm_blockrank_double(); }
int32_t hm_blocksize_f64(void) { return hm_blocksize_double(); }

int32_t hm_builder_create(hm_builder **out, int64_t nrows, int64_t ncols, int32_t dtype, int32_t device)
{
    return guarded([&]() -> int32_t {
        if (!out) return fail(HM_ERR_NULL, "out is NULL");
        *out = nullptr;
        if (dtype != HM_F64) return fail(HM_ERR_UNSUPPORTED, "only Float64 (HM_F64) operators are supported");
        if (nrows < 0 || ncols < 0) return fail(HM_ERR_SHAPE, "negative operator extent");
        if (nrows >= ((int64_t)1 << 31) || ncols >= ((int64_t)1 << 31))
            return fail(HM_ERR_UNSUPPORTED, "operator extent >= 2^31");
        hm_builder *b = new (std::nothrow) hm_builder;
        if (!b) return fail(HM_ERR_NOMEM, "out of host memory");
        b->nrows = nrows;
        b->ncols = ncols;
        b->device = device;
        if (device >= 0) {
            DeviceGuard g(device);
            if (!g.ok) {
                delete b;
                cudaGetLastError();
                return fail(HM_ERR_CUDA, "cannot select CUDA device %d: %s", device, cudaGetErrorString(g.err));
            }
            b->up = new (std::nothrow) Uploader;
            cudaError_t e = b->up ? b->up->init() : cudaErrorMemoryAllocation;
            if (e != cudaSuccess) {
                delete b;
                cudaGetLastError();
                return fail(HM_ERR_CUDA, "staging setup failed: %s", cudaGetErrorString(e));
            }
        }
        *out = b;
        return HM_OK;
    });
}

int32_t hm_builder_destroy(hm_builder *b)
{
    return guarded([&]() -> int32_t {
        if (!b) return HM_OK;
        if (b->device >= 0) {
            DeviceGuard g(b->device);
            delete b;
        } else {
            delete b;
        }
        return HM_OK;
    });
}

int32_t hm_builder_add_dense(hm_builder *b, const double *A, int64_t m, int64_t n, int64_t lda, int64_t row0,
                             int64_t col0)
{
    return guarded([&]() -> int32_t {
        if (!b) return fail(HM_ERR_NULL, "builder is NULL");
        if (int32_t st = check_block(b, m, n, row0, col0)) return st;
        HmLeaf l{};
        l.kind = HM_LEAF_DENSE;
        l.source = b->device >= 0 ? HM_SRC_COPY : HM_SRC_NONE;
        l.row0 = row0;
        l.col0 = col0;
        l.m = m;
        l.n = n;
        if (b->device >= 0) {
            HM_DEVICE(b->device);
            if (int32_t st = stage(b, A, m, n, lda, "A", &l.dU)) return st;
        }
        l.ldu = m;
        b->leaves.push_back(l);
        return HM_OK;
    });
}

static int32_t add_factored(hm_builder *b, int32_t kind, const double *U, int64_t ldu, const double *C,
                            int64_t ldc, const double *V, int64_t ldv, int64_t m, int64_t n, int64_t r,
                            int64_t row0, int64_t col0)
{
    return guarded([&]() -> int32_t {
        if (!b) return fail(HM_ERR_NULL, "builder is NULL");
        if (int32_t st = check_block(b, m, n, row0, col0)) return st;
        if (r < 0) return fail(HM_ERR_SHAPE, "negative rank");
        if (r > 2048) return fail(HM_ERR_UNSUPPORTED, "rank %lld > 2048", (long long)r);
        HmLeaf l{};
        l.kind = kind;
        l.source = b->device >= 0 ? HM_SRC_COPY : HM_SRC_NONE;
        l.row0 = row0;
        l.col0 = col0;
        l.m = m;
        l.n = n;
        l.ru = l.rv = (int32_t)r;
        if (b->device >= 0) {
            HM_DEVICE(b->device);
            if (int32_t st = stage(b, U, m, r, ldu, "U", &l.dU)) return st;
            if (kind == HM_LEAF_LOWRANK) {
                if (int32_t st = stage(b, C, r, 1, r, "S", &l.dC)) return st;
            } else {
                if (int32_t st = stage(b, C, r, r, ldc, "F", &l.dC)) return st;
            }
            if (int32_t st = stage(b, V, n, r, ldv, "V", &l.dV)) return st;
        }
        l.ldu = m;
        l.ldc = r;
        l.ldv = n;
        b->leaves.push_back(l);
        return HM_OK;
    });
}

int32_t hm_builder_add_lowrank(hm_builder *b, const double *U, int64_t ldu, const double *S, const double *V,
                               int64_t ldv, int64_t m, int64_t n, int64_t r, int64_t row0, int64_t col0)
{
    return guarded([&]() -> int32_t {
        return add_factored(b, HM_LEAF_LOWRANK, U, ldu, S, r, V, ldv, m, n, r, row0, col0);
    });
}

int32_t hm_builder_add_bary2d(hm_builder *b, const double *U, int64_t ldu, const double *F, int64_t ldf,
                              const double *V, int64_t ldv, int64_t m, int64_t n, int64_t r, int64_t row0,
                              int64_t col0)
{
    return guarded([&]() -> int32_t {
        return add_factored(b, HM_LEAF_BARY2D, U, ldu, F, ldf, V, ldv, m, n, r, row0, col0);
    });
}

// EvenBarycentricMatrix (reference src/BarycentricMatrix.jl:5-45, apply src/algebra.jl:168-239):
// entry (i, j) = sum_k F[j,k] W[k,i] when (shift + row0 + i + col0 + j) is even, else 0.  Packed as a
// rank-2r LowRankMatrix leaf with unit Sigma: factor columns [0, r) carry the rows with row0+i even
// and the columns they pair with, [r, 2r) the other class; the masked-out entries are stored zeros.
int32_t hm_builder_add_evenbary(hm_builder *b, const double *W, int64_t ldw, const double *F, int64_t ldf,
                                int64_t m, int64_t n, int64_t r, int64_t row0, int64_t col0,
                                int32_t shift_parity)
{
    return guarded([&]() -> int32_t {
        if (!b) return fail(HM_ERR_NULL, "builder is NULL");
        if (int32_t st = check_block(b, m, n, row0, col0)) return st;
        if (r < 0) return fail(HM_ERR_SHAPE, "negative rank");
        if (r > 1024) return fail(HM_ERR_UNSUPPORTED, "rank %lld > 1024", (long long)r);
        if (shift_parity != 0 && shift_parity != 1) return fail(HM_ERR_INVALID, "shift_parity must be 0 or 1");
        if (b->device >= 0 && m * r > 0 && !W) return fail(HM_ERR_NULL, "W is NULL");
        if (b->device >= 0 && n * r > 0 && !F) return fail(HM_ERR_NULL, "F is NULL");
        if (m * r > 0 && ldw < r) return fail(HM_ERR_SHAPE, "W: leading dimension %lld < rank %lld", (long long)ldw, (long long)r);
        if (n * r > 0 && ldf < n) return fail(HM_ERR_SHAPE, "F: leading dimension %lld < rows %lld", (long long)ldf, (long long)n);
        const int64_t r2 = 2 * r;
        std::vector<double> U, V, S;
        if (b->device >= 0) {
            try {
                U.assign((size_t)(m * r2), 0.0);
                V.assign((size_t)(n * r2), 0.0);
                S.assign((size_t)r2, 1.0);
            } catch (const std::bad_alloc &) {
                return fail(HM_ERR_NOMEM, "out of host memory");
            }
            for (int64_t i = 0; i < m; i++) {
                const int64_t cls = (row0 + i) & 1;
                for (int64_t k = 0; k < r; k++) U[(size_t)(i + (cls * r + k) * m)] = W[k + i * ldw];
            }
            for (int64_t j = 0; j < n; j++) {
                const int64_t cls = (shift_parity + col0 + j) & 1;
                for (int64_t k = 0; k < r; k++) V[(size_t)(j + (cls * r + k) * n)] = F[j + k * ldf];
            }
        }
        int32_t st = add_factored(b, HM_LEAF_LOWRANK, U.data(), std::max<int64_t>(m, 1), S.data(), r2, V.data(),
                                  std::max<int64_t>(n, 1), m, n, r2, row0, col0);
        if (st == HM_OK) b->leaves.back().extra_words = (m + n) * r + r2;
        return st;
    });
}

int32_t hm_builder_layout_stats(hm_builder *b, int32_t part, int32_t nparts, hm_stats *out)
{
    return guarded([&]() -> int32_t {
        if (!b || !out) return fail(HM_ERR_NULL, "NULL argument");
        HmLayout L;
        std::string err = hm_build_layout(b->leaves, b->nrows, b->ncols, part, nparts, layout_params(), L);
        if (!err.empty()) return fail(HM_ERR_INVALID, "%s", err.c_str());
        fill_stats(L, out);
        return HM_OK;
    });
}

int32_t hm_plan_finalize_part(hm_builder *b, int32_t part, int32_t nparts, hm_plan **out)
{
    return guarded([&]() -> int32_t {
        if (!b || !out) return fail(HM_ERR_NULL, "NULL argument");
        *out = nullptr;
        if (b->device < 0) return fail(HM_ERR_STATE, "structure-only builder (device = -1) cannot be finalized");
        HM_DEVICE(b->device);
        HM_CUDA(b->up->finish());
        hm_plan *P = new (std::nothrow) hm_plan;
        if (!P) return fail(HM_ERR_NOMEM, "out of host memory");
        P->device = b->device;
        std::string err = hm_build_layout(b->leaves, b->nrows, b->ncols, part, nparts, layout_params(), P->L);
        if (!err.empty()) {
            delete P;
            return fail(HM_ERR_INVALID, "%s", err.c_str());
        }
        P->cheb.r = 0;
        int32_t st = materialize(P, nullptr, nullptr);
        if (st != HM_OK) {
            delete P;
            return st;
        }
        *out = P;
        return HM_OK;
    });
}

int32_t hm_plan_finalize(hm_builder *b, const int32_t *devices, int32_t ndev, hm_plan **out)
{
    return guarded([&]() -> int32_t {
        if (!b || !out) return fail(HM_ERR_NULL, "NULL argument");
        if (ndev != 1)
            return fail(HM_ERR_UNSUPPORTED,
                        "one plan drives one GPU; for multi-GPU run one process per GPU with hm_plan_finalize_part");
        if (devices && devices[0] != b->device)
            return fail(HM_ERR_INVALID, "devices[0] = %d but the builder staged its data on device %d", devices[0],
                        b->device);
        return hm_plan_finalize_part(b, 0, 1, out);
    });
}

int32_t hm_plan_destroy(hm_plan *p)
{
    return guarded([&]() -> int32_t {
        if (!p) return HM_OK;
        DeviceGuard g(p->device);
        delete p;
        return HM_OK;
    });
}

int32_t hm_plan_stats(const hm_plan *p, hm_stats *out)
{
    return guarded([&]() -> int32_t {
        if (!p || !out) return fail(HM_ERR_NULL, "NULL argument");
        fill_stats(p->L, out);
        if (p->matrix_free) { // nothing but the cores is stored; the apply reads tables and points
            out->v_stream_bytes = out->u_stream_bytes = 0;
            out->stored_bytes = 8 * p->L.core_words;
        }
        return HM_OK;
    });
}

int32_t hm_plan_scale(hm_plan *p, const double *b, int64_t incb, int32_t side)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        if (p->matrix_free) return fail(HM_ERR_UNSUPPORTED, "not available on a matrix-free plan (hm_assemble_kernel_free)");
        if (side != 0 && side != 1) return fail(HM_ERR_INVALID, "side must be 0 (columns) or 1 (rows)");
        if (incb <= 0) return fail(HM_ERR_INVALID, "stride must be positive");
        const HmLayout &L = p->L;
        const int64_t n = side == 0 ? L.ncols : L.nrows;
        if (n == 0) return HM_OK;
        if (!b) return fail(HM_ERR_NULL, "b is NULL");
        std::lock_guard<std::mutex> lock(p->mu);
        HM_DEVICE(p->device);
        std::vector<double> hb((size_t)n);
        for (int64_t i = 0; i < n; i++) hb[(size_t)i] = b[i * incb];
        DevBuf<double> db;
        HM_CUDA(db.alloc((size_t)n));
        cudaStream_t st = p->stream;
        HM_CUDA(cudaMemcpyAsync(db.p, hb.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        if (side == 1)
            HM_CUDA(hm_launch_scale_rows(p->items3.p, (int64_t)L.items3.size(), p->ustream.p, db.p, st));
        else
            HM_CUDA(hm_launch_scale_cols(p->items1.p, (int64_t)L.items1.size(), p->vstream.p, p->items3.p,
                                         (int64_t)L.items3.size(), p->runs.p, p->ustream.p, db.p, st));
        HM_CUDA(cudaStreamSynchronize(st));
        return HM_OK;
    });
}

int32_t hm_plan_timing_begin(hm_plan *p, int32_t max_calls)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        if (max_calls < 0 || max_calls > 100000) return fail(HM_ERR_INVALID, "max_calls out of range");
        HM_DEVICE(p->device);
        while ((int)p->tev.size() < 4 * max_calls) {
            cudaEvent_t e;
            HM_CUDA(cudaEventCreate(&e));
            p->tev.push_back(e);
        }
        p->tcap = max_calls;
        p->tcount = 0;
        return HM_OK;
    });
}

int32_t hm_plan_timing_end(hm_plan *p, double *stage_ms3, int64_t *ncalls)
{
    return guarded([&]() -> int32_t {
        if (!p || !stage_ms3 || !ncalls) return fail(HM_ERR_NULL, "NULL argument");
        HM_DEVICE(p->device);
        stage_ms3[0] = stage_ms3[1] = stage_ms3[2] = 0.0;
        for (int c = 0; c < p->tcount; c++) {
            cudaEvent_t *ev = &p->tev[(size_t)c * 4];
            HM_CUDA(cudaEventSynchronize(ev[3]));
            for (int s = 0; s < 3; s++) {
                float ms = 0.f;
                HM_CUDA(cudaEventElapsedTime(&ms, ev[s], ev[s + 1]));
                stage_ms3[s] += ms;
            }
        }
        *ncalls = p->tcount;
        p->tcap = 0;
        p->tcount = 0;
        return HM_OK;
    });
}

int32_t hm_plan_launches_per_matvec(const hm_plan *p)
{
    if (!p) return 0;
    int n = 0;
    if (p->nested) { // up (subtrees, top), cores, down (top, subtrees), dense rounds
        n = (p->n_rows.nnodes > 0) + (p->n_cols.nbase > 0);
        for (const HmNestDev *T : {&p->n_cols, &p->n_rows})
            for (int k = 0; k < T->ntiers; k++) n += T->tier_sub0[k + 1] > T->tier_sub0[k];
        for (size_t r = 0; r + 1 < p->n_round_begin.size(); r++)
            if (p->n_round_begin[r + 1] > p->n_round_begin[r]) n++;
        return n;
    }
    if (!p->L.items1.empty()) n++;
    if (!p->fuse && !p->L.cores.empty()) n++;
    if (!p->fuse && p->nbig > 0) n++;
    for (size_t r = 0; r + 1 < p->L.round_begin.size(); r++)
        if (p->L.round_begin[r + 1] > p->L.round_begin[r]) n++;
    return n;
}

// transposed: (x, a, b) are the COLUMN points and box of the reference's tree, (y, c, d) its row points and
// box -- the leaves of KernelMatrix(f, y, x, c, d, a, b) with rows and columns exchanged, i.e. the block
// structure of the adjoint operator (whose rows are the points x)
static int32_t kernel_tree_layout(const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                                  double c, double d, int32_t part, int32_t nparts, HmLayout &L,
                                  bool transposed = false)
{
    return guarded([&]() -> int32_t {
        if (!x || !y) return fail(HM_ERR_NULL, "point set is NULL");
        if (nx < 0 || ny < 0) return fail(HM_ERR_SHAPE, "negative point count");
        std::vector<HmLeaf> leaves;
        int64_t nrows = 0, ncols = 0;
        std::string err = transposed ? hm_kernel_tree(y, ny, x, nx, c, d, a, b, leaves, ncols, nrows)
                                     : hm_kernel_tree(x, nx, y, ny, a, b, c, d, leaves, nrows, ncols);
        if (!err.empty()) return fail(HM_ERR_REFERENCE, "%s", err.c_str());
        if (transposed)
            for (HmLeaf &l : leaves) {
                std::swap(l.row0, l.col0);
                std::swap(l.m, l.n);
                std::swap(l.ru, l.rv);
                std::swap(l.xi0, l.yj0);
                std::swap(l.a, l.c);
                std::swap(l.b, l.d);
            }
        err = hm_build_layout(leaves, nrows, ncols, part, nparts, layout_params(), L);
        if (!err.empty()) return fail(HM_ERR_INVALID, "%s", err.c_str());
        return HM_OK;
    });
}

int32_t hm_kernel_tree_leaves(const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                              double c, double d, hm_tree_leaf *out, int64_t cap, int64_t *count)
{
    return guarded([&]() -> int32_t {
        if (!x || !y || !count) return fail(HM_ERR_NULL, "NULL argument");
        if (nx < 0 || ny < 0) return fail(HM_ERR_SHAPE, "negative point count");
        std::vector<HmLeaf> leaves;
        int64_t nrows = 0, ncols = 0;
        std::string err = hm_kernel_tree(x, nx, y, ny, a, b, c, d, leaves, nrows, ncols);
        if (!err.empty()) return fail(HM_ERR_REFERENCE, "%s", err.c_str());
        *count = (int64_t)leaves.size();
        for (int64_t i = 0; out && i < cap && i < *count; i++) {
            const HmLeaf &l = leaves[(size_t)i];
            out[i] = hm_tree_leaf{l.kind, l.ru, l.row0, l.col0, l.m, l.n, l.xi0, l.yj0, l.a, l.b, l.c, l.d};
        }
        return HM_OK;
    });
}

int32_t hm_assemble_kernel_stats(const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                                 double c, double d, int32_t part, int32_t nparts, hm_stats *out)
{
    return guarded([&]() -> int32_t {
        if (!out) return fail(HM_ERR_NULL, "out is NULL");
        HmLayout L;
        if (int32_t st = kernel_tree_layout(x, nx, y, ny, a, b, c, d, part, nparts, L)) return st;
        fill_stats(L, out);
        return HM_OK;
    });
}

// hm_assemble_kernel_fn: the parts of the operator that depend on f -- the r x r cores
// F[m,n] = f(x_m, y_n) at the mapped Chebyshev nodes (BarycentricMatrix.jl:159-175) and the dense
// leaves T[f(x[i], y[j])] (KernelMatrix.jl:57-60) -- are evaluated by the caller's f on the host, in
// large batches, and copied into the packed streams; U and V (88 % of the bytes) do not depend on f
// and were filled by the device kernels.
static int32_t host_fill(hm_plan *P, const double *x, const double *y, hm_kernel_fn fn, void *user)
{
    HmLayout &L = P->L;
    cudaStream_t st = P->stream;
    constexpr size_t BATCH = (size_t)1 << 21; // evaluations per callback
    std::vector<double> xs, ys;
    xs.reserve(BATCH + 4096);
    ys.reserve(BATCH + 4096);
    double *pin = nullptr; // results of one batch, in stream layout
    size_t pin_cap = 0;
    struct Piece {
        double *dst;
        size_t off, words;
    };
    std::vector<Piece> pieces;
    auto ensure = [&](size_t words) -> cudaError_t {
        if (words <= pin_cap) return cudaSuccess;
        if (pin) cudaFreeHost(pin);
        pin = nullptr;
        pin_cap = 0;
        cudaError_t e = cudaMallocHost((void **)&pin, words * 8);
        if (e == cudaSuccess) pin_cap = words;
        return e;
    };
    struct PinGuard {
        double *&p;
        ~PinGuard()
        {
            if (p) cudaFreeHost(p);
        }
    } pin_guard{pin};
    // ---- cores ----
    {
        size_t c = 0;
        const size_t nc = L.cores.size();
        while (c < nc) {
            xs.clear();
            ys.clear();
            pieces.clear();
            size_t words = 0;
            const size_t c_begin = c;
            for (; c < nc && xs.size() < BATCH; c++) {
                const HmCoreBlock &cb = L.cores[c];
                const HmLeaf &l = L.leaves[(size_t)L.core_leaf[c]];
                if (l.source != HM_SRC_KERNEL || cb.kind != HM_LEAF_BARY2D) continue;
                const double xm = 0.5 * (l.a + l.b), xh = 0.5 * (l.b - l.a);
                const double ym = 0.5 * (l.c + l.d), yh = 0.5 * (l.d - l.c);
                for (int n = 0; n < cb.rv; n++)
                    for (int m = 0; m < cb.ru; m++) {
                        xs.push_back(xm + xh * P->cheb.node[m]);
                        ys.push_back(ym + yh * P->cheb.node[n]);
                    }
                pieces.push_back(Piece{P->core.p + cb.core, words, (size_t)cb.ru * cb.rv});
                words += (size_t)cb.ru * cb.rv;
            }
            if (words == 0) continue;
            (void)c_begin;
            HM_CUDA(ensure(words));
            fn(xs.data(), ys.data(), (int64_t)words, pin, user);
            for (const Piece &pc : pieces)
                HM_CUDA(cudaMemcpyAsync(pc.dst, pin + pc.off, pc.words * 8, cudaMemcpyHostToDevice, st));
            HM_CUDA(cudaStreamSynchronize(st));
        }
    }
    // ---- dense tiles (stage-3 fill entries of dense leaves: kn slab columns of F rows, pitch Fp) ----
    {
        size_t i = 0;
        const size_t nf = L.fill3.size();
        std::vector<double> vals;
        while (i < nf) {
            xs.clear();
            ys.clear();
            pieces.clear();
            size_t words = 0;
            const size_t i_begin = i;
            for (; i < nf && xs.size() < BATCH; i++) {
                const HmFill &f = L.fill3[i];
                const HmLeaf &l = L.leaves[(size_t)f.leaf];
                if (l.source != HM_SRC_KERNEL || l.kind != HM_LEAF_DENSE) continue;
                const double *xr = x + l.xi0 + f.off, *yc = y + l.yj0 + f.k0;
                for (int k = 0; k < f.kn; k++)
                    for (int r = 0; r < f.F; r++) {
                        xs.push_back(xr[r]);
                        ys.push_back(yc[k]);
                    }
                pieces.push_back(Piece{P->ustream.p + f.dst, words, (size_t)f.kn * f.Fp});
                words += (size_t)f.kn * f.Fp;
            }
            if (xs.empty()) continue;
            vals.resize(xs.size());
            fn(xs.data(), ys.data(), (int64_t)xs.size(), vals.data(), user);
            HM_CUDA(ensure(words));
            size_t src = 0, pi = 0;
            for (size_t j = i_begin; j < i; j++) {
                const HmFill &f = L.fill3[j];
                const HmLeaf &l = L.leaves[(size_t)f.leaf];
                if (l.source != HM_SRC_KERNEL || l.kind != HM_LEAF_DENSE) continue;
                double *dstp = pin + pieces[pi++].off;
                for (int k = 0; k < f.kn; k++) {
                    for (int r = 0; r < f.F; r++) dstp[(size_t)k * f.Fp + r] = vals[src++];
                    for (int r = f.F; r < f.Fp; r++) dstp[(size_t)k * f.Fp + r] = 0.0;
                }
            }
            for (const Piece &pc : pieces)
                HM_CUDA(cudaMemcpyAsync(pc.dst, pin + pc.off, pc.words * 8, cudaMemcpyHostToDevice, st));
            HM_CUDA(cudaStreamSynchronize(st));
        }
    }
    return HM_OK;
}

// Nested-basis form of a matrix-free plan (hm_nest.h): the default when the operator has the dyadic
// structure of KernelMatrix(f, x, y, a, b, c, d) with descending points; HMB200_FREE_FORM=cheb keeps the
// per-leaf Chebyshev form.  A structure the builder declines is not an error: the plan stays as it is.
static int32_t build_nested(hm_plan *P, const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                            double c, double d)
{
    const char *e = getenv("HMB200_FREE_FORM");
    if (e && e[0] == 'c') return HM_OK;
    const HmLayout &L = P->L;
    std::vector<HmFreeRun> run3(L.fill3.size());
    for (size_t i = 0; i < L.fill3.size(); i++) {
        const HmFill &f = L.fill3[i];
        const HmLeaf &l = L.leaves[(size_t)f.leaf];
        run3[i] = HmFreeRun{0.5 * (l.a + l.b), 0.5 * (l.b - l.a), l.xi0 + f.off, l.yj0 + f.k0, f.k0, f.kn};
    }
    HmNest N;
    const std::string why = hm_nest_build(L, x, nx, y, ny, a, b, c, d, P->kernel_id, run3, N);
    if (!why.empty()) return HM_OK;
    if ((int)N.rows.tier_sub0.size() - 1 > HM_NEST_MAXTIERS || (int)N.cols.tier_sub0.size() - 1 > HM_NEST_MAXTIERS) return HM_OK;
    for (const HmItem &it : N.items3)
        if (it.F > 128) return HM_OK; // hm_nest_dense_kernel: four rows per lane
    cudaStream_t st = P->stream;
    HM_CUDA(P->nr_nodes.upload(N.rows.nodes, st));
    HM_CUDA(P->nr_order.upload(N.rows.order, st));
    HM_CUDA(P->nr_grp.upload(N.rows.grp, st));
    HM_CUDA(P->nr_sub.upload(N.rows.sub_g0, st));
    HM_CUDA(P->nc_nodes.upload(N.cols.nodes, st));
    HM_CUDA(P->nc_order.upload(N.cols.order, st));
    HM_CUDA(P->nc_grp.upload(N.cols.grp, st));
    HM_CUDA(P->nc_sub.upload(N.cols.sub_g0, st));
    HM_CUDA(P->nr_base.upload(N.rows.base, st));
    HM_CUDA(P->nc_base.upload(N.cols.base, st));
    HM_CUDA(P->n_item_box.upload(N.item_box, st));
    HM_CUDA(P->n_fin.upload(N.fin, st));
    HM_CUDA(P->n_items3p.upload(N.items3p, st));
    HM_CUDA(P->n_runsp.upload(N.runsp, st));
    HM_CUDA(P->n_frunp.upload(N.frunp, st));
    P->n_fin_rows = (int64_t)N.rows.base.size() * HM_NEST_R;
    P->n_fused_eval = N.fused_eval;
    {
        const char *eo = getenv("HMB200_NEST_OVERLAP");
        P->n_overlap = !(eo && eo[0] == '0');
    }
    HM_CUDA(P->n_rleaf_begin.upload(N.rleaf_begin, st));
    HM_CUDA(P->n_rleaf.upload(N.rleaf, st));
    HM_CUDA(P->n_cores.upload(N.cores, st));
    HM_CUDA(P->n_M.upload(N.M, st));
    HM_CUDA(P->n_MU.alloc(N.cols.nodes.size() * HM_NEST_R));
    HM_CUDA(P->n_LAM.alloc(N.rows.nodes.size() * HM_NEST_R));
    HM_CUDA(P->n_items3.upload(N.items3, st));
    HM_CUDA(P->n_runs.upload(N.runs, st));
    HM_CUDA(P->n_frun.upload(N.frun, st));
    HM_CUDA(cudaStreamSynchronize(st));
    auto dev = [](const HmNestTree &T, const DevBuf<HmNestNode> &nodes, const DevBuf<int32_t> &order,
                  const DevBuf<int32_t> &grp, const DevBuf<int32_t> &sub, const DevBuf<int32_t> &base) {
        HmNestDev D;
        D.base = base.p;
        D.nbase = (int)T.base.size();
        D.nodes = nodes.p;
        D.order = order.p;
        D.grp = grp.p;
        D.sub_g0 = sub.p;
        D.ntiers = (int)T.tier_sub0.size() - 1;
        for (int k = 0; k <= D.ntiers; k++) D.tier_sub0[k] = T.tier_sub0[(size_t)k];
        D.nnodes = (int)T.nodes.size();
        return D;
    };
    P->n_rows = dev(N.rows, P->nr_nodes, P->nr_order, P->nr_grp, P->nr_sub, P->nr_base);
    P->n_cols = dev(N.cols, P->nc_nodes, P->nc_order, P->nc_grp, P->nc_sub, P->nc_base);
    P->n_round_begin = N.round_begin;
    P->n_zcap = (N.zcap + 1) & ~1;
    P->n_distinct_cores = (int64_t)(N.cores.size() / (HM_NEST_R * HM_NEST_R));
    P->nested = true;
    return HM_OK;
}

static int32_t assemble_kernel_impl(const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                                    double c, double d, int32_t kernel_id, int32_t device, int32_t part,
                                    int32_t nparts, bool matrix_free, hm_plan **out, hm_kernel_fn fn = nullptr,
                                    void *user = nullptr, bool transposed = false)
{
    return guarded([&]() -> int32_t {
        if (!out) return fail(HM_ERR_NULL, "out is NULL");
        *out = nullptr;
        if (fn) kernel_id = HM_KERNEL_HOST_FN;
        else if (kernel_id < 0 || kernel_id > 3) return fail(HM_ERR_INVALID, "unknown kernel id %d", kernel_id);
        if (hm_blockrank_double() != 20) return fail(HM_ERR_UNSUPPORTED, "BLOCKRANK(Float64) != 20");
        HM_DEVICE(device);
        hm_plan *P = new (std::nothrow) hm_plan;
        if (!P) return fail(HM_ERR_NOMEM, "out of host memory");
        P->device = device;
        P->kernel_id = kernel_id;
        P->matrix_free = matrix_free;
        P->box[0] = a;
        P->box[1] = b;
        P->box[2] = c;
        P->box[3] = d;
        if (int32_t st = kernel_tree_layout(x, nx, y, ny, a, b, c, d, part, nparts, P->L, transposed)) {
            delete P;
            return st;
        }
        P->cheb.r = hm_blockrank_double();
        hm_cheb_nodes_weights(P->cheb.r, P->cheb.node, P->cheb.lam);
        cudaError_t e = P->f_px.alloc((size_t)std::max<int64_t>(nx, 1));
        if (e == cudaSuccess) e = P->f_py.alloc((size_t)std::max<int64_t>(ny, 1));
        if (e == cudaSuccess && nx) e = cudaMemcpy(P->f_px.p, x, (size_t)nx * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && ny) e = cudaMemcpy(P->f_py.p, y, (size_t)ny * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            delete P;
            cudaGetLastError();
            return fail(HM_ERR_CUDA, "point upload failed: %s", cudaGetErrorString(e));
        }
        hm_trace_point("assemble: tree + layout + point upload");
        int32_t st = materialize(P, P->f_px.p, P->f_py.p);
        hm_trace_point("assemble: materialize");
        if (st != HM_OK) {
            delete P;
            return st;
        }
        if (matrix_free && P->free_cheb) {
            st = build_nested(P, x, nx, y, ny, a, b, c, d);
            hm_trace_point("assemble: nested form");
            if (st != HM_OK) {
                delete P;
                return st;
            }
        }
        if (fn) {
            st = host_fill(P, x, y, fn, user);
            if (st != HM_OK) {
                delete P;
                return st;
            }
        }
        if (!matrix_free) { // the stored operator no longer needs the points
            P->f_px.release();
            P->f_py.release();
        }
        *out = P;
        return HM_OK;
    });
}

int32_t hm_assemble_kernel(const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                           double c, double d, int32_t kernel_id, int32_t device, int32_t part, int32_t nparts,
                           hm_plan **out)
{
    return guarded([&]() -> int32_t {
        return assemble_kernel_impl(x, nx, y, ny, a, b, c, d, kernel_id, device, part, nparts, false, out);
    });
}

int32_t hm_assemble_kernel_fn(const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                              double c, double d, hm_kernel_fn f, void *user, int32_t device, int32_t part,
                              int32_t nparts, hm_plan **out)
{
    return guarded([&]() -> int32_t {
        if (!f) return fail(HM_ERR_NULL, "kernel function is NULL");
        return assemble_kernel_impl(x, nx, y, ny, a, b, c, d, HM_KERNEL_HOST_FN, device, part, nparts, false, out, f,
                                    user);
    });
}

int32_t hm_plan_form(const hm_plan *p, int32_t *form)
{
    return guarded([&]() -> int32_t {
        if (!p || !form) return fail(HM_ERR_NULL, "NULL argument");
        *form = !p->matrix_free ? 0 : p->nested ? 3 : p->free_cheb ? 2 : 1;
        return HM_OK;
    });
}

int32_t hm_assemble_kernel_free(const double *x, int64_t nx, const double *y, int64_t ny, double a, double b,
                                double c, double d, int32_t kernel_id, int32_t device, int32_t part,
                                int32_t nparts, hm_plan **out)
{
    return guarded([&]() -> int32_t {
        return assemble_kernel_impl(x, nx, y, ny, a, b, c, d, kernel_id, device, part, nparts, true, out);
    });
}

// ---------------------------------------------------------------------------
// mul!
// ---------------------------------------------------------------------------

int32_t hm_matvec_device(hm_plan *p, const double *dx, double *dy, int32_t accumulate, void *stream)
{
    return guarded([&]() -> int32_t {
        return hm_matvec_device_peers(p, dx, dy, accumulate, stream, nullptr);
    });
}

int32_t hm_matvec_device_allgather(hm_plan *p, const double *dx, const uint64_t *ypeers, int32_t npeers,
                                   int32_t self, int32_t accumulate, void *stream)
{
    return guarded([&]() -> int32_t {
        if (!p || !ypeers) return fail(HM_ERR_NULL, "NULL argument");
        if (npeers < 1 || npeers > HM_MAX_PEERS) return fail(HM_ERR_INVALID, "npeers must be 1..%d", HM_MAX_PEERS);
        if (self < 0 || self >= npeers) return fail(HM_ERR_INVALID, "self out of range");
        HmPeers pe;
        pe.n = npeers;
        for (int i = 0; i < npeers; i++) {
            if (!ypeers[i] && p->L.nrows > 0) return fail(HM_ERR_NULL, "peer pointer %d is NULL", i);
            pe.y[i] = reinterpret_cast<double *>(ypeers[i]);
        }
        return hm_matvec_device_peers(p, dx, pe.y[self], accumulate, stream, &pe);
    });
}

int32_t hm_matvec_device_peers(hm_plan *p, const double *dx, double *dy, int32_t accumulate, void *stream,
                                  const HmPeers *peers)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        if ((!dx && p->L.ncols > 0) || (!dy && p->L.nrows > 0)) return fail(HM_ERR_NULL, "vector pointer is NULL");
        HM_DEVICE(p->device);
        cudaStream_t st = (cudaStream_t)stream;
        const HmLayout &L = p->L;
        cudaEvent_t *ev = p->tcount < p->tcap ? &p->tev[(size_t)p->tcount * 4] : nullptr;
        if (ev) HM_CUDA(cudaEventRecord(ev[0], st));
        HmFuse fz;
        if (p->fuse) {
            fz.s1ent = p->s1ent.p;
            fz.counters = p->counters.p;
            fz.blocks = p->cores.p;
            fz.plist = p->plist.p;
            fz.core = p->core.p;
            fz.svec = p->svec.p;
            fz.max_r = std::max(L.max_r, 1);
        }
        if (p->nested) {
            // nested-basis form: moments up the column boxes, cores, coefficients down the row boxes and
            // their evaluation (writes every owned row), then the dense leaves on top
            const double *M = p->n_M.p, *Mt = p->n_M.p + 2 * HM_NEST_R * HM_NEST_R;
            // The dense leaves (FP64-bound, 0.12 ms at N = 2^20) do not depend on the tree passes (short,
            // latency-bound launches that leave most SMs idle): the passes run beside them on a second,
            // high-priority stream (fork / join by events, so the call is still one unit of work on the
            // caller's stream and can be captured into a graph), and the finest tier of the downward pass
            // adds the series to the rows the dense kernel wrote.  HMB200_NEST_OVERLAP=0: one stream,
            // evaluation fused into the dense pass.
            if (p->n_fused_eval && !peers && p->n_overlap && p->n_round_begin.size() == 2) {
                if (!p->side_stream) {
                    // the tree passes get the highest priority: their few CTAs take the slots the dense kernel's
                    // retiring CTAs free, ahead of its pending ones
                    int lo = 0, hi = 0;
                    HM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                    HM_CUDA(cudaStreamCreateWithPriority(&p->side_stream, cudaStreamNonBlocking, hi));
                    HM_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
                    HM_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
                    HM_CUDA(cudaEventCreateWithFlags(&p->ev_dense, cudaEventDisableTiming));
                }
                cudaStream_t hs = p->side_stream;
                HM_CUDA(cudaEventRecord(p->ev_fork, st));
                HM_CUDA(cudaStreamWaitEvent(hs, p->ev_fork, 0));
                HM_CUDA(hm_launch_nest_up(p->n_cols, p->f_py.p, dx, Mt, p->n_MU.p, hs));
                HM_CUDA(hm_launch_nest_dense(p->n_items3.p, p->n_round_begin[1], p->n_runs.p, p->n_frun.p, p->f_px.p,
                                             p->f_py.p, dx, dy, accumulate != 0, p->kernel_id, nullptr, p->nr_nodes.p,
                                             p->n_LAM.p, nullptr, st));
                HM_CUDA(cudaEventRecord(p->ev_dense, st));
                if (ev) HM_CUDA(cudaEventRecord(ev[1], st));
                HM_CUDA(hm_launch_nest_core(p->n_rows.nnodes, p->n_rleaf_begin.p, p->n_rleaf.p, p->n_cores.p, p->n_MU.p,
                                            p->n_LAM.p, hs));
                if (ev) HM_CUDA(cudaEventRecord(ev[2], st));
                // translations of every tier without waiting for the dense kernel; only the evaluation adds to
                // the rows it wrote
                HM_CUDA(hm_launch_nest_down(p->n_rows, p->f_px.p, M, p->n_LAM.p, dy, 1, L.row_begin, L.row_end, false, hs));
                HM_CUDA(cudaStreamWaitEvent(hs, p->ev_dense, 0));
                HM_CUDA(hm_launch_nest_eval(p->n_rows, p->f_px.p, p->n_LAM.p, dy, 1, L.row_begin, L.row_end, hs));
                HM_CUDA(cudaEventRecord(p->ev_join, hs));
                HM_CUDA(cudaStreamWaitEvent(st, p->ev_join, 0));
                if (ev) {
                    HM_CUDA(cudaEventRecord(ev[3], st));
                    p->tcount++;
                }
                return HM_OK;
            }
            HM_CUDA(hm_launch_nest_up(p->n_cols, p->f_py.p, dx, Mt, p->n_MU.p, st));
            if (ev) HM_CUDA(cudaEventRecord(ev[1], st));
            HM_CUDA(hm_launch_nest_core(p->n_rows.nnodes, p->n_rleaf_begin.p, p->n_rleaf.p, p->n_cores.p, p->n_MU.p,
                                        p->n_LAM.p, st));
            if (ev) HM_CUDA(cudaEventRecord(ev[2], st));
            const bool fused = p->n_fused_eval; // the dense pass evaluates the low-rank part of its rows too
            HM_CUDA(hm_launch_nest_down(p->n_rows, p->f_px.p, M, p->n_LAM.p, dy, accumulate != 0, L.row_begin, L.row_end,
                                        !fused, st));
            for (size_t r = 0; r + 1 < p->n_round_begin.size(); r++) {
                const int64_t i0 = p->n_round_begin[r], i1 = p->n_round_begin[r + 1];
                HM_CUDA(hm_launch_nest_dense(p->n_items3.p + i0, i1 - i0, p->n_runs.p, p->n_frun.p, p->f_px.p, p->f_py.p,
                                             dx, dy, fused ? (accumulate != 0) : 1, p->kernel_id,
                                             fused ? p->n_item_box.p : nullptr, p->nr_nodes.p, p->n_LAM.p, peers, st));
            }
            if (ev) {
                HM_CUDA(cudaEventRecord(ev[3], st));
                p->tcount++;
            }
            return HM_OK;
        }
        if (p->matrix_free)
            HM_CUDA(hm_launch_free1(p->items1.p, (int64_t)L.items1.size(), p->f_ent1.p, p->f_py.p, dx, p->partial.p,
                                    p->cheb, p->free1_units, p->free_cheb, st));
        else
            HM_CUDA(hm_launch_stage1(p->items1.p, (int64_t)L.items1.size(), p->vstream.p, dx, p->partial.p,
                                     p->fuse ? &fz : nullptr, st));
        if (ev) HM_CUDA(cudaEventRecord(ev[1], st));
        if (!p->fuse || p->matrix_free) {
            HM_CUDA(hm_launch_stage2(p->cores.p, (int64_t)L.cores.size(), p->plist.p, p->partial.p, p->core.p,
                                     p->svec.p, std::max(L.max_r, 1), st));
            HM_CUDA(hm_launch_stage2_big(p->cores.p, p->bigcores.p, p->nbig, p->plist.p, p->partial.p, p->core.p,
                                         p->svec.p, std::max(L.max_r, 1), st));
        }
        if (ev) HM_CUDA(cudaEventRecord(ev[2], st));
        for (size_t r = 0; r + 1 < L.round_begin.size(); r++) {
            int64_t i0 = L.round_begin[r], i1 = L.round_begin[r + 1];
            if (p->matrix_free)
                HM_CUDA(hm_launch_free3(p->items3.p + i0, i1 - i0, p->runs.p, p->f_run3.p, p->f_px.p, p->f_py.p, dx,
                                        p->svec.p, dy, r == 0 ? (accumulate != 0) : 1, p->cheb, p->kernel_id, peers,
                                        p->free3_zcap, p->free_cheb, st));
            else
                HM_CUDA(hm_launch_stage3(p->items3.p + i0, i1 - i0, p->runs.p, p->ustream.p, dx, p->svec.p, dy,
                                         r == 0 ? (accumulate != 0) : 1, peers, st));
        }
        if (ev) {
            HM_CUDA(cudaEventRecord(ev[3], st));
            p->tcount++;
        }
        return HM_OK;
    });
}

int32_t hm_matvec(hm_plan *p, const double *x, int64_t incx, double *y, int64_t incy, int32_t accumulate)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        const HmLayout &L = p->L;
        if ((!x && L.ncols > 0) || (!y && L.nrows > 0)) return fail(HM_ERR_NULL, "vector pointer is NULL");
        if (incx <= 0 || incy <= 0) return fail(HM_ERR_INVALID, "strides must be positive (INCX, INCY >= 1 as in the reference)");
        std::lock_guard<std::mutex> lock(p->mu);
        HM_DEVICE(p->device);
        const int64_t nc = L.ncols, r0 = L.row_begin, nr = L.row_end - L.row_begin;
        if (!p->dx.p) HM_CUDA(p->dx.alloc((size_t)std::max<int64_t>(nc, 1)));
        if (!p->dy.p) HM_CUDA(p->dy.alloc((size_t)std::max<int64_t>(L.nrows, 1)));
        cudaStream_t st = p->stream;
        // Unit strides, one stage-3 round: pipeline the host copies against the kernels.  x goes up
        // in HM_NCHUNK chunks on a copy stream and stage 1 starts on the items whose columns have
        // arrived; y comes down in row chunks while stage 3 still computes the following ones.
        // Below ~2 MB of vector data the extra launches and events cost more than the overlap gains
        // (measured: N = 4096 48 vs 100 us, N = 65 536 175 vs 230 us, N = 262 144 600 vs 577 us).
        if (incx == 1 && incy == 1 && nc > 0 && nr > 0 && nc + nr >= 400000 && L.round_begin.size() == 2 &&
            !L.items3c.empty() && p->tcap == 0 && !p->nested && !getenv("HMB200_NO_COPY_PIPELINE")) {
            if (!p->chunk_ready) {
                HM_CUDA(p->items1c.upload(L.items1c, st));
                HM_CUDA(p->items3c.upload(L.items3c, st));
                HM_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
                for (int k = 0; k < HM_NCHUNK; k++) {
                    HM_CUDA(cudaEventCreateWithFlags(&p->ev_x[k], cudaEventDisableTiming));
                    HM_CUDA(cudaEventCreateWithFlags(&p->ev_y[k], cudaEventDisableTiming));
                }
                HM_CUDA(cudaEventCreateWithFlags(&p->ev_y0, cudaEventDisableTiming));
                HM_CUDA(cudaStreamSynchronize(st));
                p->chunk_ready = true;
            }
            cudaStream_t cst = p->copy_stream;
            for (int k = 0; k < HM_NCHUNK; k++) {
                const int64_t c0 = L.xchunk[(size_t)k], c1 = L.xchunk[(size_t)k + 1];
                if (c1 > c0) HM_CUDA(cudaMemcpyAsync(p->dx.p + c0, x + c0, (size_t)(c1 - c0) * 8, cudaMemcpyHostToDevice, cst));
                HM_CUDA(cudaEventRecord(p->ev_x[k], cst));
            }
            if (accumulate) HM_CUDA(cudaMemcpyAsync(p->dy.p + r0, y + r0, (size_t)nr * 8, cudaMemcpyHostToDevice, cst));
            HM_CUDA(cudaEventRecord(p->ev_y0, cst));
            for (int k = 0; k < HM_NCHUNK; k++) {
                HM_CUDA(cudaStreamWaitEvent(st, p->ev_x[k], 0));
                const int64_t i0 = L.c1_begin[(size_t)k], i1 = L.c1_begin[(size_t)k + 1];
                if (p->matrix_free)
                    HM_CUDA(hm_launch_free1(p->items1c.p + i0, i1 - i0, p->f_ent1.p, p->f_py.p, p->dx.p, p->partial.p,
                                            p->cheb, p->free1_units, p->free_cheb, st));
                else
                    HM_CUDA(hm_launch_stage1(p->items1c.p + i0, i1 - i0, p->vstream.p, p->dx.p, p->partial.p, nullptr, st,
                                             false));
            }
            HM_CUDA(hm_launch_stage2(p->cores.p, (int64_t)L.cores.size(), p->plist.p, p->partial.p, p->core.p,
                                     p->svec.p, std::max(L.max_r, 1), st, false));
            HM_CUDA(hm_launch_stage2_big(p->cores.p, p->bigcores.p, p->nbig, p->plist.p, p->partial.p, p->core.p,
                                         p->svec.p, std::max(L.max_r, 1), st, false));
            HM_CUDA(cudaStreamWaitEvent(st, p->ev_y0, 0));
            for (int k = 0; k < HM_NCHUNK; k++) {
                const int64_t i0 = L.c3_begin[(size_t)k], i1 = L.c3_begin[(size_t)k + 1];
                if (p->matrix_free)
                    HM_CUDA(hm_launch_free3(p->items3c.p + i0, i1 - i0, p->runs.p, p->f_run3.p, p->f_px.p, p->f_py.p,
                                            p->dx.p, p->svec.p, p->dy.p, accumulate != 0, p->cheb, p->kernel_id,
                                            nullptr, p->free3_zcap, p->free_cheb, st));
                else
                    HM_CUDA(hm_launch_stage3(p->items3c.p + i0, i1 - i0, p->runs.p, p->ustream.p, p->dx.p, p->svec.p,
                                             p->dy.p, accumulate != 0, nullptr, st, false));
                HM_CUDA(cudaEventRecord(p->ev_y[k], st));
            }
            // all launches are queued before the first copy back: with pageable y the copies block
            // the host, and must not hold up the launch of the following chunks
            for (int k = 0; k < HM_NCHUNK; k++) {
                HM_CUDA(cudaStreamWaitEvent(cst, p->ev_y[k], 0));
                const int64_t a0 = L.ychunk[(size_t)k], a1 = L.ychunk[(size_t)k + 1];
                if (a1 > a0) HM_CUDA(cudaMemcpyAsync(y + a0, p->dy.p + a0, (size_t)(a1 - a0) * 8, cudaMemcpyDeviceToHost, cst));
            }
            HM_CUDA(cudaStreamSynchronize(cst));
            HM_CUDA(cudaStreamSynchronize(st));
            return HM_OK;
        }
        // x -> device
        if (nc > 0) {
            if (incx == 1) {
                HM_CUDA(cudaMemcpyAsync(p->dx.p, x, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
            } else {
                if (!p->hx) HM_CUDA(cudaMallocHost((void **)&p->hx, (size_t)nc * 8));
                for (int64_t j = 0; j < nc; j++) p->hx[j] = x[j * incx];
                HM_CUDA(cudaMemcpyAsync(p->dx.p, p->hx, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
            }
        }
        // y (owned rows) -> device when accumulating
        if (nr > 0 && accumulate) {
            if (incy == 1) {
                HM_CUDA(cudaMemcpyAsync(p->dy.p + r0, y + r0, (size_t)nr * 8, cudaMemcpyHostToDevice, st));
            } else {
                if (!p->hy) HM_CUDA(cudaMallocHost((void **)&p->hy, (size_t)L.nrows * 8));
                for (int64_t i = 0; i < nr; i++) p->hy[r0 + i] = y[(r0 + i) * incy];
                HM_CUDA(cudaMemcpyAsync(p->dy.p + r0, p->hy + r0, (size_t)nr * 8, cudaMemcpyHostToDevice, st));
            }
        }
        if (int32_t rc = hm_matvec_device(p, p->dx.p, p->dy.p, accumulate, st)) return rc;
        if (nr > 0) {
            if (incy == 1) {
                HM_CUDA(cudaMemcpyAsync(y + r0, p->dy.p + r0, (size_t)nr * 8, cudaMemcpyDeviceToHost, st));
                HM_CUDA(cudaStreamSynchronize(st));
            } else {
                if (!p->hy) HM_CUDA(cudaMallocHost((void **)&p->hy, (size_t)L.nrows * 8));
                HM_CUDA(cudaMemcpyAsync(p->hy + r0, p->dy.p + r0, (size_t)nr * 8, cudaMemcpyDeviceToHost, st));
                HM_CUDA(cudaStreamSynchronize(st));
                for (int64_t i = 0; i < nr; i++) y[(r0 + i) * incy] = p->hy[r0 + i];
            }
        } else {
            HM_CUDA(cudaStreamSynchronize(st));
        }
        return HM_OK;
    });
}

// Adjoint of a matrix-free plan.  K' is again a kernel operator: its rows are the points y, its columns the
// points x, its leaves the transposed leaves (same boxes, roles exchanged) and its entries f(x_i, y_j) =
// phi(x_i - y_j) = s phi(y_j - x_i) with s = -1 for the odd kernels (cauchy, coulomb') and +1 for the even ones
// (coulomb, log): K' w = s K~ w = K~ (s w) with K~ the matrix-free plan of (f, y, x) on the transposed
// leaves -- in nested-basis form whenever the forward plan is.  Built on first use, whole operators only.
static int32_t adjoint_matrix_free(hm_plan *p, const double *dx, double *dy, int32_t accumulate, cudaStream_t st)
{
    const HmLayout &L = p->L;
    if (L.nparts != 1) return fail(HM_ERR_UNSUPPORTED, "adjoint of a matrix-free plan: whole operators only (nparts = 1)");
    if (p->kernel_id < 0 || p->kernel_id > 3) return fail(HM_ERR_UNSUPPORTED, "adjoint of a matrix-free plan: built-in kernels only");
    if (!p->adj) {
        HM_CUDA(cudaStreamSynchronize(st));
        std::vector<double> hx((size_t)std::max<int64_t>(L.nrows, 1)), hy((size_t)std::max<int64_t>(L.ncols, 1));
        if (L.nrows) HM_CUDA(cudaMemcpy(hx.data(), p->f_px.p, (size_t)L.nrows * 8, cudaMemcpyDeviceToHost));
        if (L.ncols) HM_CUDA(cudaMemcpy(hy.data(), p->f_py.p, (size_t)L.ncols * 8, cudaMemcpyDeviceToHost));
        hm_plan *q = nullptr;
        if (int32_t rc = assemble_kernel_impl(hy.data(), L.ncols, hx.data(), L.nrows, p->box[2], p->box[3], p->box[0],
                                              p->box[1], p->kernel_id, p->device, 0, 1, true, &q, nullptr, nullptr, true))
            return rc;
        p->adj = q;
    }
    const bool odd = p->kernel_id == 0 || p->kernel_id == 2;
    const double *xin = dx;
    if (odd && L.nrows > 0) {
        if (!p->adj_x.p) {
            HM_CUDA(cudaStreamSynchronize(st));
            HM_CUDA(p->adj_x.alloc((size_t)L.nrows));
        }
        HM_CUDA(hm_launch_negate(dx, p->adj_x.p, L.nrows, st));
        xin = p->adj_x.p;
    }
    return hm_matvec_device(p->adj, xin, dy, accumulate, (void *)st);
}

// Adjoint apply y (+)= H' x.
int32_t hm_matvec_adjoint_device(hm_plan *p, const double *dx, double *dy, int32_t accumulate, void *stream)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        const HmLayout &L = p->L;
        if ((!dx && L.nrows > 0) || (!dy && L.ncols > 0)) return fail(HM_ERR_NULL, "vector pointer is NULL");
        if (p->matrix_free) {
            HM_DEVICE(p->device);
            return adjoint_matrix_free(p, dx, dy, accumulate, (cudaStream_t)stream);
        }
        if (L.adj_max_f > HM_SMAX) return fail(HM_ERR_UNSUPPORTED, "adjoint: a column segment is covered by too many ranks");
        HM_DEVICE(p->device);
        cudaStream_t st = (cudaStream_t)stream;
        if (!p->adj_ready) {
            HM_CUDA(cudaStreamSynchronize(st));
            HM_CUDA(p->pq.alloc((size_t)std::max<int64_t>(L.pq_words, 1)));
            HM_CUDA(p->qlist.upload(L.qlist, st));
            HM_CUDA(p->core_q0.upload(L.core_q0, st));
            HM_CUDA(p->core_qn.upload(L.core_qn, st));
            HM_CUDA(p->colsegs.upload(L.colsegs, st));
            HM_CUDA(p->colbases.upload(L.colbases, st));
            {
                std::vector<int32_t> big;
                for (size_t c = 0; c < L.cores.size(); c++)
                    if (L.core_qn[c] > HM_ADJ_BIG) big.push_back((int32_t)c);
                p->nadjbig = (int)big.size();
                HM_CUDA(p->adjbig.upload(big, st));
            }
            if (!p->s1ent.p) HM_CUDA(p->s1ent.upload(L.s1ent, st));
            HM_CUDA(cudaStreamSynchronize(st));
            p->adj_ready = true;
        }
        HmAdjoint A;
        A.items3 = p->items3.p;
        A.items1 = p->items1.p;
        A.n3 = (int64_t)L.items3.size();
        A.n1 = (int64_t)L.items1.size();
        A.ncores = (int64_t)L.cores.size();
        A.nsegs = (int64_t)L.colsegs.size();
        A.ustream = p->ustream.p;
        A.vstream = p->vstream.p;
        A.core = p->core.p;
        A.blocks = p->cores.p;
        A.q0 = p->core_q0.p;
        A.qn = p->core_qn.p;
        A.qlist = p->qlist.p;
        A.s1ent = p->s1ent.p;
        A.big = p->adjbig.p;
        A.nbig = p->nadjbig;
        A.segs = p->colsegs.p;
        A.bases = p->colbases.p;
        A.PQ = p->pq.p;
        A.svec = p->svec.p;
        A.max_r = std::max(L.max_r, 1);
        HM_CUDA(hm_launch_adjoint(A, dx, dy, accumulate != 0, st));
        return HM_OK;
    });
}

int32_t hm_matvec_adjoint(hm_plan *p, const double *x, int64_t incx, double *y, int64_t incy, int32_t accumulate)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        const HmLayout &L = p->L;
        if ((!x && L.nrows > 0) || (!y && L.ncols > 0)) return fail(HM_ERR_NULL, "vector pointer is NULL");
        if (incx <= 0 || incy <= 0) return fail(HM_ERR_INVALID, "strides must be positive (INCX, INCY >= 1 as in the reference)");
        std::lock_guard<std::mutex> lock(p->mu);
        HM_DEVICE(p->device);
        const int64_t nr = L.nrows, nc = L.ncols;
        // the forward path's staging buffers are reused with the roles of x and y swapped
        if (!p->dx.p) HM_CUDA(p->dx.alloc((size_t)std::max<int64_t>(nc, 1)));
        if (!p->dy.p) HM_CUDA(p->dy.alloc((size_t)std::max<int64_t>(nr, 1)));
        cudaStream_t st = p->stream;
        // unit strides: straight from / to the caller's vectors; strided arguments are packed
        // through the plan's pinned staging buffers (hy holds nrows words, hx ncols), as in hm_matvec
        if (nr > 0) {
            if (incx == 1) {
                HM_CUDA(cudaMemcpyAsync(p->dy.p, x, (size_t)nr * 8, cudaMemcpyHostToDevice, st));
            } else {
                if (!p->hy) HM_CUDA(cudaMallocHost((void **)&p->hy, (size_t)nr * 8));
                for (int64_t i = 0; i < nr; i++) p->hy[i] = x[i * incx];
                HM_CUDA(cudaMemcpyAsync(p->dy.p, p->hy, (size_t)nr * 8, cudaMemcpyHostToDevice, st));
            }
        }
        if (incy != 1 && nc > 0 && !p->hx) HM_CUDA(cudaMallocHost((void **)&p->hx, (size_t)nc * 8));
        if (accumulate && nc > 0) {
            if (incy == 1) {
                HM_CUDA(cudaMemcpyAsync(p->dx.p, y, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
            } else {
                for (int64_t j = 0; j < nc; j++) p->hx[j] = y[j * incy];
                HM_CUDA(cudaMemcpyAsync(p->dx.p, p->hx, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
            }
        }
        if (int32_t rc = hm_matvec_adjoint_device(p, p->dy.p, p->dx.p, accumulate, st)) return rc;
        if (nc > 0)
            HM_CUDA(cudaMemcpyAsync(incy == 1 ? y : p->hx, p->dx.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        if (incy != 1)
            for (int64_t j = 0; j < nc; j++) y[j * incy] = p->hx[j];
        return HM_OK;
    });
}

// Multi-RHS on the nested-basis form (hm_nest_panel.cu): moments up, cores, coefficients down, then the dense
// leaves together with the evaluation of every row's coefficients (hm_free3_panel_kernel).
static int32_t matmat_nested(hm_plan *p, const double *dX, int64_t ldx, double *dY, int64_t ldy, int64_t nrhs,
                             int32_t accumulate, cudaStream_t st)
{
    const HmLayout &L = p->L;
    for (int64_t c0 = 0; c0 < nrhs; c0 += 64) {
        const int nc = (int)std::min<int64_t>(64, nrhs - c0);
        const int CS = hm_panel_width(nc);
        if (CS > p->n_ws_cs) {
            HM_CUDA(cudaStreamSynchronize(st));
            const size_t xt_rows = ((size_t)std::max<int64_t>(L.ncols, 1) + 3) & ~(size_t)3;
            HM_CUDA(p->wXt.alloc(xt_rows * CS));
            HM_CUDA(p->wYt.alloc((size_t)std::max<int64_t>(L.nrows, 1) * CS));
            HM_CUDA(p->n_MUp.alloc((size_t)std::max(p->n_cols.nnodes, 1) * HM_NEST_R * CS));
            HM_CUDA(p->n_LAMp.alloc((size_t)std::max(p->n_rows.nnodes, 1) * HM_NEST_R * CS));
            HM_CUDA(p->n_Sp.alloc((((size_t)std::max<int64_t>(p->n_fin_rows, 1) + 3) & ~(size_t)3) * CS));
            p->n_ws_cs = CS;
            p->ws_cs = 0; // the plain panel path re-allocates its own workspace if it is ever taken
        }
        cudaEvent_t *ev = p->tcount < p->tcap ? &p->tev[(size_t)p->tcount * 4] : nullptr;
        if (ev) HM_CUDA(cudaEventRecord(ev[0], st));
        HM_CUDA(hm_launch_panel_in(dX + c0 * ldx, ldx, L.ncols, nc, CS, p->wXt.p, st, true));
        HM_CUDA(hm_launch_nest_up_panel(CS, p->n_cols, p->f_py.p, p->wXt.p, p->n_M.p, p->n_MUp.p, st));
        if (ev) HM_CUDA(cudaEventRecord(ev[1], st));
        HM_CUDA(hm_launch_nest_core_panel(CS, p->n_rows.nnodes, p->n_rleaf_begin.p, p->n_rleaf.p, p->n_cores.p,
                                          p->n_MUp.p, p->n_LAMp.p, st));
        if (ev) HM_CUDA(cudaEventRecord(ev[2], st));
        HM_CUDA(hm_launch_nest_down_panel(CS, p->n_rows, p->n_fin.p, p->n_M.p, p->n_LAMp.p, p->n_Sp.p, st));
        HM_CUDA(hm_launch_free3_panel(CS, p->n_items3p.p, (int64_t)p->n_items3p.n, p->n_runsp.p, p->n_frunp.p, p->f_px.p,
                                      p->f_py.p, p->wXt.p, p->n_Sp.p, p->wYt.p, 0, p->kernel_id, st));
        HM_CUDA(hm_launch_panel_out(p->wYt.p, CS, L.row_begin, L.row_end, nc, dY + c0 * ldy, ldy, accumulate != 0, st));
        if (ev) {
            HM_CUDA(cudaEventRecord(ev[3], st));
            p->tcount++;
        }
    }
    return HM_OK;
}

// Multi-RHS: Y[:, c] (+)= H X[:, c].  Columns are processed in panels of up to 64 with the
// FP64 tensor-core kernels of hm_panel.cu; a single column takes the matvec path.
int32_t hm_matmat_device(hm_plan *p, const double *dX, int64_t ldx, double *dY, int64_t ldy, int64_t nrhs,
                         int32_t accumulate, void *stream)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        if (nrhs < 0) return fail(HM_ERR_SHAPE, "negative nrhs");
        if (nrhs == 0) return HM_OK;
        const HmLayout &L = p->L;
        if (ldx < std::max<int64_t>(L.ncols, 1) || ldy < std::max<int64_t>(L.nrows, 1))
            return fail(HM_ERR_SHAPE, "leading dimension smaller than the vector length");
        if ((!dX && L.ncols > 0) || (!dY && L.nrows > 0)) return fail(HM_ERR_NULL, "panel pointer is NULL");
        if (nrhs == 1) return hm_matvec_device(p, dX, dY, accumulate, stream);
        // a matrix-free plan has panel kernels for its Chebyshev form (hm_free_panel.cu); the barycentric
        // form (HMB200_FREE_FORM=bary) goes column by column
        if (!hm_panel_supports_rank(L.max_r, (int)std::min<int64_t>(nrhs, 64)) || (p->matrix_free && !p->free_cheb)) {
            // ranks beyond the panel kernels' shared-memory staging: column by column
            for (int64_t c = 0; c < nrhs; c++)
                if (int32_t rc = hm_matvec_device(p, dX + c * ldx, dY + c * ldy, accumulate, stream)) return rc;
            return HM_OK;
        }
        HM_DEVICE(p->device);
        cudaStream_t st = (cudaStream_t)stream;
        if (p->nested && p->n_fused_eval) return matmat_nested(p, dX, ldx, dY, ldy, nrhs, accumulate, st);
        if (p->panel_zcap == 0) {
            int zc = 4;
            for (const HmItem &it : L.items3) zc = std::max(zc, (int)it.S);
            p->panel_zcap = zc;
        }
        for (int64_t c0 = 0; c0 < nrhs; c0 += 64) {
            const int nc = (int)std::min<int64_t>(64, nrhs - c0);
            const int CS = hm_panel_width(nc);
            if (CS > p->ws_cs) {
                HM_CUDA(cudaStreamSynchronize(st));
                // Xt and Sp share one allocation (Sp right behind Xt): the stage-3 kernel addresses the z
                // rows of both through one base pointer and a 32-bit row index
                // (rows rounded up to the 4-row tiles of the blocked layout of matrix-free plans)
                const size_t xt_rows = ((size_t)std::max<int64_t>(L.ncols, 1) + 3) & ~(size_t)3;
                HM_CUDA(p->wXt.alloc((xt_rows + (((size_t)std::max<int64_t>(L.s_words, 1) + 3) & ~(size_t)3)) * CS));
                p->wSp = p->wXt.p + xt_rows * CS;
                HM_CUDA(p->wPp.alloc((size_t)std::max<int64_t>(L.partial_words, 1) * CS));
                HM_CUDA(p->wYt.alloc((size_t)std::max<int64_t>(L.nrows, 1) * CS));
                p->ws_cs = CS;
            }
            cudaEvent_t *ev = p->tcount < p->tcap ? &p->tev[(size_t)p->tcount * 4] : nullptr;
            if (ev) HM_CUDA(cudaEventRecord(ev[0], st));
            HM_CUDA(hm_launch_panel_in(dX + c0 * ldx, ldx, L.ncols, nc, CS, p->wXt.p, st, p->matrix_free));
            if (p->matrix_free)
                HM_CUDA(hm_launch_free1_panel(CS, p->items1.p, (int64_t)L.items1.size(), p->f_ent1.p, p->f_py.p,
                                              p->wXt.p, p->wPp.p, st));
            else
                HM_CUDA(hm_launch_panel_stage1(CS, p->items1.p, (int64_t)L.items1.size(), p->vstream.p, p->wXt.p,
                                               p->wPp.p, st));
            if (ev) HM_CUDA(cudaEventRecord(ev[1], st));
            HM_CUDA(hm_launch_panel_stage2(CS, p->cores.p, (int64_t)L.cores.size(), p->plist.p, p->wPp.p, p->core.p,
                                           p->wSp, std::max(L.max_r, 1), st, p->matrix_free));
            if (ev) HM_CUDA(cudaEventRecord(ev[2], st));
            for (size_t r = 0; r + 1 < L.round_begin.size(); r++) {
                int64_t i0 = L.round_begin[r], i1 = L.round_begin[r + 1];
                if (p->matrix_free)
                    HM_CUDA(hm_launch_free3_panel(CS, p->items3.p + i0, i1 - i0, p->runs.p, p->f_run3.p, p->f_px.p,
                                                  p->f_py.p, p->wXt.p, p->wSp, p->wYt.p, r == 0 ? 0 : 1,
                                                  p->kernel_id, st));
                else
                    HM_CUDA(hm_launch_panel_stage3(CS, p->items3.p + i0, i1 - i0, p->runs.p, p->ustream.p, p->wXt.p,
                                                   p->wSp, p->wYt.p, r == 0 ? 0 : 1, p->panel_zcap, st));
            }
            HM_CUDA(hm_launch_panel_out(p->wYt.p, CS, L.row_begin, L.row_end, nc, dY + c0 * ldy, ldy,
                                        accumulate != 0, st));
            if (ev) {
                HM_CUDA(cudaEventRecord(ev[3], st));
                p->tcount++;
            }
        }
        return HM_OK;
    });
}

int32_t hm_matmat(hm_plan *p, const double *X, int64_t ldx, double *Y, int64_t ldy, int64_t nrhs,
                  int32_t accumulate)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        if (nrhs < 0) return fail(HM_ERR_SHAPE, "negative nrhs");
        if (nrhs == 0) return HM_OK;
        const HmLayout &L = p->L;
        if (ldx < std::max<int64_t>(L.ncols, 1) || ldy < std::max<int64_t>(L.nrows, 1))
            return fail(HM_ERR_SHAPE, "leading dimension smaller than the vector length");
        if ((!X && L.ncols > 0) || (!Y && L.nrows > 0)) return fail(HM_ERR_NULL, "panel pointer is NULL");
        std::lock_guard<std::mutex> lock(p->mu);
        HM_DEVICE(p->device);
        cudaStream_t st = p->stream;
        const int64_t nc = std::max<int64_t>(L.ncols, 1), nr = std::max<int64_t>(L.nrows, 1);
        if (nrhs > p->nrhs_cap) {
            HM_CUDA(p->dX.alloc((size_t)nc * nrhs));
            HM_CUDA(p->dY.alloc((size_t)nr * nrhs));
            p->nrhs_cap = nrhs;
        }
        const int64_t r0 = L.row_begin, rows = L.row_end - L.row_begin;
        if (L.ncols > 0)
            HM_CUDA(cudaMemcpy2DAsync(p->dX.p, (size_t)nc * 8, X, (size_t)ldx * 8, (size_t)L.ncols * 8, (size_t)nrhs,
                                      cudaMemcpyHostToDevice, st));
        if (accumulate && rows > 0)
            HM_CUDA(cudaMemcpy2DAsync(p->dY.p + r0, (size_t)nr * 8, Y + r0, (size_t)ldy * 8, (size_t)rows * 8,
                                      (size_t)nrhs, cudaMemcpyHostToDevice, st));
        if (int32_t rc = hm_matmat_device(p, p->dX.p, nc, p->dY.p, nr, nrhs, accumulate, st)) return rc;
        if (rows > 0)
            HM_CUDA(cudaMemcpy2DAsync(Y + r0, (size_t)ldy * 8, p->dY.p + r0, (size_t)nr * 8, (size_t)rows * 8,
                                      (size_t)nrhs, cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        return HM_OK;
    });
}

// ---------------------------------------------------------------------------
// test hooks
// ---------------------------------------------------------------------------
int32_t hm_debug_fail_alloc(int64_t nth)
{
    hm_fault_arm(nth);
    return HM_OK;
}

int32_t hm_plan_num_leaves(const hm_plan *p, int64_t *out)
{
    return guarded([&]() -> int32_t {
        if (!p || !out) return fail(HM_ERR_NULL, "NULL argument");
        *out = (int64_t)p->L.leaves.size();
        return HM_OK;
    });
}

int32_t hm_plan_leaf_info(const hm_plan *p, int64_t leaf, int32_t *kind, int64_t *row0, int64_t *col0,
                          int64_t *m, int64_t *n, int64_t *r)
{
    return guarded([&]() -> int32_t {
        if (!p) return fail(HM_ERR_NULL, "plan is NULL");
        if (leaf < 0 || leaf >= (int64_t)p->L.leaves.size()) return fail(HM_ERR_RANGE, "leaf index out of range");
        const HmLeaf &l = p->L.leaves[(size_t)leaf];
        if (kind) *kind = l.kind;
        if (row0) *row0 = l.row0;
        if (col0) *col0 = l.col0;
        if (m) *m = l.m;
        if (n) *n = l.n;
        if (r) *r = l.ru;
        return HM_OK;
    });
}

int32_t hm_plan_read_leaf(hm_plan *p, int64_t leaf, int32_t which, double *out, int64_t cap)
{
    return guarded([&]() -> int32_t {
        if (!p || !out) return fail(HM_ERR_NULL, "NULL argument");
        if (p->matrix_free) return fail(HM_ERR_UNSUPPORTED, "not available on a matrix-free plan (hm_assemble_kernel_free)");
        if (leaf < 0 || leaf >= (int64_t)p->L.leaves.size()) return fail(HM_ERR_RANGE, "leaf index out of range");
        std::lock_guard<std::mutex> lock(p->mu);
        HM_DEVICE(p->device);
        HmLayout &L = p->L;
        if (!p->indexed) {
            for (size_t i = 0; i < L.fill1.size(); i++) p->idx1.emplace(L.fill1[i].leaf, i);
            for (size_t i = 0; i < L.fill3.size(); i++) p->idx3.emplace(L.fill3[i].leaf, i);
            p->indexed = true;
        }
        const HmLeaf &l = L.leaves[(size_t)leaf];
        const bool dense = l.kind == HM_LEAF_DENSE;
        int64_t need = 0;
        if (which == 0) need = dense ? -1 : l.m * l.ru;
        else if (which == 1) need = dense ? -1 : (l.kind == HM_LEAF_BARY2D ? (int64_t)l.ru * l.rv : l.ru);
        else if (which == 2) need = dense ? -1 : l.n * l.rv;
        else if (which == 3) need = dense ? l.m * l.n : -1;
        else return fail(HM_ERR_INVALID, "which must be 0..3");
        if (need < 0) return fail(HM_ERR_INVALID, "leaf kind has no such factor");
        if (cap < need) return fail(HM_ERR_SHAPE, "output capacity %lld < %lld", (long long)cap, (long long)need);
        for (int64_t i = 0; i < need; i++) out[i] = 0.0; // rows outside this part stay 0
        if (which == 1) {
            for (size_t c = 0; c < L.cores.size(); c++)
                if (L.core_leaf[c] == (int32_t)leaf)
                    HM_CUDA(cudaMemcpy(out, p->core.p + L.cores[c].core, (size_t)need * 8, cudaMemcpyDeviceToHost));
            return HM_OK;
        }
        if (which == 0 || which == 3) {
            auto range = p->idx3.equal_range((int32_t)leaf);
            for (auto it = range.first; it != range.second; ++it) {
                const HmFill &f = L.fill3[it->second];
                // kn columns of F rows (ld Fp) -> out[(off + i) + (k0 + k) * m]
                HM_CUDA(cudaMemcpy2D(out + f.off + (int64_t)f.k0 * l.m, (size_t)l.m * 8, p->ustream.p + f.dst,
                                     (size_t)f.Fp * 8, (size_t)f.F * 8, (size_t)f.kn, cudaMemcpyDeviceToHost));
            }
            return HM_OK;
        }
        auto range = p->idx1.equal_range((int32_t)leaf);
        std::vector<double> tmp;
        for (auto it = range.first; it != range.second; ++it) {
            const HmFill &f = L.fill1[it->second];
            tmp.resize((size_t)f.S * f.kn);
            HM_CUDA(cudaMemcpy2D(tmp.data(), (size_t)f.kn * 8, p->vstream.p + f.dst, (size_t)f.Fp * 8,
                                 (size_t)f.kn * 8, (size_t)f.S, cudaMemcpyDeviceToHost));
            for (int s = 0; s < f.S; s++)
                for (int k = 0; k < f.kn; k++) out[(f.off + s) + (int64_t)(f.k0 + k) * l.n] = tmp[(size_t)s * f.kn + k];
        }
        return HM_OK;
    });
}

} // extern "C"
