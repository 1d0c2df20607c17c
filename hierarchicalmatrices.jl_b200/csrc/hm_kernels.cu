// hm_kernels.cu -- hand-written sm_100a kernels of the H-matrix matvec.
//
// The reference walks the block tree and runs one scalar loop nest per leaf
// (/root/reference/src/algebra.jl:37-48 dense, :110-131 LowRankMatrix, :243-277
// BarycentricMatrix2D).  Here all leaves are applied by three flat passes over
// packed streams (see hm_layout.cpp):
//
//   stage 1 / stage 3   hm_stream_kernel: out[f] = sum_s W[s][f] * z[s]
//        one CTA per item, W is one contiguous slab read front to back with
//        128-bit streaming loads; thread <-> fast index f, so the sum over s is
//        a private register accumulation (no atomics, no shuffles); the few
//        column groups of a CTA are combined in a fixed order through shared
//        memory, so results are run-to-run deterministic.
//   stage 2             hm_core_kernel: tiny r x r core apply per low-rank leaf.
//
// HBM-bound: 8 B of stream per FMA.  No tensor-core use on the single-vector path.
#include <algorithm>
#include <cstdlib>

#include "hm_device.cuh"
#include "hm_kernels.cuh"

namespace {

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-
// serialization attribute may begin before its predecessor in the stream has finished;
// griddepcontrol.wait blocks until that predecessor has completed and its writes are visible.
// Without the attribute both instructions do nothing.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double2 ld_stream(const double2 *p)
{
    // read-once data: streaming (evict-first) so that x, s and the partial sums
    // keep their place in the 126 MB L2
    return __ldcs(p);
}

// ---------------------------------------------------------------------------
// stage 2: the core apply of one low-rank leaf
//   t[k]  = sum over the leaf's stage-1 partial sums, in column order
//   s     = Sigma .* t            (LowRankMatrix,       algebra.jl:120)
//   s     = F * t  (l outer)      (BarycentricMatrix2D, algebra.jl:260-265)
// Written as device functions because they run in two places: fused into the tail of
// stage 1 (the CTA that delivers a leaf's last partial sum applies its core at once) and
// in the stand-alone kernels (the default; the fused form is HMB200_FUSE_STAGE2=1).  Partial sums are read with ld.cg:
// other SMs wrote them.
// ---------------------------------------------------------------------------

// one warp, at most HM_CORE_BIG partial sums; tbuf: max_r doubles of shared memory
__device__ __noinline__ void core_apply_warp(const HmCoreBlock &cb, const int32_t *__restrict__ plist,
                                                const double *partial, const double *__restrict__ core,
                                                double *__restrict__ svec, double *tbuf, int lane)
{
    const int32_t *pl = plist + cb.pl0;
    const double *c = core + cb.core;
    if (cb.kind == HM_LEAF_BARY2D && cb.ru == 20 && cb.rv == 20) {
        // the rank the assembler produces (BLOCKRANK(Float64) = 20): lane k first issues its 20
        // loads of row k of F (independent, all in flight), then walks the partial list
        constexpr int R = 20;
        double f[R];
        const bool act = lane < R;
        if (act) {
#pragma unroll
            for (int l = 0; l < R; l++) f[l] = __ldcs(c + lane + l * R);
        }
        pdl_wait(); // the core does not depend on stage 1; the partial sums below do
        double t = 0.0;
        if (act) {
            int i = 0;
            for (; i + 3 < cb.npl; i += 4) {
                int o0 = pl[i], o1 = pl[i + 1], o2 = pl[i + 2], o3 = pl[i + 3];
                double p0 = __ldcg(partial + o0 + lane), p1 = __ldcg(partial + o1 + lane);
                double p2 = __ldcg(partial + o2 + lane), p3 = __ldcg(partial + o3 + lane);
                t += p0;
                t += p1;
                t += p2;
                t += p3;
            }
            for (; i < cb.npl; i++) t += __ldcg(partial + pl[i] + lane);
        }
        double a = 0.0;
#pragma unroll
        for (int l = 0; l < R; l++) {
            double tl = __shfl_sync(0xffffffffu, t, l);
            a = fma(f[l], tl, a);
        }
        if (act) svec[cb.soff + lane] = a;
        return;
    }
    pdl_wait();
    for (int k = lane; k < cb.rv; k += 32) {
        double t = 0.0;
        int i = 0;
        for (; i + 3 < cb.npl; i += 4) {
            double p0 = __ldcg(partial + pl[i] + k), p1 = __ldcg(partial + pl[i + 1] + k);
            double p2 = __ldcg(partial + pl[i + 2] + k), p3 = __ldcg(partial + pl[i + 3] + k);
            t += p0;
            t += p1;
            t += p2;
            t += p3;
        }
        for (; i < cb.npl; i++) t += __ldcg(partial + pl[i] + k);
        tbuf[k] = t;
    }
    __syncwarp();
    if (cb.kind == HM_LEAF_LOWRANK) {
        for (int k = lane; k < cb.ru; k += 32) svec[cb.soff + k] = tbuf[k] * c[k];
    } else {
        for (int k = lane; k < cb.ru; k += 32) {
            double a = 0.0;
            for (int l = 0; l < cb.rv; l++) a = fma(c[k + (size_t)l * cb.ru], tbuf[l], a);
            svec[cb.soff + k] = a;
        }
    }
    __syncwarp();
}

// a whole CTA of 256 threads for a leaf with a long partial list: warp w adds partials
// w, w+8, ..., the eight warp sums are combined in warp order.  sm: 9*max_r doubles.
// Must be called by all threads of the CTA.
__device__ __noinline__ void core_apply_cta(const HmCoreBlock &cb, const int32_t *__restrict__ plist,
                                               const double *partial, const double *__restrict__ core,
                                               double *__restrict__ svec, double *sm, int max_r, int tid)
{
    const int lane = tid & 31, w = tid >> 5;
    const int32_t *pl = plist + cb.pl0;
    pdl_wait();
    for (int k = lane; k < cb.rv; k += 32) {
        double t = 0.0;
        int i = w;
        for (; i + 24 < cb.npl; i += 32) {
            double p0 = __ldcg(partial + pl[i] + k), p1 = __ldcg(partial + pl[i + 8] + k);
            double p2 = __ldcg(partial + pl[i + 16] + k), p3 = __ldcg(partial + pl[i + 24] + k);
            t += p0;
            t += p1;
            t += p2;
            t += p3;
        }
        for (; i < cb.npl; i += 8) t += __ldcg(partial + pl[i] + k);
        sm[w * max_r + k] = t;
    }
    __syncthreads();
    double *tbuf = sm + 8 * max_r;
    for (int k = tid; k < cb.rv; k += 256) {
        double t = 0.0;
        for (int g = 0; g < 8; g++) t += sm[g * max_r + k];
        tbuf[k] = t;
    }
    __syncthreads();
    const double *c = core + cb.core;
    if (cb.kind == HM_LEAF_LOWRANK) {
        for (int k = tid; k < cb.ru; k += 256) svec[cb.soff + k] = tbuf[k] * c[k];
    } else {
        for (int k = tid; k < cb.ru; k += 256) {
            double a = 0.0;
            for (int l = 0; l < cb.rv; l++) a = fma(c[k + (size_t)l * cb.ru], tbuf[l], a);
            svec[cb.soff + k] = a;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 4)
hm_core_kernel(const HmCoreBlock *__restrict__ blocks, int64_t nblocks,
               const int32_t *__restrict__ plist, const double *__restrict__ partial,
               const double *__restrict__ core, double *__restrict__ svec, int max_r)
{
    extern __shared__ double tbuf_all[];
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (b >= nblocks) return;
    const HmCoreBlock cb = blocks[b];
    if (cb.npl > HM_CORE_BIG) return; // long partial lists: hm_core_big_kernel
    core_apply_warp(cb, plist, partial, core, svec, tbuf_all + (size_t)wib * max_r, lane);
}

__global__ void __launch_bounds__(256)
hm_core_big_kernel(const HmCoreBlock *__restrict__ blocks, const int32_t *__restrict__ big,
                   const int32_t *__restrict__ plist, const double *__restrict__ partial,
                   const double *__restrict__ core, double *__restrict__ svec, int max_r)
{
    extern __shared__ double sm[]; // [8][max_r] warp sums, then [max_r] t
    pdl_launch_dependents();
    core_apply_cta(blocks[big[blockIdx.x]], plist, partial, core, svec, sm, max_r, threadIdx.x);
}

// ---------------------------------------------------------------------------
// stage 1 / stage 3
// ---------------------------------------------------------------------------
template <bool GATHER, bool FUSE, bool PEERS>
__global__ void __launch_bounds__(HM_THREADS, 4)
hm_stream_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                 const double *__restrict__ W, const double *__restrict__ x,
                 const double *__restrict__ svec, double *out, int accumulate, HmFuse fz, HmPeers pe)
{
    constexpr int T = HM_THREADS;
    __shared__ double zs[HM_SMAX];
    __shared__ double2 red[T];
    __shared__ int rpos[GATHER ? HM_MAXRUNS + 1 : 1];
    __shared__ int rsrc[GATHER ? HM_MAXRUNS : 1];

    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x;
    const int S = it.S, F = it.F, L = it.Fp >> 1;
    // programmatic dependent launch: the next kernel of the stream may start its own prologue while
    // this grid runs (it blocks in griddepcontrol.wait until this grid has completed)
    pdl_launch_dependents();

    const double2 *__restrict__ W2 = reinterpret_cast<const double2 *>(W + it.slab);

    // ---- first batch of the slab: issued before z is staged.  The slab does not depend on the
    // previous kernel, so these loads are in flight during the prologue (and, under programmatic
    // dependent launch, while the previous kernel is still finishing) ----
    const bool narrow = L <= T;
    const int ncg = narrow ? (L > 0 ? T / L : 1) : 1; // column groups of L threads (narrow items)
    const int TA = ncg * L;
    const int nvec = S * L;
    double2 w0[8];
    bool pre = false;
    if (narrow) {
        if (t < TA && t + 7 * TA < nvec) {
            pre = true;
#pragma unroll
            for (int u = 0; u < 8; u++) w0[u] = ld_stream(W2 + t + u * TA);
        }
    } else if (S >= 8) {
        pre = true; // t < T < L: the thread's first column exists
#pragma unroll
        for (int u = 0; u < 8; u++) w0[u] = ld_stream(W2 + t + (size_t)u * L);
    }

    // everything below reads what the previous kernel wrote (x / the stage-2 vector) or writes
    // what it may still be reading
    pdl_wait();

    // ---- stage z in shared memory ----
    if (GATHER) {
        for (int r = t; r < it.nrun; r += T) {
            HmRun rr = runs[it.run0 + r];
            rpos[r] = rr.pos;
            rsrc[r] = rr.src;
        }
        if (t == 0) rpos[it.nrun] = S;
        __syncthreads();
        for (int e = t; e < S; e += T) {
            int lo = 0, hi = it.nrun; // rpos[lo] <= e < rpos[hi]
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (rpos[mid] <= e)
                    lo = mid;
                else
                    hi = mid;
            }
            int src = rsrc[lo], off = e - rpos[lo];
            zs[e] = src >= 0 ? x[src + off] : svec[(~src) + off];
        }
    } else {
        for (int e = t; e < S; e += T) zs[e] = x[it.zoff + e];
    }
    __syncthreads();

    // one writer per output element.  Stage 3 with PEERS: the all-gather of y is fused into
    // the kernel -- every owned row is stored straight into each rank's (symmetric, NVLink
    // peer-mapped) y buffer instead of a local buffer followed by a collective.
    auto put = [&](int f, double v) {
        double *o = out + it.out + f;
        if (GATHER) {
            const double r = (accumulate ? *o : 0.0) + v;
            if (PEERS) {
                for (int q = 0; q < pe.n; q++) pe.y[q][it.out + f] = r;
            } else {
                *o = r;
            }
        } else {
            *o = v;
        }
    };

    if (narrow) {
        // ncg column groups of L threads; thread t reads double2 number t, t+TA, ...
        double2 acc = make_double2(0.0, 0.0);
        if (t < TA) {
            const int cg = t / L;
            int i = t, si = cg;
            if (pre) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    double z = zs[si + u * ncg];
                    acc.x = fma(w0[u].x, z, acc.x);
                    acc.y = fma(w0[u].y, z, acc.y);
                }
                i += 8 * TA;
                si += 8 * ncg;
            }
            for (; i + 7 * TA < nvec; i += 8 * TA, si += 8 * ncg) {
                double2 w[8];
#pragma unroll
                for (int u = 0; u < 8; u++) w[u] = ld_stream(W2 + i + u * TA);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    double z = zs[si + u * ncg];
                    acc.x = fma(w[u].x, z, acc.x);
                    acc.y = fma(w[u].y, z, acc.y);
                }
            }
            for (; i < nvec; i += TA, si += ncg) {
                double2 w = ld_stream(W2 + i);
                double z = zs[si];
                acc.x = fma(w.x, z, acc.x);
                acc.y = fma(w.y, z, acc.y);
            }
        }
        if (ncg > 1) {
            red[t] = acc;
            __syncthreads();
            if (t < L) {
                double2 a = red[t];
                for (int g = 1; g < ncg; g++) {
                    double2 b = red[g * L + t];
                    a.x += b.x;
                    a.y += b.y;
                }
                acc = a;
            }
        }
        if (t < L) {
            int f = 2 * t;
            if (f < F) put(f, acc.x);
            if (f + 1 < F) put(f + 1, acc.y);
        }
    } else {
        // wide items: each thread owns whole columns of the slab
        for (int f2 = t; f2 < L; f2 += T) {
            double2 acc = make_double2(0.0, 0.0);
            const double2 *p = W2 + f2;
            int s = 0;
            if (pre && f2 == t) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    double z = zs[u];
                    acc.x = fma(w0[u].x, z, acc.x);
                    acc.y = fma(w0[u].y, z, acc.y);
                }
                s = 8;
            }
            for (; s + 7 < S; s += 8) {
                double2 w[8];
#pragma unroll
                for (int u = 0; u < 8; u++) w[u] = ld_stream(p + (size_t)(s + u) * L);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    double z = zs[s + u];
                    acc.x = fma(w[u].x, z, acc.x);
                    acc.y = fma(w[u].y, z, acc.y);
                }
            }
            for (; s < S; s++) {
                double2 w = ld_stream(p + (size_t)s * L);
                double z = zs[s];
                acc.x = fma(w.x, z, acc.x);
                acc.y = fma(w.y, z, acc.y);
            }
            int f = 2 * f2;
            if (f < F) put(f, acc.x);
            if (f + 1 < F) put(f + 1, acc.y);
        }
    }

    if (FUSE) {
        // Stage 2 fused into the tail of stage 1: count this item's arrival on every leaf it
        // contributes to; whoever delivers a leaf's last partial sum applies its core.
        __shared__ int nlast, nbig;
        __shared__ int lastlist[HM_MAXRUNS], biglist[8];
        __threadfence(); // this thread's partial sums are visible device-wide before the count
        __syncthreads();
        for (int e0 = 0; e0 < it.nrun; e0 += HM_MAXRUNS) {
            if (t == 0) nlast = nbig = 0;
            __syncthreads();
            for (int e = e0 + t; e < min(it.nrun, e0 + HM_MAXRUNS); e += T) {
                const int c = fz.s1ent[it.run0 + e];
                const int npl = fz.blocks[c].npl;
                if (atomicAdd(fz.counters + c, 1) == npl - 1) {
                    fz.counters[c] = 0; // re-armed for the next matvec
                    if (npl > HM_CORE_BIG && fz.max_r * 9 <= HM_SMAX) {
                        int k = atomicAdd(&nbig, 1);
                        if (k < 8) biglist[k] = c; else lastlist[atomicAdd(&nlast, 1)] = c;
                    } else {
                        lastlist[atomicAdd(&nlast, 1)] = c;
                    }
                }
            }
            __syncthreads();
            if (nlast > 0 || nbig > 0) {
                __threadfence();
                const int lane = t & 31, wib = t >> 5;
                for (int i = wib; i < nlast; i += T / 32)
                    core_apply_warp(fz.blocks[lastlist[i]], fz.plist, out, fz.core, fz.svec,
                                    zs + (size_t)wib * fz.max_r, lane);
                __syncthreads();
                for (int i = 0; i < min(nbig, 8); i++)
                    core_apply_cta(fz.blocks[biglist[i]], fz.plist, out, fz.core, fz.svec, zs, fz.max_r, t);
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// plan construction: fill the streams
// ---------------------------------------------------------------------------

// examples/Kernel.jl:34-37
__device__ __forceinline__ double kernel_eval(int id, double x, double y)
{
    double d = __dsub_rn(x, y);
    switch (id) {
    case 0: return __drcp_rn(d);
    case 1: return __drcp_rn(__dmul_rn(d, d));
    case 2: return __drcp_rn(__dmul_rn(__dmul_rn(d, d), d));
    default: return log(fabs(d));
    }
}

// Row of a barycentric factor, src/BarycentricMatrix.jl:256-287:
//   w[k] = lam[k] * inv(p - node_k), node_k = mid + half*cheb_k (:159-167);
//   row  = w / (w[0] + w[1] + ... sequentially).
// Explicit _rn intrinsics: no FMA contraction, so the factors are bit-identical
// to the two-rounding arithmetic of the reference.
template <int R>
__device__ __forceinline__ void bary_row(const HmCheb &cheb, double lo, double hi, double p,
                                         double (&w)[R])
{
    const double mid = __dmul_rn(0.5, __dadd_rn(lo, hi));
    const double half = __dmul_rn(0.5, __dsub_rn(hi, lo));
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < R; k++) {
        double node = __dadd_rn(mid, __dmul_rn(half, cheb.node[k]));
        w[k] = __dmul_rn(cheb.lam[k], __drcp_rn(__dsub_rn(p, node)));
        sum = __dadd_rn(sum, w[k]);
    }
#pragma unroll
    for (int k = 0; k < R; k++) w[k] = __ddiv_rn(w[k], sum);
}

template <int R>
__global__ void __launch_bounds__(128)
hm_fill3_kernel(const HmFill *__restrict__ fills, const HmLeaf *__restrict__ leaves,
                double *__restrict__ W, const double *__restrict__ px,
                const double *__restrict__ py, const HmCheb cheb, int kernel_id)
{
    const HmFill f = fills[blockIdx.x];
    const HmLeaf *l = leaves + f.leaf;
    const int t = threadIdx.x, T = blockDim.x;
    double *dst = W + f.dst;
    if (l->source == HM_SRC_COPY) {
        const double *src = l->dU + f.off + (int64_t)f.k0 * l->ldu;
        const int64_t ld = l->ldu;
        for (int idx = t; idx < f.kn * f.F; idx += T) {
            int k = idx / f.F, i = idx - k * f.F;
            dst[(int64_t)k * f.Fp + i] = src[i + k * ld];
        }
    } else if (l->kind == HM_LEAF_DENSE) {
        // T[f(x[i], y[j]) for i in ir, j in jr] -- src/KernelMatrix.jl:57-60
        if (kernel_id == HM_KERNEL_HOST_FN) return; // evaluated on the host (hm_assemble_kernel_fn)
        const double *xr = px + l->xi0 + f.off;
        const double *yc = py + l->yj0 + f.k0;
        for (int idx = t; idx < f.kn * f.F; idx += T) {
            int k = idx / f.F, i = idx - k * f.F;
            dst[(int64_t)k * f.Fp + i] = kernel_eval(kernel_id, xr[i], yc[k]);
        }
    } else {
        const double *xr = px + l->xi0 + f.off;
        const double a = l->a, b = l->b;
        for (int i = t; i < f.F; i += T) {
            double w[R];
            bary_row<R>(cheb, a, b, xr[i], w);
#pragma unroll
            for (int k = 0; k < R; k++)
                if (k >= f.k0 && k < f.k0 + f.kn) dst[(int64_t)(k - f.k0) * f.Fp + i] = w[k];
        }
    }
}

template <int R>
__global__ void __launch_bounds__(128)
hm_fill1_kernel(const HmFill *__restrict__ fills, const HmLeaf *__restrict__ leaves,
                double *__restrict__ W, const double *__restrict__ py, const HmCheb cheb)
{
    const HmFill f = fills[blockIdx.x];
    const HmLeaf *l = leaves + f.leaf;
    const int t = threadIdx.x, T = blockDim.x;
    double *dst = W + f.dst;
    if (l->source == HM_SRC_COPY) {
        const double *src = l->dV + f.off;
        const int64_t ld = l->ldv;
        for (int idx = t; idx < f.kn * f.S; idx += T) {
            int k = idx / f.S, s = idx - k * f.S;
            dst[(int64_t)s * f.Fp + k] = src[s + k * ld];
        }
    } else {
        const double *yc = py + l->yj0 + f.off;
        const double c = l->c, d = l->d;
        for (int s = t; s < f.S; s += T) {
            double w[R];
            bary_row<R>(cheb, c, d, yc[s], w);
#pragma unroll
            for (int k = 0; k < R; k++) dst[(int64_t)s * f.Fp + k] = w[k];
        }
    }
}

__global__ void __launch_bounds__(128)
hm_fillcore_kernel(const HmCoreBlock *__restrict__ blocks, const int32_t *__restrict__ core_leaf,
                   const HmLeaf *__restrict__ leaves, double *__restrict__ core, const HmCheb cheb,
                   int kernel_id)
{
    const HmCoreBlock cb = blocks[blockIdx.x];
    const HmLeaf *l = leaves + core_leaf[blockIdx.x];
    const int t = threadIdx.x, T = blockDim.x;
    double *dst = core + cb.core;
    if (l->source == HM_SRC_COPY) {
        if (cb.kind == HM_LEAF_LOWRANK) {
            for (int k = t; k < cb.ru; k += T) dst[k] = l->dC[k];
        } else {
            for (int idx = t; idx < cb.ru * cb.rv; idx += T) {
                int n = idx / cb.ru, m = idx - n * cb.ru;
                dst[idx] = l->dC[m + n * l->ldc];
            }
        }
    } else {
        // F[m,n] = f(x_m, y_n) at the mapped Chebyshev nodes -- BarycentricMatrix.jl:159-175
        if (kernel_id == HM_KERNEL_HOST_FN) return; // evaluated on the host (hm_assemble_kernel_fn)
        const double xm = __dmul_rn(0.5, __dadd_rn(l->a, l->b)), xh = __dmul_rn(0.5, __dsub_rn(l->b, l->a));
        const double ym = __dmul_rn(0.5, __dadd_rn(l->c, l->d)), yh = __dmul_rn(0.5, __dsub_rn(l->d, l->c));
        for (int idx = t; idx < cb.ru * cb.rv; idx += T) {
            int n = idx / cb.ru, m = idx - n * cb.ru;
            double xn = __dadd_rn(xm, __dmul_rn(xh, cheb.node[m]));
            double yn = __dadd_rn(ym, __dmul_rn(yh, cheb.node[n]));
            dst[idx] = kernel_eval(kernel_id, xn, yn);
        }
    }
}

// Matrix-free plans in Chebyshev form: core <- C F~ C' per BarycentricMatrix2D leaf, where
// C[q,k] = (2/R) T_q(cheb_k) (1/R for q = 0) maps values at the R first-kind Chebyshev nodes to
// Chebyshev coefficients (discrete orthogonality).  With it stage 2 turns the moments mu of stage 1
// straight into the coefficients c of stage 3:  c = C F~ C' mu.
//
// F~ is F moved from the nodes the reference really uses to the exact Chebyshev nodes.  The reference
// evaluates f at the *rounded* nodes n_k = fl(mid + fl(half cheb_k)) (BarycentricMatrix.jl:159-167);
// in the box variable that is xi_k = cheb_k + eps_k with eps_k up to ulp(mid) / half -- 1e-6 for the
// tiny boxes of clustered point sets.  Treating those values as values at cheb_k would cost that
// much relative accuracy on the block; one Newton-like step with the Chebyshev differentiation
// matrix D removes it to second order:  g(cheb) = f - diag(eps) D f, for rows and for columns.
__global__ void __launch_bounds__(128)
hm_core_cheb_kernel(const HmCoreBlock *__restrict__ blocks, const int32_t *__restrict__ core_leaf,
                    const HmLeaf *__restrict__ leaves, double *__restrict__ core, const double *__restrict__ Cm,
                    const double *__restrict__ Dm, const HmCheb cheb)
{
    constexpr int R = 20;
    __shared__ double Fs[R * R], Gs[R * R], Cs[R * R], Ds[R * R], ex[R], ey[R];
    const HmCoreBlock cb = blocks[blockIdx.x];
    if (cb.kind != HM_LEAF_BARY2D || cb.ru != R || cb.rv != R) return;
    const HmLeaf *l = leaves + core_leaf[blockIdx.x];
    double *F = core + cb.core;
    const int t = threadIdx.x, T = blockDim.x;
    for (int i = t; i < R * R; i += T) {
        Fs[i] = F[i];   // F[m + l*R]
        Cs[i] = Cm[i];  // C[q + k*R]
        Ds[i] = Dm[i];  // D[i + j*R]
    }
    if (t < 2 * R) {
        const bool xs = t < R;
        const int k = xs ? t : t - R;
        const double lo = xs ? l->a : l->c, hi = xs ? l->b : l->d;
        const double mid = __dmul_rn(0.5, __dadd_rn(lo, hi)), half = __dmul_rn(0.5, __dsub_rn(hi, lo));
        const double node = __dadd_rn(mid, __dmul_rn(half, cheb.node[k])); // the node the reference uses
        const double e = __ddiv_rn(__dsub_rn(node, mid), half) - cheb.node[k];
        (xs ? ex : ey)[k] = e;
    }
    __syncthreads();
    for (int i = t; i < R * R; i += T) { // rows: G = F - diag(ex) D F
        const int m = i % R, c = i / R;
        double a = 0.0;
        for (int j = 0; j < R; j++) a = fma(Ds[m + j * R], Fs[j + c * R], a);
        Gs[i] = Fs[i] - ex[m] * a;
    }
    __syncthreads();
    for (int i = t; i < R * R; i += T) { // columns: F~ = G - (D G')' diag(ey)
        const int m = i % R, c = i / R;
        double a = 0.0;
        for (int j = 0; j < R; j++) a = fma(Ds[c + j * R], Gs[m + j * R], a);
        Fs[i] = Gs[i] - ey[c] * a;
    }
    __syncthreads();
    for (int i = t; i < R * R; i += T) { // G[m][q] = sum_l F~[m][l] C[q][l]
        const int m = i % R, q = i / R;
        double a = 0.0;
        for (int c = 0; c < R; c++) a = fma(Fs[m + c * R], Cs[q + c * R], a);
        Gs[m + q * R] = a;
    }
    __syncthreads();
    for (int i = t; i < R * R; i += T) { // out[p][q] = sum_m C[p][m] G[m][q]
        const int pp = i % R, q = i / R;
        double a = 0.0;
        for (int m = 0; m < R; m++) a = fma(Cs[pp + m * R], Gs[m + q * R], a);
        F[pp + q * R] = a;
    }
}

// ---------------------------------------------------------------------------
// matrix-free apply ("fused assemble + apply")
//
// The same items, runs and fill tables as the stored operator, but the slabs are never
// written: every U, V and dense entry is evaluated from the point sets at the moment it is
// used (the arithmetic of the fill kernels above: src/BarycentricMatrix.jl:248-297,
// src/KernelMatrix.jl:57-60) and consumed at once.  A barycentric row is used in its
// unnormalised form, y_i += (sum_k w_k s_k) / (sum_k w_k), one divide per (row, leaf) instead of
// one per entry.  The bound moves from HBM to the FP64 pipe (one reciprocal per entry); the
// operator occupies no memory beyond its tables and the r x r cores, so N = 2^24 fits one GPU.
// ---------------------------------------------------------------------------
// stage 1: partial[item.out + (leaf, k)] = sum_s V_leaf[s, k] x[zoff + s] with V evaluated on the
// fly: V[s,k] = lam_k r_sk / sigma_s, r_sk = 1/(y_s - node_k), sigma_s = sum_k lam_k r_sk.  A warp
// owns a unit = (leaf of the item, chunk of its columns); the leaf's mapped nodes (exact
// two-rounding form of BarycentricMatrix.jl:159-167) sit in shared memory, lane pairs stride the
// columns with private accumulators of r_sk x_s / sigma_s, the warp sum goes through shared
// memory in lane order, and the chunks of a leaf are combined in chunk order -- deterministic.
//
// CHEB = true (the default of matrix-free plans): the same interpolant in the
// Chebyshev basis.  A barycentric row is the vector of Lagrange cardinal functions l_k(eta) of the
// R first-kind Chebyshev nodes, l_k(eta) = sum_q C[q,k] T_q(eta), so  t = C' mu  with the moments
// mu_q = sum_s x_s T_q(eta_s), eta_s = (y_s - mid) / half.  The lane that owns a column runs the
// three-term recurrence T_{q+1} = 2 eta T_q - T_{q-1} and accumulates x_s T_q: 2 FP64 operations
// per (column, q) and no reciprocal, against 6 + MUFU for the barycentric entry.  C' (and C for
// stage 3) are folded into the leaf's core at plan time (hm_core_cheb_kernel), so stage 2 is unchanged.
template <int R, bool CHEB>
__global__ void __launch_bounds__(HM_FREE1_THREADS, CHEB ? 6 : 8)
hm_free1_kernel(const HmItem *__restrict__ items, const HmFreeEnt *__restrict__ ents,
                const double *__restrict__ py, const double *__restrict__ x, double *__restrict__ partial,
                const HmCheb cheb)
{
    // Two lanes share a column, each taking half of the R ranks: half the accumulator registers per
    // thread (more warps per SM) for one shuffle per column (the two halves of sigma).
    constexpr int T = HM_FREE1_THREADS, H = R / 2;
    static_assert(R % 2 == 0, "rank must be even");
    extern __shared__ double ures[]; // [units][R]
    __shared__ double2 nodeW[CHEB ? 1 : T / 32][CHEB ? 1 : R]; // (node_k, lam_k) of the warp's current leaf
    __shared__ double wred[T / 32][32][(CHEB ? R : H) + 1];
    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int h = lane & 1, cl = lane >> 1;
    const int S = it.S;
    int nch, CH;
    hm_free1_split(S, it.nrun, nch, CH);
    const int U = it.nrun * nch;
    const double *__restrict__ xs = x + it.zoff;
    if (CHEB && S <= 256) {
        // short column segments (the joint slabs of the small leaves: 40-79 columns, a dozen leaves):
        // eight lanes per leaf, four leaves per warp, so that a lane sums several columns before the
        // 20 accumulators go through the shared-memory reduction
        for (int e0 = warp * 4; e0 < it.nrun; e0 += 4 * (T / 32)) {
            const int sub = lane >> 3, cl = lane & 7;
            const int e = e0 + sub;
            double acc[R];
#pragma unroll
            for (int k = 0; k < R; k++) acc[k] = 0.0;
            if (e < it.nrun) {
                const HmFreeEnt en = ents[it.run0 + e];
                const double *__restrict__ yc = py + en.yoff;
                const double ih = __drcp_rn(en.half);
                for (int s = cl; s < S; s += 8) {
                    const double eta = (yc[s] - en.mid) * ih, two = eta + eta, xv = xs[s];
                    double tm2 = 1.0, tm1 = eta;
                    acc[0] += xv;
                    acc[1] = fma(xv, eta, acc[1]);
#pragma unroll
                    for (int k = 2; k < R; k++) {
                        const double tk = fma(two, tm1, -tm2);
                        acc[k] = fma(xv, tk, acc[k]);
                        tm2 = tm1;
                        tm1 = tk;
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < R; k++) wred[warp][lane][k] = acc[k];
            __syncwarp();
            for (int idx = lane; idx < 4 * R; idx += 32) {
                const int sb = idx / R, k = idx - sb * R;
                if (e0 + sb < it.nrun) {
                    double tsum = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; j++) tsum += wred[warp][sb * 8 + j][k];
                    partial[it.out + ents[it.run0 + e0 + sb].fofs + k] = tsum;
                }
            }
        }
        return;
    }
    for (int u = warp; u < U; u += T / 32) {
        const int e = u / nch, c = u - e * nch;
        const HmFreeEnt en = ents[it.run0 + e];
        if (CHEB) {
            const double *__restrict__ yc = py + en.yoff;
            const double ih = __drcp_rn(en.half);
            double acc[R];
#pragma unroll
            for (int k = 0; k < R; k++) acc[k] = 0.0;
            const int s1 = min(S, (c + 1) * CH);
            for (int s = c * CH + lane; s < s1; s += 32) {
                const double eta = (yc[s] - en.mid) * ih, two = eta + eta, xv = xs[s];
                double tm2 = 1.0, tm1 = eta;
                acc[0] += xv;
                acc[1] = fma(xv, eta, acc[1]);
#pragma unroll
                for (int k = 2; k < R; k++) {
                    const double tk = fma(two, tm1, -tm2);
                    acc[k] = fma(xv, tk, acc[k]);
                    tm2 = tm1;
                    tm1 = tk;
                }
            }
            // warp sum through shared memory, in lane order
            __syncwarp();
#pragma unroll
            for (int k = 0; k < R; k++) wred[warp][lane][k] = acc[k];
            __syncwarp();
            if (lane < R) {
                double tsum = 0.0;
#pragma unroll 8
                for (int j = 0; j < 32; j++) tsum += wred[warp][j][lane];
                ures[u * R + lane] = tsum;
            }
            continue;
        }
        __syncwarp();
        if (lane < R)
            nodeW[warp][lane] = make_double2(__dadd_rn(en.mid, __dmul_rn(en.half, cheb.node[lane])), cheb.lam[lane]);
        __syncwarp();
        const double2 *__restrict__ nw = nodeW[warp] + h * H;
        const double *__restrict__ yc = py + en.yoff;
        double acc[H];
#pragma unroll
        for (int k = 0; k < H; k++) acc[k] = 0.0;
        const int s1 = min(S, (c + 1) * CH);
        for (int s0 = c * CH; s0 < s1; s0 += 16) { // uniform trip count: every lane joins the shuffle
            const int s = s0 + cl;
            const bool valid = s < s1;
            const double q = yc[valid ? s : s1 - 1];
            const double xv = valid ? xs[s] : 0.0;
            double r[H], sum = 0.0;
#pragma unroll
            for (int k = 0; k < H; k++) {
                const double2 nl = nw[k];
                r[k] = frcp(__dsub_rn(q, nl.x));
                sum = fma(nl.y, r[k], sum);
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            const double cf = xv * frcp(sum);
#pragma unroll
            for (int k = 0; k < H; k++) acc[k] = fma(r[k], cf, acc[k]);
        }
        // warp sum through shared memory, in column-lane order
        __syncwarp();
#pragma unroll
        for (int k = 0; k < H; k++) wred[warp][lane][k] = acc[k];
        __syncwarp();
        if (lane < R) {
            const int kh = lane / H, kk = lane - kh * H;
            double tsum = 0.0;
#pragma unroll 8
            for (int j = 0; j < 16; j++) tsum += wred[warp][2 * j + kh][kk];
            ures[u * R + lane] = cheb.lam[lane] * tsum;
        }
    }
    __syncthreads();
    for (int idx = t; idx < it.nrun * R; idx += T) {
        const int e = idx / R, k = idx - e * R;
        double sum = 0.0;
        for (int c = 0; c < nch; c++) sum += ures[(e * nch + c) * R + k];
        partial[it.out + ents[it.run0 + e].fofs + k] = sum;
    }
}

// stage 3: y[item.out + f] (+)= sum over the item's runs, entries evaluated on the fly.  G = T / F
// thread groups share the rows.  Dense runs: the columns are dealt out over the groups.  Low-rank
// runs go through shared-memory tables in batches of 24: per run the R exact nodes and lam_k s_k,
// so that a thread spends one subtraction, one reciprocal and two FMAs per entry and one divide
// per (row, leaf):  y_i += (sum_k lam_k s_k r_ik) / (sum_k lam_k r_ik).  The group sums are
// combined in group order -- deterministic.
//
// CHEB = true: the stage-2 vector holds the Chebyshev coefficients c = C s of every leaf's row
// interpolant, and a low-rank run is the Clenshaw sum  y_i += sum_q c_q T_q(xi_i),
// xi_i = (x_i - mid) / half: 2 FP64 operations per (row, q), no reciprocal.  A thread keeps four
// rows going against one broadcast coefficient (every operand from shared memory would otherwise
// make the shared-memory port, not the FP64 pipe, the limit).
template <int R, bool PEERS, bool CHEB>
__global__ void __launch_bounds__(HM_THREADS, CHEB ? 4 : 6)
hm_free3_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                const HmFreeRun *__restrict__ frun,
                const double *__restrict__ px, const double *__restrict__ py,
                const double *__restrict__ x, const double *__restrict__ svec, double *y, int accumulate,
                const HmCheb cheb, int kernel_id, HmPeers pe)
{
    constexpr int T = HM_THREADS, B = 24;
    extern __shared__ double zs[]; // the widest z of the launch (host-sized), not the 32 KB worst case
    __shared__ double red[T];
    __shared__ double2 tab[B][R];
    __shared__ int64_t xoff[B];
    __shared__ int rpos[HM_MAXRUNS + 1];
    __shared__ int rsrc[HM_MAXRUNS];
    __shared__ int lrlist[HM_MAXRUNS];
    __shared__ double2 rbox[HM_MAXRUNS]; // (mid, half) of every low-rank run: read once, up front
    __shared__ int64_t rxo[HM_MAXRUNS];
    __shared__ int2 rk[HM_MAXRUNS];
    __shared__ int nlr_s, ndn_s, samex_s;
    __shared__ int dnlist[HM_MAXRUNS];
    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31;
    const int S = it.S, F = it.F;

    for (int r = t; r < it.nrun; r += T) {
        HmRun rr = runs[it.run0 + r];
        rpos[r] = rr.pos;
        rsrc[r] = rr.src;
        if (rr.src < 0) {
            const HmFreeRun fr = frun[it.run0 + r];
            rbox[r] = make_double2(fr.mid, fr.half);
            rxo[r] = fr.xoff;
            rk[r] = make_int2(fr.k0, fr.kn);
        }
    }
    if (t == 0) rpos[it.nrun] = S;
    __syncthreads();
    if (t < 32) { // the low-rank runs (src < 0: they read the stage-2 vector) and the dense ones, in order
        int cnt = 0, dcnt = 0;
        bool same = true;
        for (int base = 0; base < it.nrun; base += 32) {
            const int r = base + lane;
            const bool lrr = r < it.nrun && rsrc[r] < 0, dnr = r < it.nrun && rsrc[r] >= 0;
            const unsigned m = __ballot_sync(0xffffffffu, lrr), md = __ballot_sync(0xffffffffu, dnr);
            if (lrr) lrlist[cnt + __popc(m & ((1u << lane) - 1u))] = r;
            if (dnr) dnlist[dcnt + __popc(md & ((1u << lane) - 1u))] = r;
            cnt += __popc(m);
            dcnt += __popc(md);
        }
        __syncwarp();
        // do all low-rank runs address the same points for the item's rows?
        for (int i = lane; i < cnt; i += 32) same = same && rxo[lrlist[i]] == rxo[lrlist[0]];
        same = __all_sync(0xffffffffu, same);
        if (lane == 0) {
            nlr_s = cnt;
            ndn_s = dcnt;
            samex_s = same && cnt > 0;
        }
    }
    // z: 32 lanes per run (the low-rank runs are R <= 32 long, the dense ones a few times that)
    for (int idx = t; idx < it.nrun * 32; idx += T) {
        const int r = idx >> 5;
        const int pos = rpos[r], len = rpos[r + 1] - pos, src = rsrc[r];
        for (int k = idx & 31; k < len; k += 32) zs[pos + k] = src >= 0 ? x[src + k] : svec[(~src) + k];
    }
    __syncthreads();
    const int nlr = nlr_s;

    const int G = F > 0 ? T / F : 1; // F <= T (checked when the plan is built)
    const int g = F > 0 ? t / F : G, f = t - g * F;
    const bool active = g < G;
    double acc = 0.0;
    if (active) {
        const int ndn = ndn_s;
        for (int ri = 0; ri < ndn; ri++) {
            const int r = dnlist[ri];
            const HmFreeRun fr = frun[it.run0 + r];
            const double p = px[fr.xoff + f];
            const double *__restrict__ yc = py + fr.yoff;
            const double *__restrict__ z = zs + rpos[r];
            if (kernel_id == 0) { // Cauchy: the branch-free form of the common case, four columns in flight
                int j = g;
                for (; j + 3 * G < fr.kn; j += 4 * G) {
                    const double y0 = yc[j], y1 = yc[j + G], y2 = yc[j + 2 * G], y3 = yc[j + 3 * G];
                    const double r0 = frcp(__dsub_rn(p, y0)), r1 = frcp(__dsub_rn(p, y1));
                    const double r2 = frcp(__dsub_rn(p, y2)), r3 = frcp(__dsub_rn(p, y3));
                    acc = fma(r0, z[j], acc);
                    acc = fma(r1, z[j + G], acc);
                    acc = fma(r2, z[j + 2 * G], acc);
                    acc = fma(r3, z[j + 3 * G], acc);
                }
                for (; j < fr.kn; j += G) acc = fma(frcp(__dsub_rn(p, yc[j])), z[j], acc);
            } else {
                for (int j = g; j < fr.kn; j += G) acc = fma(kernel_eval_fast(kernel_id, p, yc[j]), z[j], acc);
            }
        }
    }
    // CHEB: row quads.  Thread (g4, fq) owns rows fq, fq + FQ, fq + 2 FQ, fq + 3 FQ of the item
    const int FQ = (F + 3) >> 2;
    const int G4 = FQ > 0 ? T / FQ : 1;
    const int g4 = FQ > 0 ? t / FQ : G4, fq = t - g4 * FQ;
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};
    constexpr int BC = 2 * B; // coefficient tables are half the size of the barycentric ones
    __shared__ double2 tbox[CHEB ? BC : 1];
    __shared__ int64_t xoffc[CHEB ? BC : 1];
    // the rows of an item are the same points for every run when rows and points are numbered alike
    // (always so for KernelMatrix): then a thread loads its four points once
    const bool same_x = CHEB && samex_s;
    double p4[4] = {0.0, 0.0, 0.0, 0.0};
    if (same_x && g4 < G4) {
        const double *__restrict__ pp = px + rxo[lrlist[0]] + fq;
#pragma unroll
        for (int j = 0; j < 4; j++) p4[j] = fq + j * FQ < F ? pp[j * FQ] : 0.0;
    }
    for (int b0 = 0; CHEB && b0 < nlr; b0 += BC) {
        const int nb = min(BC, nlr - b0);
        __syncthreads(); // the previous batch's tables are no longer read
        double *tabc = reinterpret_cast<double *>(&tab[0][0]); // [BC][R] coefficients
        int64_t *xoff = xoffc;
        for (int idx = t; idx < nb * R; idx += T) {
            const int b = idx / R, k = idx - b * R;
            const int r = lrlist[b0 + b];
            const int2 kr = rk[r];
            tabc[b * R + k] = (k >= kr.x && k < kr.x + kr.y) ? zs[rpos[r] + k - kr.x] : 0.0;
            if (k == 0) {
                const double2 box = rbox[r];
                tbox[b] = make_double2(box.x, __drcp_rn(box.y));
                xoff[b] = rxo[r];
            }
        }
        __syncthreads();
        if (g4 < G4) {
            for (int b = g4; b < nb; b += G4) {
                const double2 box = tbox[b];
                const double *__restrict__ pp = px + xoff[b] + fq;
                const double *__restrict__ cf = tabc + b * R;
                double two[4], b1[4], b2[4];
                const double ih2 = box.y + box.y;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double p = same_x ? p4[j] : (fq + j * FQ < F ? pp[j * FQ] : box.x);
                    two[j] = (p - box.x) * ih2; // 2 xi
                    b1[j] = b2[j] = 0.0;
                }
#pragma unroll
                for (int k = R - 1; k >= 1; k--) {
                    const double ck = cf[k];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const double nb1 = fma(two[j], b1[j], ck) - b2[j];
                        b2[j] = b1[j];
                        b1[j] = nb1;
                    }
                }
                const double c0 = cf[0];
#pragma unroll
                for (int j = 0; j < 4; j++) acc4[j] += fma(0.5 * two[j], b1[j], c0) - b2[j];
            }
        }
    }
    for (int b0 = 0; !CHEB && b0 < nlr; b0 += B) {
        const int nb = min(B, nlr - b0);
        __syncthreads(); // the previous batch's tables are no longer read
        for (int idx = t; idx < nb * R; idx += T) {
            const int b = idx / R, k = idx - b * R;
            const int r = lrlist[b0 + b];
            const double2 box = rbox[r];
            const int2 kr = rk[r];
            const double zk = (k >= kr.x && k < kr.x + kr.y) ? zs[rpos[r] + k - kr.x] : 0.0;
            tab[b][k] = make_double2(__dadd_rn(box.x, __dmul_rn(box.y, cheb.node[k])), cheb.lam[k] * zk);
            if (k == 0) xoff[b] = rxo[r];
        }
        __syncthreads();
        if (active) {
            for (int b = g; b < nb; b += G) {
                const double p = px[xoff[b] + f];
                double dot = 0.0, sum = 0.0;
#pragma unroll
                for (int k = 0; k < R; k++) {
                    const double2 nz = tab[b][k];
                    const double r = frcp(__dsub_rn(p, nz.x));
                    dot = fma(nz.y, r, dot);
                    sum = fma(cheb.lam[k], r, sum);
                }
                acc = fma(dot, frcp(sum), acc);
            }
        }
    }
    red[t] = acc;
    if (CHEB) {
        __syncthreads(); // z is no longer read: the row-quad sums go through its storage (>= 4 T words)
#pragma unroll
        for (int j = 0; j < 4; j++) zs[j * T + t] = acc4[j];
    }
    __syncthreads();
    if (t < F) {
        double v = red[t];
        for (int gg = 1; gg < G; gg++) v += red[gg * F + t];
        if (CHEB) {
            const int j = t / FQ, q = t - j * FQ;
            for (int gg = 0; gg < G4; gg++) v += zs[j * T + gg * FQ + q];
        }
        double *o = y + it.out + t;
        const double r = (accumulate ? *o : 0.0) + v;
        if (PEERS) {
            for (int q = 0; q < pe.n; q++) pe.y[q][it.out + t] = r;
        } else {
            *o = r;
        }
    }
}

// ---------------------------------------------------------------------------
// operator updates on the packed streams: H <- H*Diagonal(b) and
// H <- Diagonal(b)*H without re-planning.  Reference: rmul!/lmul! ->
// scale! (/root/reference/src/HierarchicalMatrix.jl:15-16, 54-108) and its leaf methods
// (src/algebra.jl:280-315): dense A[i,j] *= b_j | b_i, low rank V[j,:] *= b_j | U[i,:] *= b_i.
// ---------------------------------------------------------------------------
// every element of a stage-3 slab by its row: W[s][f] *= b[row0 + f]
__global__ void __launch_bounds__(256)
hm_scale_rows_kernel(const HmItem *__restrict__ items, double *__restrict__ W, const double *__restrict__ b)
{
    const HmItem it = items[blockIdx.x];
    double *w = W + it.slab;
    const int n = it.Fp * it.S;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        int f = idx % it.Fp;
        if (f < it.F) w[idx] *= b[it.out + f];
    }
}

// every element of a stage-1 slab by its column of the operator: W[s][f] *= b[col0 + s]
__global__ void __launch_bounds__(256)
hm_scale_cols_v_kernel(const HmItem *__restrict__ items, double *__restrict__ W, const double *__restrict__ b)
{
    const HmItem it = items[blockIdx.x];
    double *w = W + it.slab;
    const int n = it.Fp * it.S;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) w[idx] *= b[it.zoff + idx / it.Fp];
}

// the dense-tile columns of a stage-3 slab (runs gathered from x): W[s][f] *= b[col(s)]
__global__ void __launch_bounds__(256)
hm_scale_cols_dense_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                           double *__restrict__ W, const double *__restrict__ b)
{
    const HmItem it = items[blockIdx.x];
    double *w = W + it.slab;
    for (int r = 0; r < it.nrun; r++) {
        const HmRun rr = runs[it.run0 + r];
        if (rr.src < 0) continue; // low-rank columns: their V rows carry the scaling
        const int n = rr.len * it.Fp;
        double *wr = w + (size_t)rr.pos * it.Fp;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) wr[idx] *= b[rr.src + idx / it.Fp];
    }
}

// ---------------------------------------------------------------------------
// adjoint apply y = H' x.  No hierarchical adjoint exists in the
// reference; the leaf rules are those of its Transpose/Adjoint leaves
// (/root/reference/src/algebra.jl:52-82 dense, :138-159 LowRankMatrix): y_j += sum_i A[i,j] x_i,
// temp = Sigma .* (U' x), y += V temp; BarycentricMatrix2D likewise with F'.
// The packed streams are read exactly as in the forward product, but reduced over the
// fast index: a group of lanes owns one slab row and combines with shuffles.
// ---------------------------------------------------------------------------
// Eight row sums held per lane (a[r] = this lane's share of row r) are reduced across a
// group of LR lanes by folding: each step halves the number of live values instead of
// running one butterfly per row (9 shuffles per 8 rows at LR = 32 instead of 40).  The
// order of the additions is fixed.  Returns the row (0..7) whose total ends up in a[0];
// `owner` tells whether this lane holds a complete total.
template <int LR>
__device__ __forceinline__ int fold8(double (&a)[8], int lg, bool &owner)
{
    static_assert(LR == 8 || LR == 16 || LR == 32, "group width");
    constexpr int B0 = LR / 2, B1 = LR / 4, B2 = LR / 8;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool up = lg & B0;
        double send = up ? a[i] : a[i + 4];
        double keep = up ? a[i + 4] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, B0);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const bool up = lg & B1;
        double send = up ? a[i] : a[i + 2];
        double keep = up ? a[i + 2] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, B1);
    }
    {
        const bool up = lg & B2;
        double send = up ? a[0] : a[1];
        double keep = up ? a[1] : a[0];
        a[0] = keep + __shfl_xor_sync(0xffffffffu, send, B2);
    }
#pragma unroll
    for (int d = B2 / 2; d > 0; d >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], d);
    owner = (lg & (B2 - 1)) == 0;
    return ((lg & B0) ? 4 : 0) + ((lg & B1) ? 2 : 0) + ((lg & B2) ? 1 : 0);
}

// groups of LR lanes, 8 rows per group and step; rows longer than 2*LR double2 words are
// walked in chunks of 2*LR (two loads per row and lane in flight), folded once at the end
template <int LR>
__device__ __forceinline__ void rowdot_groups(const double2 *__restrict__ W2, const double *zs, double *__restrict__ o,
                                              int S, int L, int t)
{
    constexpr int G = HM_THREADS / LR; // groups per CTA
    const int gi = t / LR, lg = t - gi * LR;
    for (int base = 0; base < S; base += G * 8) { // uniform over the CTA: every lane joins the shuffles
        const int row0 = base + gi * 8;
        double a[8];
#pragma unroll
        for (int r = 0; r < 8; r++) a[r] = 0.0;
        for (int c0 = 0; c0 < L; c0 += 2 * LR) {
            const int i0 = c0 + lg, i1 = c0 + lg + LR;
            const bool act0 = i0 < L, act1 = i1 < L;
            const double z0 = act0 ? zs[2 * i0] : 0.0, z1 = act0 ? zs[2 * i0 + 1] : 0.0;
            const double z2 = act1 ? zs[2 * i1] : 0.0, z3 = act1 ? zs[2 * i1 + 1] : 0.0;
            double2 w[8], w2[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const bool v = row0 + r < S;
                w[r] = (v && act0) ? __ldcs(W2 + (size_t)(row0 + r) * L + i0) : make_double2(0.0, 0.0);
                w2[r] = (v && act1) ? __ldcs(W2 + (size_t)(row0 + r) * L + i1) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int r = 0; r < 8; r++) {
                double q = fma(w[r].x, z0, a[r]);
                q = fma(w[r].y, z1, q);
                q = fma(w2[r].x, z2, q);
                a[r] = fma(w2[r].y, z3, q);
            }
        }
        bool owner;
        const int r = fold8<LR>(a, lg, owner);
        if (owner && row0 + r < S) o[row0 + r] = a[0];
    }
}

// MODE 0 (stage A'): items of the U-stream, z = x[rows of the item]
// MODE 1 (stage C'): items of the V-stream, z gathered from the adjoint stage-2 vector
template <int MODE>
__global__ void __launch_bounds__(HM_THREADS, 3)
hm_rowdot_kernel(const HmItem *__restrict__ items, const double *__restrict__ W,
                 const double *__restrict__ zsrc, const int32_t *__restrict__ s1ent,
                 const HmCoreBlock *__restrict__ blocks, double *__restrict__ PQ)
{
    constexpr int T = HM_THREADS;
    __shared__ double zs[HM_SMAX + 2];
    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int S = it.S, F = it.F, L = it.Fp >> 1;
    if (MODE == 0) {
        for (int f = t; f < it.Fp; f += T) zs[f] = f < F ? zsrc[it.out + f] : 0.0;
    } else {
        // z[f], f = (leaf, k) in the order of the item's leaf list.  The leaf records are fetched by
        // all threads at once (one dependent round trip for the whole list instead of one per leaf:
        // with a single warp walking it the other seven waited at the barrier, 8.4 stalled warps per
        // issue in round 1's profile), offsets by a short scan, the copies spread over the warps.
        __shared__ int e_soff[T], e_fofs[T + 1];
        int fbase = 0;
        for (int e0 = 0; e0 < it.nrun; e0 += T) {
            const int ne = min(T, it.nrun - e0);
            int rv = 0;
            if (t < ne) {
                const HmCoreBlock cb = blocks[s1ent[it.run0 + e0 + t]];
                e_soff[t] = cb.soff;
                rv = cb.rv;
            }
            // inclusive scan of rv over the chunk (warp scans + warp totals)
            int incl = rv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            __shared__ int wtot[T / 32];
            if (lane == 31) wtot[warp] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < warp; w++) woff += wtot[w];
            if (t < ne) e_fofs[t] = fbase + woff + incl - rv;
            int total = 0;
            for (int w = 0; w < T / 32; w++) total += wtot[w];
            if (t == 0) e_fofs[ne] = fbase + total;
            __syncthreads();
            for (int e = warp; e < ne; e += T / 32) {
                const int f0 = e_fofs[e], n = e_fofs[e + 1] - f0, so = e_soff[e];
                for (int k = lane; k < n; k += 32) zs[f0 + k] = zsrc[so + k];
            }
            fbase += total;
            __syncthreads();
        }
        if (t == 0 && fbase < it.Fp) zs[fbase] = 0.0;
    }
    __syncthreads();
    const double2 *__restrict__ W2 = reinterpret_cast<const double2 *>(W + it.slab);
    double *__restrict__ o = PQ + it.aux;
    // a group walks its rows in chunks of 2*LR words: the narrowest group that covers a row in one
    // chunk keeps the most rows (and loads) in flight per warp -- rank-20 slabs (L = 10) run four
    // groups of 8 lanes per warp
    if (L <= 16) {
        rowdot_groups<8>(W2, zs, o, S, L, t);
    } else if (L <= 32) {
        rowdot_groups<16>(W2, zs, o, S, L, t);
    } else {
        rowdot_groups<32>(W2, zs, o, S, L, t);
    }
}

// stage B': t'[k] = sum of the leaf's q pieces (row order); s' = F' t' | Sigma .* t'.
// CTAs [0, nbig) take one leaf with a long piece list each (a leaf of m rows has m/64
// pieces: thousands for the top levels) and split the list over 12 thread groups, combined
// in group order; the remaining CTAs take one leaf per warp.
__global__ void __launch_bounds__(256)
hm_core_adj_kernel(const HmCoreBlock *__restrict__ blocks, int64_t nblocks, const int32_t *__restrict__ q0,
                   const int32_t *__restrict__ qn, const int32_t *__restrict__ qlist,
                   const double *__restrict__ PQ, const double *__restrict__ core, double *__restrict__ svec,
                   int max_r, const int32_t *__restrict__ big, int nbig)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    if ((int)blockIdx.x < nbig) {
        const HmCoreBlock cb = blocks[big[blockIdx.x]];
        const int32_t *ql = qlist + q0[big[blockIdx.x]];
        const int n = qn[big[blockIdx.x]];
        const int ngrp = max(1, 256 / max(cb.ru, 1));
        const int g = tid / max(cb.ru, 1), k = tid - g * cb.ru;
        double *part = sm; // [ngrp][ru] group sums, then t' at part + ngrp*ru
        if (cb.ru <= 256) {
            if (g < ngrp) {
                double t = 0.0;
                int i = g;
                for (; i + 3 * ngrp < n; i += 4 * ngrp) {
                    double p0 = PQ[ql[i] + k], p1 = PQ[ql[i + ngrp] + k];
                    double p2 = PQ[ql[i + 2 * ngrp] + k], p3 = PQ[ql[i + 3 * ngrp] + k];
                    t += p0;
                    t += p1;
                    t += p2;
                    t += p3;
                }
                for (; i < n; i += ngrp) t += PQ[ql[i] + k];
                part[g * cb.ru + k] = t;
            }
            __syncthreads();
            double *tb = part + ngrp * cb.ru;
            for (int kk = tid; kk < cb.ru; kk += 256) {
                double t = 0.0;
                for (int gg = 0; gg < ngrp; gg++) t += part[gg * cb.ru + kk];
                tb[kk] = t;
            }
            __syncthreads();
            const double *c = core + cb.core;
            if (cb.kind == HM_LEAF_LOWRANK) {
                for (int kk = tid; kk < cb.rv; kk += 256) svec[cb.soff + kk] = tb[kk] * c[kk];
            } else {
                for (int l = tid; l < cb.rv; l += 256) {
                    double a = 0.0;
                    const double *col = c + (size_t)l * cb.ru;
                    for (int kk = 0; kk < cb.ru; kk++) a = fma(col[kk], tb[kk], a);
                    svec[cb.soff + l] = a;
                }
            }
            return;
        }
        // ranks beyond 256: fall through to the serial warp path on warp 0
        if (wib != 0) return;
    }
    int64_t b;
    if ((int)blockIdx.x < nbig) {
        b = big[blockIdx.x];
    } else {
        b = (int64_t)(blockIdx.x - nbig) * (blockDim.x >> 5) + wib;
        if (b >= nblocks) return;
        if (qn[b] > HM_ADJ_BIG) return; // handled by a "big" CTA
    }
    const int fs_words = max_r <= 20 ? 20 * 21 : 32 * 32; // must match hm_launch_adjoint
    double *tbuf = sm + (size_t)wib * (max_r + fs_words);
    double *Fs = tbuf + max_r; // F staged in shared memory when it fits (ru, rv <= 32)
    const HmCoreBlock cb = blocks[b];
    const double *c = core + cb.core;
    const int32_t *ql = qlist + q0[b];
    const int n = qn[b];
    if (cb.kind == HM_LEAF_BARY2D && cb.ru == 20 && cb.rv == 20) {
        // the rank the assembler produces: F goes to registers with coalesced loads that stay in
        // flight while lane k walks the piece list; then F is laid out in shared memory with an odd
        // pitch and lane l reads column l without bank conflicts
        constexpr int R = 20, RP = 21, NJ = (R * R + 31) / 32;
        double f[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j++) f[j] = lane + 32 * j < R * R ? __ldcs(c + lane + 32 * j) : 0.0;
        double t = 0.0;
        if (lane < R) {
            int i = 0;
            for (; i + 3 < n; i += 4) {
                double p0 = PQ[ql[i] + lane], p1 = PQ[ql[i + 1] + lane];
                double p2 = PQ[ql[i + 2] + lane], p3 = PQ[ql[i + 3] + lane];
                t += p0;
                t += p1;
                t += p2;
                t += p3;
            }
            for (; i < n; i++) t += PQ[ql[i] + lane];
        }
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            const int idx = lane + 32 * j; // F[k + l*R] -> Fs[l*RP + k]
            if (idx < R * R) Fs[(idx / R) * RP + idx % R] = f[j];
        }
        __syncwarp();
        double a = 0.0;
        const double *col = Fs + (lane < R ? lane : 0) * RP;
#pragma unroll
        for (int k = 0; k < R; k++) a = fma(col[k], __shfl_sync(0xffffffffu, t, k), a);
        if (lane < R) svec[cb.soff + lane] = a;
        return;
    }
    const bool staged = cb.kind == HM_LEAF_BARY2D && cb.ru <= 32 && cb.rv <= 32;
    if (staged)
        for (int i = lane; i < cb.ru * cb.rv; i += 32) Fs[i] = __ldcs(c + i); // coalesced
    for (int k = lane; k < cb.ru; k += 32) {
        double t = 0.0;
        int i = 0;
        for (; i + 3 < n; i += 4) {
            double p0 = PQ[ql[i] + k], p1 = PQ[ql[i + 1] + k], p2 = PQ[ql[i + 2] + k], p3 = PQ[ql[i + 3] + k];
            t += p0;
            t += p1;
            t += p2;
            t += p3;
        }
        for (; i < n; i++) t += PQ[ql[i] + k];
        tbuf[k] = t;
    }
    __syncwarp();
    if (cb.kind == HM_LEAF_LOWRANK) {
        for (int k = lane; k < cb.rv; k += 32) svec[cb.soff + k] = tbuf[k] * c[k];
    } else {
        for (int l = lane; l < cb.rv; l += 32) {
            double a = 0.0;
            const double *col = staged ? Fs + l * cb.ru : c + (size_t)l * cb.ru; // F[:, l]
            for (int k = 0; k < cb.ru; k++) a = fma(col[k], tbuf[k], a);
            svec[cb.soff + l] = a;
        }
    }
}

// stage D': y[j] = (accumulate ? y[j] : 0) + sum_i PQ[base_i + j], one writer per column
__global__ void __launch_bounds__(256)
hm_colsum_kernel(const HmColSeg *__restrict__ segs, const int64_t *__restrict__ bases,
                 const double *__restrict__ PQ, double *__restrict__ y, int accumulate)
{
    const HmColSeg sg = segs[blockIdx.x];
    const int j = sg.c0 + threadIdx.x;
    if (j >= sg.c1) return;
    double a = 0.0;
    const int64_t *b = bases + sg.b0;
    int i = 0;
    for (; i + 3 < sg.nb; i += 4) {
        double p0 = PQ[b[i] + j], p1 = PQ[b[i + 1] + j], p2 = PQ[b[i + 2] + j], p3 = PQ[b[i + 3] + j];
        a += p0;
        a += p1;
        a += p2;
        a += p3;
    }
    for (; i < sg.nb; i++) a += PQ[b[i] + j];
    y[j] = (accumulate ? y[j] : 0.0) + a;
}

} // namespace

namespace {
__global__ void hm_negate_kernel(const double *__restrict__ in, double *__restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = -in[i];
}
} // namespace

cudaError_t hm_launch_negate(const double *in, double *out, int64_t n, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    hm_negate_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(in, out, n);
    return cudaGetLastError();
}

cudaError_t hm_launch_adjoint(const HmAdjoint &A, const double *x, double *y, int accumulate, cudaStream_t st)
{
    if (A.n3 > 0)
        hm_rowdot_kernel<0><<<(unsigned)A.n3, HM_THREADS, 0, st>>>(A.items3, A.ustream, x, nullptr, nullptr, A.PQ);
    if (A.ncores > 0) {
        // per warp: t' (max_r) + staged F (32 x 32, or 20 x 21 when no rank exceeds 20: the kernel is
        // latency-bound and shared memory limits its occupancy); big CTAs: 13 * max_r at most
        const size_t fs_words = A.max_r <= 20 ? 20 * 21 : 32 * 32;
        size_t smem = (size_t)8 * ((size_t)A.max_r + fs_words) * sizeof(double);
        smem = std::max(smem, (size_t)(256 + A.max_r) * 2 * sizeof(double));
        if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(hm_core_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        unsigned grid = (unsigned)(A.nbig + (A.ncores + 7) / 8);
        hm_core_adj_kernel<<<grid, 256, smem, st>>>(A.blocks, A.ncores, A.q0, A.qn, A.qlist, A.PQ, A.core, A.svec,
                                                   A.max_r, A.big, A.nbig);
    }
    if (A.n1 > 0)
        hm_rowdot_kernel<1><<<(unsigned)A.n1, HM_THREADS, 0, st>>>(A.items1, A.vstream, A.svec, A.s1ent, A.blocks,
                                                                    A.PQ);
    if (A.nsegs > 0) hm_colsum_kernel<<<(unsigned)A.nsegs, 256, 0, st>>>(A.segs, A.bases, A.PQ, y, accumulate);
    return cudaGetLastError();
}

cudaError_t hm_launch_scale_rows(const HmItem *items3, int64_t n3, double *ustream, const double *b, cudaStream_t st)
{
    if (n3 <= 0) return cudaSuccess;
    hm_scale_rows_kernel<<<(unsigned)n3, 256, 0, st>>>(items3, ustream, b);
    return cudaGetLastError();
}

cudaError_t hm_launch_scale_cols(const HmItem *items1, int64_t n1, double *vstream, const HmItem *items3,
                                 int64_t n3, const HmRun *runs, double *ustream, const double *b,
                                 cudaStream_t st)
{
    if (n1 > 0) hm_scale_cols_v_kernel<<<(unsigned)n1, 256, 0, st>>>(items1, vstream, b);
    if (n3 > 0) hm_scale_cols_dense_kernel<<<(unsigned)n3, 256, 0, st>>>(items3, runs, ustream, b);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------
// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may start while the
// previous kernel of the stream is still running; it synchronises itself with griddepcontrol.wait.
// Only kernels that contain pdl_wait() may be launched this way.  HMB200_PDL=0 disables it.
static bool pdl_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HMB200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

template <class... KArgs, class... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                            bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

cudaError_t hm_launch_stage1(const HmItem *items, int64_t nitems, const double *vstream,
                             const double *x, double *partial, const HmFuse *fuse, cudaStream_t st, bool pdl)
{
    if (nitems <= 0) return cudaSuccess;
    if (fuse && fuse->counters && (size_t)fuse->max_r * 8 <= HM_SMAX)
        return launch_k(hm_stream_kernel<false, true, false>, (unsigned)nitems, HM_THREADS, 0, st, pdl, items,
                        (const HmRun *)nullptr, vstream, x, (const double *)nullptr, partial, 0, *fuse, HmPeers{});
    return launch_k(hm_stream_kernel<false, false, false>, (unsigned)nitems, HM_THREADS, 0, st, pdl, items,
                    (const HmRun *)nullptr, vstream, x, (const double *)nullptr, partial, 0, HmFuse{}, HmPeers{});
}

cudaError_t hm_launch_stage2_big(const HmCoreBlock *blocks, const int32_t *big, int64_t nbig,
                                 const int32_t *plist, const double *partial, const double *core,
                                 double *svec, int max_r, cudaStream_t st, bool pdl)
{
    if (nbig <= 0) return cudaSuccess;
    size_t smem = (size_t)9 * max_r * sizeof(double);
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
        {
            cudaError_t e = cudaFuncSetAttribute(hm_core_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            if (e != cudaSuccess) return e;
        }
    }
    return launch_k(hm_core_big_kernel, (unsigned)nbig, 256, smem, st, pdl, blocks, big, plist, partial, core, svec,
                    max_r);
}

cudaError_t hm_launch_stage2(const HmCoreBlock *blocks, int64_t nblocks, const int32_t *plist,
                             const double *partial, const double *core, double *svec, int max_r,
                             cudaStream_t st, bool pdl)
{
    if (nblocks <= 0) return cudaSuccess;
    int threads = 256;
    size_t smem = (size_t)(threads / 32) * (size_t)max_r * sizeof(double);
    while (smem > 48 * 1024 && threads > 32) {
        threads >>= 1;
        smem = (size_t)(threads / 32) * (size_t)max_r * sizeof(double);
    }
    if (smem > 48 * 1024) return cudaErrorInvalidConfiguration;
    int wpb = threads / 32;
    unsigned grid = (unsigned)((nblocks + wpb - 1) / wpb);
    // Two restructurings were measured in round 2 and dropped: a persistent kernel with two leaves in
    // flight per warp (0.204 vs 0.108 ms at N = 2^20: 16 warps per SM hide less latency than 32 one-leaf
    // warps), and the rank-20 core staged in shared memory by cp.async at 32 registers / 64 warps per SM
    // (0.110 vs 0.101 ms alone, and 0.76 ms slower per matvec inside a PDL graph: its early-launched CTAs
    // hold 215 KB of shared memory per SM while they wait for stage 1).
    return launch_k(hm_core_kernel, grid, (unsigned)threads, smem, st, pdl, blocks, nblocks, plist, partial, core,
                    svec, max_r);
}

cudaError_t hm_launch_stage3(const HmItem *items, int64_t nitems, const HmRun *runs,
                             const double *ustream, const double *x, const double *svec, double *y,
                             int accumulate, const HmPeers *peers, cudaStream_t st, bool pdl)
{
    if (nitems <= 0) return cudaSuccess;
    if (peers && peers->n > 0)
        return launch_k(hm_stream_kernel<true, false, true>, (unsigned)nitems, HM_THREADS, 0, st, pdl, items, runs,
                        ustream, x, svec, y, accumulate, HmFuse{}, *peers);
    return launch_k(hm_stream_kernel<true, false, false>, (unsigned)nitems, HM_THREADS, 0, st, pdl, items, runs,
                    ustream, x, svec, y, accumulate, HmFuse{}, HmPeers{});
}

cudaError_t hm_launch_core_cheb(const HmCoreBlock *blocks, const int32_t *core_leaf, int64_t nblocks,
                                const HmLeaf *leaves, double *core, const double *Cm, const double *Dm,
                                const HmCheb &cheb, cudaStream_t st)
{
    if (nblocks <= 0) return cudaSuccess;
    hm_core_cheb_kernel<<<(unsigned)nblocks, 128, 0, st>>>(blocks, core_leaf, leaves, core, Cm, Dm, cheb);
    return cudaGetLastError();
}

cudaError_t hm_launch_free1(const HmItem *items, int64_t nitems, const HmFreeEnt *ents, const double *py,
                            const double *x, double *partial, const HmCheb &cheb, int max_units, bool cheb_form,
                            cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    const size_t smem = (size_t)std::max(max_units, 1) * 20 * sizeof(double);
    if (smem > 160 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cheb_form ? cudaFuncSetAttribute(hm_free1_kernel<20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : cudaFuncSetAttribute(hm_free1_kernel<20, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (cheb_form)
        hm_free1_kernel<20, true><<<(unsigned)nitems, HM_FREE1_THREADS, smem, st>>>(items, ents, py, x, partial, cheb);
    else
        hm_free1_kernel<20, false><<<(unsigned)nitems, HM_FREE1_THREADS, smem, st>>>(items, ents, py, x, partial, cheb);
    return cudaGetLastError();
}

template <bool PEERS, bool CHEB>
static cudaError_t launch_free3(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun,
                                const double *px, const double *py, const double *x, const double *svec, double *y,
                                int accumulate, const HmCheb &cheb, int kernel_id, const HmPeers &pe, size_t smem,
                                cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(hm_free3_kernel<20, PEERS, CHEB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    hm_free3_kernel<20, PEERS, CHEB><<<(unsigned)nitems, HM_THREADS, smem, st>>>(items, runs, frun, px, py, x, svec, y,
                                                                               accumulate, cheb, kernel_id, pe);
    return cudaGetLastError();
}

cudaError_t hm_launch_free3(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun,
                            const double *px, const double *py, const double *x,
                            const double *svec, double *y, int accumulate, const HmCheb &cheb, int kernel_id,
                            const HmPeers *peers, int zcap, bool cheb_form, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    if (zcap < 2 || zcap > HM_SMAX) zcap = HM_SMAX;
    if (cheb_form) zcap = std::max(zcap, 4 * HM_THREADS); // the row-quad sums are combined through z's storage
    const size_t smem = (size_t)zcap * sizeof(double);
    const bool pe = peers && peers->n > 0;
    const HmPeers none{};
    if (cheb_form)
        return pe ? launch_free3<true, true>(items, nitems, runs, frun, px, py, x, svec, y, accumulate, cheb, kernel_id, *peers, smem, st)
                  : launch_free3<false, true>(items, nitems, runs, frun, px, py, x, svec, y, accumulate, cheb, kernel_id, none, smem, st);
    return pe ? launch_free3<true, false>(items, nitems, runs, frun, px, py, x, svec, y, accumulate, cheb, kernel_id, *peers, smem, st)
              : launch_free3<false, false>(items, nitems, runs, frun, px, py, x, svec, y, accumulate, cheb, kernel_id, none, smem, st);
}

cudaError_t hm_launch_fill3(const HmFill *fills, int64_t nfills, const HmLeaf *leaves, double *ustream,
                            const double *px, const double *py, const HmCheb &cheb, int kernel_id,
                            cudaStream_t st)
{
    if (nfills <= 0) return cudaSuccess;
    hm_fill3_kernel<20><<<(unsigned)nfills, 128, 0, st>>>(fills, leaves, ustream, px, py, cheb, kernel_id);
    return cudaGetLastError();
}

cudaError_t hm_launch_fill1(const HmFill *fills, int64_t nfills, const HmLeaf *leaves, double *vstream,
                            const double *py, const HmCheb &cheb, cudaStream_t st)
{
    if (nfills <= 0) return cudaSuccess;
    hm_fill1_kernel<20><<<(unsigned)nfills, 128, 0, st>>>(fills, leaves, vstream, py, cheb);
    return cudaGetLastError();
}

cudaError_t hm_launch_fillcore(const HmCoreBlock *blocks, const int32_t *core_leaf, int64_t nblocks,
                               const HmLeaf *leaves, double *core, const HmCheb &cheb, int kernel_id,
                               cudaStream_t st)
{
    if (nblocks <= 0) return cudaSuccess;
    hm_fillcore_kernel<<<(unsigned)nblocks, 128, 0, st>>>(blocks, core_leaf, leaves, core, cheb, kernel_id);
    return cudaGetLastError();
}
