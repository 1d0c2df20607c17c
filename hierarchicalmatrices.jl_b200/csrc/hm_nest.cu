// hm_nest.cu -- kernels of the nested-basis form of a matrix-free plan (hm_nest.h): the low-rank
// part of y (+)= K x as one upward pass (moments), one pass over the leaves (cores) and one
// downward pass (coefficients, evaluation); the dense leaves are applied by hm_free3_kernel /
// hm_free3_panel_kernel on the item list without low-rank runs.
//
// A CTA owns a subtree of the box tree and walks it depth by depth (a warp per box, __syncthreads
// between depths).  Subtrees are cut by point count into tiers (hm_nest_host.cpp), one launch per tier:
// three launches per pass at N = 2^20.  Every box value has one writer and a fixed summation order:
// deterministic.
#include <cuda_runtime.h>

#include <cstdint>

#include "hm_device.cuh"
#include "hm_kernels.cuh"
#include "hm_nest.h"
#include "hm_nest_dev.cuh"

namespace {

constexpr int R = HM_NEST_R;
constexpr int NT = 256; // threads of a subtree CTA


// ---------------------------------------------------------------------------
// upward pass: MU[box][q] = sum over the box's columns of T_q(eta_s) x_s
//   finest boxes from the points (flat kernel, a warp per box), the others from their two halves
//   (M0, M1), subtree by subtree
// ---------------------------------------------------------------------------
// moments of one finest box from its points (a warp)
__device__ __forceinline__ void base_moments(const HmNestNode &nd, int id, int lane, const double *__restrict__ pts,
                                             const double *__restrict__ x, double *MU)
{
    double acc[R];
#pragma unroll
    for (int k = 0; k < R; k++) acc[k] = 0.0;
    const double *__restrict__ pp = pts + nd.p0;
    const double *__restrict__ xs = x + nd.p0;
    for (int s = lane; s < nd.np; s += 32) {
        const double eta = (pp[s] - nd.mid) * nd.ih, two = eta + eta, xv = xs[s];
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
    // lane sums in a fixed (butterfly) order; lane k keeps moment k
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < R; k++) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == k) mine = acc[k];
    }
    if (lane < R) MU[(size_t)id * R + lane] = mine;
}

__global__ void __launch_bounds__(NT)
hm_nest_base_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ base, int nbase,
                    const double *__restrict__ pts, const double *__restrict__ x, double *__restrict__ MU)
{
    hm_pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (b >= nbase) return;
    const int id = base[b];
    const HmNestNode nd = nodes[id];
    hm_pdl_wait(); // (x may be the previous kernel's output; MU may still be read by it)
    base_moments(nd, id, lane, pts, x, MU);
}

// BASE (measured slower than the separate flat launch above -- 71 vs 28 + 35 us at N = 2^20: subtrees with
// many finest boxes hold up their CTA -- and not used): the CTA first forms the moments of
// the subtree's finest boxes from the points, then translates upwards -- no separate base launch.
// NTH: 256 threads for the many small subtrees of the finest tier, 1024 for the few of the upper tiers,
// whose depths are latency-bound (a warp per box, so a level of 32 boxes is one round instead of four).
template <int NTH, bool BASE>
__global__ void __launch_bounds__(NTH)
hm_nest_up_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ order,
                  const int32_t *__restrict__ grp, const int32_t *__restrict__ sub_g0, int sub0,
                  const double *__restrict__ Mt, const double *__restrict__ pts, const double *__restrict__ x,
                  double *MU)
{
    __shared__ double sMt[2][R * R]; // transposed maps: [p][q]
    __shared__ HmSubSched S;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    const int sub = sub0 + blockIdx.x;
    const int g0 = sub_g0[sub], g1 = sub_g0[sub + 1];
    hm_pdl_launch_dependents();
    hm_stage_schedule<false>(S, nodes, order, grp, nullptr, g0, g1);
    for (int i = t; i < 2 * R * R; i += blockDim.x) sMt[0][i] = Mt[i];
    __syncthreads();
    hm_pdl_wait();
    const bool cached = S.cached;
    const int eb = cached ? S.grp[0] : 0;
    if (BASE) {
        const int ea = grp[g0], ez = grp[g1];
        for (int e = ea + warp; e < ez; e += nw) {
            const int id = cached ? S.id[e - eb] : order[e];
            const int c0 = cached ? S.aux[e - eb] : nodes[id].child0;
            if (c0 < 0) base_moments(nodes[id], id, lane, pts, x, MU); // (warp-uniform)
        }
        __syncthreads();
    }
    for (int g = g0; g < g1; g++) {
        const int e0 = cached ? S.grp[g - g0] : grp[g], e1 = cached ? S.grp[g - g0 + 1] : grp[g + 1];
        for (int e = e0 + warp; e < e1; e += nw) {
            const int id = cached ? S.id[e - eb] : order[e];
            const int c0 = cached ? S.aux[e - eb] : nodes[id].child0;
            if (c0 >= 0) { // (warp-uniform)
                // the 40 moments of the two halves in two coalesced loads, handed round by shuffles (40
                // broadcast loads are 40 memory requests; the pass is bound by requests in flight)
                const int q = min(lane, R - 1);
                const double *m0 = MU + (size_t)c0 * R;
                const double v0 = m0[q], v1 = m0[R + q];
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0; // four short chains instead of one of 40
#pragma unroll
                for (int p = 0; p < R; p += 2) {
                    s0 = fma(sMt[0][p * R + q], __shfl_sync(0xffffffffu, v0, p), s0);
                    s1 = fma(sMt[0][(p + 1) * R + q], __shfl_sync(0xffffffffu, v0, p + 1), s1);
                    s2 = fma(sMt[1][p * R + q], __shfl_sync(0xffffffffu, v1, p), s2);
                    s3 = fma(sMt[1][(p + 1) * R + q], __shfl_sync(0xffffffffu, v1, p + 1), s3);
                }
                if (lane < R) MU[(size_t)id * R + lane] = (s0 + s1) + (s2 + s3);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// the leaves: LAM[row box][q] = sum over the box's leaves of G_leaf[q][:] . MU[column box of the leaf]
// ---------------------------------------------------------------------------
// Lanes (r, j) = (lane / 10, lane % 10), r < 3: lane (r, j) multiplies the rows p = r, r + 3, r + 6, ... of the
// core (stored [p][q], q fastest: 160 bytes per row) with 16-byte loads of (G[p][2j], G[p][2j + 1]) -- one load
// instruction of the warp covers three rows, seven cover the core -- against mu[p], which the warp fetches with
// one coalesced load and hands round by shuffles: 8 memory requests per leaf instead of 40 (the kernel is bound
// by outstanding requests, not by bytes or flops).  The three partial sums of a coefficient pair are added in
// the order r = 0, 1, 2: deterministic.
__global__ void __launch_bounds__(NT)
hm_nest_core_kernel(int nboxes, const int32_t *__restrict__ rleaf_begin, const HmNestLeaf *__restrict__ rleaf,
                    const double *__restrict__ cores, const double *__restrict__ MU, double *__restrict__ LAM)
{
    static_assert(R == 20, "lane mapping of hm_nest_core_kernel is written for rank 20");
    hm_pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int box = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (box >= nboxes) return;
    const int l0 = rleaf_begin[box], l1 = rleaf_begin[box + 1];
    hm_pdl_wait();
    const int r = lane / 10, j = lane - 10 * r; // lanes 30, 31: r = 3, idle
    const bool act = r < 3;
    double ax = 0.0, ay = 0.0;
    for (int lb = l0; lb < l1; lb += 32) {
        const int nl = min(32, l1 - lb);
        HmNestLeaf mine{0, 0};
        if (lane < nl) mine = rleaf[lb + lane];
#pragma unroll 1
        for (int l = 0; l < nl; l++) {
            const int core = __shfl_sync(0xffffffffu, mine.core, l), cnode = __shfl_sync(0xffffffffu, mine.cnode, l);
            const double2 *__restrict__ G2 = reinterpret_cast<const double2 *>(cores + (size_t)core * (R * R)) + j;
            const double mv = MU[(size_t)cnode * R + min(lane, R - 1)];
            double2 g[7];
#pragma unroll
            for (int i = 0; i < 7; i++) {
                const int p = min(3 * i + (act ? r : 0), R - 1);
                g[i] = __ldg(G2 + p * (R / 2));
            }
#pragma unroll
            for (int i = 0; i < 7; i++) {
                const int p = 3 * i + r;
                const double m = __shfl_sync(0xffffffffu, mv, min(p, R - 1));
                if (act && p < R) {
                    ax = fma(g[i].x, m, ax);
                    ay = fma(g[i].y, m, ay);
                }
            }
        }
    }
    const double bx = __shfl_sync(0xffffffffu, ax, min(lane + 10, 31)), by = __shfl_sync(0xffffffffu, ay, min(lane + 10, 31));
    const double cx = __shfl_sync(0xffffffffu, ax, min(lane + 20, 31)), cy = __shfl_sync(0xffffffffu, ay, min(lane + 20, 31));
    if (lane < 10)
        reinterpret_cast<double2 *>(LAM + (size_t)box * R)[lane] = make_double2((ax + bx) + cx, (ay + by) + cy);
}

// ---------------------------------------------------------------------------
// downward pass: LAM[box] += M_which' LAM[parent]; at a finest box the series is evaluated at its
// rows (Clenshaw):  y_i = (accumulate ? y_i : 0) + sum_q LAM[box][q] T_q(xi_i)
// ---------------------------------------------------------------------------
template <int NTH, bool EVAL>
__global__ void __launch_bounds__(NTH)
hm_nest_down_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ order,
                    const int32_t *__restrict__ grp, const int32_t *__restrict__ sub_g0, int sub0,
                    const double *__restrict__ pts, const double *__restrict__ M, double *LAM, double *y,
                    int accumulate, int row_begin, int row_end)
{
    __shared__ double sM[2][R * R]; // maps: [q][p]
    __shared__ HmSubSched S;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    const int sub = sub0 + blockIdx.x;
    const int g0 = sub_g0[sub], g1 = sub_g0[sub + 1];
    hm_pdl_launch_dependents();
    hm_stage_schedule<true>(S, nodes, order, grp, nullptr, g0, g1);
    for (int i = t; i < 2 * R * R; i += blockDim.x) sM[0][i] = M[i];
    __syncthreads();
    hm_pdl_wait();
    const bool cached = S.cached;
    const int eb = cached ? S.grp[0] : 0;
    for (int g = g1 - 1; g >= g0; g--) { // shallowest depth first
        const int e0 = cached ? S.grp[g - g0] : grp[g], e1 = cached ? S.grp[g - g0 + 1] : grp[g + 1];
        for (int e = e0 + warp; e < e1; e += nw) {
            const int id = cached ? S.id[e - eb] : order[e];
            int pw;
            if (cached) {
                pw = S.aux[e - eb];
            } else {
                const int par = nodes[id].parent;
                pw = par >= 0 ? par * 2 + nodes[id].which : -1;
            }
            double *lam = LAM + (size_t)id * R;
            if (pw >= 0) { // (warp-uniform)
                const int ql = min(lane, R - 1);
                const double pv = LAM[(size_t)(pw >> 1) * R + ql]; // the parent's coefficients: one coalesced load
                const double *m = sM[pw & 1];
                double s0 = lam[ql], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int q = 0; q < R; q += 4) {
                    s0 = fma(m[q * R + ql], __shfl_sync(0xffffffffu, pv, q), s0);
                    s1 = fma(m[(q + 1) * R + ql], __shfl_sync(0xffffffffu, pv, q + 1), s1);
                    s2 = fma(m[(q + 2) * R + ql], __shfl_sync(0xffffffffu, pv, q + 2), s2);
                    s3 = fma(m[(q + 3) * R + ql], __shfl_sync(0xffffffffu, pv, q + 3), s3);
                }
                if (lane < R) lam[lane] = (s0 + s1) + (s2 + s3);
            }
            HmNestNode nd;
            if (EVAL) nd = nodes[id];
            if (EVAL && nd.child0 < 0) {
                __syncwarp();
                double cf[R];
#pragma unroll
                for (int k = 0; k < R; k++) cf[k] = lam[k];
                const double *__restrict__ pp = pts + nd.p0;
                for (int i = lane; i < nd.np; i += 32) {
                    const int row = nd.p0 + i;
                    if (row < row_begin || row >= row_end) continue;
                    const double xi = (pp[i] - nd.mid) * nd.ih, two = xi + xi;
                    double b1 = 0.0, b2 = 0.0;
#pragma unroll
                    for (int k = R - 1; k >= 1; k--) {
                        const double nb = fma(two, b1, cf[k]) - b2;
                        b2 = b1;
                        b1 = nb;
                    }
                    const double v = fma(xi, b1, cf[0]) - b2;
                    y[row] = (accumulate ? y[row] : 0.0) + v;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// evaluation alone: y_i (+)= sum_q LAM[box][q] T_q(xi_i) over the rows of every finest box (a warp per box).
// When the dense leaves run beside the tree passes, the translations of the finest tier need not wait for
// them -- only this short kernel does, so the tail after the dense kernel is the evaluation, not the tier.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
hm_nest_eval_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ base, int nbase,
                    const double *__restrict__ pts, const double *__restrict__ LAM, double *y, int accumulate,
                    int row_begin, int row_end)
{
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (b >= nbase) return;
    const int id = base[b];
    const HmNestNode nd = nodes[id];
    const double mine = LAM[(size_t)id * R + min(lane, R - 1)]; // the 20 coefficients: one coalesced load
    double cf[R];
#pragma unroll
    for (int k = 0; k < R; k++) cf[k] = __shfl_sync(0xffffffffu, mine, k);
    const double *__restrict__ pp = pts + nd.p0;
    for (int i = lane; i < nd.np; i += 32) {
        const int row = nd.p0 + i;
        if (row < row_begin || row >= row_end) continue;
        const double xi = (pp[i] - nd.mid) * nd.ih, two = xi + xi;
        double b1 = 0.0, b2 = 0.0;
#pragma unroll
        for (int k = R - 1; k >= 1; k--) {
            const double nb = fma(two, b1, cf[k]) - b2;
            b2 = b1;
            b1 = nb;
        }
        const double v = fma(xi, b1, cf[0]) - b2;
        y[row] = (accumulate ? y[row] : 0.0) + v;
    }
}

// ---------------------------------------------------------------------------
// dense leaves: y[rows of the item] += sum over its dense runs of K(x_i, y_j) v_j, entries evaluated on
// the fly (src/KernelMatrix.jl:57-60, src/algebra.jl:37-48).  A warp owns an item (a row segment of at
// most 128 rows, up to four rows per lane) and walks the columns of its runs; the column point and the
// vector entry are warp-uniform loads.  No shared memory, no block barrier: eight independent items
// per CTA keep their short dependent chains (item -> runs -> points) in flight together.
// ---------------------------------------------------------------------------
// ibox != nullptr: the low-rank part is evaluated here too -- the rows of item i lie in the finest row box
// ibox[i], whose coefficients LAM the downward pass has completed: y_i = (accumulate ? y_i : 0) +
// sum_q LAM[box][q] T_q(xi_i) + dense part, one pass over y.
template <int KID, int NRL> // NRL rows per lane
__device__ __forceinline__ void nest_dense_rows(const HmItem &it, int lane, const HmRun *__restrict__ runs,
                                                const HmFreeRun *__restrict__ frun, const double *__restrict__ px,
                                                const double *__restrict__ py, const double *__restrict__ x,
                                                const int32_t *__restrict__ ibox, int idx,
                                                const HmNestNode *__restrict__ nodes, const double *__restrict__ LAM,
                                                double (&out)[4])
{
    const int F = it.F;
    double acc[NRL][2];
#pragma unroll
    for (int k = 0; k < NRL; k++) acc[k][0] = acc[k][1] = 0.0;
    if (ibox) {
        const int box = ibox[idx];
        const HmNestNode nd = nodes[box];
        const double *__restrict__ lam = LAM + (size_t)box * R;
        double cf[R];
#pragma unroll
        for (int k = 0; k < R; k++) cf[k] = lam[k];
        const double *__restrict__ pp = px + it.out; // rows are numbered like the points
#pragma unroll
        for (int k = 0; k < NRL; k++) {
            const double xi = (pp[min(lane + 32 * k, F - 1)] - nd.mid) * nd.ih, two = xi + xi;
            double b1 = 0.0, b2 = 0.0;
#pragma unroll
            for (int q = R - 1; q >= 1; q--) {
                const double nb = fma(two, b1, cf[q]) - b2;
                b2 = b1;
                b1 = nb;
            }
            acc[k][0] = fma(xi, b1, cf[0]) - b2;
        }
    }
    int64_t xo_prev = -1;
    double p[NRL];
#pragma unroll
    for (int k = 0; k < NRL; k++) p[k] = 0.0;
    for (int r = 0; r < it.nrun; r++) {
        const HmRun rr = runs[it.run0 + r];
        const HmFreeRun fr = frun[it.run0 + r];
        if (fr.xoff != xo_prev) { // the rows' points (the same for every run of a KernelMatrix item)
            xo_prev = fr.xoff;
#pragma unroll
            for (int k = 0; k < NRL; k++) p[k] = px[fr.xoff + min(lane + 32 * k, F - 1)];
        }
        const double *__restrict__ yc = py + fr.yoff;
        const double *__restrict__ xs = x + rr.src;
        const int kn = fr.kn;
        // two columns per step, the next two already in flight (warp-uniform loads)
        double y0 = 0.0, y1 = 0.0, v0 = 0.0, v1 = 0.0;
        if (kn > 0) {
            y0 = yc[0];
            v0 = xs[0];
        }
        if (kn > 1) {
            y1 = yc[1];
            v1 = xs[1];
        }
        int j = 0;
        for (; j + 1 < kn; j += 2) {
            const int jn0 = min(j + 2, kn - 1), jn1 = min(j + 3, kn - 1);
            const double ny0 = yc[jn0], nv0 = xs[jn0], ny1 = yc[jn1], nv1 = xs[jn1];
#pragma unroll
            for (int k = 0; k < NRL; k++) {
                acc[k][0] = fma(kernel_eval_fast(KID, p[k], y0), v0, acc[k][0]);
                acc[k][1] = fma(kernel_eval_fast(KID, p[k], y1), v1, acc[k][1]);
            }
            y0 = ny0;
            v0 = nv0;
            y1 = ny1;
            v1 = nv1;
        }
        if (j < kn) {
#pragma unroll
            for (int k = 0; k < NRL; k++) acc[k][0] = fma(kernel_eval_fast(KID, p[k], y0), v0, acc[k][0]);
        }
    }
#pragma unroll
    for (int k = 0; k < NRL; k++) out[k] = acc[k][0] + acc[k][1];
}

// The same with 16-lane groups: lane (h, l) = (lane >> 4, lane & 15) holds rows l, l + 16, ... and the two halves
// of the warp split the columns of every run (h takes columns h, h + 2, ...).  Used when it wastes fewer lanes:
// a segment of 40 rows occupies 48 row slots this way instead of 64, one of 70 rows 80 instead of 96 (segments of
// graded point sets have 40 .. 79 rows).  The halves' sums are added in the order h = 0, 1.
template <int KID, int NRH> // NRH = ceil(F / 16) rows per lane
__device__ __forceinline__ void nest_dense_rows16(const HmItem &it, int lane, const HmRun *__restrict__ runs,
                                                  const HmFreeRun *__restrict__ frun, const double *__restrict__ px,
                                                  const double *__restrict__ py, const double *__restrict__ x,
                                                  const int32_t *__restrict__ ibox, int idx,
                                                  const HmNestNode *__restrict__ nodes, const double *__restrict__ LAM,
                                                  double (&out)[8])
{
    const int F = it.F, h = lane >> 4, hl = lane & 15;
    double acc[NRH][2];
#pragma unroll
    for (int k = 0; k < NRH; k++) acc[k][0] = acc[k][1] = 0.0;
    if (ibox) { // (warp-uniform) the series of the segment's finest box, counted once: by half 0
        const int box = ibox[idx];
        const HmNestNode nd = nodes[box];
        const double *__restrict__ lam = LAM + (size_t)box * R;
        double cf[R];
#pragma unroll
        for (int k = 0; k < R; k++) cf[k] = lam[k];
        const double *__restrict__ pp = px + it.out;
#pragma unroll
        for (int k = 0; k < NRH; k++) {
            const double xi = (pp[min(hl + 16 * k, F - 1)] - nd.mid) * nd.ih, two = xi + xi;
            double b1 = 0.0, b2 = 0.0;
#pragma unroll
            for (int q = R - 1; q >= 1; q--) {
                const double nb = fma(two, b1, cf[q]) - b2;
                b2 = b1;
                b1 = nb;
            }
            acc[k][0] = h == 0 ? fma(xi, b1, cf[0]) - b2 : 0.0;
        }
    }
    int64_t xo_prev = -1;
    double p[NRH];
#pragma unroll
    for (int k = 0; k < NRH; k++) p[k] = 0.0;
    for (int r = 0; r < it.nrun; r++) {
        const HmRun rr = runs[it.run0 + r];
        const HmFreeRun fr = frun[it.run0 + r];
        if (fr.xoff != xo_prev) {
            xo_prev = fr.xoff;
#pragma unroll
            for (int k = 0; k < NRH; k++) p[k] = px[fr.xoff + min(hl + 16 * k, F - 1)];
        }
        const double *__restrict__ yc = py + fr.yoff;
        const double *__restrict__ xs = x + rr.src;
        const int kn = fr.kn;
        // this half's columns h, h + 2, ...: two per step (an odd one out gets a zero vector entry)
        for (int j = h; j < kn; j += 4) {
            const int j1 = j + 2 < kn ? j + 2 : j;
            const double y0 = yc[j], v0 = xs[j], y1 = yc[j1], v1 = j + 2 < kn ? xs[j1] : 0.0;
#pragma unroll
            for (int k = 0; k < NRH; k++) {
                acc[k][0] = fma(kernel_eval_fast(KID, p[k], y0), v0, acc[k][0]);
                acc[k][1] = fma(kernel_eval_fast(KID, p[k], y1), v1, acc[k][1]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NRH; k++) {
        const double mine = acc[k][0] + acc[k][1];
        const double other = __shfl_xor_sync(0xffffffffu, mine, 16);
        out[k] = h == 0 ? mine + other : other + mine;
    }
}

template <int KID, bool PEERS>
__global__ void __launch_bounds__(NT)
hm_nest_dense_kernel(const HmItem *__restrict__ items, int nitems, const HmRun *__restrict__ runs,
                     const HmFreeRun *__restrict__ frun, const double *__restrict__ px,
                     const double *__restrict__ py, const double *__restrict__ x, double *y, int accumulate,
                     const int32_t *__restrict__ ibox, const HmNestNode *__restrict__ nodes,
                     const double *__restrict__ LAM, HmPeers pe)
{
    const int lane = threadIdx.x & 31;
    const int idx = __shfl_sync(0xffffffffu, blockIdx.x * (NT / 32) + (threadIdx.x >> 5), 0);
    if (idx >= nitems) return;
    const HmItem it = items[idx];
    const int F = it.F;
    double out[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int nrl = __shfl_sync(0xffffffffu, (F + 31) >> 5, 0); // warp-uniform, and the compiler knows it
    const int nrh = __shfl_sync(0xffffffffu, (F + 15) >> 4, 0);
    // 16-lane groups occupy fewer row slots when F mod 32 is in 1 .. 16 (33 .. 48 and 65 .. 80 rows; longer segments
    // would cost registers and are rare)
    const bool half = nrh == 3 || nrh == 5;
    if (half) {
        switch (nrh) {
        case 3: nest_dense_rows16<KID, 3>(it, lane, runs, frun, px, py, x, ibox, idx, nodes, LAM, out); break;
        default: nest_dense_rows16<KID, 5>(it, lane, runs, frun, px, py, x, ibox, idx, nodes, LAM, out); break;
        }
    } else {
        double o4[4] = {0.0, 0.0, 0.0, 0.0};
        switch (nrl) {
        case 1: nest_dense_rows<KID, 1>(it, lane, runs, frun, px, py, x, ibox, idx, nodes, LAM, o4); break;
        case 2: nest_dense_rows<KID, 2>(it, lane, runs, frun, px, py, x, ibox, idx, nodes, LAM, o4); break;
        case 3: nest_dense_rows<KID, 3>(it, lane, runs, frun, px, py, x, ibox, idx, nodes, LAM, o4); break;
        default: nest_dense_rows<KID, 4>(it, lane, runs, frun, px, py, x, ibox, idx, nodes, LAM, o4); break;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] = o4[k];
    }
    const int stride = half ? 16 : 32, first = half ? (lane & 15) : lane;
    const bool writer = !half || lane < 16;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int f = first + stride * k;
        if (writer && f < F && (half || k < 4)) {
            double *o = y + it.out + f;
            const double v = (accumulate ? *o : 0.0) + out[k];
            if (PEERS) {
                for (int q = 0; q < pe.n; q++) pe.y[q][it.out + f] = v;
            } else {
                *o = v;
            }
        }
    }
}

template <int KID>
cudaError_t launch_dense(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun, const double *px,
                         const double *py, const double *x, double *y, int accumulate, const int32_t *ibox,
                         const HmNestNode *nodes, const double *LAM, const HmPeers *peers, cudaStream_t st)
{
    const unsigned grid = (unsigned)((nitems + NT / 32 - 1) / (NT / 32));
    const HmPeers none{};
    if (peers && peers->n > 0)
        hm_nest_dense_kernel<KID, true><<<grid, NT, 0, st>>>(items, (int)nitems, runs, frun, px, py, x, y, accumulate, ibox,
                                                             nodes, LAM, *peers);
    else
        hm_nest_dense_kernel<KID, false><<<grid, NT, 0, st>>>(items, (int)nitems, runs, frun, px, py, x, y, accumulate, ibox,
                                                              nodes, LAM, none);
    return cudaGetLastError();
}

} // namespace

cudaError_t hm_launch_nest_dense(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun,
                                 const double *px, const double *py, const double *x, double *y, int accumulate,
                                 int kernel_id, const int32_t *ibox, const HmNestNode *nodes, const double *LAM,
                                 const HmPeers *peers, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    switch (kernel_id) {
    case 0: return launch_dense<0>(items, nitems, runs, frun, px, py, x, y, accumulate, ibox, nodes, LAM, peers, st);
    case 1: return launch_dense<1>(items, nitems, runs, frun, px, py, x, y, accumulate, ibox, nodes, LAM, peers, st);
    case 2: return launch_dense<2>(items, nitems, runs, frun, px, py, x, y, accumulate, ibox, nodes, LAM, peers, st);
    case 3: return launch_dense<3>(items, nitems, runs, frun, px, py, x, y, accumulate, ibox, nodes, LAM, peers, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_nest_up(const HmNestDev &T, const double *pts, const double *x, const double *Mt, double *MU,
                              cudaStream_t st)
{
    if (T.nbase > 0) {
        cudaError_t e = hm_launch_pdl(hm_nest_base_kernel, (unsigned)((T.nbase + NT / 32 - 1) / (NT / 32)), NT, st, T.nodes,
                                      T.base, T.nbase, pts, x, MU);
        if (e != cudaSuccess) return e;
    }
    constexpr int NTU = 512; // (16 K registers: fits beside three resident CTAs of the dense kernel)
    for (int k = 0; k < T.ntiers; k++) { // finest tier first
        const int n = T.tier_sub0[k + 1] - T.tier_sub0[k];
        if (n <= 0) continue;
        cudaError_t e = k == 0 ? hm_launch_pdl(hm_nest_up_kernel<NT, false>, (unsigned)n, NT, st, T.nodes, T.order, T.grp,
                                               T.sub_g0, T.tier_sub0[k], Mt, pts, x, MU)
                               : hm_launch_pdl(hm_nest_up_kernel<NTU, false>, (unsigned)n, NTU, st, T.nodes, T.order, T.grp,
                                               T.sub_g0, T.tier_sub0[k], Mt, pts, x, MU);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t hm_launch_nest_core(int nboxes, const int32_t *rleaf_begin, const HmNestLeaf *rleaf, const double *cores,
                                const double *MU, double *LAM, cudaStream_t st)
{
    if (nboxes <= 0) return cudaSuccess;
    return hm_launch_pdl(hm_nest_core_kernel, (unsigned)((nboxes + NT / 32 - 1) / (NT / 32)), NT, st, nboxes, rleaf_begin,
                         rleaf, cores, MU, LAM);
}

// eval: the finest boxes also evaluate their series into y (otherwise hm_launch_nest_dense does, fused
// with the dense leaves)
cudaError_t hm_launch_nest_down(const HmNestDev &T, const double *pts, const double *M, double *LAM, double *y,
                                int accumulate, int64_t row_begin, int64_t row_end, bool eval, cudaStream_t st,
                                cudaEvent_t before_finest)
{
    constexpr int NTU = 512; // (16 K registers: fits beside three resident CTAs of the dense kernel)
    for (int k = T.ntiers - 1; k >= 0; k--) { // coarsest tier first
        const int n = T.tier_sub0[k + 1] - T.tier_sub0[k];
        if (k == 0 && before_finest) { // the finest tier evaluates into y: after whoever else writes y
            cudaError_t e = cudaStreamWaitEvent(st, before_finest, 0);
            if (e != cudaSuccess) return e;
        }
        if (n <= 0) continue;
        const int sub0 = T.tier_sub0[k];
        cudaError_t e;
        if (k > 0) // (no finest box up here: nothing to evaluate)
            e = hm_launch_pdl(hm_nest_down_kernel<NTU, false>, (unsigned)n, NTU, st, T.nodes, T.order, T.grp, T.sub_g0, sub0, pts, M,
                              LAM, y, accumulate, (int)row_begin, (int)row_end);
        else if (eval)
            e = hm_launch_pdl(hm_nest_down_kernel<NT, true>, (unsigned)n, NT, st, T.nodes, T.order, T.grp, T.sub_g0, sub0, pts, M,
                              LAM, y, accumulate, (int)row_begin, (int)row_end);
        else
            e = hm_launch_pdl(hm_nest_down_kernel<NT, false>, (unsigned)n, NT, st, T.nodes, T.order, T.grp, T.sub_g0, sub0, pts, M,
                              LAM, y, accumulate, (int)row_begin, (int)row_end);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t hm_launch_nest_eval(const HmNestDev &T, const double *pts, const double *LAM, double *y, int accumulate,
                                int64_t row_begin, int64_t row_end, cudaStream_t st)
{
    if (T.nbase <= 0) return cudaSuccess;
    hm_nest_eval_kernel<<<(unsigned)((T.nbase + NT / 32 - 1) / (NT / 32)), NT, 0, st>>>(T.nodes, T.base, T.nbase, pts, LAM, y,
                                                                                     accumulate, (int)row_begin, (int)row_end);
    return cudaGetLastError();
}
