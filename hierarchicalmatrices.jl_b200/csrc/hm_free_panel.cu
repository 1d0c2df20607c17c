// hm_free_panel.cu -- many right-hand sides on a matrix-free plan (hm_assemble_kernel_free in its
// Chebyshev form): Y[:, c] (+)= H X[:, c] for a panel of 16 / 32 / 64 columns
// with the operator evaluated on the fly and consumed by FP64 tensor-core MMAs
// (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA; tcgen05 has no f64 kind).
//
// The reference reaches several right-hand sides only by repeating its scalar leaf loops per
// column (/root/reference/src/HierarchicalMatrix.jl:9-12, src/KernelMatrix.jl:9-12, and the leaf
// methods src/algebra.jl:37-48, 243-277).  With a panel the cost of evaluating an entry is shared
// by all columns, so the operand that the stored path streams from HBM (8 bytes per entry) is
// instead *generated in registers in MMA fragment layout*:
//
//   stage 1  Pp[(leaf, q)][c] = sum_s T_q(eta_s) Xt[s][c]      moments of every leaf's columns
//            a lane evaluates the Chebyshev recurrence of one column point (1 FMA per entry), the
//            warp's 20 x 32 tile goes through a private shared-memory tile into A fragments
//   stage 2  unchanged (hm_panel.cu): Sp = (C F C') Pp per leaf
//   stage 3  Yt[i][c] (+)= sum_leaf sum_q T_q(xi_i) Sp[(leaf, q)][c] + sum_j K(x_i, y_j) Xt[j][c]
//            lane (g, t) of an 8 x 4 A fragment needs T_{4j+t}(xi_g), j = 0..4: the stride-4
//            recurrence T_{n+4} = 2 T_4 T_n - T_{|n-4|} seeded with (T_t, T_{4-t}) gives them with one
//            FMA each and no shared memory; dense entries are one reciprocal per lane and k-step.
//
// B fragments (rows of Xt / Sp) come through L1: the warps of a CTA read the same rows.  Both panels
// are kept fragment-major (hm_panel_blocked_index: 4 x 8 tiles in lane order), so a warp's fragment
// load is one contiguous 256-byte piece; Pp and Yt stay row-major with pitch CS.  Per MMA the kernels execute about two instructions (the stored
// panel kernels: 8-12), so the FP64 pipe, not instruction issue, is what bounds them.
// Deterministic: fixed k order per accumulator, k-groups combined in group order.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "hm_device.cuh"
#include "hm_kernels.cuh"

namespace {

constexpr int FT = 256;          // threads per CTA
constexpr int FW = FT / 32;      // warps
constexpr int TPITCH = 36;       // stage-1 tile pitch: 4 (mod 16) words -> conflict-free A fragments
constexpr int TROWS = 24;        // 20 ranks padded to three 8-row MMA blocks

// D(8x8) += A(8x4, row) * B(4x8, col): lane (g = lane / 4, t = lane % 4) holds A[g][t], B[t][g],
// D[g][2t], D[g][2t + 1]
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// stage 1.  Item = column segment [zoff, zoff + S) covered by nrun leaves (HmFreeEnt: box of the
// leaf's columns, point offset, offset of its 20 sums).  Unit = (leaf, half of the panel columns
// when CS = 64); a warp owns a unit and walks the segment in chunks of 32 points; items with fewer
// than 8 units split the chunks over warp groups, combined in group order.
// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(FT, 2)
hm_free1_panel_kernel(const HmItem *__restrict__ items, const HmFreeEnt *__restrict__ ents,
                      const double *__restrict__ py, const double *__restrict__ Xt, double *__restrict__ Pp)
{
    constexpr int R = 20, CS = NB * 8, NBW = NB > 4 ? 4 : NB, NCH = NB / NBW, CP = CS + 8;
    extern __shared__ __align__(16) double fsm[]; // [FW][TROWS][TPITCH] tiles, then the combine buffer
    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int S = it.S;
    const int nu = it.nrun * NCH;
    if (nu == 0) return;
    const int kgroups = nu >= FW ? 1 : FW / nu;
    const int ntw = nu >= FW ? FW : nu;
    const int kg = warp / ntw;
    const bool active = kg < kgroups;
    const int nchunks = (S + 31) >> 5;
    double *Tw = fsm + (size_t)warp * TROWS * TPITCH;
    for (int i = lane; i < (TROWS - R) * TPITCH; i += 32) Tw[R * TPITCH + i] = 0.0; // pad rows: finite
    __syncwarp();

    double acc[3][NBW][2]; // nu < FW: a warp has a single unit, whose sums outlive the loop for the combine
    int keep_e = 0, keep_ch = 0;

    for (int u = warp % ntw; u < nu && active; u += ntw) {
        const int e = u / NCH, ch = u - e * NCH;
        const HmFreeEnt en = ents[it.run0 + e];
        const double ih = __drcp_rn(en.half);
        const double *__restrict__ yc = py + en.yoff;
        // Xt is fragment-major (hm_panel_blocked_index): the lane's element of k-step j, column block n
        // of its chunk is xb[(row0 >> 2 + j) tiles rows + n tiles], one contiguous 256-byte piece per warp
        const int row00 = it.zoff + tig;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int n = 0; n < NBW; n++) acc[a][n][0] = acc[a][n][1] = 0.0;

        for (int c = kg; c < nchunks; c += kgroups) {
            const int s0 = c << 5;
            {
                // the lane's point: T_0 .. T_19 by the three-term recurrence, into the warp's tile
                const int s = s0 + lane;
                const double eta = s < S ? (yc[s] - en.mid) * ih : 0.0, two = eta + eta;
                double tm2 = 1.0, tm1 = eta;
                Tw[lane] = 1.0;
                Tw[TPITCH + lane] = eta;
#pragma unroll
                for (int k = 2; k < R; k++) {
                    const double tk = fma(two, tm1, -tm2);
                    Tw[k * TPITCH + lane] = tk;
                    tm2 = tm1;
                    tm1 = tk;
                }
            }
            __syncwarp();
            const bool full = s0 + 32 <= S;
            const int row0 = row00 + s0;
            const double *__restrict__ xb = Xt + ((size_t)(row0 >> 2) * NB + ch * NBW) * 32 + gid * 4 + (row0 & 3);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int s = s0 + 4 * j + tig;
                const bool v = full || s < S;
                double b[NBW];
#pragma unroll
                for (int n = 0; n < NBW; n++) b[n] = v ? __ldg(xb + (j * NB + n) * 32) : 0.0;
                const double a0 = Tw[gid * TPITCH + 4 * j + tig];
                const double a1 = Tw[(8 + gid) * TPITCH + 4 * j + tig];
                const double a2 = Tw[(16 + gid) * TPITCH + 4 * j + tig];
#pragma unroll
                for (int n = 0; n < NBW; n++) {
                    dmma884(acc[0][n][0], acc[0][n][1], a0, b[n]);
                    dmma884(acc[1][n][0], acc[1][n][1], a1, b[n]);
                    dmma884(acc[2][n][0], acc[2][n][1], a2, b[n]);
                }
            }
            __syncwarp(); // the tile is rewritten by the next chunk
        }
        if (kgroups == 1) {
            // sole owner of the leaf's sums: rows q = 8a + gid < 20
            double *o = Pp + (size_t)(it.out + en.fofs) * CS + ch * (NBW * 8) + 2 * tig;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int q = 8 * a + gid;
                if (q < R) {
#pragma unroll
                    for (int n = 0; n < NBW; n++)
                        *reinterpret_cast<double2 *>(o + (size_t)q * CS + n * 8) = make_double2(acc[a][n][0], acc[a][n][1]);
                }
            }
        } else {
            keep_e = e;
            keep_ch = ch;
        }
    }
    if (kgroups > 1) {
        // nu < 8: every active warp holds exactly one unit; sum the groups in group order
        __syncthreads(); // the tiles are no longer read: their storage becomes the combine buffer
        double *csm = fsm; // [nrun * 24][CP]
        for (int g = 0; g < kgroups; g++) {
            if (active && kg == g) {
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    double *row = csm + (size_t)(keep_e * TROWS + 8 * a + gid) * CP + keep_ch * (NBW * 8) + 2 * tig;
#pragma unroll
                    for (int n = 0; n < NBW; n++) {
                        if (g == 0) {
                            row[n * 8] = acc[a][n][0];
                            row[n * 8 + 1] = acc[a][n][1];
                        } else {
                            row[n * 8] += acc[a][n][0];
                            row[n * 8 + 1] += acc[a][n][1];
                        }
                    }
                }
            }
            __syncthreads();
        }
        constexpr int hz = CS / 2;
        for (int idx = t; idx < it.nrun * R * hz; idx += FT) {
            const int p = idx % hz, eq = idx / hz;
            const int e = eq / R, q = eq - e * R;
            const double2 v = *reinterpret_cast<const double2 *>(csm + (size_t)(e * TROWS + q) * CP + 2 * p);
            *reinterpret_cast<double2 *>(Pp + (size_t)(it.out + ents[it.run0 + e].fofs + q) * CS + 2 * p) = v;
        }
    }
}

// ---------------------------------------------------------------------------
// stage 3.  Item = row segment of F rows with its runs (low-rank: 20 rows of Sp and the box of the
// leaf's rows; dense: kn rows of Xt and the column points).  Unit = (16-row tile, half of the panel
// columns when CS = 64); a warp owns a unit and walks the runs; items with fewer than 8 units
// split the runs over warp groups, combined in group order.
//
// Everything that steers control flow (run kind, rank range, column count) is made warp-uniform in
// the compiler's eyes with a shuffle from lane 0, and the lane-dependent seeds of the recurrence
// are selected with selp, so that no MMA sits behind a possibly divergent branch (ptxas otherwise
// guards every mma.sync with WARPSYNC + NOP: three instructions per MMA instead of one).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double selp(double a, double b, bool p)
{
    double d;
    asm("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\nselp.f64 %0, %1, %2, q;\n}\n" : "=d"(d) : "d"(a), "d"(b), "r"((unsigned)p));
    return d;
}
__device__ __forceinline__ int uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }

template <int NB, int KID> // KID: kernel id 0..3 of the dense entries (hm_assemble_kernel)
__global__ void __launch_bounds__(FT, 3)
hm_free3_panel_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                      const HmFreeRun *__restrict__ frun, const double *__restrict__ px,
                      const double *__restrict__ py, const double *__restrict__ Xt,
                      const double *__restrict__ Sp, double *__restrict__ Yt, int accumulate)
{
    constexpr int R = 20, CS = NB * 8, NBW = NB > 4 ? 4 : NB, NCH = NB / NBW, CP = CS + 8;
    extern __shared__ __align__(16) double csm[]; // [<= 7 / NCH tiles][16][CP] combine buffer
    __shared__ int rsrc[HM_MAXRUNS];
    __shared__ int2 rk[HM_MAXRUNS];
    __shared__ double2 rbox[HM_MAXRUNS]; // low-rank: (mid, 1 / half)
    __shared__ int64_t rxo[HM_MAXRUNS], ryo[HM_MAXRUNS];

    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = uniform(t >> 5);
    const int gid = lane >> 2, tig = lane & 3;
    const int F = it.F;
    bool same = true; // do all runs address the same points for the item's rows? (always so for KernelMatrix)
    const int64_t xo0 = it.nrun > 0 ? frun[it.run0].xoff : 0;
    for (int r = t; r < it.nrun; r += FT) {
        const HmRun rr = runs[it.run0 + r];
        const HmFreeRun fr = frun[it.run0 + r];
        rsrc[r] = rr.src;
        rk[r] = make_int2(fr.k0, fr.kn);
        rbox[r] = make_double2(fr.mid, rr.src < 0 ? __drcp_rn(fr.half) : 0.0);
        rxo[r] = fr.xoff;
        ryo[r] = fr.yoff;
        same = same && fr.xoff == xo0;
    }
    const bool same_x = __syncthreads_and(same) != 0;
    const int nfb = (F + 7) >> 3, nft = (nfb + 1) >> 1;
    const int nu = nft * NCH;
    if (nu == 0) return;
    const int kgroups = nu >= FW ? 1 : FW / nu;
    const int ntw = nu >= FW ? FW : nu;
    const int kg = warp / ntw;
    const bool active = kg < kgroups;

    double acc[2][NBW][2]; // nu < FW: a warp has a single unit, whose sums outlive the loop for the combine
    int keep_ft = 0, keep_ch = 0;

    for (int u = warp % ntw; u < nu && active; u += ntw) {
        const int ft = u / NCH, ch = u - ft * NCH;
        const bool two_blocks = 2 * ft + 1 < nfb;
        const int i0 = min(ft * 16 + gid, F - 1), i1 = min(ft * 16 + 8 + gid, F - 1); // pad rows repeat the last one
        const int cb0 = ch * NBW; // first column block of the unit
        double p0 = 0.0, p1 = 0.0;
        if (same_x) {
            p0 = px[xo0 + i0];
            p1 = px[xo0 + i1];
        }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int n = 0; n < NBW; n++) acc[a][n][0] = acc[a][n][1] = 0.0;

        for (int r = kg; r < it.nrun; r += kgroups) {
            const int src = uniform(rsrc[r]);
            const int k0 = uniform(rk[r].x), kn = uniform(rk[r].y);
            if (!same_x) {
                const double *__restrict__ pp = px + rxo[r];
                p0 = pp[i0];
                p1 = pp[i1];
            }
            if (src < 0) {
                // low-rank run: rows [k0, k0 + kn) of the leaf's 20 coefficients are rows of Sp
                // (Sp and Xt are fragment-major: hm_panel_blocked_index)
                const int64_t row0 = (int64_t)(~src) - k0 + tig; // row of rank k = tig
                const double *__restrict__ bp = Sp + ((row0 >> 2) * NB + cb0) * 32 + gid * 4 + (row0 & 3);
                const bool whole = k0 == 0 && kn == R;
                const double2 box = rbox[r];
                double cur[2], prev[2], t4x2[2];
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const double xi = ((a ? p1 : p0) - box.x) * box.y, two = xi + xi;
                    const double T2 = fma(two, xi, -1.0), T3 = fma(two, T2, -xi), T4 = fma(two, T3, -T2);
                    // (cur, prev) = (T_t, T_{4-t}): t = 0: (1, T4), 1: (xi, T3), 2: (T2, T2), 3: (T3, xi)
                    cur[a] = selp(selp(T3, T2, tig & 1), selp(xi, 1.0, tig & 1), tig & 2);
                    prev[a] = selp(selp(xi, T2, tig & 1), selp(T3, T4, tig & 1), tig & 2);
                    t4x2[a] = T4 + T4;
                }
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    // (fragments loaded step by step: all five at once cost 40 registers and a third CTA per SM)
                    const int k = 4 * j + tig;
                    const bool v = whole || (k >= k0 && k < k0 + kn);
                    double b[NBW];
#pragma unroll
                    for (int n = 0; n < NBW; n++) b[n] = v ? __ldg(bp + (j * NB + n) * 32) : 0.0;
#pragma unroll
                    for (int n = 0; n < NBW; n++) {
                        dmma884(acc[0][n][0], acc[0][n][1], cur[0], b[n]);
                        if (two_blocks) dmma884(acc[1][n][0], acc[1][n][1], cur[1], b[n]);
                    }
                    if (j < 4) {
#pragma unroll
                        for (int a = 0; a < 2; a++) {
                            const double nx = fma(t4x2[a], cur[a], -prev[a]);
                            prev[a] = cur[a];
                            cur[a] = nx;
                        }
                    }
                }
            } else {
                // dense run: kn columns, entries K(x_i, y_j) evaluated by the lane that holds them
                const int64_t row0 = (int64_t)src + tig;
                const double *__restrict__ bp = Xt + ((row0 >> 2) * NB + cb0) * 32 + gid * 4 + (row0 & 3);
                const double *__restrict__ yc = py + ryo[r] + tig;
                int j0 = 0;
                if (kn >= 4) {
                    // the next step's column point and B fragments are in flight during this step's MMAs
                    double yv = yc[0];
                    double b[NBW];
#pragma unroll
                    for (int n = 0; n < NBW; n++) b[n] = __ldg(bp + n * 32);
                    for (; j0 + 4 <= kn; j0 += 4) {
                        const int jn = j0 + 8 <= kn ? j0 + 4 : j0; // (last full step: reload, unused)
                        const double yn = yc[jn];
                        double bn[NBW];
#pragma unroll
                        for (int n = 0; n < NBW; n++) bn[n] = __ldg(bp + ((jn >> 2) * NB + n) * 32);
                        const double a0 = kernel_eval_fast(KID, p0, yv);
                        const double a1 = kernel_eval_fast(KID, p1, yv);
#pragma unroll
                        for (int n = 0; n < NBW; n++) {
                            dmma884(acc[0][n][0], acc[0][n][1], a0, b[n]);
                            if (two_blocks) dmma884(acc[1][n][0], acc[1][n][1], a1, b[n]);
                        }
                        yv = yn;
#pragma unroll
                        for (int n = 0; n < NBW; n++) b[n] = bn[n];
                    }
                }
                if (j0 < kn) {
                    const bool v = j0 + tig < kn;
                    const double yv = v ? yc[j0] : 0.0;
                    double b[NBW];
#pragma unroll
                    for (int n = 0; n < NBW; n++) b[n] = v ? __ldg(bp + ((j0 >> 2) * NB + n) * 32) : 0.0;
                    const double a0 = selp(kernel_eval_fast(KID, p0, yv), 0.0, v);
                    const double a1 = selp(kernel_eval_fast(KID, p1, yv), 0.0, v);
#pragma unroll
                    for (int n = 0; n < NBW; n++) {
                        dmma884(acc[0][n][0], acc[0][n][1], a0, b[n]);
                        if (two_blocks) dmma884(acc[1][n][0], acc[1][n][1], a1, b[n]);
                    }
                }
            }
        }
        if (kgroups == 1) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int f = ft * 16 + a * 8 + gid;
                if (f < F) {
                    double2 *g = reinterpret_cast<double2 *>(Yt + (size_t)(it.out + f) * CS + ch * (NBW * 8)) + tig;
#pragma unroll
                    for (int n = 0; n < NBW; n++) {
                        double2 v = make_double2(acc[a][n][0], acc[a][n][1]);
                        if (accumulate) {
                            const double2 o = g[n * 4];
                            v.x += o.x;
                            v.y += o.y;
                        }
                        g[n * 4] = v;
                    }
                }
            }
        } else {
            keep_ft = ft;
            keep_ch = ch;
        }
    }
    if (kgroups > 1) {
        for (int g = 0; g < kgroups; g++) {
            if (active && kg == g) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    double *row = csm + (size_t)(keep_ft * 16 + a * 8 + gid) * CP + keep_ch * (NBW * 8) + 2 * tig;
#pragma unroll
                    for (int n = 0; n < NBW; n++) {
                        if (g == 0) {
                            row[n * 8] = acc[a][n][0];
                            row[n * 8 + 1] = acc[a][n][1];
                        } else {
                            row[n * 8] += acc[a][n][0];
                            row[n * 8 + 1] += acc[a][n][1];
                        }
                    }
                }
            }
            __syncthreads();
        }
        constexpr int hz = CS / 2;
        for (int idx = t; idx < F * hz; idx += FT) {
            const int r = idx / hz, p = idx - r * hz;
            double2 v = *reinterpret_cast<const double2 *>(csm + (size_t)r * CP + 2 * p);
            double2 *g = reinterpret_cast<double2 *>(Yt + (size_t)(it.out + r) * CS) + p;
            if (accumulate) {
                const double2 o = *g;
                v.x += o.x;
                v.y += o.y;
            }
            *g = v;
        }
    }
}

template <int NB>
cudaError_t launch1(const HmItem *items, int64_t nitems, const HmFreeEnt *ents, const double *py, const double *Xt,
                    double *Pp, cudaStream_t st)
{
    constexpr int CS = NB * 8, CP = CS + 8;
    // tiles of the eight warps; the combine buffer (<= 7 leaves x 24 rows) reuses the storage
    constexpr int NCH = NB > 4 ? 2 : 1;
    const size_t smem = std::max((size_t)FW * TROWS * TPITCH, (size_t)((FW - 1) / NCH) * TROWS * CP) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(hm_free1_panel_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hm_free1_panel_kernel<NB><<<(unsigned)nitems, FT, smem, st>>>(items, ents, py, Xt, Pp);
    return cudaGetLastError();
}

template <int NB, int KID>
cudaError_t launch3k(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun, const double *px,
                     const double *py, const double *Xt, const double *Sp, double *Yt, int accumulate, cudaStream_t st)
{
    constexpr int CS = NB * 8, CP = CS + 8;
    constexpr int NCH = NB > 4 ? 2 : 1;
    const size_t smem = (size_t)((FW - 1) / NCH) * 16 * CP * sizeof(double); // only items with < 8 units are combined
    cudaError_t e = cudaFuncSetAttribute(hm_free3_panel_kernel<NB, KID>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    hm_free3_panel_kernel<NB, KID><<<(unsigned)nitems, FT, smem, st>>>(items, runs, frun, px, py, Xt, Sp, Yt, accumulate);
    return cudaGetLastError();
}

template <int NB>
cudaError_t launch3(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun, const double *px,
                    const double *py, const double *Xt, const double *Sp, double *Yt, int accumulate, int kernel_id,
                    cudaStream_t st)
{
    switch (kernel_id) {
    case 0: return launch3k<NB, 0>(items, nitems, runs, frun, px, py, Xt, Sp, Yt, accumulate, st);
    case 1: return launch3k<NB, 1>(items, nitems, runs, frun, px, py, Xt, Sp, Yt, accumulate, st);
    case 2: return launch3k<NB, 2>(items, nitems, runs, frun, px, py, Xt, Sp, Yt, accumulate, st);
    case 3: return launch3k<NB, 3>(items, nitems, runs, frun, px, py, Xt, Sp, Yt, accumulate, st);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace

cudaError_t hm_launch_free1_panel(int CS, const HmItem *items, int64_t nitems, const HmFreeEnt *ents, const double *py,
                                  const double *Xt, double *Pp, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    switch (CS) {
    case 16: return launch1<2>(items, nitems, ents, py, Xt, Pp, st);
    case 32: return launch1<4>(items, nitems, ents, py, Xt, Pp, st);
    case 64: return launch1<8>(items, nitems, ents, py, Xt, Pp, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_free3_panel(int CS, const HmItem *items, int64_t nitems, const HmRun *runs,
                                  const HmFreeRun *frun, const double *px, const double *py, const double *Xt,
                                  const double *Sp, double *Yt, int accumulate, int kernel_id, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    switch (CS) {
    case 16: return launch3<2>(items, nitems, runs, frun, px, py, Xt, Sp, Yt, accumulate, kernel_id, st);
    case 32: return launch3<4>(items, nitems, runs, frun, px, py, Xt, Sp, Yt, accumulate, kernel_id, st);
    case 64: return launch3<8>(items, nitems, runs, frun, px, py, Xt, Sp, Yt, accumulate, kernel_id, st);
    default: return cudaErrorInvalidValue;
    }
}
