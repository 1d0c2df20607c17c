// hm_nest_panel.cu -- many right-hand sides on the nested-basis form (hm_nest.h): the passes of hm_nest.cu
// applied to a panel of CS = 16 / 32 / 64 columns.
//
//   MUp[box][q][c], LAMp[box][q][c]   row-major panels of 20 x CS words per box
//   base   MUp[finest column box] = T(eta)' Xt[columns of the box]      FP64 tensor cores (DMMA), the 20 x 32
//          Chebyshev tile generated per chunk of 32 points into a warp-private shared-memory tile
//   up     MUp[box] = [M0 M1] [MUp[half 0]; MUp[half 1]]                DMMA; the maps are triangular, their
//   down   LAMp[box] += M_which' LAMp[parent]                           zero k-steps skipped
//   cores  LAMp[row box] = sum over its leaves of G_leaf MUp[column box of the leaf]           DMMA
//   the finest row boxes leave their coefficients fragment-major in Sp, where the panel kernel of the dense
//   leaves (hm_free3_panel_kernel, hm_free_panel.cu) picks them up as one 20-term "low-rank run" per item
//   and evaluates them together with the dense entries.
#include <cuda_runtime.h>

#include <cstdint>

#include "hm_kernels.cuh"
#include "hm_nest.h"
#include "hm_nest_dev.cuh"

namespace {

constexpr int R = HM_NEST_R;
constexpr int NT = 256;
constexpr int TPITCH = 36; // tile pitch of the base kernel: 4 (mod 16) words -> conflict-free A fragments
constexpr int TROWS = 24;  // 20 moments padded to three 8-row MMA blocks

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// base: a warp owns (finest box, half of the panel columns when CS = 64)
// ---------------------------------------------------------------------------
constexpr int NTB = 128; // threads of a base-kernel CTA (four warp-private tiles: 27 KB of shared memory)

template <int NB>
__global__ void __launch_bounds__(NTB, 4)
hm_nest_base_panel_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ base, int nbase,
                          const double *__restrict__ pts, const double *__restrict__ Xt, double *__restrict__ MUp)
{
    constexpr int CS = NB * 8, NBW = NB > 4 ? 4 : NB, NCH = NB / NBW;
    __shared__ double tiles[NTB / 32][TROWS][TPITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int u = __shfl_sync(0xffffffffu, blockIdx.x * (NTB / 32) + warp, 0);
    if (u >= nbase * NCH) return;
    const int b = u / NCH, ch = u - b * NCH;
    const int id = base[b];
    const HmNestNode nd = nodes[id];
    double(*Tw)[TPITCH] = tiles[warp];
    for (int i = lane; i < (TROWS - R) * TPITCH; i += 32) Tw[R][i] = 0.0; // pad rows: finite
    double acc[3][NBW][2];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int n = 0; n < NBW; n++) acc[a][n][0] = acc[a][n][1] = 0.0;
    const double *__restrict__ pp = pts + nd.p0;
    const int S = nd.np;
    for (int s0 = 0; s0 < S; s0 += 32) {
        __syncwarp();
        {
            const int s = s0 + lane;
            const double eta = s < S ? (pp[s] - nd.mid) * nd.ih : 0.0, two = eta + eta;
            double tm2 = 1.0, tm1 = eta;
            Tw[0][lane] = 1.0;
            Tw[1][lane] = eta;
#pragma unroll
            for (int k = 2; k < R; k++) {
                const double tk = fma(two, tm1, -tm2);
                Tw[k][lane] = tk;
                tm2 = tm1;
                tm1 = tk;
            }
        }
        __syncwarp();
        const int row0 = nd.p0 + s0 + tig; // Xt is fragment-major (hm_panel_blocked_index)
        const double *__restrict__ xb = Xt + ((size_t)(row0 >> 2) * NB + ch * NBW) * 32 + gid * 4 + (row0 & 3);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const bool v = s0 + 4 * j + tig < S;
            double bf[NBW];
#pragma unroll
            for (int n = 0; n < NBW; n++) bf[n] = v ? __ldg(xb + (j * NB + n) * 32) : 0.0;
            const double a0 = Tw[gid][4 * j + tig], a1 = Tw[8 + gid][4 * j + tig], a2 = Tw[16 + gid][4 * j + tig];
#pragma unroll
            for (int n = 0; n < NBW; n++) {
                dmma884(acc[0][n][0], acc[0][n][1], a0, bf[n]);
                dmma884(acc[1][n][0], acc[1][n][1], a1, bf[n]);
                dmma884(acc[2][n][0], acc[2][n][1], a2, bf[n]);
            }
        }
    }
    double *o = MUp + (size_t)id * R * CS + ch * (NBW * 8) + 2 * tig;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int q = 8 * a + gid;
        if (q < R) {
#pragma unroll
            for (int n = 0; n < NBW; n++)
                *reinterpret_cast<double2 *>(o + (size_t)q * CS + n * 8) = make_double2(acc[a][n][0], acc[a][n][1]);
        }
    }
}

// ---------------------------------------------------------------------------
// up / down over the subtree schedule of hm_nest_host.cpp, on the FP64 tensor cores: a warp owns
// (box, 32 columns).  The transfer maps are the A operands, staged once per CTA in shared memory with
// pitches of 4 (mod 16) words (conflict-free m8n8k4 fragment reads); they are triangular, so the 4-wide
// k-steps that lie entirely in the zero part are skipped at compile time (22 of 30 remain going up,
// 9 of 15 going down).  B fragments are rows of the children's / the parent's panel through L1 (written
// by this CTA or by an earlier launch).  The first version kept lane = column with 20 + 20 FMA
// registers and broadcast the maps from shared memory: one LDS.128 per two DFMAs, 128 registers with
// spills, 0.58 ms per 64-column pass over the column tree at N = 2^20.
// ---------------------------------------------------------------------------
constexpr int UPITCH = 44; // [q][20 w + p], 40 + 4

template <int NB, int NTH> // NTH: 256 threads in the finest tier, 512 in the few latency-bound subtrees above
__global__ void __launch_bounds__(NTH)
hm_nest_up_panel_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ order,
                        const int32_t *__restrict__ grp, const int32_t *__restrict__ sub_g0, int sub0,
                        const double *__restrict__ M, double *MUp)
{
    constexpr int CS = NB * 8, NBW = NB > 4 ? 4 : NB, NCH = NB / NBW;
    __shared__ double sA[TROWS][UPITCH]; // MUp[box][q] = sum_w sum_p M_w[q][p] MUp[half w][p]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    for (int i = t; i < TROWS * UPITCH; i += blockDim.x) {
        const int q = i / UPITCH, k = i - q * UPITCH;
        sA[q][k] = (q < R && k < 2 * R) ? M[(k / R) * (R * R) + q * R + (k % R)] : 0.0;
    }
    __shared__ HmSubSched S;
    const int sub = sub0 + blockIdx.x;
    const int g0 = sub_g0[sub], g1 = sub_g0[sub + 1];
    hm_stage_schedule<false>(S, nodes, order, grp, nullptr, g0, g1);
    __syncthreads();
    const bool cached = S.cached;
    const int eb = cached ? S.grp[0] : 0;
    for (int g = g0; g < g1; g++) {
        const int e0 = cached ? S.grp[g - g0] : grp[g];
        const int nu = ((cached ? S.grp[g - g0 + 1] : grp[g + 1]) - e0) * NCH;
        for (int u = warp; u < nu; u += nw) {
            const int e = e0 + u / NCH, ch = u % NCH;
            const int id = cached ? S.id[e - eb] : order[e];
            const int c0 = cached ? S.aux[e - eb] : nodes[id].child0;
            if (c0 < 0) continue; // (warp-uniform)
            double acc[3][NBW][2];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int n = 0; n < NBW; n++) acc[a][n][0] = acc[a][n][1] = 0.0;
            const double *B = MUp + ((size_t)c0 * R + tig) * CS + ch * (NBW * 8) + gid; // 40 rows: both halves
#pragma unroll
            for (int w = 0; w < 2; w++) {
                // the five B fragment rows of one half in flight together (one L2 round trip, not five)
                double bf[R / 4][NBW];
#pragma unroll
                for (int jj = 0; jj < R / 4; jj++)
#pragma unroll
                    for (int n = 0; n < NBW; n++) bf[jj][n] = B[(size_t)(R * w + 4 * jj) * CS + n * 8];
#pragma unroll
                for (int jj = 0; jj < R / 4; jj++) { // k-step inside its map: columns p = 4 jj .. 4 jj + 3
                    const int k = R * w + 4 * jj + tig;
                    // row block a (q = 8 a ..) meets columns p <= q: active iff 4 jj <= 8 a + 7
                    if (jj <= 1) {
                        const double a0 = sA[gid][k];
#pragma unroll
                        for (int n = 0; n < NBW; n++) dmma884(acc[0][n][0], acc[0][n][1], a0, bf[jj][n]);
                    }
                    if (jj <= 3) {
                        const double a1 = sA[8 + gid][k];
#pragma unroll
                        for (int n = 0; n < NBW; n++) dmma884(acc[1][n][0], acc[1][n][1], a1, bf[jj][n]);
                    }
                    {
                        const double a2 = sA[16 + gid][k];
#pragma unroll
                        for (int n = 0; n < NBW; n++) dmma884(acc[2][n][0], acc[2][n][1], a2, bf[jj][n]);
                    }
                }
            }
            double *o = MUp + (size_t)id * R * CS + ch * (NBW * 8) + 2 * tig;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int q = 8 * a + gid;
                if (q < R) {
#pragma unroll
                    for (int n = 0; n < NBW; n++)
                        *reinterpret_cast<double2 *>(o + (size_t)q * CS + n * 8) = make_double2(acc[a][n][0], acc[a][n][1]);
                }
            }
        }
        __syncthreads();
    }
}

// fin[box] >= 0: a finest box; its completed coefficients go fragment-major into Sp at rows 20 fin[box] ..
template <int NB, int NTH>
__global__ void __launch_bounds__(NTH)
hm_nest_down_panel_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ order,
                          const int32_t *__restrict__ grp, const int32_t *__restrict__ sub_g0, int sub0,
                          const int32_t *__restrict__ fin, const double *__restrict__ M, double *LAMp,
                          double *__restrict__ Sp)
{
    constexpr int CS = NB * 8, NBW = NB > 4 ? 4 : NB, NCH = NB / NBW;
    __shared__ double sA[2][TROWS][R]; // LAMp[box][p] += sum_q M_which[q][p] LAMp[parent][q]: A = M_which'
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    for (int i = t; i < 2 * TROWS * R; i += blockDim.x) {
        const int w = i / (TROWS * R), pq = i - w * (TROWS * R), pr = pq / R, q = pq - pr * R;
        sA[w][pr][q] = pr < R ? M[w * (R * R) + q * R + pr] : 0.0;
    }
    __shared__ HmSubSched S;
    const int sub = sub0 + blockIdx.x;
    const int g0 = sub_g0[sub], g1 = sub_g0[sub + 1];
    hm_stage_schedule<true>(S, nodes, order, grp, fin, g0, g1);
    __syncthreads();
    const bool cached = S.cached;
    const int eb = cached ? S.grp[0] : 0;
    for (int g = g1 - 1; g >= g0; g--) { // shallowest depth first
        const int e0 = cached ? S.grp[g - g0] : grp[g];
        const int nu = ((cached ? S.grp[g - g0 + 1] : grp[g + 1]) - e0) * NCH;
        for (int u = warp; u < nu; u += nw) {
            const int e = e0 + u / NCH, ch = u % NCH;
            const int id = cached ? S.id[e - eb] : order[e];
            int pw, f;
            if (cached) {
                pw = S.aux[e - eb];
                f = S.fin[e - eb];
            } else {
                const int par = nodes[id].parent;
                pw = par >= 0 ? par * 2 + nodes[id].which : -1;
                f = fin[id];
            }
            double acc[3][NBW][2];
            double *o = LAMp + (size_t)id * R * CS + ch * (NBW * 8) + 2 * tig;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int pr = min(8 * a + gid, R - 1);
#pragma unroll
                for (int n = 0; n < NBW; n++) {
                    const double2 v = *reinterpret_cast<const double2 *>(o + (size_t)pr * CS + n * 8);
                    acc[a][n][0] = v.x;
                    acc[a][n][1] = v.y;
                }
            }
            if (pw >= 0) { // (warp-uniform)
                const double *B = LAMp + ((size_t)(pw >> 1) * R + tig) * CS + ch * (NBW * 8) + gid;
                const double(*A)[R] = sA[pw & 1];
                double bf[R / 4][NBW];
#pragma unroll
                for (int j = 0; j < R / 4; j++)
#pragma unroll
                    for (int n = 0; n < NBW; n++) bf[j][n] = B[(size_t)(4 * j) * CS + n * 8];
#pragma unroll
                for (int j = 0; j < R / 4; j++) {
                    // row block a (p = 8 a ..) meets q >= p: active iff 4 j + 3 >= 8 a
                    {
                        const double a0 = A[gid][4 * j + tig];
#pragma unroll
                        for (int n = 0; n < NBW; n++) dmma884(acc[0][n][0], acc[0][n][1], a0, bf[j][n]);
                    }
                    if (j >= 2) {
                        const double a1 = A[8 + gid][4 * j + tig];
#pragma unroll
                        for (int n = 0; n < NBW; n++) dmma884(acc[1][n][0], acc[1][n][1], a1, bf[j][n]);
                    }
                    if (j >= 4) {
                        const double a2 = A[16 + gid][4 * j + tig];
#pragma unroll
                        for (int n = 0; n < NBW; n++) dmma884(acc[2][n][0], acc[2][n][1], a2, bf[j][n]);
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int pr = 8 * a + gid;
                if (pr < R) {
#pragma unroll
                    for (int n = 0; n < NBW; n++) {
                        if (f >= 0) {
                            const int c = ch * (NBW * 8) + n * 8 + 2 * tig;
                            Sp[hm_panel_blocked_index((int64_t)f * R + pr, c, NB)] = acc[a][n][0];
                            Sp[hm_panel_blocked_index((int64_t)f * R + pr, c + 1, NB)] = acc[a][n][1];
                        } else {
                            *reinterpret_cast<double2 *>(o + (size_t)pr * CS + n * 8) =
                                make_double2(acc[a][n][0], acc[a][n][1]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// cores: LAMp[row box] = sum over its leaves of G_leaf (20 x 20) MUp[column box] (20 x CS), on the FP64
// tensor cores.  A warp owns (row box, half of the columns when CS = 64); A fragments (G, column-major,
// one of a few hundred distinct cores: L1-resident) and B fragments (rows of MUp) through L1.
// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(NT)
hm_nest_core_panel_kernel(int nboxes, const int32_t *__restrict__ rleaf_begin, const HmNestLeaf *__restrict__ rleaf,
                          const double *__restrict__ cores, const double *__restrict__ MUp, double *__restrict__ LAMp)
{
    constexpr int CS = NB * 8, NBW = NB > 4 ? 4 : NB, NCH = NB / NBW;
    const int lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int u = __shfl_sync(0xffffffffu, blockIdx.x * (NT / 32) + (threadIdx.x >> 5), 0);
    if (u >= nboxes * NCH) return;
    const int box = u / NCH, ch = u - box * NCH;
    double acc[3][NBW][2];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int n = 0; n < NBW; n++) acc[a][n][0] = acc[a][n][1] = 0.0;
    const int l0 = __shfl_sync(0xffffffffu, rleaf_begin[box], 0), l1 = __shfl_sync(0xffffffffu, rleaf_begin[box + 1], 0);
    const int q2 = min(16 + gid, R - 1); // rows 20 .. 23 of the third block repeat row 19 and are dropped
    for (int l = l0; l < l1; l++) {
        const HmNestLeaf lf = rleaf[l];
        const double *__restrict__ G = cores + (size_t)lf.core * (R * R) + (size_t)tig * R; // G[q + p R], p = 4 j + tig
        const double *__restrict__ mu = MUp + ((size_t)lf.cnode * R + tig) * CS + ch * (NBW * 8) + gid;
#pragma unroll
        for (int j = 0; j < R / 4; j++) {
            double bf[NBW];
#pragma unroll
            for (int n = 0; n < NBW; n++) bf[n] = mu[(size_t)(4 * j) * CS + n * 8];
            const double a0 = __ldg(G + 4 * j * R + gid), a1 = __ldg(G + 4 * j * R + 8 + gid), a2 = __ldg(G + 4 * j * R + q2);
#pragma unroll
            for (int n = 0; n < NBW; n++) {
                dmma884(acc[0][n][0], acc[0][n][1], a0, bf[n]);
                dmma884(acc[1][n][0], acc[1][n][1], a1, bf[n]);
                dmma884(acc[2][n][0], acc[2][n][1], a2, bf[n]);
            }
        }
    }
    double *o = LAMp + (size_t)box * R * CS + ch * (NBW * 8) + 2 * tig;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int q = 8 * a + gid;
        if (q < R) {
#pragma unroll
            for (int n = 0; n < NBW; n++)
                *reinterpret_cast<double2 *>(o + (size_t)q * CS + n * 8) = make_double2(acc[a][n][0], acc[a][n][1]);
        }
    }
}

template <int NB>
cudaError_t run_up(const HmNestDev &T, const double *pts, const double *Xt, const double *M, double *MUp, cudaStream_t st)
{
    constexpr int NCH = NB > 4 ? 2 : 1;
    if (T.nbase > 0) {
        const int units = T.nbase * NCH;
        hm_nest_base_panel_kernel<NB><<<(unsigned)((units + NTB / 32 - 1) / (NTB / 32)), NTB, 0, st>>>(T.nodes, T.base,
                                                                                                 T.nbase, pts, Xt, MUp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    for (int k = 0; k < T.ntiers; k++) {
        const int n = T.tier_sub0[k + 1] - T.tier_sub0[k];
        if (n <= 0) continue;
        if (k == 0)
            hm_nest_up_panel_kernel<NB, NT><<<(unsigned)n, NT, 0, st>>>(T.nodes, T.order, T.grp, T.sub_g0, T.tier_sub0[k], M, MUp);
        else
            hm_nest_up_panel_kernel<NB, 512><<<(unsigned)n, 512, 0, st>>>(T.nodes, T.order, T.grp, T.sub_g0, T.tier_sub0[k], M, MUp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int CS>
cudaError_t run_down(const HmNestDev &T, const int32_t *fin, const double *M, double *LAMp, double *Sp, cudaStream_t st)
{
    for (int k = T.ntiers - 1; k >= 0; k--) {
        const int n = T.tier_sub0[k + 1] - T.tier_sub0[k];
        if (n <= 0) continue;
        if (k == 0)
            hm_nest_down_panel_kernel<CS / 8, NT><<<(unsigned)n, NT, 0, st>>>(T.nodes, T.order, T.grp, T.sub_g0, T.tier_sub0[k],
                                                                             fin, M, LAMp, Sp);
        else
            hm_nest_down_panel_kernel<CS / 8, 512><<<(unsigned)n, 512, 0, st>>>(T.nodes, T.order, T.grp, T.sub_g0, T.tier_sub0[k],
                                                                               fin, M, LAMp, Sp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int NB>
cudaError_t run_core(int nboxes, const int32_t *rleaf_begin, const HmNestLeaf *rleaf, const double *cores,
                     const double *MUp, double *LAMp, cudaStream_t st)
{
    constexpr int NCH = NB > 4 ? 2 : 1;
    const int units = nboxes * NCH;
    if (units <= 0) return cudaSuccess;
    hm_nest_core_panel_kernel<NB><<<(unsigned)((units + NT / 32 - 1) / (NT / 32)), NT, 0, st>>>(nboxes, rleaf_begin, rleaf,
                                                                                            cores, MUp, LAMp);
    return cudaGetLastError();
}

} // namespace

cudaError_t hm_launch_nest_up_panel(int CS, const HmNestDev &T, const double *pts, const double *Xt, const double *M,
                                    double *MUp, cudaStream_t st)
{
    switch (CS) {
    case 16: return run_up<2>(T, pts, Xt, M, MUp, st);
    case 32: return run_up<4>(T, pts, Xt, M, MUp, st);
    case 64: return run_up<8>(T, pts, Xt, M, MUp, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_nest_core_panel(int CS, int nboxes, const int32_t *rleaf_begin, const HmNestLeaf *rleaf,
                                      const double *cores, const double *MUp, double *LAMp, cudaStream_t st)
{
    switch (CS) {
    case 16: return run_core<2>(nboxes, rleaf_begin, rleaf, cores, MUp, LAMp, st);
    case 32: return run_core<4>(nboxes, rleaf_begin, rleaf, cores, MUp, LAMp, st);
    case 64: return run_core<8>(nboxes, rleaf_begin, rleaf, cores, MUp, LAMp, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_nest_down_panel(int CS, const HmNestDev &T, const int32_t *fin, const double *M, double *LAMp,
                                      double *Sp, cudaStream_t st)
{
    switch (CS) {
    case 16: return run_down<16>(T, fin, M, LAMp, Sp, st);
    case 32: return run_down<32>(T, fin, M, LAMp, Sp, st);
    case 64: return run_down<64>(T, fin, M, LAMp, Sp, st);
    default: return cudaErrorInvalidValue;
    }
}
