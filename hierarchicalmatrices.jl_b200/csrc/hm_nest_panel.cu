// hm_nest_panel.cu -- many right-hand sides on the nested-basis form (hm_nest.h): the passes of hm_nest.cu
// applied to a panel of CS = 16 / 32 / 64 columns.
//
//   MUp[box][q][c], LAMp[box][q][c]   row-major panels of 20 x CS words per box
//   base   MUp[finest column box] = T(eta)' Xt[columns of the box]      FP64 tensor cores (DMMA), the 20 x 32
//          Chebyshev tile generated per chunk of 32 points into a warp-private shared-memory tile
//   up     MUp[box] = M0 MUp[half 0] + M1 MUp[half 1]                   lane = column; the maps are lower
//   down   LAMp[box] += M_which' LAMp[parent]                           triangular constants, read as constant-
//          bank operands of the DFMAs: no load instructions for them at all
//   cores  LAMp[row box] = sum over its leaves of G_leaf MUp[column box of the leaf]
//   the finest row boxes leave their coefficients fragment-major in Sp, where the panel kernel of the dense
//   leaves (hm_free3_panel_kernel, hm_free_panel.cu) picks them up as one 20-term "low-rank run" per item
//   and evaluates them together with the dense entries.
#include <cuda_runtime.h>

#include <cstdint>

#include "hm_kernels.cuh"
#include "hm_nest.h"

namespace {

constexpr int R = HM_NEST_R;
constexpr int NT = 256;
constexpr int TPITCH = 36; // tile pitch of the base kernel: 4 (mod 16) words -> conflict-free A fragments
constexpr int TROWS = 24;  // 20 moments padded to three 8-row MMA blocks

__constant__ double cM[2][R * R]; // the two transfer maps, [w][q * R + p], zero above the diagonal

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

// out[q] += sum_{p <= q} M_w[q][p] in[p] (W = 0, 1) or, transposed, out[p] += sum_{q >= p} M_w[q][p] in[q]
template <int W, bool TRANSPOSED>
__device__ __forceinline__ void apply_map(double (&out)[R], const double (&in)[R])
{
#pragma unroll
    for (int q = 0; q < R; q++)
#pragma unroll
        for (int p = 0; p <= q; p++) {
            if (TRANSPOSED)
                out[p] = fma(cM[W][q * R + p], in[q], out[p]);
            else
                out[q] = fma(cM[W][q * R + p], in[p], out[q]);
        }
}

// ---------------------------------------------------------------------------
// up / down over the subtree schedule of hm_nest_host.cpp; a warp owns (box, 32 columns), lane = column
// ---------------------------------------------------------------------------
template <int CS>
__global__ void __launch_bounds__(NT)
hm_nest_up_panel_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ order,
                        const int32_t *__restrict__ grp, const int32_t *__restrict__ sub_g0, int sub0, double *MUp)
{
    constexpr int CG = CS > 32 ? CS / 32 : 1;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    const int sub = sub0 + blockIdx.x;
    for (int g = sub_g0[sub]; g < sub_g0[sub + 1]; g++) {
        const int e0 = grp[g], nu = (grp[g + 1] - e0) * CG;
        for (int u = warp; u < nu; u += nw) {
            const int id = order[e0 + u / CG], c = (u % CG) * 32 + lane;
            const int c0 = nodes[id].child0;
            if (c0 < 0 || c >= CS) continue;
            double out[R], in[R];
#pragma unroll
            for (int q = 0; q < R; q++) out[q] = 0.0;
            const double *m0 = MUp + (size_t)c0 * R * CS + c;
#pragma unroll
            for (int p = 0; p < R; p++) in[p] = m0[(size_t)p * CS];
            apply_map<0, false>(out, in);
#pragma unroll
            for (int p = 0; p < R; p++) in[p] = m0[(size_t)(R + p) * CS];
            apply_map<1, false>(out, in);
            double *o = MUp + (size_t)id * R * CS + c;
#pragma unroll
            for (int q = 0; q < R; q++) o[(size_t)q * CS] = out[q];
        }
        __syncthreads();
    }
}

// fin[box] >= 0: a finest box; its completed coefficients go fragment-major into Sp at rows 20 fin[box] ..
template <int CS>
__global__ void __launch_bounds__(NT)
hm_nest_down_panel_kernel(const HmNestNode *__restrict__ nodes, const int32_t *__restrict__ order,
                          const int32_t *__restrict__ grp, const int32_t *__restrict__ sub_g0, int sub0,
                          const int32_t *__restrict__ fin, double *LAMp, double *__restrict__ Sp)
{
    constexpr int CG = CS > 32 ? CS / 32 : 1, NB = CS / 8;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    const int sub = sub0 + blockIdx.x;
    for (int g = sub_g0[sub + 1] - 1; g >= sub_g0[sub]; g--) { // shallowest depth first
        const int e0 = grp[g], nu = (grp[g + 1] - e0) * CG;
        for (int u = warp; u < nu; u += nw) {
            const int id = order[e0 + u / CG], c = (u % CG) * 32 + lane;
            if (c >= CS) continue;
            const HmNestNode nd = nodes[id];
            double out[R], in[R];
            double *o = LAMp + (size_t)id * R * CS + c;
#pragma unroll
            for (int p = 0; p < R; p++) out[p] = o[(size_t)p * CS];
            if (nd.parent >= 0) {
                const double *lp = LAMp + (size_t)nd.parent * R * CS + c;
#pragma unroll
                for (int q = 0; q < R; q++) in[q] = lp[(size_t)q * CS];
                if (nd.which == 0)
                    apply_map<0, true>(out, in);
                else
                    apply_map<1, true>(out, in);
            }
            const int f = fin[id];
            if (f >= 0) {
#pragma unroll
                for (int p = 0; p < R; p++) Sp[hm_panel_blocked_index((int64_t)f * R + p, c, NB)] = out[p];
            } else {
#pragma unroll
                for (int p = 0; p < R; p++) o[(size_t)p * CS] = out[p];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// cores: a warp owns (row box, 32 columns); G is read with warp-uniform 128-bit loads
// ---------------------------------------------------------------------------
template <int CS>
__global__ void __launch_bounds__(NT)
hm_nest_core_panel_kernel(int nboxes, const int32_t *__restrict__ rleaf_begin, const HmNestLeaf *__restrict__ rleaf,
                          const double *__restrict__ cores, const double *__restrict__ MUp, double *__restrict__ LAMp)
{
    constexpr int CG = CS > 32 ? CS / 32 : 1;
    const int lane = threadIdx.x & 31;
    const int u = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (u >= nboxes * CG) return;
    const int box = u / CG, c = (u % CG) * 32 + lane;
    if (c >= CS) return;
    double out[R];
#pragma unroll
    for (int q = 0; q < R; q++) out[q] = 0.0;
    for (int l = rleaf_begin[box]; l < rleaf_begin[box + 1]; l++) {
        const HmNestLeaf lf = rleaf[l];
        const double2 *__restrict__ G = reinterpret_cast<const double2 *>(cores + (size_t)lf.core * (R * R));
        const double *__restrict__ mu = MUp + (size_t)lf.cnode * R * CS + c;
#pragma unroll 4
        for (int p = 0; p < R; p++) {
            const double m = mu[(size_t)p * CS];
#pragma unroll
            for (int q2 = 0; q2 < R / 2; q2++) {
                const double2 gg = __ldg(G + p * (R / 2) + q2); // G[2 q2 + p * R], G[2 q2 + 1 + p * R]
                out[2 * q2] = fma(gg.x, m, out[2 * q2]);
                out[2 * q2 + 1] = fma(gg.y, m, out[2 * q2 + 1]);
            }
        }
    }
    double *o = LAMp + (size_t)box * R * CS + c;
#pragma unroll
    for (int q = 0; q < R; q++) o[(size_t)q * CS] = out[q];
}

template <int NB>
cudaError_t run_up(const HmNestDev &T, const double *pts, const double *Xt, double *MUp, cudaStream_t st)
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
        hm_nest_up_panel_kernel<NB * 8><<<(unsigned)n, NT, 0, st>>>(T.nodes, T.order, T.grp, T.sub_g0, T.tier_sub0[k], MUp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int CS>
cudaError_t run_down(const HmNestDev &T, const int32_t *fin, double *LAMp, double *Sp, cudaStream_t st)
{
    for (int k = T.ntiers - 1; k >= 0; k--) {
        const int n = T.tier_sub0[k + 1] - T.tier_sub0[k];
        if (n <= 0) continue;
        hm_nest_down_panel_kernel<CS><<<(unsigned)n, NT, 0, st>>>(T.nodes, T.order, T.grp, T.sub_g0, T.tier_sub0[k], fin, LAMp,
                                                                 Sp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int CS>
cudaError_t run_core(int nboxes, const int32_t *rleaf_begin, const HmNestLeaf *rleaf, const double *cores,
                     const double *MUp, double *LAMp, cudaStream_t st)
{
    constexpr int CG = CS > 32 ? CS / 32 : 1;
    const int units = nboxes * CG;
    if (units <= 0) return cudaSuccess;
    hm_nest_core_panel_kernel<CS><<<(unsigned)((units + NT / 32 - 1) / (NT / 32)), NT, 0, st>>>(nboxes, rleaf_begin, rleaf,
                                                                                            cores, MUp, LAMp);
    return cudaGetLastError();
}

} // namespace

cudaError_t hm_nest_panel_init(const double *M_host)
{
    return cudaMemcpyToSymbol(cM, M_host, sizeof(double) * 2 * R * R);
}

cudaError_t hm_launch_nest_up_panel(int CS, const HmNestDev &T, const double *pts, const double *Xt, double *MUp,
                                    cudaStream_t st)
{
    switch (CS) {
    case 16: return run_up<2>(T, pts, Xt, MUp, st);
    case 32: return run_up<4>(T, pts, Xt, MUp, st);
    case 64: return run_up<8>(T, pts, Xt, MUp, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_nest_core_panel(int CS, int nboxes, const int32_t *rleaf_begin, const HmNestLeaf *rleaf,
                                      const double *cores, const double *MUp, double *LAMp, cudaStream_t st)
{
    switch (CS) {
    case 16: return run_core<16>(nboxes, rleaf_begin, rleaf, cores, MUp, LAMp, st);
    case 32: return run_core<32>(nboxes, rleaf_begin, rleaf, cores, MUp, LAMp, st);
    case 64: return run_core<64>(nboxes, rleaf_begin, rleaf, cores, MUp, LAMp, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_nest_down_panel(int CS, const HmNestDev &T, const int32_t *fin, double *LAMp, double *Sp,
                                      cudaStream_t st)
{
    switch (CS) {
    case 16: return run_down<16>(T, fin, LAMp, Sp, st);
    case 32: return run_down<32>(T, fin, LAMp, Sp, st);
    case 64: return run_down<64>(T, fin, LAMp, Sp, st);
    default: return cudaErrorInvalidValue;
    }
}
