// hm_panel.cu -- many-right-hand-side (matmat) kernels: the same packed streams as the
// matvec, applied to a panel of up to 64 vectors with FP64 tensor-core MMAs
// (mma.sync.aligned.m8n8k4 .f64 -> SASS DMMA; tcgen05 has no f64 kind).
//
// The reference reaches multiple right-hand sides only by repeating its scalar leaf
// loops per column (/root/reference/src/HierarchicalMatrix.jl:24-52 with the stride
// pair, test/runtests.jl:23-25).  Here every stream item is a small dense contraction
//      out[f][c] = sum_s W[s][f] * Z[s][c],   f < F <= 128 per pass, c < 8*NB <= 64
// so the slab W is still read from HBM exactly once while the flops grow with the
// panel width: HBM-bound up to ~16 columns, FP64-pipe-bound at 64.
//
// All panels are kept "row-major, column index fastest" (row pitch CS = 8*NB doubles):
//   Xt[j][c]  transposed copy of the user's X       Pp[p][c]  stage-1 partial sums
//   Sp[k][c]  stage-2 results                        Yt[i][c]  stage-3 results
// so that the z rows an item needs are contiguous CS*8-byte pieces (cp.async friendly).
#include "hm_kernels.cuh"

namespace {

constexpr int PT = 256;  // threads per CTA
constexpr int NST = 3;   // cp.async pipeline stages
constexpr int MT = 128;  // rows of the fast dimension per pass

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// stage 1 / stage 3 on a panel
// ---------------------------------------------------------------------------
template <bool GATHER, int NB>
__global__ void __launch_bounds__(PT, 2)
hm_panel_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                const double *__restrict__ W, const double *__restrict__ Xt,
                const double *__restrict__ Sp, double *__restrict__ out, int accumulate,
                int budget_words)
{
    constexpr int CS = NB * 8, ZP = CS + 8;
    extern __shared__ __align__(16) double dsm[];
    __shared__ int zrow[GATHER ? HM_SMAX : 1];
    __shared__ int rpos[GATHER ? HM_MAXRUNS + 1 : 1];
    __shared__ int rsrc[GATHER ? HM_MAXRUNS : 1];

    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int S = it.S, F = it.F, Fp = it.Fp;

    if (GATHER) {
        // source row of every z entry: >= 0 row of Xt, < 0 row ~v of Sp
        for (int r = t; r < it.nrun; r += PT) {
            HmRun rr = runs[it.run0 + r];
            rpos[r] = rr.pos;
            rsrc[r] = rr.src;
        }
        if (t == 0) rpos[it.nrun] = S;
        __syncthreads();
        for (int e = t; e < S; e += PT) {
            int lo = 0, hi = it.nrun;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (rpos[mid] <= e)
                    lo = mid;
                else
                    hi = mid;
            }
            int src = rsrc[lo], off = e - rpos[lo];
            zrow[e] = src >= 0 ? src + off : ~((~src) + off);
        }
        __syncthreads();
    }

    const double *__restrict__ Wg = W + it.slab;

    for (int f0 = 0; f0 < Fp; f0 += MT) {
        const int mt = min(MT, Fp - f0);         // even
        const int WP = ((mt + 15) & ~15) + 8;    // smem pitch: = 8 (mod 16) -> conflict-free fragments
        int KC = (budget_words / NST / (WP + ZP)) & ~3;
        KC = max(4, min(32, KC));
        const int stage_words = KC * (WP + ZP);
        const int nfb = (mt + 7) >> 3;           // 8-row blocks
        const int fwarps = (nfb + 1) >> 1;       // a warp owns up to two of them
        const int kgroups = 8 / fwarps;          // split-K groups
        const int fw = warp % fwarps, kg = warp / fwarps;
        const bool active = kg < kgroups;
        const int fb0 = fw * 2;
        const bool two = fb0 + 1 < nfb;
        const int nchunks = (S + KC - 1) / KC;

        double acc[2][NB][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int n = 0; n < NB; n++) acc[a][n][0] = acc[a][n][1] = 0.0;

        auto issue = [&](int ch) {
            double *Wsm = dsm + (size_t)(ch % NST) * stage_words;
            double *Zsm = Wsm + KC * WP;
            const int s0 = ch * KC, rows = min(KC, S - s0);
            const int hw = mt >> 1;
            for (int idx = t; idx < rows * hw; idx += PT) {
                int r = idx / hw, p = idx - r * hw;
                cp_async16(Wsm + r * WP + 2 * p, Wg + (size_t)(s0 + r) * Fp + f0 + 2 * p);
            }
            constexpr int hz = CS / 2;
            for (int idx = t; idx < rows * hz; idx += PT) {
                int r = idx / hz, p = idx - r * hz;
                const double *src;
                if (GATHER) {
                    int zr = zrow[s0 + r];
                    src = zr >= 0 ? Xt + (size_t)zr * CS : Sp + (size_t)(~zr) * CS;
                } else {
                    src = Xt + (size_t)(it.zoff + s0 + r) * CS;
                }
                cp_async16(Zsm + r * ZP + 2 * p, src + 2 * p);
            }
            if (rows < KC) { // zero the tail of the last chunk (0 * stale NaN would poison the sums)
                for (int idx = t; idx < (KC - rows) * WP; idx += PT) Wsm[rows * WP + idx] = 0.0;
                for (int idx = t; idx < (KC - rows) * ZP; idx += PT) Zsm[rows * ZP + idx] = 0.0;
            }
        };

        for (int ch = 0; ch < NST - 1; ch++) {
            if (ch < nchunks) issue(ch);
            cp_async_commit();
        }
        for (int ch = 0; ch < nchunks; ch++) {
            if (ch + NST - 1 < nchunks) issue(ch + NST - 1);
            cp_async_commit();
            cp_async_wait<NST - 1>();
            __syncthreads();
            if (active) {
                const double *Wsm = dsm + (size_t)(ch % NST) * stage_words;
                const double *Zsm = Wsm + KC * WP;
                const double *ap = Wsm + tig * WP + fb0 * 8 + gid;
                const double *bp = Zsm + tig * ZP + gid;
#pragma unroll 2
                for (int ks = kg; ks < (KC >> 2); ks += kgroups) {
                    const double a0 = ap[ks * 4 * WP];
                    const double a1 = two ? ap[ks * 4 * WP + 8] : 0.0;
#pragma unroll
                    for (int n = 0; n < NB; n++) {
                        const double b = bp[ks * 4 * ZP + n * 8];
                        dmma884(acc[0][n][0], acc[0][n][1], a0, b);
                        if (two) dmma884(acc[1][n][0], acc[1][n][1], a1, b);
                    }
                }
            }
            __syncthreads();
        }
        cp_async_wait<0>();

        // combine the split-K groups in a fixed order through shared memory, then write
        double *Csm = dsm;
        for (int g = 0; g < kgroups; g++) {
            if (active && kg == g) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    if (a == 1 && !two) break;
                    double *row = Csm + ((fb0 + a) * 8 + gid) * ZP + 2 * tig;
#pragma unroll
                    for (int n = 0; n < NB; n++) {
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
        const int rows_out = min(mt, F - f0);
        constexpr int hz = CS / 2;
        for (int idx = t; idx < rows_out * hz; idx += PT) {
            int r = idx / hz, p = idx - r * hz;
            double2 v = *reinterpret_cast<const double2 *>(Csm + r * ZP + 2 * p);
            double2 *g = reinterpret_cast<double2 *>(out + (size_t)(it.out + f0 + r) * CS) + p;
            if (GATHER && accumulate) {
                double2 o = *g;
                v.x += o.x;
                v.y += o.y;
            }
            *g = v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// stage 2 on a panel: one CTA per low-rank leaf
//   T[k][c] = sum of the leaf's partial panels (column order); S = F T | Sigma .* T
// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256)
hm_core_panel_kernel(const HmCoreBlock *__restrict__ blocks, const int32_t *__restrict__ plist,
                     const double *__restrict__ Pp, const double *__restrict__ core,
                     double *__restrict__ Sp, int max_r)
{
    constexpr int CS = NB * 8;
    extern __shared__ double sm[];
    double *Tsm = sm;                         // [rv][CS]
    const HmCoreBlock cb = blocks[blockIdx.x];
    const double *__restrict__ Fg = core + cb.core; // ru x rv (bary) or r (low rank); L1-resident
    (void)max_r;
    const int t = threadIdx.x, T = blockDim.x;
    const int32_t *pl = plist + cb.pl0;
    const int nT = cb.rv * CS;
    for (int e = t; e < nT; e += T) {
        double a = 0.0;
        int i = 0;
        for (; i + 3 < cb.npl; i += 4) {
            double p0 = Pp[(size_t)pl[i] * CS + e], p1 = Pp[(size_t)pl[i + 1] * CS + e];
            double p2 = Pp[(size_t)pl[i + 2] * CS + e], p3 = Pp[(size_t)pl[i + 3] * CS + e];
            a += p0;
            a += p1;
            a += p2;
            a += p3;
        }
        for (; i < cb.npl; i++) a += Pp[(size_t)pl[i] * CS + e];
        Tsm[e] = a;
    }
    __syncthreads();
    double *o = Sp + (size_t)cb.soff * CS;
    const int nS = cb.ru * CS;
    if (cb.kind == HM_LEAF_LOWRANK) {
        for (int e = t; e < nS; e += T) o[e] = Tsm[e] * __ldg(Fg + e / CS);
    } else {
        for (int e = t; e < nS; e += T) {
            int k = e / CS, c = e - k * CS;
            double a = 0.0;
            for (int l = 0; l < cb.rv; l++) a = fma(__ldg(Fg + k + (size_t)l * cb.ru), Tsm[l * CS + c], a);
            o[e] = a;
        }
    }
}

// ---------------------------------------------------------------------------
// panel transposes (32 x 32 tiles through shared memory)
//   in : Xt[j][c] = c < nrhs ? X[j + c*ldx] : 0
//   out: Y[i + c*ldy] = (accumulate ? Y : 0) + Yt[i][c],  rows [r0, r1)
// ---------------------------------------------------------------------------
__global__ void hm_panel_in_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, int nrhs, int CS,
                                   double *__restrict__ Xt)
{
    __shared__ double tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        int64_t j = j0 + threadIdx.x;
        int c = c0 + cc;
        tile[cc][threadIdx.x] = (j < n && c < nrhs) ? X[j + (int64_t)c * ldx] : 0.0;
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        int64_t j = j0 + jj;
        int c = c0 + threadIdx.x;
        if (j < n && c < CS) Xt[j * CS + c] = tile[threadIdx.x][jj];
    }
}

__global__ void hm_panel_out_kernel(const double *__restrict__ Yt, int CS, int64_t r0, int64_t r1, int nrhs,
                                    double *__restrict__ Y, int64_t ldy, int accumulate)
{
    __shared__ double tile[32][33];
    const int64_t i0 = r0 + (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
        int64_t i = i0 + ii;
        int c = c0 + threadIdx.x;
        tile[ii][threadIdx.x] = (i < r1 && c < CS) ? Yt[i * CS + c] : 0.0;
    }
    __syncthreads();
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        int64_t i = i0 + threadIdx.x;
        int c = c0 + cc;
        if (i < r1 && c < nrhs) {
            double *y = Y + i + (int64_t)c * ldy;
            *y = (accumulate ? *y : 0.0) + tile[threadIdx.x][cc];
        }
    }
}

template <bool GATHER, int NB>
cudaError_t launch_panel(const HmItem *items, int64_t nitems, const HmRun *runs, const double *W,
                         const double *Xt, const double *Sp, double *out, int accumulate, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    const int budget_words = 84 * 1024 / 8;
    const size_t smem = (size_t)budget_words * sizeof(double);
    static bool configured = false; // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(hm_panel_kernel<GATHER, NB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    hm_panel_kernel<GATHER, NB><<<(unsigned)nitems, PT, smem, st>>>(items, runs, W, Xt, Sp, out, accumulate,
                                                                    budget_words);
    return cudaGetLastError();
}

template <int NB>
cudaError_t launch_core_panel(const HmCoreBlock *blocks, int64_t nblocks, const int32_t *plist, const double *Pp,
                              const double *core, double *Sp, int max_r, cudaStream_t st)
{
    if (nblocks <= 0) return cudaSuccess;
    const size_t smem = (size_t)max_r * NB * 8 * sizeof(double);
    if (smem > 160 * 1024) return cudaErrorInvalidConfiguration;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(hm_core_panel_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    hm_core_panel_kernel<NB><<<(unsigned)nblocks, 256, smem, st>>>(blocks, plist, Pp, core, Sp, max_r);
    return cudaGetLastError();
}

} // namespace

int hm_panel_width(int nrhs) { return nrhs <= 16 ? 16 : nrhs <= 32 ? 32 : 64; }

cudaError_t hm_launch_panel_in(const double *X, int64_t ldx, int64_t n, int nrhs, int CS, double *Xt,
                               cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((CS + 31) / 32)), block(32, 8);
    hm_panel_in_kernel<<<grid, block, 0, st>>>(X, ldx, n, nrhs, CS, Xt);
    return cudaGetLastError();
}

cudaError_t hm_launch_panel_out(const double *Yt, int CS, int64_t r0, int64_t r1, int nrhs, double *Y,
                                int64_t ldy, int accumulate, cudaStream_t st)
{
    if (r1 <= r0) return cudaSuccess;
    dim3 grid((unsigned)((r1 - r0 + 31) / 32), (unsigned)((CS + 31) / 32)), block(32, 8);
    hm_panel_out_kernel<<<grid, block, 0, st>>>(Yt, CS, r0, r1, nrhs, Y, ldy, accumulate);
    return cudaGetLastError();
}

cudaError_t hm_launch_panel_stage1(int CS, const HmItem *items, int64_t nitems, const double *vstream,
                                   const double *Xt, double *Pp, cudaStream_t st)
{
    switch (CS) {
    case 16: return launch_panel<false, 2>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
    case 32: return launch_panel<false, 4>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
    case 64: return launch_panel<false, 8>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_panel_stage2(int CS, const HmCoreBlock *blocks, int64_t nblocks, const int32_t *plist,
                                   const double *Pp, const double *core, double *Sp, int max_r, cudaStream_t st)
{
    switch (CS) {
    case 16: return launch_core_panel<2>(blocks, nblocks, plist, Pp, core, Sp, max_r, st);
    case 32: return launch_core_panel<4>(blocks, nblocks, plist, Pp, core, Sp, max_r, st);
    case 64: return launch_core_panel<8>(blocks, nblocks, plist, Pp, core, Sp, max_r, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_panel_stage3(int CS, const HmItem *items, int64_t nitems, const HmRun *runs,
                                   const double *ustream, const double *Xt, const double *Sp, double *Yt,
                                   int accumulate, cudaStream_t st)
{
    switch (CS) {
    case 16: return launch_panel<true, 2>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
    case 32: return launch_panel<true, 4>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
    case 64: return launch_panel<true, 8>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
    default: return cudaErrorInvalidValue;
    }
}
