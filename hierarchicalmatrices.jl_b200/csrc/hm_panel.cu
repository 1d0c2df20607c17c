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
#include <cstdlib>
#include <type_traits>

#include "hm_kernels.cuh"

namespace {

constexpr int PT = 256; // threads per CTA

constexpr int PD = 2; // L2 prefetch distance (load batches) of the streaming panel kernel at 64 columns
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// stage 1 / stage 3 on a panel.
//
// The fast dimension is cut into tiles of 16 rows (two 8-row MMA blocks); a warp owns
// a tile and streams its 128-byte row pieces of the slab straight from HBM into A
// fragments (lane (g, t) reads W[s0 + t][f0 + g]: four 64-byte pieces per load
// instruction, every sector fully used), U k-steps in flight.  B fragments (the z rows)
// come through L1: all warps of the CTA read the same rows.  No shared-memory staging
// and no barriers in the main loop.  Items with fewer than 8 tiles split the slow
// dimension over warp groups, combined in a fixed order through shared memory.
// ---------------------------------------------------------------------------
template <bool GATHER, int NB, int U>
__global__ void __launch_bounds__(PT, 2)
hm_panel_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                const double *__restrict__ W, const double *__restrict__ Xt,
                const double *__restrict__ Sp, double *__restrict__ out, int accumulate)
{
    constexpr int CS = NB * 8, CP = CS + 8;
    extern __shared__ __align__(16) double csm[]; // [<= 8 tiles][16][CP] split-K combine buffer
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
    const int nfb = (Fp + 7) >> 3;  // 8-row blocks
    // 16-row tiles.  With TB3 an odd block count gives the last tile three blocks instead of
    // opening a half-empty tile: the rank-20 slabs of stage 1 (most of its bytes) become one
    // 24-row tile per warp, all eight warps splitting the slow dimension -- one B fragment load
    // per three MMAs instead of per one and two, and no 2:1 imbalance between the warps.
    constexpr bool TB3 = NB <= 4 && !GATHER; // stage 1 only: stage-3 row segments gain nothing (measured)
    constexpr int TB = TB3 ? 3 : 2;
    const int nft = TB3 ? max(nfb >> 1, nfb > 0 ? 1 : 0) : (nfb + 1) >> 1;
    if (nft == 0) return;
    const int kgroups = nft >= 8 ? 1 : 8 / nft;
    const int ntw = nft >= 8 ? 8 : nft; // warps per k-group
    const int kg = warp / ntw;
    const bool active = kg < kgroups;
    int s_lo = 0, s_hi = S;
    if (kgroups > 1) {
        int sg = (((S + kgroups - 1) / kgroups) + 3) & ~3;
        s_lo = min(S, kg * sg);
        s_hi = min(S, s_lo + sg);
    }

    for (int ft = warp % ntw; ft < nft && active; ft += ntw) {
        const int nblk = min(nfb - 2 * ft, (TB3 && ft == nft - 1) ? 3 : 2); // blocks of this tile
        const int col0 = ft * 16 + gid;
        bool cok[TB];
#pragma unroll
        for (int a = 0; a < TB; a++) cok[a] = a < nblk && col0 + 8 * a < Fp;
        const double *__restrict__ ap = Wg + col0;

        double acc[TB][NB][2];
#pragma unroll
        for (int a = 0; a < TB; a++)
#pragma unroll
            for (int n = 0; n < NB; n++) acc[a][n][0] = acc[a][n][1] = 0.0;

        // One load batch = U k-steps (4 slab rows each).  Full batches run without row masks and
        // with incrementing pointers (no 64-bit multiplies in the loop); the ragged end of the
        // range takes the masked form once.
        auto batch = [&](int k0, auto masked) {
            constexpr bool MASK = decltype(masked)::value;
            double av[U][TB], b[U][NB];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int row = k0 + 4 * u + tig;
                const bool v = !MASK || row < s_hi;
                if (NB == 8) { // software prefetch into L2, PD batches ahead (helps the MMA-bound case only)
                    const int prow = row + 4 * U * PD;
                    if (prow < s_hi) {
#pragma unroll
                        for (int a = 0; a < TB; a++)
                            if (cok[a]) prefetch_l2(ap + (size_t)prow * Fp + 8 * a);
                        const double *pz;
                        if (GATHER) {
                            int zr = zrow[prow];
                            pz = zr >= 0 ? Xt + (size_t)zr * CS : Sp + (size_t)(~zr) * CS;
                        } else {
                            pz = Xt + (size_t)(it.zoff + prow) * CS;
                        }
                        if (gid * 8 < CS) prefetch_l2(pz + gid * 8);
                    }
                }
                const double *arow = ap + (size_t)row * Fp;
#pragma unroll
                for (int a = 0; a < TB; a++) av[u][a] = (v && cok[a]) ? __ldcs(arow + 8 * a) : 0.0;
                const double *zp = Xt;
                if (v) {
                    if (GATHER) {
                        int zr = zrow[row];
                        zp = zr >= 0 ? Xt + (size_t)zr * CS : Sp + (size_t)(~zr) * CS;
                    } else {
                        zp = Xt + (size_t)(it.zoff + row) * CS;
                    }
                }
#pragma unroll
                for (int n = 0; n < NB; n++) b[u][n] = v ? __ldg(zp + n * 8 + gid) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int n = 0; n < NB; n++) {
                    dmma884(acc[0][n][0], acc[0][n][1], av[u][0], b[u][n]);
                    if (nblk > 1) dmma884(acc[1][n][0], acc[1][n][1], av[u][1], b[u][n]);
                    if (TB3 && nblk > 2) dmma884(acc[TB - 1][n][0], acc[TB - 1][n][1], av[u][TB - 1], b[u][n]);
                }
            }
        };
        int k0 = s_lo;
        for (; k0 + 4 * U <= s_hi; k0 += 4 * U) batch(k0, std::false_type{});
        if (k0 < s_hi) batch(k0, std::true_type{});

        if (kgroups == 1) {
            // sole owner of the tile: write the C fragments straight out
#pragma unroll
            for (int a = 0; a < TB; a++) {
                const int f = ft * 16 + a * 8 + gid;
                if (f < F && a < nblk) {
                    double2 *g = reinterpret_cast<double2 *>(out + (size_t)(it.out + f) * CS) + tig;
#pragma unroll
                    for (int n = 0; n < NB; n++) {
                        double2 v = make_double2(acc[a][n][0], acc[a][n][1]);
                        if (GATHER && accumulate) {
                            double2 o = g[n * 4];
                            v.x += o.x;
                            v.y += o.y;
                        }
                        g[n * 4] = v;
                    }
                }
            }
        } else {
            // keep for the combine below (nft < 8: exactly one tile per warp)
            for (int g = 0; g < kgroups; g++) {
                if (kg == g) {
#pragma unroll
                    for (int a = 0; a < TB; a++) {
                        if (a < nblk) {
                            double *row = csm + ((size_t)ft * 16 + a * 8 + gid) * CP + 2 * tig;
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
                }
                // all 8 warps pass here the same number of times (see below)
                asm volatile("bar.sync 1, %0;\n" ::"r"(kgroups * ntw * 32) : "memory");
            }
        }
    }
    if (kgroups > 1) {
        __syncthreads();
        constexpr int hz = CS / 2;
        for (int idx = t; idx < F * hz; idx += PT) {
            int r = idx / hz, p = idx - r * hz;
            double2 v = *reinterpret_cast<const double2 *>(csm + (size_t)r * CP + 2 * p);
            double2 *g = reinterpret_cast<double2 *>(out + (size_t)(it.out + r) * CS) + p;
            if (GATHER && accumulate) {
                double2 o = *g;
                v.x += o.x;
                v.y += o.y;
            }
            *g = v;
        }
    }
}

// ---------------------------------------------------------------------------
// stage 1 / stage 3 on a panel, warp-specialised bulk-copy pipeline (HMB200_PANEL=tma).
//
// Warp 8 is the producer: per chunk of KC slab rows it issues one cp.async.bulk (TMA 1-D,
// SASS UBLKCP) per W row piece and per z row into padded shared-memory rows and lets an
// mbarrier count the bytes.  Warps 0-7 consume: fragments by LDS (row pitches = 8 mod 16
// doubles, conflict-free), FP64 MMAs, one mbarrier arrive per warp to hand the stage
// back.  TS stages keep both operands >= TS-1 chunks ahead of the math, so neither HBM nor
// L2 latency is exposed, and there is no block-wide barrier in the main loop.
// ---------------------------------------------------------------------------
constexpr int TS = 5;             // pipeline stages
constexpr int TSTAGE = 2304;      // doubles per stage (18 KB): KC rows of W (pitch WP) + of z (pitch ZP)
constexpr int TMT = 128;          // fast-dimension rows per pass

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// geometry of one pass over rows [f0, f0 + mt) of the fast dimension
struct PassGeom {
    int mt, WP, KC, nchunks;
};
template <int ZP> __device__ __forceinline__ PassGeom pass_geom(int Fp, int f0, int S)
{
    PassGeom g;
    g.mt = min(TMT, Fp - f0);
    g.WP = ((g.mt + 15) & ~15) + 8; // = 8 (mod 16): conflict-free fragment reads
    g.KC = min(32, (TSTAGE / (g.WP + ZP)) & ~3);
    g.nchunks = (S + g.KC - 1) / g.KC;
    return g;
}

template <bool GATHER, int NB>
__global__ void __launch_bounds__(288, 2)
hm_panel_tma_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                    const double *__restrict__ W, const double *__restrict__ Xt,
                    const double *__restrict__ Sp, double *__restrict__ out, int accumulate)
{
    constexpr int CS = NB * 8, ZP = CS + 8;
    extern __shared__ __align__(128) double dsm[]; // [TS][TSTAGE] pipeline stages
    __shared__ __align__(8) uint64_t full[TS], empty[TS];
    __shared__ int rpos[GATHER ? HM_MAXRUNS + 1 : 1];
    __shared__ int rsrc[GATHER ? HM_MAXRUNS : 1];

    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int S = it.S, F = it.F, Fp = it.Fp;

    if (t == 0) {
        for (int i = 0; i < TS; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (GATHER) {
        for (int r = t; r < it.nrun; r += blockDim.x) {
            HmRun rr = runs[it.run0 + r];
            rpos[r] = rr.pos;
            rsrc[r] = rr.src;
        }
        if (t == 0) rpos[it.nrun] = S;
    }
    __syncthreads();

    const double *__restrict__ Wg = W + it.slab;
    const int npass = (Fp + TMT - 1) / TMT;

    if (warp == 8) {
        // ------------------------------ producer ------------------------------
        int q = 0; // running chunk counter across passes -> stage index and mbarrier phase
        for (int pass = 0; pass < npass; pass++) {
            const int f0 = pass * TMT;
            const PassGeom g = pass_geom<ZP>(Fp, f0, S);
            for (int ch = 0; ch < g.nchunks; ch++, q++) {
                const int st = q % TS;
                if (q >= TS) mbar_wait(&empty[st], ((q / TS) - 1) & 1);
                double *Wsm = dsm + (size_t)st * TSTAGE;
                double *Zsm = Wsm + g.KC * g.WP;
                const int s0 = ch * g.KC, rows = min(g.KC, S - s0);
                if (lane == 0) mbar_arrive_expect_tx(&full[st], (unsigned)(rows * (g.mt + CS) * 8));
                __syncwarp();
                if (lane < rows) {
                    const int row = s0 + lane;
                    bulk_g2s(Wsm + lane * g.WP, Wg + (size_t)row * Fp + f0, (unsigned)(g.mt * 8), &full[st]);
                    const double *zp;
                    if (GATHER) {
                        int lo = 0, hi = it.nrun; // rpos[lo] <= row < rpos[hi]
                        while (hi - lo > 1) {
                            int mid = (lo + hi) >> 1;
                            if (rpos[mid] <= row)
                                lo = mid;
                            else
                                hi = mid;
                        }
                        const int src = rsrc[lo], off = row - rpos[lo];
                        zp = src >= 0 ? Xt + (size_t)(src + off) * CS : Sp + (size_t)((~src) + off) * CS;
                    } else {
                        zp = Xt + (size_t)(it.zoff + row) * CS;
                    }
                    bulk_g2s(Zsm + lane * ZP, zp, (unsigned)(CS * 8), &full[st]);
                }
            }
        }
        return;
    }

    // -------------------------------- consumers --------------------------------
    // 8 warps as fwarps (16-row tiles of the pass) x cwarps (groups of panel columns):
    // every warp owns its outputs, so there is no split-K and no combine step.
    int q = 0;
    for (int pass = 0; pass < npass; pass++) {
        const int f0 = pass * TMT;
        const PassGeom g = pass_geom<ZP>(Fp, f0, S);
        const int nfb = (g.mt + 7) >> 3;
        const int ntile = (nfb + 1) >> 1;                                  // 1..8 tiles of 16 rows
        const int fwarps = ntile > 4 ? 8 : ntile > 2 ? 4 : ntile > 1 ? 2 : 1;
        const int cwarps = 8 / fwarps;                                     // 1, 2, 4, 8
        const int nbw = (NB + cwarps - 1) / cwarps;                        // MMA column blocks per warp
        const int fw = warp % fwarps, cw = warp / fwarps;
        const int n0 = cw * nbw;                                           // first column block of this warp
        const int fb0 = fw * 2;
        const bool active = fb0 < nfb && n0 < NB;
        const bool two = fb0 + 1 < nfb;

        double acc[2][NB][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int n = 0; n < NB; n++) acc[a][n][0] = acc[a][n][1] = 0.0;

        for (int ch = 0; ch < g.nchunks; ch++, q++) {
            const int st = q % TS;
            mbar_wait(&full[st], (q / TS) & 1);
            if (active) {
                const double *Wsm = dsm + (size_t)st * TSTAGE;
                const double *Zsm = Wsm + g.KC * g.WP;
                const int rows = min(g.KC, S - ch * g.KC);
                const double *ap = Wsm + tig * g.WP + fb0 * 8 + gid;
                const double *bp = Zsm + tig * ZP + n0 * 8 + gid;
                for (int ks = 0; ks * 4 < rows; ks++) {
                    const bool v = ks * 4 + tig < rows; // rows past the end of the slab hold stale data
                    const double a0 = v ? ap[ks * 4 * g.WP] : 0.0;
                    const double a1 = (v && two) ? ap[ks * 4 * g.WP + 8] : 0.0;
#pragma unroll
                    for (int n = 0; n < NB; n++) {
                        if (n < nbw && n0 + n < NB) {
                            const double b = v ? bp[ks * 4 * ZP + n * 8] : 0.0;
                            dmma884(acc[0][n][0], acc[0][n][1], a0, b);
                            if (two) dmma884(acc[1][n][0], acc[1][n][1], a1, b);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }

        if (active) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int f = f0 + (fb0 + a) * 8 + gid;
                if (f < F && (a == 0 || two)) {
                    double2 *gp = reinterpret_cast<double2 *>(out + (size_t)(it.out + f) * CS) + n0 * 4 + tig;
#pragma unroll
                    for (int n = 0; n < NB; n++) {
                        if (n < nbw && n0 + n < NB) {
                            double2 v = make_double2(acc[a][n][0], acc[a][n][1]);
                            if (GATHER && accumulate) {
                                double2 o = gp[n * 4];
                                v.x += o.x;
                                v.y += o.y;
                            }
                            gp[n * 4] = v;
                        }
                    }
                }
            }
        }
    }
}

template <bool GATHER, int NB>
cudaError_t launch_panel_tma(const HmItem *items, int64_t nitems, const HmRun *runs, const double *W,
                             const double *Xt, const double *Sp, double *out, int accumulate, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    const size_t smem = (size_t)TS * TSTAGE * sizeof(double);
    { // the attribute is per device: set it on every call (cheap) rather than once per process
        cudaError_t e = cudaFuncSetAttribute(hm_panel_tma_kernel<GATHER, NB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    hm_panel_tma_kernel<GATHER, NB><<<(unsigned)nitems, 288, smem, st>>>(items, runs, W, Xt, Sp, out, accumulate);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// stage 1 / stage 3 on a panel as a tiled GEMM (HMB200_PANEL=mm; default at 64 columns).
//
// One CTA of 4 warps per item and pass of 64 slab rows: C[64 x CS] += W'[64 x S] Z[S x CS].
// Both operands are staged through a 4-deep cp.async pipeline of 16-row chunks (padded
// pitches, conflict-free fragment reads); a warp owns a 32-row x (CS/2 or CS/4)-column
// tile, i.e. up to 16 independent MMA accumulators, so the tensor pipe stays busy across
// the fragment loads; no split-K, C fragments go straight to global memory.
// ---------------------------------------------------------------------------
constexpr int MM_T = 128;   // threads
constexpr int MM_ST = 3;    // stages
constexpr int MM_KC = 16;   // slab rows per stage
constexpr int MM_MT = 64;   // fast-dimension rows per pass
constexpr int MM_WP = MM_MT + 8;

__device__ __forceinline__ void cp_async16_plain(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <bool GATHER, int NB>
__global__ void __launch_bounds__(MM_T, 3)
hm_panel_mm_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                   const double *__restrict__ W, const double *__restrict__ Xt,
                   const double *__restrict__ Sp, double *__restrict__ out, int accumulate)
{
    constexpr int CS = NB * 8, ZP = CS + 8;
    constexpr int STAGE = MM_KC * (MM_WP + ZP);
    constexpr int NBW = NB >= 4 ? NB / 2 : NB; // column blocks per warp in the 2 x 2 warp grid
    extern __shared__ __align__(16) double dsm[];
    __shared__ int zrow[GATHER ? HM_SMAX : 1];
    __shared__ int rpos[GATHER ? HM_MAXRUNS + 1 : 1];
    __shared__ int rsrc[GATHER ? HM_MAXRUNS : 1];

    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int S = it.S, F = it.F, Fp = it.Fp;

    if (GATHER) {
        for (int r = t; r < it.nrun; r += MM_T) {
            HmRun rr = runs[it.run0 + r];
            rpos[r] = rr.pos;
            rsrc[r] = rr.src;
        }
        if (t == 0) rpos[it.nrun] = S;
        __syncthreads();
        for (int e = t; e < S; e += MM_T) {
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
    const int nchunks = (S + MM_KC - 1) / MM_KC;

    for (int f0 = 0; f0 < Fp; f0 += MM_MT) {
        const int mt = min(MM_MT, Fp - f0); // even
        const int nfb = (mt + 7) >> 3;      // 8-row blocks in this pass (<= 8)
        // warp grid: rows beyond 32 exist -> 2 (rows) x 2 (columns); else 1 x 4
        const bool tall = nfb > 4;
        const int wr = tall ? (warp >> 1) : 0;                 // row half
        const int wc = tall ? (warp & 1) : warp;               // column group
        const int nbw = tall ? NBW : max(1, NB / 4);           // column blocks of this warp
        const int n0 = wc * nbw;
        const int fb0 = wr * 4;
        const int na = min(4, nfb - fb0);                      // row blocks of this warp (<= 0: idle)
        const bool active = na > 0 && n0 < NB;

        double acc[4][NBW][2];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int n = 0; n < NBW; n++) acc[a][n][0] = acc[a][n][1] = 0.0;

        // Each thread copies the same pieces of every chunk: up to 4 sixteen-byte pieces of W
        // (row wr, column pair wp of the pass) and CS/64 * 4 of Z.  Their offsets are computed
        // once per pass; per chunk only the base moves (keeps the copy code off the
        // instruction budget: it used to cost 4x the instructions of the math).
        const int hw = mt >> 1;
        int w_s[4], w_r[4];
        long long w_g[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int idx = t + j * MM_T;
            const int r = idx / hw, pp = idx - r * hw;
            w_r[j] = r < MM_KC ? r : MM_KC; // MM_KC: no such piece
            w_s[j] = r * MM_WP + 2 * pp;
            w_g[j] = (long long)r * Fp + f0 + 2 * pp;
        }
        constexpr int hz = CS / 2;                    // 16-byte pieces per z row
        constexpr int ZPT = (MM_KC * hz + MM_T - 1) / MM_T; // z pieces per thread
        auto issue = [&](int ch) {
            double *Wsm = dsm + (size_t)(ch % MM_ST) * STAGE;
            double *Zsm = Wsm + MM_KC * MM_WP;
            const int s0 = ch * MM_KC, rows = min(MM_KC, S - s0);
            const double *wsrc = Wg + (size_t)s0 * Fp;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (w_r[j] < rows) cp_async16_plain(Wsm + w_s[j], wsrc + w_g[j]);
#pragma unroll
            for (int j = 0; j < ZPT; j++) {
                const int idx = t + j * MM_T;
                const int r = idx / hz, pp = idx % hz; // hz is a compile-time power of two
                if (r < rows) {
                    const double *src;
                    if (GATHER) {
                        const int zr = zrow[s0 + r];
                        src = zr >= 0 ? Xt + (size_t)zr * CS : Sp + (size_t)(~zr) * CS;
                    } else {
                        src = Xt + (size_t)(it.zoff + s0 + r) * CS;
                    }
                    cp_async16_plain(Zsm + r * ZP + 2 * pp, src + 2 * pp);
                }
            }
            if (rows < MM_KC) { // zero the tail rows of the last chunk: 0 * stale NaN would poison the sums
                for (int idx = t; idx < (MM_KC - rows) * MM_WP; idx += MM_T) Wsm[rows * MM_WP + idx] = 0.0;
                for (int idx = t; idx < (MM_KC - rows) * ZP; idx += MM_T) Zsm[rows * ZP + idx] = 0.0;
            }
        };

        __syncthreads(); // the previous pass is done with the stages
        for (int ch = 0; ch < MM_ST - 1; ch++) {
            if (ch < nchunks) issue(ch);
            cp_async_commit_group();
        }
        for (int ch = 0; ch < nchunks; ch++) {
            cp_async_wait_group<MM_ST - 2>(); // chunk ch has landed (this thread's pieces)
            __syncthreads();                  // ... everyone's; and stage (ch-1) % ST is free again
            if (ch + MM_ST - 1 < nchunks) issue(ch + MM_ST - 1);
            cp_async_commit_group();
            if (active) {
                const double *Wsm = dsm + (size_t)(ch % MM_ST) * STAGE;
                const double *Zsm = Wsm + MM_KC * MM_WP;
                const double *ap = Wsm + tig * MM_WP + fb0 * 8 + gid;
                const double *bp = Zsm + tig * ZP + n0 * 8 + gid;
                if (na == 4 && nbw == NBW) { // full tile: no masking in the inner loop
#pragma unroll
                    for (int ks = 0; ks < MM_KC / 4; ks++) {
                        double a[4], b[NBW];
#pragma unroll
                        for (int i = 0; i < 4; i++) a[i] = ap[ks * 4 * MM_WP + i * 8];
#pragma unroll
                        for (int n = 0; n < NBW; n++) b[n] = bp[ks * 4 * ZP + n * 8];
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int n = 0; n < NBW; n++) dmma884(acc[i][n][0], acc[i][n][1], a[i], b[n]);
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < MM_KC / 4; ks++) {
                        double a[4], b[NBW];
#pragma unroll
                        for (int i = 0; i < 4; i++) a[i] = i < na ? ap[ks * 4 * MM_WP + i * 8] : 0.0;
#pragma unroll
                        for (int n = 0; n < NBW; n++) b[n] = n < nbw ? bp[ks * 4 * ZP + n * 8] : 0.0;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            if (i < na) {
#pragma unroll
                                for (int n = 0; n < NBW; n++)
                                    if (n < nbw) dmma884(acc[i][n][0], acc[i][n][1], a[i], b[n]);
                            }
                        }
                    }
                }
            }
        }
        cp_async_wait_group<0>();

        if (active) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int f = f0 + (fb0 + i) * 8 + gid;
                if (i < na && f < F) {
                    double2 *gp = reinterpret_cast<double2 *>(out + (size_t)(it.out + f) * CS) + n0 * 4 + tig;
#pragma unroll
                    for (int n = 0; n < NBW; n++) {
                        if (n < nbw) {
                            double2 v = make_double2(acc[i][n][0], acc[i][n][1]);
                            if (GATHER && accumulate) {
                                double2 o = gp[n * 4];
                                v.x += o.x;
                                v.y += o.y;
                            }
                            gp[n * 4] = v;
                        }
                    }
                }
            }
        }
    }
}

template <bool GATHER, int NB>
cudaError_t launch_panel_mm(const HmItem *items, int64_t nitems, const HmRun *runs, const double *W,
                            const double *Xt, const double *Sp, double *out, int accumulate, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    const size_t smem = (size_t)MM_ST * MM_KC * (MM_WP + NB * 8 + 8) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(hm_panel_mm_kernel<GATHER, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    hm_panel_mm_kernel<GATHER, NB><<<(unsigned)nitems, MM_T, smem, st>>>(items, runs, W, Xt, Sp, out, accumulate);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// stage 2 on a panel: one CTA per low-rank leaf
//   T[k][c] = sum of the leaf's partial panels (column order); S = F T | Sigma .* T
// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(128)
hm_core_panel_kernel(const HmCoreBlock *__restrict__ blocks, const int32_t *__restrict__ plist,
                     const double *__restrict__ Pp, const double *__restrict__ core,
                     double *__restrict__ Sp, int max_r, int blocked)
{
    constexpr int CS = NB * 8, TP = CS + 4; // TP = 4 (mod 16): conflict-free B-fragment reads
    extern __shared__ double sm[];
    double *Tsm = sm;                         // [rv rounded up to 4][TP]
    const HmCoreBlock cb = blocks[blockIdx.x];
    const double *__restrict__ Fg = core + cb.core; // ru x rv (bary) or r (low rank); L1-resident
    (void)max_r;
    const int t = threadIdx.x, T = blockDim.x;
    const int32_t *pl = plist + cb.pl0;
    const int rv4 = (cb.rv + 3) & ~3;
    for (int e = t; e < rv4 * CS; e += T) {
        const int l = e / CS, c = e - l * CS;
        double a = 0.0;
        if (l < cb.rv) {
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
        }
        Tsm[l * TP + c] = a; // rows rv .. rv4-1 are zero: the K padding of the MMAs below
    }
    __syncthreads();
    double *o = Sp + (size_t)cb.soff * CS;
    if (cb.kind == HM_LEAF_LOWRANK) {
        for (int e = t; e < cb.ru * CS; e += T) {
            const int k = e / CS, c = e - k * CS;
            const double v = Tsm[k * TP + c] * __ldg(Fg + k);
            if (blocked)
                Sp[hm_panel_blocked_index(cb.soff + k, c, NB)] = v;
            else
                o[e] = v;
        }
    } else {
        // S = F T on the FP64 tensor cores: a warp owns 8-column blocks of the panel; A fragments
        // (F, column-major ru x rv) through L1, B fragments (T) from shared memory.  The plain-FMA
        // form needed two memory instructions per FMA and was bound by L1 wavefronts.
        const int lane = t & 31, warp = t >> 5, nw = T >> 5;
        const int gid = lane >> 2, tig = lane & 3;
        const int mb = (cb.ru + 7) >> 3, ks = rv4 >> 2;
        for (int n = warp; n < NB; n += nw) {
            for (int m = 0; m < mb; m++) {
                const int row = m * 8 + gid;
                double d0 = 0.0, d1 = 0.0;
                for (int k = 0; k < ks; k++) {
                    const int kk = k * 4 + tig;
                    const double a = (row < cb.ru && kk < cb.rv) ? __ldg(Fg + row + (size_t)kk * cb.ru) : 0.0;
                    dmma884(d0, d1, a, Tsm[kk * TP + n * 8 + gid]);
                }
                if (row < cb.ru) {
                    if (blocked) {
                        double *ob = Sp + hm_panel_blocked_index(cb.soff + row, n * 8 + 2 * tig, NB);
                        ob[0] = d0;
                        ob[4] = d1; // the next column of the 4 x 8 tile
                    } else {
                        *reinterpret_cast<double2 *>(o + (size_t)row * CS + n * 8 + 2 * tig) = make_double2(d0, d1);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// panel transposes (32 x 32 tiles through shared memory)
//   in : Xt[j][c] = c < nrhs ? X[j + c*ldx] : 0
//   out: Y[i + c*ldy] = (accumulate ? Y : 0) + Yt[i][c],  rows [r0, r1)
// ---------------------------------------------------------------------------
__global__ void hm_panel_in_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, int nrhs, int CS,
                                   double *__restrict__ Xt, int blocked)
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
        if (j < n && c < CS) Xt[blocked ? hm_panel_blocked_index(j, c, CS >> 3) : (size_t)(j * CS + c)] = tile[threadIdx.x][jj];
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

template <bool GATHER, int NB, int U>
cudaError_t launch_panel(const HmItem *items, int64_t nitems, const HmRun *runs, const double *W,
                         const double *Xt, const double *Sp, double *out, int accumulate, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    const size_t smem = (size_t)8 * 16 * (NB * 8 + 8) * sizeof(double); // split-K combine buffer
    { // the attribute is per device: set it on every call (cheap) rather than once per process
        cudaError_t e = cudaFuncSetAttribute(hm_panel_kernel<GATHER, NB, U>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    hm_panel_kernel<GATHER, NB, U><<<(unsigned)nitems, PT, smem, st>>>(items, runs, W, Xt, Sp, out, accumulate);
    return cudaGetLastError();
}

// shared memory of the stage-2 panel kernel: (max_r + 3) rows of NB*8 columns, pitch padded by 4
constexpr size_t HM_PANEL_SMEM_CAP = 160 * 1024;
inline size_t core_panel_smem(int max_r, int NB) { return (size_t)(max_r + 3) * (NB * 8 + 4) * sizeof(double); }

template <int NB>
cudaError_t launch_core_panel(const HmCoreBlock *blocks, int64_t nblocks, const int32_t *plist, const double *Pp,
                              const double *core, double *Sp, int max_r, cudaStream_t st, bool blocked)
{
    if (nblocks <= 0) return cudaSuccess;
    const size_t smem = core_panel_smem(max_r, NB);
    if (smem > HM_PANEL_SMEM_CAP) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(hm_core_panel_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    // small CTAs: the kernel is latency-bound, more leaves in flight per SM
    hm_core_panel_kernel<NB><<<(unsigned)nblocks, NB >= 8 ? 128 : 64, smem, st>>>(blocks, plist, Pp, core, Sp, max_r,
                                                                                  blocked ? 1 : 0);
    return cudaGetLastError();
}

} // namespace

int hm_panel_width(int nrhs) { return nrhs <= 16 ? 16 : nrhs <= 32 ? 32 : 64; }

// The panel kernels stage a leaf's rank in shared memory; beyond this the caller applies the
// columns one by one (the same formula the stage-2 launcher checks).
bool hm_panel_supports_rank(int max_r, int nrhs)
{
    return core_panel_smem(std::max(max_r, 1), hm_panel_width(nrhs) / 8) <= HM_PANEL_SMEM_CAP;
}

cudaError_t hm_launch_panel_in(const double *X, int64_t ldx, int64_t n, int nrhs, int CS, double *Xt,
                               cudaStream_t st, bool blocked)
{
    if (n <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((CS + 31) / 32)), block(32, 8);
    hm_panel_in_kernel<<<grid, block, 0, st>>>(X, ldx, n, nrhs, CS, Xt, blocked ? 1 : 0);
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

// Two implementations of the panel kernels are kept:
//   stream (default)  hm_panel_kernel: A fragments straight from HBM into registers, B through L1
//   tma               hm_panel_tma_kernel: warp-specialised cp.async.bulk + mbarrier pipeline
// Measured on one B200 at N = 2^20 (ms per product, 16 / 64 right-hand sides): stream 5.2 / 15.0,
// tma 11.4 / 16.6.  HMB200_PANEL=tma selects the second one.
// Which implementation runs: HMB200_PANEL = stream | tma | mm forces one for both stages;
// by default stage 1 (short slow dimension: 64-column slabs, rank-20 leaves) takes the
// register-streaming kernel and stage 3 (slow dimension ~1000-2000) the tiled GEMM --
// measured per stage at N = 2^20, ms for 16 / 64 columns:
//              stage 1        stage 3
//   stream   1.92 / 5.50    2.48 / 6.17
//   mm       2.64 / 6.39    2.10 / 5.19
//   tma      5.70 / 7.79    5.00 / 6.85
//   pipe     hm_panelm_kernel (hm_panel_mma.cu, HMB200_PANEL=pipe): 32 x 32 DMMA warp tiles behind a
//            multi-stage cp.async ring with conflict-free padded pitches.  Round 2, measured at N = 2^20:
//            12.0 ms (64 columns) and 5.7 ms (16) against 11.65 / 4.3 ms of the default pair -- its DMMA
//            pipe is 58-67 % active at 8-12 executed instructions per MMA (per-chunk barrier + copy
//            issue), i.e. the same plateau as the kernels above; not the default.
//   dmma     (default) the per-stage choice measured above: stage 1 register streaming, stage 3 tiled GEMM
static int panel_variant(int stage)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HMB200_PANEL");
        v = !e ? 3 : e[0] == 't' ? 1 : e[0] == 'm' ? 2 : e[0] == 's' ? 0 : e[0] == 'p' ? 4 : 3;
    }
    if (v == 3) return stage == 3 ? 2 : 0;
    return v;
}

cudaError_t hm_launch_panel_stage1(int CS, const HmItem *items, int64_t nitems, const double *vstream,
                                   const double *Xt, double *Pp, cudaStream_t st)
{
    const int variant = panel_variant(1);
    if (variant == 4) return hm_launch_panelm_stage1(CS, items, nitems, vstream, Xt, Pp, st);
    if (variant == 1) switch (CS) {
        case 16: return launch_panel_tma<false, 2>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
        case 32: return launch_panel_tma<false, 4>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
        case 64: return launch_panel_tma<false, 8>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
        default: return cudaErrorInvalidValue;
        }
    if (variant == 2) switch (CS) {
        case 16: return launch_panel_mm<false, 2>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
        case 32: return launch_panel_mm<false, 4>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
        case 64: return launch_panel_mm<false, 8>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
        default: return cudaErrorInvalidValue;
        }
    switch (CS) {
    case 16: return launch_panel<false, 2, 8>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
    case 32: return launch_panel<false, 4, 4>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
    case 64: return launch_panel<false, 8, 2>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_panel_stage2(int CS, const HmCoreBlock *blocks, int64_t nblocks, const int32_t *plist,
                                   const double *Pp, const double *core, double *Sp, int max_r, cudaStream_t st,
                                   bool blocked)
{
    switch (CS) {
    case 16: return launch_core_panel<2>(blocks, nblocks, plist, Pp, core, Sp, max_r, st, blocked);
    case 32: return launch_core_panel<4>(blocks, nblocks, plist, Pp, core, Sp, max_r, st, blocked);
    case 64: return launch_core_panel<8>(blocks, nblocks, plist, Pp, core, Sp, max_r, st, blocked);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_panel_stage3(int CS, const HmItem *items, int64_t nitems, const HmRun *runs,
                                   const double *ustream, const double *Xt, const double *Sp, double *Yt,
                                   int accumulate, int zcap, cudaStream_t st)
{
    const int variant = panel_variant(3);
    if (variant == 4) return hm_launch_panelm_stage3(CS, items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, zcap, st);
    if (variant == 1) switch (CS) {
        case 16: return launch_panel_tma<true, 2>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
        case 32: return launch_panel_tma<true, 4>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
        case 64: return launch_panel_tma<true, 8>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
        default: return cudaErrorInvalidValue;
        }
    if (variant == 2) switch (CS) {
        case 16: return launch_panel_mm<true, 2>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
        case 32: return launch_panel_mm<true, 4>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
        case 64: return launch_panel_mm<true, 8>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
        default: return cudaErrorInvalidValue;
        }
    switch (CS) {
    case 16: return launch_panel<true, 2, 8>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
    case 32: return launch_panel<true, 4, 4>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
    case 64: return launch_panel<true, 8, 2>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, st);
    default: return cudaErrorInvalidValue;
    }
}
