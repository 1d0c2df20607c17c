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

constexpr int PT = 256;   // threads per CTA
constexpr int RD = 8;     // depth of the per-warp A ring (k-steps in flight)
constexpr int RP = 24;    // ring row pitch in doubles: = 8 (mod 16) -> conflict-free fragment reads
constexpr int RSLOT = 4 * RP; // one k-step: 4 slab rows x 16 columns

// 16-byte async copy global -> shared; bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int bytes)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
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
// stage 1 / stage 3 on a panel.
//
// The fast dimension is cut into tiles of 16 rows (two 8-row MMA blocks); a warp owns a
// tile and streams its 128-byte row pieces of the slab from HBM through a private
// cp.async ring (RD k-steps = 4 KB in flight per warp, no block-wide barrier in the main
// loop).  B fragments (the z rows, shared by all warps of the CTA) come through L1 with
// 128-bit loads: lane (g, t) reads columns 16j+2g, 16j+2g+1 of row t and uses them as the
// B values of MMA column blocks 2j and 2j+1, i.e. the 64 panel columns are permuted
// inside each group of 16 so that one load feeds two MMAs and a lane ends up holding four
// consecutive output columns.  Items with fewer than 8 tiles split the slow dimension
// over warp groups, combined in a fixed order through shared memory.
// ---------------------------------------------------------------------------
template <bool GATHER, int NB>
__global__ void __launch_bounds__(PT, 2)
hm_panel_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                const double *__restrict__ W, const double *__restrict__ Xt,
                const double *__restrict__ Sp, double *__restrict__ out, int accumulate)
{
    constexpr int CS = NB * 8, CP = CS + 8, NP = NB / 2;
    extern __shared__ __align__(16) double dsm[]; // A rings [8 warps][RD][RSLOT]; later the combine buffer
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
    const int nft = (nfb + 1) >> 1; // 16-row tiles
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
    const int nk = (s_hi - s_lo + 3) >> 2; // k-steps of this warp
    double *ring = dsm + (size_t)warp * RD * RSLOT;
    // this lane's 16-byte piece of a k-step: slab row (lane / 8), columns 2*(lane % 8), +1 of the tile
    const int prow = lane >> 3, pcol = (lane & 7) * 2;

    auto z_row = [&](int row) -> const double * {
        if (GATHER) {
            int zr = zrow[row];
            return zr >= 0 ? Xt + (size_t)zr * CS : Sp + (size_t)(~zr) * CS;
        }
        return Xt + (size_t)(it.zoff + row) * CS;
    };

    for (int ft = warp % ntw; ft < nft && active; ft += ntw) {
        const bool two = ft * 2 + 1 < nfb;
        const int fcol = ft * 16 + pcol;
        const bool colok = fcol < Fp;
        const double *__restrict__ wsrc = Wg + fcol;

        double acc[2][NB][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int n = 0; n < NB; n++) acc[a][n][0] = acc[a][n][1] = 0.0;

        auto issue = [&](int ks) {
            const int row = s_lo + 4 * ks + prow;
            const bool ok = colok && row < s_hi;
            cp_async16(ring + (ks % RD) * RSLOT + prow * RP + pcol, ok ? wsrc + (size_t)row * Fp : Wg, ok ? 16 : 0);
        };
        auto load_b = [&](int ks, double2 (&b)[NP]) {
            const int row = s_lo + 4 * ks + tig;
            if (row < s_hi) {
                const double2 *zp = reinterpret_cast<const double2 *>(z_row(row)) + gid;
#pragma unroll
                for (int j = 0; j < NP; j++) b[j] = __ldg(zp + 8 * j);
            } else {
#pragma unroll
                for (int j = 0; j < NP; j++) b[j] = make_double2(0.0, 0.0);
            }
        };

        __syncwarp();
        for (int ks = 0; ks < RD - 1; ks++) {
            if (ks < nk) issue(ks);
            cp_async_commit();
        }
        double2 bcur[NP], bnext[NP];
        if (nk > 0) load_b(0, bcur);
        for (int ks = 0; ks < nk; ks++) {
            __syncwarp(); // every lane is done reading the slot that is refilled next
            if (ks + RD - 1 < nk) issue(ks + RD - 1);
            cp_async_commit();
            if (ks + 1 < nk) load_b(ks + 1, bnext);
            cp_async_wait<RD - 1>();
            __syncwarp(); // the other lanes' pieces of this k-step have landed too
            const double *ap = ring + (ks % RD) * RSLOT + tig * RP + gid;
            const double a0 = ap[0];
            const double a1 = ap[8];
#pragma unroll
            for (int j = 0; j < NP; j++) {
                dmma884(acc[0][2 * j][0], acc[0][2 * j][1], a0, bcur[j].x);
                dmma884(acc[0][2 * j + 1][0], acc[0][2 * j + 1][1], a0, bcur[j].y);
                if (two) {
                    dmma884(acc[1][2 * j][0], acc[1][2 * j][1], a1, bcur[j].x);
                    dmma884(acc[1][2 * j + 1][0], acc[1][2 * j + 1][1], a1, bcur[j].y);
                }
            }
#pragma unroll
            for (int j = 0; j < NP; j++) bcur[j] = bnext[j];
        }
        cp_async_wait<0>();

        // C fragment of MMA column block 2j (2j+1) holds out columns 16j + 4t + {0,2} ({1,3})
        if (kgroups == 1) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int f = ft * 16 + a * 8 + gid;
                if (f < F && (a == 0 || two)) {
                    double2 *g = reinterpret_cast<double2 *>(out + (size_t)(it.out + f) * CS) + 2 * tig;
#pragma unroll
                    for (int j = 0; j < NP; j++) {
                        double2 v0 = make_double2(acc[a][2 * j][0], acc[a][2 * j + 1][0]);
                        double2 v1 = make_double2(acc[a][2 * j][1], acc[a][2 * j + 1][1]);
                        if (GATHER && accumulate) {
                            double2 o0 = g[8 * j], o1 = g[8 * j + 1];
                            v0.x += o0.x;
                            v0.y += o0.y;
                            v1.x += o1.x;
                            v1.y += o1.y;
                        }
                        g[8 * j] = v0;
                        g[8 * j + 1] = v1;
                    }
                }
            }
        } else {
            // nft < 8: every active warp owns exactly one tile and runs this body once.
            // The combine buffer [tile rows][CP] aliases the A rings: wait until all
            // active warps have drained theirs, then add the k-groups in a fixed order.
            const int nact = kgroups * ntw * 32;
            asm volatile("bar.sync 1, %0;\n" ::"r"(nact) : "memory");
            for (int g = 0; g < kgroups; g++) {
                if (kg == g) {
#pragma unroll
                    for (int a = 0; a < 2; a++) {
                        double *row = dsm + ((size_t)ft * 16 + a * 8 + gid) * CP + 4 * tig;
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            double *q = row + 16 * j;
                            if (g == 0) {
                                q[0] = acc[a][2 * j][0];
                                q[1] = acc[a][2 * j + 1][0];
                                q[2] = acc[a][2 * j][1];
                                q[3] = acc[a][2 * j + 1][1];
                            } else {
                                q[0] += acc[a][2 * j][0];
                                q[1] += acc[a][2 * j + 1][0];
                                q[2] += acc[a][2 * j][1];
                                q[3] += acc[a][2 * j + 1][1];
                            }
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;\n" ::"r"(nact) : "memory");
            }
        }
    }

    if (kgroups > 1) {
        __syncthreads();
        constexpr int hz = CS / 2;
        for (int idx = t; idx < F * hz; idx += PT) {
            int r = idx / hz, p = idx - r * hz;
            double2 v = *reinterpret_cast<const double2 *>(dsm + (size_t)r * CP + 2 * p);
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
    const size_t rings = (size_t)8 * RD * RSLOT * sizeof(double);             // 48 KB
    const size_t comb = (size_t)7 * 16 * (NB * 8 + 8) * sizeof(double);       // split-K combine buffer
    const size_t smem = rings > comb ? rings : comb;
    static bool configured = false; // per instantiation; static + dynamic smem exceeds 48 KB
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(hm_panel_kernel<GATHER, NB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    hm_panel_kernel<GATHER, NB><<<(unsigned)nitems, PT, smem, st>>>(items, runs, W, Xt, Sp, out, accumulate);
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
