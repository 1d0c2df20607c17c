// hm_panel_mma.cu -- many-right-hand-side stage 1 / stage 3: pipelined FP64 tensor-core GEMM.
//
// Same contraction as hm_panel.cu,  out[f][c] = sum_s W[s][f] * Z[s][c]  over one stream item,
// organised like a library GEMM main loop:
//   * the slab rows W[s][.] and the z rows are staged through a multi-stage cp.async ring in
//     shared memory (16-byte copies, L2 only: every slab byte comes from HBM once); row pitches
//     are padded to 4 (mod 8) words, which makes every m8n8k4 fragment read conflict-free
//     (lane (g, t) reads word (s0 + t) * pitch + f0 + g: the 16 lanes of a half-warp hit 16
//     different 8-byte banks);
//   * eight warps; a warp owns up to 4 x 4 MMA blocks (32 rows x 32 columns, 64 accumulator
//     registers): 8 fragment loads per 16 DMMA.8x8x4, so shared memory runs at a quarter of what
//     the tensor pipe needs (a register-tiled FMA kernel was built first and measured at 6.6
//     TFLOP/s: every DFMA operand crosses the 128 B/clk shared-memory port);
//   * tiles of up to 128 rows (64 columns) or 256 rows (16 / 32 columns); the 40-128-row segments
//     of stage 3 and the rank-20 slabs of the big leaves are single tiles whose spare warps split
//     the s range (k-groups), combined at the end in group order through the drained ring.
// One barrier per chunk of 8-32 slab rows.  Results are run-to-run deterministic.
#include <algorithm>
#include <type_traits>

#include "hm_kernels.cuh"

namespace {

constexpr int FT = 256;           // threads per CTA (8 warps)
constexpr int RING_WORDS = 12288; // 96 KB of stages
constexpr int CHUNK_WORDS = 2048; // slab words per chunk (about)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// all but the newest n groups of this thread have completed (n is uniform over the CTA)
__device__ __forceinline__ void cp_async_wait_dyn(int n)
{
    switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    default: cp_async_wait<6>(); break;
    }
}

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <bool GATHER, int CS>
__global__ void __launch_bounds__(FT, 2)
hm_panelm_kernel(const HmItem *__restrict__ items, const HmRun *__restrict__ runs,
                 const double *__restrict__ W, const double *__restrict__ Xt,
                 const double *__restrict__ Sp, double *__restrict__ out, int accumulate)
{
    constexpr int NB = CS / 8;            // 8-column MMA blocks
    constexpr int NR = NB < 4 ? NB : 4;   // column blocks per warp
    constexpr int NCW = NB / NR;          // warps side by side over the columns (2 at 64 columns)
    constexpr int MR = 4;                 // row blocks per warp (at most)
    constexpr int NRW_MAX = 8 / NCW;      // warps over the rows
    constexpr int TWMAX = NRW_MAX * MR * 8; // rows of a full tile: 128 (64 columns) or 256
    constexpr int PZ = CS + 4;            // z row pitch
    extern __shared__ __align__(16) double ring[]; // RING_WORDS words, then zrow[] (GATHER)
    __shared__ int rpos[GATHER ? HM_MAXRUNS + 1 : 1];
    __shared__ int rsrc[GATHER ? HM_MAXRUNS : 1];
    int *zrow = reinterpret_cast<int *>(ring + RING_WORDS);

    const HmItem it = items[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int S = it.S, F = it.F, Fp = it.Fp;
    if (S <= 0 || F <= 0) return;

    // Sp lies behind Xt in one allocation (hm_matmat_device): one base, 32-bit row indices
    const int sp_row0 = GATHER ? (int)((Sp - Xt) / CS) : 0;
    if (GATHER) {
        // source row of every z entry in the combined panel: rows of Xt, then rows of Sp
        for (int r = t; r < it.nrun; r += FT) {
            HmRun rr = runs[it.run0 + r];
            rpos[r] = rr.pos;
            rsrc[r] = rr.src;
        }
        if (t == 0) rpos[it.nrun] = S;
        __syncthreads();
        for (int e = t; e < S; e += FT) {
            int lo = 0, hi = it.nrun;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (rpos[mid] <= e)
                    lo = mid;
                else
                    hi = mid;
            }
            int src = rsrc[lo], off = e - rpos[lo];
            zrow[e] = src >= 0 ? src + off : sp_row0 + (~src) + off;
        }
    }

    // ---- tiling of this item: ntiles tiles of TW rows (multiple of 8), all but the last full ----
    const int ntiles = (Fp + TWMAX - 1) / TWMAX;
    const int TW = (((Fp + ntiles - 1) / ntiles) + 7) & ~7;
    const int PW = TW + 4;                       // slab row pitch, 4 (mod 8)
    int KC = (CHUNK_WORDS / PW) & ~3;
    KC = KC < 8 ? 8 : KC > 32 ? 32 : KC;
    const int nchunks = (S + KC - 1) / KC;
    const int NQ = ntiles * nchunks;
    const int stage_words = KC * (PW + PZ);
    int nst = RING_WORDS / stage_words;          // >= 3
    nst = nst > 8 ? 8 : nst;

    // ---- warp roles ----
    const int MB = TW >> 3;                      // row blocks of a tile
    const int nrw = (MB + MR - 1) / MR;          // warps over the rows (<= NRW_MAX)
    const int mrw = (MB + nrw - 1) / nrw;        // row blocks per warp
    const int wpg = nrw * NCW;                   // warps per k-group
    const int kgroups = 8 / wpg;
    const int kgrp = warp / wpg;
    const int ww = warp - kgrp * wpg;
    // warps w and w + 4 share an SM sub-partition: consecutive k-groups take the row groups in rotated
    // order so that a sub-partition does not collect only the short (or only the full) row groups
    const int cw = ww % NCW, rw = (ww / NCW + kgrp) % nrw;
    const int rb0 = rw * mrw;
    const int nmr = kgrp < kgroups ? max(0, min(mrw, MB - rb0)) : 0; // row blocks of this warp (0: idle)

    // ---- copy roles: a thread owns one 16-byte column unit and every RPP-th row of a chunk, so
    // that its source and destination advance by constant strides (no index arithmetic per copy) ----
    const int UPW = TW <= 32 ? 16 : TW <= 64 ? 32 : TW <= 128 ? 64 : 128; // units per slab row, padded to a power of two
    const int cuW = t & (UPW - 1), rW = t / UPW, RPPW = FT / UPW;
    constexpr int ZU = CS / 2, RPPZ = FT / ZU;
    const int cuZ = t & (ZU - 1), rZ = t / ZU;
    const double *__restrict__ Wg = W + it.slab;
    const size_t wsrc_step = (size_t)RPPW * Fp;
    const int wdst_step = RPPW * PW;

    // issue state (chunk q_i = tile ti, chunk ci of the tile, ring stage si)
    int qi = 0, ti = 0, ci = 0, si = 0;
    auto issue = [&]() {
        if (qi < NQ) {
            const int s0 = ci * KC, kc = min(KC, S - s0);
            double *sw = ring + si * stage_words;
            double *sz = sw + KC * PW;
            const int f0 = ti * TW;
            if (f0 + 2 * cuW < Fp && 2 * cuW < TW) {
                const double *src = Wg + (size_t)(s0 + rW) * Fp + f0 + 2 * cuW;
                double *dst = sw + rW * PW + 2 * cuW;
#pragma unroll 4
                for (int r = rW; r < kc; r += RPPW, src += wsrc_step, dst += wdst_step) cp_async16(dst, src);
            }
            {
                double *dst = sz + rZ * PZ + 2 * cuZ;
                const double *zbase = Xt + 2 * cuZ;
                if (GATHER) {
                    const int *zr = zrow + s0;
#pragma unroll 4
                    for (int r = rZ; r < kc; r += RPPZ, dst += RPPZ * PZ) cp_async16(dst, zbase + (size_t)zr[r] * CS);
                } else {
                    const double *zsrc = zbase + (size_t)(it.zoff + s0 + rZ) * CS;
#pragma unroll 4
                    for (int r = rZ; r < kc; r += RPPZ, dst += RPPZ * PZ, zsrc += RPPZ * CS) cp_async16(dst, zsrc);
                }
            }
            if (kc & 3) {
                // ragged end of the s range: the MMA k-step reads whole groups of four rows
                const int k4 = (kc + 3) & ~3;
                for (int i = t; i < (k4 - kc) * PW; i += FT) sw[kc * PW + i] = 0.0;
                for (int i = t; i < (k4 - kc) * PZ; i += FT) sz[kc * PZ + i] = 0.0;
            }
            qi++;
            if (++ci == nchunks) {
                ci = 0;
                ti++;
            }
            if (++si == nst) si = 0;
        }
        cp_async_commit();
    };

    double acc[MR][NR][2];
#pragma unroll
    for (int i = 0; i < MR; i++)
#pragma unroll
        for (int j = 0; j < NR; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (GATHER) __syncthreads(); // zrow complete before the first z copies
    // prologue: nst - 1 chunks in flight, one commit group per chunk (chunk q = group q)
    for (int q = 0; q < nst - 1; q++) issue();

    // rows gid of the warp's blocks, columns 2 tig, 2 tig + 1 of its column blocks
    auto store_tile = [&](int tile) {
#pragma unroll
        for (int i = 0; i < MR; i++) {
            const int f = tile * TW + (rb0 + i) * 8 + gid;
            if (i < nmr && f < F) {
                double *o = out + (size_t)(it.out + f) * CS + (cw * NR) * 8 + 2 * tig;
#pragma unroll
                for (int j = 0; j < NR; j++) {
                    double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                    double2 *p = reinterpret_cast<double2 *>(o + j * 8);
                    if (GATHER && accumulate) {
                        const double2 old = *p;
                        v.x += old.x;
                        v.y += old.y;
                    }
                    *p = v;
                }
            }
        }
    };

    // the warp's MMAs over one chunk, NMR row blocks (compile-time: no predicated MMAs)
    auto chunk_mma = [&](auto nmr_c, const double *wa, const double *zb, int nk4) {
        constexpr int NMR = decltype(nmr_c)::value;
        const int wstep = kgroups * 4 * PW, zstep = kgroups * 4 * PZ;
#pragma unroll 2
        for (int k4 = kgrp; k4 < nk4; k4 += kgroups, wa += wstep, zb += zstep) {
            double a[NMR], b[NR];
#pragma unroll
            for (int i = 0; i < NMR; i++) a[i] = wa[i * 8];
#pragma unroll
            for (int j = 0; j < NR; j++) b[j] = zb[j * 8];
#pragma unroll
            for (int i = 0; i < NMR; i++)
#pragma unroll
                for (int j = 0; j < NR; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    };

    const int woff = rb0 * 8 + gid + tig * PW + kgrp * 4 * PW;
    const int zoff = cw * NR * 8 + gid + tig * PZ + kgrp * 4 * PZ;
    int tc = 0, cc = 0, sc = 0; // compute state: tile, chunk of the tile, ring stage
    for (int q = 0; q < NQ; q++) {
        cp_async_wait_dyn(nst - 2); // this thread's copies of chunk q have landed
        __syncthreads();            // everybody's have; everybody is done with chunk q - 1
        issue();                    // refill the stage chunk q - 1 occupied
        const int nk4 = (min(KC, S - cc * KC) + 3) >> 2;
        const double *sw = ring + sc * stage_words;
        const double *sz = sw + KC * PW;
        switch (nmr) {
        case 4: chunk_mma(std::integral_constant<int, 4>{}, sw + woff, sz + zoff, nk4); break;
        case 3: chunk_mma(std::integral_constant<int, 3>{}, sw + woff, sz + zoff, nk4); break;
        case 2: chunk_mma(std::integral_constant<int, 2>{}, sw + woff, sz + zoff, nk4); break;
        case 1: chunk_mma(std::integral_constant<int, 1>{}, sw + woff, sz + zoff, nk4); break;
        default: break;
        }
        if (++sc == nst) sc = 0;
        if (++cc == nchunks) {
            cc = 0;
            if (kgroups == 1) {
                // every warp owns its outputs
                store_tile(tc);
#pragma unroll
                for (int i = 0; i < MR; i++)
#pragma unroll
                    for (int j = 0; j < NR; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            }
            tc++;
        }
    }
    if (kgroups > 1) {
        // a single narrow tile: combine the k-groups in group order through the (drained) ring
        cp_async_wait<0>();
        __syncthreads();
        constexpr int PER = MR * NR * 2; // words per thread
        double *red = ring;
        if (kgrp > 0 && kgrp < kgroups) {
            double *r = red + ((size_t)((kgrp - 1) * wpg + rw * NCW + cw) * 32 + lane) * PER;
#pragma unroll
            for (int i = 0; i < MR; i++)
#pragma unroll
                for (int j = 0; j < NR; j++)
                    *reinterpret_cast<double2 *>(r + (i * NR + j) * 2) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
        __syncthreads();
        if (kgrp == 0) {
            for (int g = 1; g < kgroups; g++) {
                const double *r = red + ((size_t)((g - 1) * wpg + rw * NCW + cw) * 32 + lane) * PER;
#pragma unroll
                for (int i = 0; i < MR; i++)
#pragma unroll
                    for (int j = 0; j < NR; j++) {
                        const double2 v = *reinterpret_cast<const double2 *>(r + (i * NR + j) * 2);
                        acc[i][j][0] += v.x;
                        acc[i][j][1] += v.y;
                    }
            }
            store_tile(0);
        }
    }
}

template <bool GATHER, int CS>
cudaError_t launch_panelm(const HmItem *items, int64_t nitems, const HmRun *runs, const double *W, const double *Xt,
                          const double *Sp, double *out, int accumulate, int zcap, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    const size_t smem = (size_t)RING_WORDS * sizeof(double) + (GATHER ? (size_t)zcap * sizeof(int) : 0);
    static size_t configured_smem = 0;
    if (smem > configured_smem) {
        cudaError_t e = cudaFuncSetAttribute(hm_panelm_kernel<GATHER, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        configured_smem = smem;
    }
    hm_panelm_kernel<GATHER, CS><<<(unsigned)nitems, FT, smem, st>>>(items, runs, W, Xt, Sp, out, accumulate);
    return cudaGetLastError();
}

} // namespace

cudaError_t hm_launch_panelm_stage1(int CS, const HmItem *items, int64_t nitems, const double *vstream,
                                    const double *Xt, double *Pp, cudaStream_t st)
{
    switch (CS) {
    case 16: return launch_panelm<false, 16>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, 0, st);
    case 32: return launch_panelm<false, 32>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, 0, st);
    case 64: return launch_panelm<false, 64>(items, nitems, nullptr, vstream, Xt, nullptr, Pp, 0, 0, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t hm_launch_panelm_stage3(int CS, const HmItem *items, int64_t nitems, const HmRun *runs,
                                    const double *ustream, const double *Xt, const double *Sp, double *Yt,
                                    int accumulate, int zcap, cudaStream_t st)
{
    zcap = (zcap + 3) & ~3;
    switch (CS) {
    case 16: return launch_panelm<true, 16>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, zcap, st);
    case 32: return launch_panelm<true, 32>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, zcap, st);
    case 64: return launch_panelm<true, 64>(items, nitems, runs, ustream, Xt, Sp, Yt, accumulate, zcap, st);
    default: return cudaErrorInvalidValue;
    }
}
