// hm_kernels.cuh -- launch wrappers of the sm_100a kernels (hm_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "hm_types.h"

#define HM_THREADS 256
#define HM_SMAX 4096    // z staging capacity of a stream item (words)
#define HM_MAXRUNS 256  // run-table capacity of a stage-3 item
#define HM_RMAX_ASM 32  // max interpolation rank of on-device assembly
#define HM_CORE_BIG 128 // stage 2: leaves with more partial sums than this get a whole CTA
#define HM_ADJ_BIG 32   // adjoint stage B': leaves with more q pieces than this get a whole CTA
#define HM_KERNEL_HOST_FN 4 // kernel id of hm_assemble_kernel_fn: f is a host callback

// Chebyshev nodes / barycentric weights of the reference's BarycentricPoly2D
// (src/BarycentricMatrix.jl:147-156), computed once on the host.
struct HmCheb {
    int r;
    double node[HM_RMAX_ASM];
    double lam[HM_RMAX_ASM];
};

// Stage 2 fused into the tail of stage 1 (see hm_kernels.cu): per-leaf arrival counters and
// what the core apply needs.  counters == nullptr: not fused.
struct HmFuse {
    const int32_t *s1ent = nullptr; // core index of every (stage-1 item, leaf) entry; item.run0/nrun index it
    int *counters = nullptr;        // one per low-rank leaf, zero between matvecs
    const HmCoreBlock *blocks = nullptr;
    const int32_t *plist = nullptr;
    const double *core = nullptr;
    double *svec = nullptr;
    int max_r = 0;
};

// Stage 3 fused with the all-gather of y: every owned row is also stored into the y buffers of
// the other ranks (peer-mapped device pointers, NVLink).  n == 0: plain local store.
#define HM_MAX_PEERS 16
struct HmPeers {
    double *y[HM_MAX_PEERS] = {};
    int n = 0;
};

// stage 1: partial[item.out + f] = sum_s V-slab[s][f] * x[item.zoff + s]
// pdl: launch with the programmatic-stream-serialization attribute (the kernel may start while the
// previous kernel of the stream still runs and synchronises itself); false where a launch follows
// an event wait on another stream's copy.
cudaError_t hm_launch_stage1(const HmItem *items, int64_t nitems, const double *vstream,
                             const double *x, double *partial, const HmFuse *fuse, cudaStream_t st, bool pdl = true);
// stage 2: s_b = F_b * (sum of partials) | Sigma_b .* (sum of partials)
cudaError_t hm_launch_stage2(const HmCoreBlock *blocks, int64_t nblocks, const int32_t *plist,
                             const double *partial, const double *core, double *svec, int max_r,
                             cudaStream_t st, bool pdl = true);
cudaError_t hm_launch_stage2_big(const HmCoreBlock *blocks, const int32_t *big, int64_t nbig,
                                 const int32_t *plist, const double *partial, const double *core,
                                 double *svec, int max_r, cudaStream_t st, bool pdl = true);
// stage 3: y[item.out + f] (+)= sum_s U-slab[s][f] * z[s],  z gathered from x and s
cudaError_t hm_launch_stage3(const HmItem *items, int64_t nitems, const HmRun *runs,
                             const double *ustream, const double *x, const double *svec, double *y,
                             int accumulate, const HmPeers *peers, cudaStream_t st, bool pdl = true);

// adjoint apply y (+)= H' x: four launches over the same streams (hm_kernels.cu)
struct HmAdjoint {
    const HmItem *items3 = nullptr, *items1 = nullptr;
    int64_t n3 = 0, n1 = 0, ncores = 0, nsegs = 0;
    const double *ustream = nullptr, *vstream = nullptr, *core = nullptr;
    const HmCoreBlock *blocks = nullptr;
    const int32_t *q0 = nullptr, *qn = nullptr, *qlist = nullptr, *s1ent = nullptr;
    const int32_t *big = nullptr; // leaves with more than HM_ADJ_BIG q pieces
    int nbig = 0;
    const HmColSeg *segs = nullptr;
    const int64_t *bases = nullptr;
    double *PQ = nullptr, *svec = nullptr;
    int max_r = 1;
};
cudaError_t hm_launch_adjoint(const HmAdjoint &A, const double *x, double *y, int accumulate, cudaStream_t st);
// out = -in (the adjoint of a matrix-free plan with an odd kernel applies K~(-x))
cudaError_t hm_launch_negate(const double *in, double *out, int64_t n, cudaStream_t st);

// operator updates in place: H <- Diagonal(b) H (rows) and H <- H Diagonal(b) (columns)
cudaError_t hm_launch_scale_rows(const HmItem *items3, int64_t n3, double *ustream, const double *b,
                                 cudaStream_t st);
cudaError_t hm_launch_scale_cols(const HmItem *items1, int64_t n1, double *vstream, const HmItem *items3,
                                 int64_t n3, const HmRun *runs, double *ustream, const double *b,
                                 cudaStream_t st);

// plan construction
cudaError_t hm_launch_fill3(const HmFill *fills, int64_t nfills, const HmLeaf *leaves, double *ustream,
                            const double *px, const double *py, const HmCheb &cheb, int kernel_id,
                            cudaStream_t st);
cudaError_t hm_launch_fill1(const HmFill *fills, int64_t nfills, const HmLeaf *leaves, double *vstream,
                            const double *py, const HmCheb &cheb, cudaStream_t st);
cudaError_t hm_launch_fillcore(const HmCoreBlock *blocks, const int32_t *core_leaf, int64_t nblocks,
                               const HmLeaf *leaves, double *core, const HmCheb &cheb, int kernel_id,
                               cudaStream_t st);

// matrix-free apply (the operator is evaluated on the fly from the point sets; hm_kernels.cu)
// Work split of a stage-1 item: every leaf's S columns are cut into nch chunks of CH columns so
// that the warps of the CTA have about two units each (host and device must agree).
#define HM_FREE1_THREADS 128
__host__ __device__ inline void hm_free1_split(int S, int nrun, int &nch, int &CH)
{
    constexpr int NW = HM_FREE1_THREADS / 32;
    int want = nrun > 1 ? (2 * NW + nrun - 1) / nrun : NW; // one leaf: one long chunk per warp
    if (want < 1) want = 1;
    CH = (S + want - 1) / want;
    CH = (CH + 31) & ~31;
    if (CH < 32) CH = 32;
    nch = S > 0 ? (S + CH - 1) / CH : 1;
}
// cheb_form: the plan's cores were transformed by hm_launch_core_cheb (Chebyshev-series kernels)
cudaError_t hm_launch_core_cheb(const HmCoreBlock *blocks, const int32_t *core_leaf, int64_t nblocks,
                                const HmLeaf *leaves, double *core, const double *Cm, const double *Dm,
                                const HmCheb &cheb, cudaStream_t st);
cudaError_t hm_launch_free1(const HmItem *items, int64_t nitems, const HmFreeEnt *ents, const double *py,
                            const double *x, double *partial, const HmCheb &cheb, int max_units, bool cheb_form,
                            cudaStream_t st);
cudaError_t hm_launch_free3(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun,
                            const double *px, const double *py, const double *x,
                            const double *svec, double *y, int accumulate, const HmCheb &cheb, int kernel_id,
                            const HmPeers *peers, int zcap, bool cheb_form, cudaStream_t st);

// many right-hand sides on a matrix-free plan in Chebyshev form (hm_free_panel.cu): stage 1 and stage 3
// with the operator generated in MMA fragment layout; stage 2 is hm_launch_panel_stage2
cudaError_t hm_launch_free1_panel(int CS, const HmItem *items, int64_t nitems, const HmFreeEnt *ents, const double *py,
                                  const double *Xt, double *Pp, cudaStream_t st);
cudaError_t hm_launch_free3_panel(int CS, const HmItem *items, int64_t nitems, const HmRun *runs,
                                  const HmFreeRun *frun, const double *px, const double *py, const double *Xt,
                                  const double *Sp, double *Yt, int accumulate, int kernel_id, cudaStream_t st);

// many right-hand sides (hm_panel.cu): panels are row-major with pitch CS = hm_panel_width(nrhs)
// blocked = true (matrix-free plans): Xt and Sp are kept "fragment-major" instead -- tiles of 4 rows x
// 8 columns (256 bytes) in the lane order of an m8n8k4 B fragment, tile rows first -- so that a warp's
// fragment load is one contiguous 256-byte piece (two L1 wavefronts instead of one per row and more)
__host__ __device__ inline size_t hm_panel_blocked_index(int64_t k, int c, int NB)
{
    return ((size_t)(k >> 2) * NB + (c >> 3)) * 32 + (size_t)((c & 7) * 4) + (size_t)(k & 3);
}
int hm_panel_width(int nrhs);
bool hm_panel_supports_rank(int max_r, int nrhs);
cudaError_t hm_launch_panel_in(const double *X, int64_t ldx, int64_t n, int nrhs, int CS, double *Xt,
                               cudaStream_t st, bool blocked = false);
cudaError_t hm_launch_panel_out(const double *Yt, int CS, int64_t r0, int64_t r1, int nrhs, double *Y,
                                int64_t ldy, int accumulate, cudaStream_t st);
cudaError_t hm_launch_panel_stage1(int CS, const HmItem *items, int64_t nitems, const double *vstream,
                                   const double *Xt, double *Pp, cudaStream_t st);
cudaError_t hm_launch_panel_stage2(int CS, const HmCoreBlock *blocks, int64_t nblocks, const int32_t *plist,
                                   const double *Pp, const double *core, double *Sp, int max_r, cudaStream_t st,
                                   bool blocked = false);
// zcap: largest S of the items (sizes the z row table of the pipelined kernel)
cudaError_t hm_launch_panel_stage3(int CS, const HmItem *items, int64_t nitems, const HmRun *runs,
                                   const double *ustream, const double *Xt, const double *Sp, double *Yt,
                                   int accumulate, int zcap, cudaStream_t st);
// pipelined DMMA implementation of the two panel stages (hm_panel_mma.cu)
cudaError_t hm_launch_panelm_stage1(int CS, const HmItem *items, int64_t nitems, const double *vstream,
                                    const double *Xt, double *Pp, cudaStream_t st);
cudaError_t hm_launch_panelm_stage3(int CS, const HmItem *items, int64_t nitems, const HmRun *runs,
                                    const double *ustream, const double *Xt, const double *Sp, double *Yt,
                                    int accumulate, int zcap, cudaStream_t st);
