// hm_nest.h -- nested-basis form of a matrix-free plan (hm_assemble_kernel_free).
//
// Every low-rank leaf of KernelMatrix(f, x, y, a, b, c, d) (/root/reference/src/KernelMatrix.jl:49-116)
// interpolates f on a pair of boxes obtained from the root boxes by repeated exact halving
// (ab2 = half(T)*(a+b), indsplit: src/BarycentricMatrix.jl:299-307), always with the same 20
// Chebyshev nodes (src/BarycentricMatrix.jl:147-178).  Two consequences, both exact polynomial
// identities (degree <= 19 on the parent box is degree <= 19 on either half):
//   * the Chebyshev moments of a column box are a fixed 20 x 20 map of the moments of its two halves,
//     mu_q(parent) = sum_p M0[q][p] mu_p(first half) + M1[q][p] mu_p(second half),
//     T_q((eta -+ 1) / 2) = sum_p M0|1[q][p] T_p(eta);
//   * the Chebyshev coefficients that leaves deposit on a row box can be pushed to its halves with
//     the transposed maps and evaluated once per row at the finest box.
// So the 3 x depth leaves over a row (column) share ONE pass over the points: moments are formed
// from the points at the finest boxes only and translated upwards, cores act on boxes, coefficients
// are translated downwards and evaluated at the finest boxes.  The U and V factors of the
// reference (90 % of its bytes, (m + n) r words per leaf) are never formed; the work of the
// low-rank part drops from O(N r depth) to O(N r) + O(boxes r^2).  For the translation-invariant
// kernels of hm_assemble_kernel the r x r cores depend only on the relative position of the two
// boxes, so a few hundred distinct cores (computed in extended precision on the host) serve all leaves.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "hm_layout.h"
#include "hm_types.h"

#define HM_NEST_R 20       // BLOCKRANK(Float64)
#define HM_NEST_BASE 96    // boxes with more points than this are halved further
#define HM_NEST_TIER0 2048 // points of a bottom subtree ...
#define HM_NEST_GROWTH 16  // ... every tier above holds this many times more (measured: 1024 / 32 is 4 % slower)
#define HM_NEST_MAXTIERS 8

// One box of the row or column cluster tree.  40 bytes.
struct HmNestNode {
    double mid, ih;  // eta = (p - mid) * ih maps the box to [-1, 1]
    int32_t p0, np;  // its points (= rows / columns of the operator)
    int32_t child0;  // halves child0, child0 + 1; < 0: a finest box
    int32_t parent;  // < 0: root
    int32_t which;   // 0: first half of the parent, eta_parent = (eta - 1) / 2; 1: second, (eta + 1) / 2
    int32_t pad;
};

// A leaf of the operator as seen by stage 2: coefficients(row box) += G * moments(column box)
struct HmNestLeaf {
    int32_t core;  // index of the 20 x 20 core (column-major, C F C')
    int32_t cnode; // column box
};

struct HmNestTree {
    std::vector<HmNestNode> nodes;
    // schedule: boxes are tiered by point count; a maximal connected set of boxes of one tier is a
    // subtree handled by one CTA, one launch per tier.  order = box ids subtree after subtree (tier
    // after tier), inside a subtree by depth, deepest first; grp = starts of the runs of equal depth
    // (plus end); sub_g0 = first group of every subtree (plus end); tier_sub0 = first subtree of every
    // tier (plus end).
    std::vector<int32_t> order, grp, sub_g0, tier_sub0;
    std::vector<int32_t> base; // the finest boxes
    int max_group = 0;
};

struct HmNest {
    HmNestTree rows, cols;
    std::vector<int32_t> rleaf_begin; // per row box: its leaves [begin, end) in rleaf
    std::vector<HmNestLeaf> rleaf;
    std::vector<double> cores;        // distinct cores, 400 words each
    std::vector<double> M;            // M0, M1 (row-major [q][p]), then their transposes: 4 x 400 words
    // dense part: the stage-3 items of the layout with their low-rank runs removed
    std::vector<HmItem> items3;
    std::vector<HmRun> runs;
    std::vector<HmFreeRun> frun;
    std::vector<int64_t> round_begin;
    int zcap = 2;
    // finest row box of every item, when every item lies inside one and the items of the single round
    // cover all owned rows (then the evaluation of the low-rank part is fused into the dense pass)
    std::vector<int32_t> item_box;
    bool fused_eval = false;
    // many right-hand sides (fused_eval only): the same items with one more run each, the 20 completed
    // coefficients of the item's finest row box (rows 20 fin[box] .. of the fragment-major panel Sp)
    std::vector<int32_t> fin; // per row box: index among the finest boxes, or -1
    std::vector<HmItem> items3p;
    std::vector<HmRun> runsp;
    std::vector<HmFreeRun> frunp;
};

// Builds the nested form of a kernel-assembled layout.  x, y: the point sets; (a, b), (c, d): the root
// boxes; kernel_id 0..3.  Returns "" on success; a non-empty reason when the operator does not have
// the structure (the caller then keeps the plain matrix-free form).
std::string hm_nest_build(const HmLayout &L, const double *x, int64_t nx, const double *y, int64_t ny, double a,
                          double b, double c, double d, int kernel_id, const std::vector<HmFreeRun> &frun3,
                          HmNest &out);

// device side (hm_nest.cu)
struct HmNestDev {
    const HmNestNode *nodes = nullptr;
    const int32_t *order = nullptr, *grp = nullptr, *sub_g0 = nullptr;
    int ntiers = 0;
    int tier_sub0[HM_NEST_MAXTIERS + 1] = {}; // subtrees [tier_sub0[k], tier_sub0[k + 1]) form launch k
    int nnodes = 0;
    const int32_t *base = nullptr; // the finest boxes
    int nbase = 0;
};
// M: the two maps [q][p]; Mt: their transposes [p][q] (2 x 400 words each)
cudaError_t hm_launch_nest_up(const HmNestDev &T, const double *pts, const double *x, const double *Mt, double *MU,
                              cudaStream_t st);
cudaError_t hm_launch_nest_core(int nboxes, const int32_t *rleaf_begin, const HmNestLeaf *rleaf, const double *cores,
                                const double *MU, double *LAM, cudaStream_t st);
// before_finest: an event the launch of the finest tier (the one that evaluates into y) waits for
cudaError_t hm_launch_nest_down(const HmNestDev &T, const double *pts, const double *M, double *LAM, double *y,
                                int accumulate, int64_t row_begin, int64_t row_end, bool eval, cudaStream_t st,
                                cudaEvent_t before_finest = nullptr);
// evaluation of the finished series of every finest box at its rows (after hm_launch_nest_down(..., eval = false))
cudaError_t hm_launch_nest_eval(const HmNestDev &T, const double *pts, const double *LAM, double *y, int accumulate,
                                int64_t row_begin, int64_t row_end, cudaStream_t st);
// the dense leaves (items with dense runs only, F <= 128 rows): y (+)= K(rows, columns of the runs) x;
// with ibox (finest row box of every item) the low-rank part is evaluated in the same pass
struct HmPeers;
cudaError_t hm_launch_nest_dense(const HmItem *items, int64_t nitems, const HmRun *runs, const HmFreeRun *frun,
                                 const double *px, const double *py, const double *x, double *y, int accumulate,
                                 int kernel_id, const int32_t *ibox, const HmNestNode *nodes, const double *LAM,
                                 const HmPeers *peers, cudaStream_t st);

// many right-hand sides (hm_nest_panel.cu); MUp / LAMp: 20 x CS words per box, row-major; Xt, Sp fragment-major
cudaError_t hm_launch_nest_up_panel(int CS, const HmNestDev &T, const double *pts, const double *Xt, const double *M,
                                    double *MUp, cudaStream_t st);
cudaError_t hm_launch_nest_core_panel(int CS, int nboxes, const int32_t *rleaf_begin, const HmNestLeaf *rleaf,
                                      const double *cores, const double *MUp, double *LAMp, cudaStream_t st);
cudaError_t hm_launch_nest_down_panel(int CS, const HmNestDev &T, const int32_t *fin, const double *M, double *LAMp,
                                      double *Sp, cudaStream_t st);
