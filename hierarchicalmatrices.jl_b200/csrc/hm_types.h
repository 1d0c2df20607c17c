// hm_types.h -- plain structs shared by the host planner and the CUDA kernels.
//
// Vocabulary:
//   leaf     one dense / LowRankMatrix / BarycentricMatrix2D block of the tree
//   stream   one contiguous packed array of Float64 words that a kernel reads
//            front to back: the V-stream (stage 1) and the U-stream (stage 3,
//            which also carries the dense tiles)
//   item     the work of one CTA: a slab W[S][Fp] of a stream (fast index f,
//            slow index s) applied as  out[f] = sum_s W[s][f] * z[s]
//   run      a contiguous piece of an item's z vector gathered from x or from
//            the stage-2 vector s
#pragma once
#include <cstdint>

enum HmLeafKind : int32_t { HM_LEAF_LOWRANK = 2, HM_LEAF_DENSE = 3, HM_LEAF_BARY2D = 4 };

enum HmLeafSource : int32_t {
    HM_SRC_NONE = 0,  // structure only (planner dry run)
    HM_SRC_COPY = 1,  // raw column-major copies staged on the device by the builder
    HM_SRC_KERNEL = 2 // evaluated on the device from the point sets (hm_assemble_kernel)
};

struct HmLeaf {
    int32_t kind;
    int32_t source;
    int64_t row0, col0, m, n;
    int32_t ru, rv; // dense: 0; LowRankMatrix: r, r; BarycentricMatrix2D: r, r
    // HM_SRC_COPY
    const double *dU; // dense: A
    const double *dC; // Sigma (r) or F (ru x rv)
    const double *dV;
    int64_t ldu, ldc, ldv;
    // HM_SRC_KERNEL
    int64_t xi0, yj0;  // first point index of the block's row / column range
    double a, b, c, d; // interpolation box of the block
    // words stored here beyond what the reference stores for this leaf (EvenBarycentricMatrix
    // is packed zero-interleaved at twice its rank); subtracted from the algorithmic count
    int64_t extra_words;
};

// One CTA's work.  48 bytes.
struct HmItem {
    int64_t slab; // word offset of W in its stream (multiple of 16)
    int64_t out;  // stage 1: word offset into the partial-sum array; stage 3: first row of y
    int32_t F;    // fast extent (stage 1: sum of ranks; stage 3: rows)
    int32_t Fp;   // F rounded up to even = leading dimension of W
    int32_t S;    // slow extent (stage 1: columns; stage 3: z length)
    int32_t zoff; // stage 1: first column of x
    int32_t run0; // stage 3: first run (stage 1: first entry of the leaf list s1ent)
    int32_t nrun;
    int64_t aux;  // adjoint apply: word offset of this item's S row-dot results in the buffer PQ
};

// Adjoint apply, final gather: y[j] = sum_i PQ[base_i + j] for j in [c0, c1)
struct HmColSeg {
    int32_t c0, c1;
    int32_t b0, nb; // bases [b0, b0 + nb) in the base list
};

struct HmRun {
    int32_t src; // >= 0: index into x; < 0: index ~src into the stage-2 vector
    int32_t len;
    int32_t pos; // position of the run inside the item's z
};

// Stage 2: one low-rank leaf.
struct HmCoreBlock {
    int64_t core; // word offset of Sigma / F (tight, ru x rv column-major) in the core array
    int32_t kind;
    int32_t ru, rv;
    int32_t soff; // offset of this block's ru outputs in the stage-2 vector
    int32_t pl0;  // first entry in the partial list
    int32_t npl;  // number of stage-1 partial sums to add (in column order)
};

// Slab filling (plan construction only): entry e of an item.
struct HmFill {
    int64_t dst;  // word offset in the stream of the entry's first element
    int32_t leaf; // index into the (device) leaf table
    int32_t off;  // stage 3: first row inside the leaf; stage 1: first column inside the leaf
    int32_t k0;   // first factor column (dense: matrix column) of the entry
    int32_t kn;   // number of factor columns
    int32_t F;    // stage 3: rows of the item; stage 1: unused
    int32_t Fp;   // leading dimension of the slab
    int32_t S;    // stage 1: columns of the item
    int32_t pad;
};

// Matrix-free apply: what a kernel needs about one stage-3 run / one (stage-1 item, leaf) entry,
// precomputed at plan time so that the kernels do not chase fill -> leaf records.
struct HmFreeRun {
    double mid, half; // low-rank run: node_k = mid + half * cheb_k (box of the leaf's rows)
    int64_t xoff;     // point index of the item's first row: leaf.xi0 + (item row - leaf.row0)
    int64_t yoff;     // dense run: point index of its first column: leaf.yj0 + k0
    int32_t k0, kn;   // low-rank: factor columns [k0, k0 + kn); dense: kn columns
};

struct HmFreeEnt {
    double mid, half; // box of the leaf's columns
    int64_t yoff;     // point index of the item's first column: leaf.yj0 + (item column - leaf.col0)
    int64_t fofs;     // offset of the leaf's ranks inside the item's partial sums
};

