// hm_layout.h -- host planner: flattens a leaf list into packed streams + work items.
// Pure C++ (no CUDA), so it is testable on a machine without a GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "hm_types.h"

struct HmLayoutParams {
    int rmax = 128;    // max rows of a stage-3 item (atomic row segments longer than this are split)
    int cmax0 = 256;   // max columns of a joint stage-1 item (low-rank leaves with n < nbig)
    int nbig = 1024;   // low-rank leaves with n >= nbig get their own stage-1 items
    int cmax1 = 4096;  // max columns of such an item (must be <= smax)
    int smax = 4096;   // max z length of an item (shared-memory staging capacity)
    int maxruns = 256; // max runs of a stage-3 item
};

constexpr int HM_NCHUNK = 4;

struct HmLayout {
    int64_t nrows = 0, ncols = 0;
    int64_t row_begin = 0, row_end = 0; // owned rows
    int part = 0, nparts = 1;           // which block-row part this is
    // leaves of this part (those intersecting the owned rows), in walk order
    std::vector<HmLeaf> leaves;
    std::vector<int64_t> leaf_global; // index in the full leaf list
    // stage 1
    std::vector<HmItem> items1;
    std::vector<HmFill> fill1;
    std::vector<int32_t> s1ent; // core index of every (stage-1 item, leaf) entry (items1[i].run0/nrun)
    int64_t vstream_words = 0;
    int64_t partial_words = 0;
    // stage 2
    std::vector<HmCoreBlock> cores; // one per low-rank leaf of this part
    std::vector<int32_t> core_leaf; // leaf index (into `leaves`) of each core block
    std::vector<int32_t> plist;     // partial-sum offsets
    int64_t core_words = 0;
    int64_t s_words = 0;
    int max_r = 0;
    // stage 3 (items sorted by round, then by size descending)
    std::vector<HmItem> items3;
    std::vector<HmRun> runs;
    std::vector<HmFill> fill3;
    std::vector<int64_t> round_begin; // items3 index of each round start, plus end
    int64_t ustream_words = 0;
    // host-pointer path (hm_matvec): the same items regrouped so that copies overlap compute --
    // stage-1 items by the last chunk of x they read, stage-3 items (single round only) by
    // the row chunk of y they write; HM_NCHUNK chunks each, sizes descending inside a chunk
    std::vector<HmItem> items1c, items3c;
    std::vector<int64_t> c1_begin, c3_begin; // item ranges per chunk (NCHUNK + 1 entries)
    std::vector<int64_t> xchunk, ychunk;     // column / row boundaries of the chunks (NCHUNK + 1)
    // adjoint apply y = H' x: the same two streams, reduced over the fast index
    //   A'  q = (row dots of every U-stream slab row with x)      -> PQ[item3.aux + s]
    //   B'  t'_b = sum of the leaf's q pieces; s'_b = F_b' t'_b | Sigma_b .* t'_b
    //   C'  r = (row dots of every V-stream slab row with s')     -> PQ[item1.aux + s]
    //   D'  y[j] = sum of the q (dense tiles) and r entries that belong to column j
    int64_t pq_words = 0;
    std::vector<int32_t> qlist;            // per core: PQ offsets of its q pieces, in row order
    std::vector<int32_t> core_q0, core_qn; // parallel to `cores`
    std::vector<HmColSeg> colsegs;
    std::vector<int64_t> colbases;
    int adj_max_f = 0; // widest stage-1 item (the z staging of stage C')
    // accounting
    int64_t n_dense = 0, n_lowrank = 0, n_bary2d = 0; // whole operator
    int64_t dense_words = 0, lowrank_words = 0, core_words_all = 0;
    int64_t part_words = 0; // unpadded words this part stores
    int64_t part_v_words = 0, part_core_words = 0, part_u_words = 0, part_dense_words = 0;
};

// Row cut points [0 = c_0 <= c_1 <= ... <= c_nparts = nrows], on block-row
// boundaries, minimising the largest number of words any part streams per matvec: its U
// rows and dense tile rows, plus the whole V and core of every low-rank leaf it touches
// (a leaf that straddles a cut has V and F replicated on both sides), plus x and its y rows.
std::vector<int64_t> hm_partition_rows(const std::vector<HmLeaf> &leaves, int64_t nrows,
                                       int nparts);

// Fault injection for the C ABI's exception barrier (tests only): after hm_fault_arm(n) the
// n-th checkpoint passed on the calling thread (n = 0: the next one) throws std::bad_alloc, as a
// failed container allocation at that point would; n < 0 disarms.  Checkpoints sit at the
// allocation-heavy steps of the planner and the tree builder.
void hm_fault_arm(int64_t nth);
void hm_fault_checkpoint();

// Returns "" on success, else an error message.
std::string hm_build_layout(const std::vector<HmLeaf> &all_leaves, int64_t nrows, int64_t ncols,
                            int part, int nparts, const HmLayoutParams &prm, HmLayout &out);

// HMB200_PLAN_TRACE=1: time since the previous trace point, on stderr (plan-time diagnostics)
void hm_trace_point(const char *name);
