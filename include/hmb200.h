/*
 * hmb200.h -- C ABI of the B200-native hierarchical-matrix matvec engine.
 *
 * This is the drop-in boundary for the `mul!` hot path of
 * JuliaLinearAlgebra/HierarchicalMatrices.jl (citations: /root/reference/...).
 * The reference has no FFI table for this path: it is Julia multiple dispatch
 * (src/KernelMatrix.jl:14-45, src/HierarchicalMatrix.jl:14-52).  Its only FFI
 * precedent is the (disabled) BLAS binding in src/blas.jl:6-14 -- column-major
 * arrays, explicit leading dimensions, start offsets turned into pointer
 * offsets, element strides, alpha = beta = 1 (accumulate).  The entry points
 * below keep exactly those conventions; a Julia shim (`ccall`) that overrides
 * the reference methods with them is in
 * hierarchicalmatrices.jl_b200/julia/HierarchicalMatricesB200.jl and
 * INTEGRATION.md.
 *
 * Conventions
 *   - all matrices column-major with explicit leading dimension, Float64;
 *   - row/column offsets are 0-based, strides are in elements;
 *   - every function returns an hm_status (0 = ok); hm_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - the library copies what it is given (device-resident snapshot) and never
 *     keeps a host pointer;
 *   - no CPU fallback exists: without a CUDA device every compute entry point
 *     fails with HM_ERR_CUDA.
 */
#ifndef HMB200_H
#define HMB200_H

#include <stdint.h>

#if defined(__GNUC__)
#define HM_API __attribute__((visibility("default")))
#else
#define HM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hm_builder hm_builder;
typedef struct hm_plan hm_plan;

typedef enum hm_status {
    HM_OK = 0,
    HM_ERR_INVALID = 1,     /* bad argument value */
    HM_ERR_NULL = 2,        /* null pointer */
    HM_ERR_SHAPE = 3,       /* negative extent, ld < rows, rank mismatch */
    HM_ERR_RANGE = 4,       /* block does not fit inside nrows x ncols */
    HM_ERR_STATE = 5,       /* call not valid in this object state */
    HM_ERR_NOMEM = 6,       /* host or device allocation failed */
    HM_ERR_CUDA = 7,        /* CUDA runtime error (or no device) */
    HM_ERR_UNSUPPORTED = 8, /* valid request this build cannot serve */
    HM_ERR_REFERENCE = 9    /* the reference itself would throw here (e.g. BoundsError) */
} hm_status;

enum { HM_F64 = 0 };

/* kernel ids of examples/Kernel.jl:34-37 */
enum { HM_KERNEL_CAUCHY = 0, HM_KERNEL_COULOMB = 1, HM_KERNEL_COULOMBPRIME = 2, HM_KERNEL_LOG = 3 };

/* What the roofline is computed from.  "words" are Float64 words. */
typedef struct hm_stats {
    int64_t nrows, ncols;
    int64_t n_dense, n_lowrank, n_bary2d;          /* leaves of the whole operator */
    int64_t dense_words;                           /* sum m*n */
    int64_t lowrank_words;                         /* sum (m+n)*r + (r*r | r) */
    int64_t core_words;                            /* the (r*r | r) part of the line above */
    int64_t algorithmic_bytes;                     /* 8*(dense+lowrank words) + 8*ncols + 8*nrows */
    /* this plan (one row part) */
    int64_t row_begin, row_end;                    /* owned rows [begin, end) */
    int64_t part_words;                            /* unpadded words this part stores */
    int64_t stored_bytes;                          /* device bytes of the packed streams (padded) */
    int64_t v_stream_bytes, u_stream_bytes;        /* stage-1 / stage-3 stream sizes (padded) */
    int64_t partial_bytes;                         /* stage-1 partial sums written per matvec */
    int64_t n_stage1_items, n_stage2_blocks, n_stage3_items, n_stage3_rounds;
    int64_t part_algorithmic_bytes;                /* 8*part_words + 8*ncols + 8*(row_end-row_begin) */
    /* split of part_words by the stage that reads them */
    int64_t part_v_words;                          /* stage 1: sum n*r */
    int64_t part_core_words;                       /* stage 2: sum r*r | r */
    int64_t part_u_words;                          /* stage 3: sum (owned rows)*r */
    int64_t part_dense_words;                      /* stage 3: sum (owned rows)*n */
} hm_stats;

HM_API const char *hm_last_error(void);
HM_API int32_t hm_version(void);

/* BLOCKRANK(Float64) / BLOCKSIZE(Float64): src/HierarchicalMatrices.jl:5-7 */
HM_API int32_t hm_blockrank_f64(void);
HM_API int32_t hm_blocksize_f64(void);

/* ------------------------------------------------------------------------
 * Builder: the flattened block tree.  The host-side planner walks `assigned`
 * (src/hierarchical.jl:49-52, codes :84-91) with the offset rule of
 * src/KernelMatrix.jl:24-41 / src/HierarchicalMatrix.jl:30-48 and pushes each
 * leaf here with its absolute 0-based (row0, col0).  Overlapping leaves are
 * legal: contributions add, as in the reference walk.
 *
 * device >= 0: leaf data is staged onto that CUDA device as it is added.
 * device = -1: structure-only builder (no data is read; pointers may be NULL);
 *              only hm_builder_layout_stats() can be called on it.
 * ------------------------------------------------------------------------ */
HM_API int32_t hm_builder_create(hm_builder **out, int64_t nrows, int64_t ncols, int32_t dtype,
                          int32_t device);
HM_API int32_t hm_builder_destroy(hm_builder *b);

/* Matrix leaf, code 3.  Replaces src/algebra.jl:37-48 (dgemv 'N' in src/blas.jl:6-14). */
HM_API int32_t hm_builder_add_dense(hm_builder *b, const double *A, int64_t m, int64_t n, int64_t lda,
                             int64_t row0, int64_t col0);
/* LowRankMatrix leaf U*Diagonal(S)*V' (no conjugation), code 2 of HierarchicalMatrix.
 * Replaces src/algebra.jl:110-131 (src/blas.jl:42-68). U m x r, V n x r. */
HM_API int32_t hm_builder_add_lowrank(hm_builder *b, const double *U, int64_t ldu, const double *S,
                               const double *V, int64_t ldv, int64_t m, int64_t n, int64_t r,
                               int64_t row0, int64_t col0);
/* BarycentricMatrix2D leaf U*F*V', code 2 of KernelMatrix.
 * Replaces src/algebra.jl:243-277 (src/blas.jl:72-104). U m x r, F r x r, V n x r. */
HM_API int32_t hm_builder_add_bary2d(hm_builder *b, const double *U, int64_t ldu, const double *F,
                              int64_t ldf, const double *V, int64_t ldv, int64_t m, int64_t n,
                              int64_t r, int64_t row0, int64_t col0);

/* EvenBarycentricMatrix leaf.  Replaces the apply of src/algebra.jl:168-239 on
 * the factors src/BarycentricMatrix.jl:18-45 builds: W r x m (ld ldw), F n x r (ld ldf); entry
 * (i, j) of the leaf is sum_k F[j,k] W[k,i] where shift_parity + row0 + i + col0 + j is even and 0
 * elsewhere.  shift_parity is the parity of (istart-1)+(jstart-1) of the mul! calls the plan will
 * serve (the reference decides the active class from the absolute offsets, algebra.jl:172).
 * Stored zero-interleaved as a rank-2r LowRankMatrix leaf; hm_stats counts it at (m+n) r words. */
HM_API int32_t hm_builder_add_evenbary(hm_builder *b, const double *W, int64_t ldw, const double *F,
                                int64_t ldf, int64_t m, int64_t n, int64_t r, int64_t row0,
                                int64_t col0, int32_t shift_parity);

/* Planner only (no GPU needed): lay the operator out for row part `part` of
 * `nparts` and report the sizes. */
HM_API int32_t hm_builder_layout_stats(hm_builder *b, int32_t part, int32_t nparts, hm_stats *out);

/* ------------------------------------------------------------------------
 * Plan: immutable packed operator on one device.
 * hm_plan_finalize      -- whole operator on devices[0].  ndev must be 1: a plan lives on one
 *                          GPU; the multi-GPU form is one process per GPU, each with
 *                          hm_plan_finalize_part + hm_dist_init (below)
 * hm_plan_finalize_part -- block-row part `part` of `nparts` (rows balanced by
 *                          stored bytes); y rows outside the part are not touched.
 * The builder can be destroyed afterwards.
 * ------------------------------------------------------------------------ */
HM_API int32_t hm_plan_finalize(hm_builder *b, const int32_t *devices, int32_t ndev, hm_plan **out);
HM_API int32_t hm_plan_finalize_part(hm_builder *b, int32_t part, int32_t nparts, hm_plan **out);
HM_API int32_t hm_plan_destroy(hm_plan *p);
HM_API int32_t hm_plan_stats(const hm_plan *p, hm_stats *out);

/* KernelMatrix(f, x, y, a, b, c, d) assembled on the device
 * (src/KernelMatrix.jl:47-116, src/BarycentricMatrix.jl:147-178, 236-307):
 * the host builds only the tree of index ranges, the device fills U, V, F and
 * the dense leaves straight into the packed streams.  x, y are host pointers,
 * sorted descending as the reference requires. */
HM_API int32_t hm_assemble_kernel(const double *x, int64_t nx, const double *y, int64_t ny, double a,
                           double b, double c, double d, int32_t kernel_id, int32_t device,
                           int32_t part, int32_t nparts, hm_plan **out);
/* KernelMatrix(f, x, y, a, b, c, d) for ANY kernel function (src/KernelMatrix.jl:47 takes any
 * `f::Function`): f is a host callback evaluating a batch, out[i] = f(x[i], y[i]), i < n.  Only the
 * r x r cores F[m,n] = f(x_m, y_n) (src/BarycentricMatrix.jl:159-175; 2.9 % of the bytes at
 * N = 2^20) and the dense leaves (src/KernelMatrix.jl:57-60; 9.3 %) depend on f: they are evaluated
 * through the callback in batches of ~2 M points and copied into the packed streams, while U and V
 * are filled on the device exactly as in hm_assemble_kernel.  The callback is used during this
 * call only. */
typedef void (*hm_kernel_fn)(const double *x, const double *y, int64_t n, double *out, void *user);
HM_API int32_t hm_assemble_kernel_fn(const double *x, int64_t nx, const double *y, int64_t ny, double a,
                              double b, double c, double d, hm_kernel_fn f, void *user, int32_t device,
                              int32_t part, int32_t nparts, hm_plan **out);

/* Matrix-free variant ("fused assemble + apply"): same arguments and the same
 * operator as hm_assemble_kernel, but U, V and the dense tiles are never stored -- hm_matvec /
 * hm_matvec_device / hm_matvec_device_allgather evaluate every entry from the point sets while
 * applying it (the arithmetic of src/BarycentricMatrix.jl:248-297 and src/KernelMatrix.jl:57-60).
 * The plan holds the r x r cores and the planner tables only, so an operator whose packed form
 * exceeds the GPU memory still fits; the apply is bound by the FP64 pipe instead of HBM.
 * hm_matmat works (panel kernels that generate the entries in tensor-core fragment layout), and so does
 * hm_matvec_adjoint on a whole operator (nparts = 1): the adjoint of a kernel operator is the kernel
 * operator of the transposed leaves with the point sets exchanged -- a second matrix-free plan built on
 * first use, applied to -x for the odd kernels.  hm_plan_scale and hm_plan_read_leaf return
 * HM_ERR_UNSUPPORTED.
 * When the points are in descending order (as KernelMatrix's indsplit assumes) the plan takes its
 * nested-basis form: all leaves interpolate on dyadic halves of the root boxes with the same 20 nodes,
 * so column moments are formed once at the finest boxes and translated up the box tree, coefficients are
 * translated down and evaluated once per row, and the few hundred distinct r x r cores of the
 * translation-invariant kernels are shared -- O(N r) work for the low-rank part instead of O(N r depth).
 * Same operator within rounding (<= 1e-12 from the reference's mul!). */
HM_API int32_t hm_assemble_kernel_free(const double *x, int64_t nx, const double *y, int64_t ny, double a,
                                double b, double c, double d, int32_t kernel_id, int32_t device,
                                int32_t part, int32_t nparts, hm_plan **out);

/* How the plan applies its operator: 0 = stored streams, 1 = matrix-free with the reference's barycentric
 * arithmetic, 2 = matrix-free in Chebyshev form leaf by leaf, 3 = matrix-free in nested-basis form.
 * HMB200_FREE_FORM=bary|cheb at plan time selects 1 or 2 instead of the default. */
HM_API int32_t hm_plan_form(const hm_plan *plan, int32_t *form);

/* One leaf of the assembled tree, as the planner sees it (device-free). */
typedef struct hm_tree_leaf {
    int32_t kind;          /* 3 = Matrix, 4 = BarycentricMatrix2D */
    int32_t rank;
    int64_t row0, col0, m, n; /* position in the operator (0-based) */
    int64_t xi0, yj0;      /* first point index of the row / column range */
    double a, b, c, d;     /* interpolation box (BarycentricMatrix.jl:147-167) */
} hm_tree_leaf;
/* Leaves of KernelMatrix(f, x, y, a, b, c, d) in the order and with the offsets of
 * the reference's mul! walk (src/KernelMatrix.jl:17-45).  Writes at most `cap`
 * records and always the total count. */
HM_API int32_t hm_kernel_tree_leaves(const double *x, int64_t nx, const double *y, int64_t ny,
                                     double a, double b, double c, double d, hm_tree_leaf *out,
                                     int64_t cap, int64_t *count);
/* Same tree, planner only (device-free): leaf counts and byte sizes. */
HM_API int32_t hm_assemble_kernel_stats(const double *x, int64_t nx, const double *y, int64_t ny,
                                 double a, double b, double c, double d, int32_t part,
                                 int32_t nparts, hm_stats *out);

/* ------------------------------------------------------------------------
 * mul!: y[i*incy] (+)= sum_j H[i,j] x[j*incx]     (accumulate != 0: +=, the
 * reference's mul!; accumulate == 0: =, the reference's `*`).
 * Replaces LinearAlgebra.mul!(u, H, v) -> mul!(u, H, v, 1, 1[, INCX, INCY])
 * (src/KernelMatrix.jl:14-45, src/HierarchicalMatrix.jl:14-52); the 1-based
 * istart/jstart of the reference become pointer offsets at the call site, as in
 * src/blas.jl:12.  Host pointers; copies are part of the call.
 * ------------------------------------------------------------------------ */
HM_API int32_t hm_matvec(hm_plan *p, const double *x, int64_t incx, double *y, int64_t incy,
                  int32_t accumulate);
/* Device pointers (contiguous), enqueued on `stream` (a cudaStream_t; NULL =
 * default stream); returns without synchronising.  The *_device entry points use the plan's
 * scratch buffers (partial sums, stage-2 vector, panel workspace) without locking: a plan admits
 * ONE in-flight device call at a time, all on one stream (or ordered by events); the host-pointer
 * entry points (hm_matvec, hm_matmat, ...) serialise themselves on the plan's mutex. */
HM_API int32_t hm_matvec_device(hm_plan *p, const double *dx, double *dy, int32_t accumulate,
                         void *stream);

/* Multi-GPU form of hm_matvec_device with the all-gather of y fused into stage 3: the rows
 * this plan owns are stored into all `npeers` y buffers (`ypeers[i]` = device address, valid
 * on this device, of rank i's full-length y: NVLink peer-mapped / symmetric memory) while
 * they are computed; `self` is this rank's index (its buffer is read when accumulating).
 * The caller synchronises the ranks afterwards (a device barrier) before anyone reads y. */
HM_API int32_t hm_matvec_device_allgather(hm_plan *p, const double *dx, const uint64_t *ypeers, int32_t npeers,
                                          int32_t self, int32_t accumulate, void *stream);

/* ------------------------------------------------------------------------
 * Multi-GPU mul!: one process per GPU, rank r holds block-row part r of nranks
 * (hm_plan_finalize_part / hm_assemble_kernel with part = r, nparts = nranks).  The plan owns the
 * exchange: an NCCL communicator (bound at run time, dlopen of libnccl.so.2) for the broadcast of
 * x, and one peer-mapped (cudaIpc, NVLink) exchange region per rank through which stage 3 stores
 * the rows it owns into the y buffer of every rank -- the all-gather of y is fused into the
 * kernel -- followed by a cross-rank barrier kernel.  After hm_dist_matvec* every rank holds the
 * whole y.
 *
 * hm_dist_get_id   one rank creates the 128-byte id (ncclGetUniqueId); the caller distributes it
 *                  to all ranks (MPI, torch.distributed, a file).
 * hm_dist_init     collective over all ranks.
 * x and y live in plan-owned, double-buffered device buffers ("slots" 0 / 1, hm_dist_buffers): a
 * call writing y slot s may overlap peers reading slot 1 - s; calls alternate slots.
 * ------------------------------------------------------------------------ */
#define HM_DIST_ID_BYTES 128
HM_API int32_t hm_dist_get_id(void *id_out);
HM_API int32_t hm_dist_init(hm_plan *p, const void *id, int32_t nranks, int32_t rank);
/* device addresses of this rank's x slots and (full-length, replicated) y slots */
HM_API int32_t hm_dist_buffers(hm_plan *p, double **x2, double **y2);
/* ncclBroadcast of x (ncols words) from `root` into x slot `slot` of every rank, enqueued on
 * `stream`.  dx_root: device pointer on the root (NULL = the root's slot already holds x). */
HM_API int32_t hm_dist_bcast_x(hm_plan *p, const double *dx_root, int32_t root, int32_t slot, void *stream);
/* The same replication without a collective kernel: the root's copy engines write x into x slot
 * `slot` of every rank through the peer mappings (one cudaMemcpyAsync per destination on internal
 * streams, forked from `stream`); no SM is taken from the matvec it overlaps with.  Completion is
 * joined into the NEXT barrier this plan enqueues on the root (hm_dist_matvec_device /
 * hm_dist_barrier): once that barrier has completed on a rank, its slot holds x.  So, pipelined:
 *   push_x(x[k+1] -> slot (k+1)&1);  matvec_device(x slot k&1 -> y slot k&1);  ...
 * Non-root ranks return immediately.  The root must not overwrite dx_root before that barrier. */
HM_API int32_t hm_dist_push_x(hm_plan *p, const double *dx_root, int32_t root, int32_t slot, void *stream);
/* y slot `yslot` (+)= H x on every rank: the three stages of this rank's part with the all-gather
 * fused into stage 3, then the barrier; enqueued on `stream`, no host synchronisation (may be
 * captured into a CUDA graph).  dx: any device vector of ncols words -- an x slot after
 * hm_dist_bcast_x, or the other y slot (x_{k+1} = y_k: a dependent iteration needs no broadcast). */
HM_API int32_t hm_dist_matvec_device(hm_plan *p, const double *dx, int32_t yslot, int32_t accumulate, void *stream);
HM_API int32_t hm_dist_barrier(hm_plan *p, void *stream);
/* HM_ERR_CUDA if a barrier timed out since hm_dist_init (a peer died); synchronous. */
HM_API int32_t hm_dist_check(hm_plan *p);
/* Host pointers: x (ncols, stride incx) is read on `root` only; y (nrows, stride incy) receives the
 * whole result on every rank that passes a non-NULL pointer.  Collective; copies are part of the call. */
HM_API int32_t hm_dist_matvec(hm_plan *p, const double *x, int64_t incx, double *y, int64_t incy,
                              int32_t root, int32_t accumulate);

/* Adjoint apply: y[j*incy] (+)= sum_i H[i,j] x[i*incx], x with nrows
 * entries, y with ncols.  The reference has no adjoint of its hierarchical types; the
 * leaf rules are those of its Transpose/Adjoint leaf methods (src/algebra.jl:52-82,
 * 138-159).  For a row part the result is the contribution of the owned rows (sum the
 * parts to get H'x); matrix-free plans: whole operators only. */
HM_API int32_t hm_matvec_adjoint(hm_plan *p, const double *x, int64_t incx, double *y, int64_t incy,
                                 int32_t accumulate);
HM_API int32_t hm_matvec_adjoint_device(hm_plan *p, const double *dx, double *dy, int32_t accumulate,
                                        void *stream);

/* Multi-right-hand-side form: Y[:, c] (+)= H X[:, c], c < nrhs; X ncols x nrhs
 * (ldx), Y nrows x nrhs (ldy), column-major.  (The reference reaches this through
 * the stride pair, test/runtests.jl:23-25.) */
HM_API int32_t hm_matmat(hm_plan *p, const double *X, int64_t ldx, double *Y, int64_t ldy, int64_t nrhs,
                  int32_t accumulate);
HM_API int32_t hm_matmat_device(hm_plan *p, const double *dX, int64_t ldx, double *dY, int64_t ldy,
                         int64_t nrhs, int32_t accumulate, void *stream);

/* ------------------------------------------------------------------------
 * Operator updates on the device, no re-planning.
 * side = 0: H <- H * Diagonal(b), b has ncols entries   (rmul!(H, b::Diagonal) ->
 *           scale!(H, b.diag, 1), src/HierarchicalMatrix.jl:15, 54-80)
 * side = 1: H <- Diagonal(b) * H, b has nrows entries   (lmul!(b::Diagonal, H) ->
 *           scale!(b.diag, H, 1), src/HierarchicalMatrix.jl:16, 82-108)
 * Leaf rule of src/algebra.jl:280-315: dense A[i,j] *= b; LowRankMatrix V[j,:] *= b_j
 * resp. U[i,:] *= b_i (BarycentricMatrix2D likewise; the reference defines no scale!
 * for it).  b is a host pointer with element stride incb; the 1-based jstart/istart of
 * the reference is a pointer offset at the call site.
 * ------------------------------------------------------------------------ */
HM_API int32_t hm_plan_scale(hm_plan *p, const double *b, int64_t incb, int32_t side);

/* Per-stage device timing (bench bookkeeping).  Between begin and end every
 * hm_matvec_device call records CUDA events around its three stages on the
 * launch stream; end synchronises and returns the summed milliseconds of
 * stage 1, 2, 3 and the number of matvecs timed (at most max_calls). */
HM_API int32_t hm_plan_timing_begin(hm_plan *p, int32_t max_calls);
HM_API int32_t hm_plan_timing_end(hm_plan *p, double *stage_ms3, int64_t *ncalls);

/* Number of kernel launches one hm_matvec_device enqueues (bench bookkeeping). */
HM_API int32_t hm_plan_launches_per_matvec(const hm_plan *p);

/* Test hook for the exception barrier: every entry point catches C++ exceptions and returns
 * HM_ERR_NOMEM (std::bad_alloc / std::length_error) or HM_ERR_INVALID.  After
 * hm_debug_fail_alloc(n) the n-th allocation checkpoint of the planner / tree builder passed on
 * the calling thread (0 = the next) fails as an exhausted host would; n < 0 disarms. */
HM_API int32_t hm_debug_fail_alloc(int64_t nth);

/* Test hooks: copy packed factors back to the host to compare the on-device
 * assembly with the oracle's factors.  which: 0 = U (m x ru), 1 = core
 * (F ru x rv | Sigma), 2 = V (n x rv), 3 = dense A (m x n); tight column-major. */
HM_API int32_t hm_plan_num_leaves(const hm_plan *p, int64_t *out);
HM_API int32_t hm_plan_leaf_info(const hm_plan *p, int64_t leaf, int32_t *kind, int64_t *row0,
                          int64_t *col0, int64_t *m, int64_t *n, int64_t *r);
HM_API int32_t hm_plan_read_leaf(hm_plan *p, int64_t leaf, int32_t which, double *out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* HMB200_H */
