/*
 * hm_oracle.h -- CPU restatement of the HierarchicalMatrices.jl matvec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the CPU baseline.
 *
 * PARITY UNPINNED: the reference is pure Julia, `julia` is not installed in the
 * build image or on the GPU box, and the reference ships no golden vectors
 * (test/runtests.jl uses Julia-RNG data and only pins the dense offset/stride
 * leaf; the example prints an error and asserts nothing).  This restatement is
 * therefore checked only (i) against the properties the reference's own tests
 * assert (runtests.jl:5-7, :15-52) and (ii) against the example's own criterion,
 * the dense kernel product (examples/Kernel.jl:78), evaluated in long double.
 *
 * All file:line citations are relative to /root/reference.
 * Indices and offsets in this C API are 0-based; strides are in elements.
 */
#ifndef HM_ORACLE_H
#define HM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kernel ids -- examples/Kernel.jl:34-37 */
enum { HMO_CAUCHY = 0, HMO_COULOMB = 1, HMO_COULOMBPRIME = 2, HMO_LOG = 3, HMO_USER = 4 };
/* any f::Function (KernelMatrix.jl:47): kernel id HMO_USER evaluates this scalar callback */
void hmo_set_user_kernel(double (*f)(double, double));

/* block kinds; the `assigned` code of the reference (hierarchical.jl:84-91) is
 * 1 for NODE, 2 for LOWRANK/BARY2D (first listed leaf type), 3 for DENSE. */
enum { HMO_NONE = 0, HMO_NODE = 1, HMO_LOWRANK = 2, HMO_DENSE = 3, HMO_BARY2D = 4, HMO_EVENBARY = 5 };

typedef struct hmo_node hmo_node;

/* ---- constants: src/HierarchicalMatrices.jl:5-7 ---- */
int hmo_blockrank_f64(void);
int hmo_blockrank_f32(void);
int hmo_blocksize_f64(void);

/* ---- nodes / weights: src/BarycentricMatrix.jl:92-136 ---- */
double hmo_sinpi(double x);
void hmo_chebyshevpoints(int64_t n, int kind, double *out);
void hmo_chebyshevbarycentricweights(int64_t n, int kind, double *out);

/* ---- split: src/BarycentricMatrix.jl:299-307 (0-based half-open ranges) ----
 * returns 0, or -1 where the reference would throw BoundsError. */
int hmo_indsplit(const double *x, int64_t nx, int64_t i0, int64_t i1, double a, double b,
                 int64_t *mid);

/* ---- kernel evaluation: examples/Kernel.jl:34-37 ---- */
double hmo_kernel_eval(int kernel, double x, double y);

/* ---- leaf applies (accumulate), src/algebra.jl:37-48, 110-131, 243-277 ---- */
void hmo_mul_dense(double *y, const double *A, int64_t m, int64_t n, int64_t lda, const double *x,
                   int64_t i0, int64_t j0, int64_t incx, int64_t incy);
/* transpose(A) leaf, src/algebra.jl:52-65 (pinned by test/runtests.jl:27-33) */
void hmo_mul_dense_t(double *y, const double *A, int64_t m, int64_t n, int64_t lda,
                     const double *x, int64_t i0, int64_t j0, int64_t incx, int64_t incy);
void hmo_mul_lowrank(double *y, const double *U, int64_t ldu, const double *S, const double *V,
                     int64_t ldv, int64_t m, int64_t n, int64_t r, const double *x, int64_t i0,
                     int64_t j0, int64_t incx, int64_t incy);
void hmo_mul_bary2d(double *u, const double *U, int64_t ldu, const double *F, int64_t ldf,
                    const double *V, int64_t ldv, int64_t m, int64_t n, int64_t r,
                    const double *v, int64_t i0, int64_t j0);

/* ---- EvenBarycentricMatrix: BarycentricMatrix.jl:5-59, algebra.jl:166-239 ----
 * 1-D barycentric interpolant on the integer grid a..b whose entries vanish on one
 * parity class.  W is r x (b-a+1) (ld r), F is (d-c+1) x r: the caller evaluates the
 * user kernel f(T, (a+b)/2 + (b-a)*x[k]/2, j) into F (BarycentricMatrix.jl:38-43). */
void hmo_evenbary_weights(int64_t a, int64_t b, double *w, double *W);
/* mul!(u, B, v, istart, jstart) -- algebra.jl:168-239; i0 = istart-1, j0 = jstart-1 */
void hmo_mul_evenbary(double *u, const double *W, int64_t ldw, const double *F, int64_t ldf,
                      int64_t m, int64_t n, int64_t r, const double *v, int64_t i0, int64_t j0);
/* getindex(B, i, j) -- BarycentricMatrix.jl:48-59 (0-based i, j here) */
double hmo_evenbary_getindex(const double *W, int64_t ldw, const double *F, int64_t ldf, int64_t m,
                             int64_t n, int64_t r, int64_t i, int64_t j);

/* ---- generic @hierarchical container: src/hierarchical.jl:49-69 ---- */
hmo_node *hmo_node_create(int M, int N);
void hmo_node_free(hmo_node *h); /* recursive; frees owned blocks */
/* setindex!(H, A, Block(m), Block(n)) -- hierarchical.jl:149-172; data is copied */
int hmo_node_set_node(hmo_node *h, int m, int n, hmo_node *child); /* takes ownership */
int hmo_node_set_dense(hmo_node *h, int m, int n, const double *A, int64_t rows, int64_t cols,
                       int64_t lda);
int hmo_node_set_lowrank(hmo_node *h, int m, int n, const double *U, int64_t ldu, const double *S,
                         const double *V, int64_t ldv, int64_t rows, int64_t cols, int64_t r);
int hmo_node_set_bary2d(hmo_node *h, int m, int n, const double *U, int64_t ldu, const double *F,
                        int64_t ldf, const double *V, int64_t ldv, int64_t rows, int64_t cols,
                        int64_t r);
int hmo_node_set_evenbary(hmo_node *h, int m, int n, const double *W, int64_t ldw, const double *F,
                          int64_t ldf, int64_t rows, int64_t cols, int64_t r);
int hmo_node_assigned(const hmo_node *h, int m, int n); /* reference code 0..3 */
/* blocksize(H,m,n,k) and size(H,k): hierarchical.jl:33-47, 76-97 (k = 1 rows, 2 cols) */
int64_t hmo_blocksize(const hmo_node *h, int m, int n, int k);
int64_t hmo_size(const hmo_node *h, int k);
/* getindex(H,i,j): hierarchical.jl:120-147 with the leaf getindex methods
 * (BarycentricMatrix.jl:222-234, LowRankMatrix.jl:50-58); 0-based i,j */
double hmo_getindex(const hmo_node *h, int64_t i, int64_t j);

/* ---- tree walks: KernelMatrix.jl:17-45 and HierarchicalMatrix.jl:24-52 ----
 * y[i0 + i*incy] += (H x)[i]; the KernelMatrix walk is the incx = incy = 1 case. */
void hmo_mul(double *y, const hmo_node *h, const double *x, int64_t i0, int64_t j0, int64_t incx,
             int64_t incy);
/* Same arithmetic per leaf, leaves distributed over OpenMP threads with
 * per-thread y accumulators summed at the end (CPU-baseline "all cores" leg). */
void hmo_mul_omp(double *y, const hmo_node *h, const double *x, int64_t i0, int64_t j0,
                 int nthreads);

/* ---- adjoint apply y += H' x with the reference's transposed leaf rules
 * (algebra.jl:52-65, 138-159); ranks up to 64 ---- */
void hmo_mul_adjoint(double *y, const hmo_node *h, const double *x, int64_t i0, int64_t j0);

/* ---- scale!: HierarchicalMatrix.jl:54-108, algebra.jl:280-315 (in place) ---- */
void hmo_scale_cols(hmo_node *h, const double *b, int64_t j0); /* H <- H*Diagonal(b[j0:]) */
void hmo_scale_rows(const double *b, hmo_node *h, int64_t i0); /* H <- Diagonal(b[i0:])*H */

/* ---- assembly: KernelMatrix.jl:47-116, BarycentricMatrix.jl:147-178, 236-297 ----
 * x, y descending; (a,b), (c,d) as in the example.  Returns NULL where the
 * reference would throw. */
hmo_node *hmo_kernelmatrix(int kernel, const double *x, int64_t nx, const double *y, int64_t ny,
                           double a, double b, double c, double d);
/* Build one BarycentricMatrix2D (for unit tests of the factor arithmetic) */
void hmo_bary2d_build(int kernel, double a, double b, double c, double d, const double *x,
                      int64_t i0, int64_t i1, const double *y, int64_t j0, int64_t j1, double *U,
                      double *F, double *V);

/* ---- leaf enumeration in walk order (feeds the GPU builder in tests) ---- */
typedef struct hmo_leaf {
    int32_t kind; /* HMO_DENSE / HMO_LOWRANK / HMO_BARY2D / HMO_EVENBARY (A = W, r x m, ld r; V = F) */
    int32_t depth;
    int64_t row0, col0, m, n, r;
    const double *A; /* dense (ld = m) or U (ld = m) */
    const double *S; /* lowrank: Sigma (r); bary2d: F (r x r, ld = r) */
    const double *V; /* n x r, ld = n */
} hmo_leaf;
int64_t hmo_count_leaves(const hmo_node *h);
int64_t hmo_list_leaves(const hmo_node *h, hmo_leaf *out, int64_t cap);
/* stored words: sum_dense m n + sum_lr ((m+n) r + r^2 | r)   */
int64_t hmo_stored_words(const hmo_node *h);
int64_t hmo_count_nodes(const hmo_node *h, int *maxdepth);

/* ---- independent check: dense kernel product in long double (Kernel.jl:73-78) ---- */
void hmo_dense_kernel_matvec_ld(int kernel, const double *x, int64_t nx, const double *y,
                                int64_t ny, const double *b, double *out);

#ifdef __cplusplus
}
#endif
#endif
