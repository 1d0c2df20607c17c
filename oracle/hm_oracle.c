/*
 * hm_oracle.c -- CPU restatement of the HierarchicalMatrices.jl matvec hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see hm_oracle.h).  PARITY UNPINNED: no Julia in this
 * image, no golden vectors in the reference; see the header for what it is
 * checked against instead.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Loop orders and the order of floating-point operations are
 * kept as in the reference; compile with -ffp-contract=off so that `a*b + c` is
 * two roundings as in Julia (which does not contract without @fastmath).
 */
#include "hm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* constants: src/HierarchicalMatrices.jl:5-7                          */
/* ------------------------------------------------------------------ */

/* BLOCKRANK(T) = 2round(Int, half(T)*log(3+sqrt(T(8)), inv(eps(T))))
 * Julia's log(b, x) is log(x)/log(b); round(Int, .) rounds half to even. */
int hmo_blockrank_f64(void)
{
    double v = 0.5 * (log(1.0 / DBL_EPSILON) / log(3.0 + sqrt(8.0)));
    return 2 * (int)nearbyint(v);
}

int hmo_blockrank_f32(void)
{
    float v = 0.5f * (logf(1.0f / FLT_EPSILON) / logf(3.0f + sqrtf(8.0f)));
    return 2 * (int)nearbyintf(v);
}

/* BLOCKSIZE(T) = 4BLOCKRANK(T) -- HierarchicalMatrices.jl:7 */
int hmo_blocksize_f64(void) { return 4 * hmo_blockrank_f64(); }

/* ------------------------------------------------------------------ */
/* sinpi, Chebyshev nodes and weights: src/BarycentricMatrix.jl:92-136 */
/* ------------------------------------------------------------------ */

/* Julia's Base.sinpi is not available; this evaluates sin(pi x) in x87 long
 * double (64-bit significand) after an exact reduction to |r| <= 1/4, then
 * rounds to double.  The result is the correctly rounded value except in the
 * rare double-rounding case; Julia's own sinpi is also only faithful (<1 ulp),
 * so agreement is to <= 1 ulp, far inside the 1e-12 parity tolerance. */
double hmo_sinpi(double x)
{
    if (!isfinite(x)) return NAN;
    double r = x - 2.0 * nearbyint(0.5 * x); /* exact, r in [-1, 1] */
    double s = r < 0 ? -1.0 : 1.0;
    double t = fabs(r);
    if (t > 0.5) t = 1.0 - t; /* exact: sin(pi t) = sin(pi (1 - t)) */
    long double v;
    if (t <= 0.25)
        v = sinl(3.14159265358979323846264338327950288L * (long double)t);
    else
        v = cosl(3.14159265358979323846264338327950288L * (long double)(0.5 - t));
    return s * (double)v;
}

/* chebyshevpoints(T, n; kind) -- BarycentricMatrix.jl:92-111 */
void hmo_chebyshevpoints(int64_t n, int kind, double *out)
{
    int64_t nd2 = n / 2;
    for (int64_t k = 0; k < n; k++) out[k] = 0.0;
    for (int64_t k = 1; k <= nd2; k++) {
        double num = (double)(n - 2 * k) + 1.0; /* (n-2k+one(T)) */
        double arg = kind == 1 ? num / (double)(2 * n) : num / (double)(2 * (n - 1));
        out[k - 1] = hmo_sinpi(arg);
    }
    for (int64_t k = 1; k <= nd2; k++) out[n - k] = -out[k - 1];
}

/* chebyshevbarycentricweights(T, n; kind) -- BarycentricMatrix.jl:114-136 */
void hmo_chebyshevbarycentricweights(int64_t n, int kind, double *out)
{
    int64_t nd2 = n / 2;
    for (int64_t k = 0; k < n; k++) out[k] = 0.0;
    if (kind == 1) {
        for (int64_t k = 1; k <= nd2 + 1 && k <= n; k++)
            out[k - 1] = hmo_sinpi(((double)(2 * k) - 1.0) / (double)(2 * n));
        for (int64_t k = 1; k <= nd2; k++) out[n - k] = out[k - 1];
        for (int64_t k = 2; k <= n; k += 2) out[k - 1] *= -1.0;
    } else {
        for (int64_t k = 0; k < n; k++) out[k] = 1.0;
        for (int64_t k = 2; k <= n; k += 2) out[k - 1] *= -1.0;
        out[0] *= 0.5;
        out[n - 1] *= 0.5;
    }
}

/* indsplit(x, ir, a, b) -- BarycentricMatrix.jl:299-307.
 * Half-open 0-based range [i0, i1); on return the two halves are [i0, *mid)
 * and [*mid, i1) (the second is empty when *mid >= i1).  Like the reference
 * the first element is read before the range is tested, so an empty input
 * range can yield a one-element first half; reading past the end of x is the
 * reference's BoundsError and is reported as -1. */
int hmo_indsplit(const double *x, int64_t nx, int64_t i0, int64_t i1, double a, double b,
                 int64_t *mid)
{
    int64_t i = i0;
    double ab2 = 0.5 * (a + b);
    for (;;) {
        if (i < 0 || i >= nx) return -1;
        if (!(x[i] >= ab2)) break;
        i += 1;
        if (i > i1 - 1) break;
    }
    *mid = i;
    return 0;
}

/* KernelMatrix.jl:47 takes any f::Function: kernel id HMO_USER calls the function registered
 * here, element by element, exactly where the reference calls f(x, y). */
static double (*hmo_user_kernel)(double, double) = 0;
void hmo_set_user_kernel(double (*f)(double, double)) { hmo_user_kernel = f; }

/* examples/Kernel.jl:34-37.  `^2`/`^3` are literal powers (x*x, x*x*x). */
double hmo_kernel_eval(int kernel, double x, double y)
{
    double d = x - y;
    if (kernel == HMO_USER) return hmo_user_kernel ? hmo_user_kernel(x, y) : NAN;
    switch (kernel) {
    case HMO_CAUCHY: return 1.0 / d;
    case HMO_COULOMB: return 1.0 / (d * d);
    case HMO_COULOMBPRIME: return 1.0 / (d * d * d);
    case HMO_LOG: return log(fabs(d));
    default: return NAN;
    }
}

/* ------------------------------------------------------------------ */
/* leaf applies                                                        */
/* ------------------------------------------------------------------ */

/* mul!(y, A::AbstractMatrix, x, istart, jstart, INCX, INCY) -- algebra.jl:37-48
 * column sweep, y[i0 + i*incy] += A[i,j]*x[j0 + j*incx]. */
void hmo_mul_dense(double *y, const double *A, int64_t m, int64_t n, int64_t lda, const double *x,
                   int64_t i0, int64_t j0, int64_t incx, int64_t incy)
{
    for (int64_t j = 0; j < n; j++) {
        double xj = x[j0 + j * incx];
        const double *Aj = A + j * lda;
        if (incy == 1) {
            double *yy = y + i0;
            for (int64_t i = 0; i < m; i++) yy[i] += Aj[i] * xj;
        } else {
            for (int64_t i = 0; i < m; i++) y[i0 + i * incy] += Aj[i] * xj;
        }
    }
}

/* mul!(y, At::Transpose, x, istart, jstart, INCX, INCY) -- algebra.jl:52-65
 * A is the m x n parent; the result has n entries. */
void hmo_mul_dense_t(double *y, const double *A, int64_t m, int64_t n, int64_t lda,
                     const double *x, int64_t i0, int64_t j0, int64_t incx, int64_t incy)
{
    for (int64_t i = 0; i < n; i++) {
        double yi = 0.0;
        for (int64_t j = 0; j < m; j++) yi += A[j + i * lda] * x[j0 + j * incx];
        y[i0 + i * incy] += yi;
    }
}

/* mul!(y, L::LowRankMatrix, x, istart, jstart, INCX, INCY) -- algebra.jl:110-131
 * temp[k] = (sum_j V[j,k] x_j) * Sigma[k];  y_i += U[i,k] temp[k], k outer. */
void hmo_mul_lowrank(double *y, const double *U, int64_t ldu, const double *S, const double *V,
                     int64_t ldv, int64_t m, int64_t n, int64_t r, const double *x, int64_t i0,
                     int64_t j0, int64_t incx, int64_t incy)
{
    double stack_tmp[64];
    double *temp = r <= 64 ? stack_tmp : (double *)malloc((size_t)r * sizeof(double));
    for (int64_t k = 0; k < r; k++) {
        double t = 0.0;
        const double *Vk = V + k * ldv;
        for (int64_t j = 0; j < n; j++) t += Vk[j] * x[j0 + j * incx];
        temp[k] = t * S[k];
    }
    for (int64_t k = 0; k < r; k++) {
        double tk = temp[k];
        const double *Uk = U + k * ldu;
        if (incy == 1) {
            double *yy = y + i0;
            for (int64_t i = 0; i < m; i++) yy[i] += Uk[i] * tk;
        } else {
            for (int64_t i = 0; i < m; i++) y[i0 + i * incy] += Uk[i] * tk;
        }
    }
    if (temp != stack_tmp) free(temp);
}

/* mul!(u, B::BarycentricMatrix2D, v, istart, jstart) -- algebra.jl:243-277
 * temp1 = V' v;  temp2 = F temp1 (l outer, k inner);  u += U temp2 (k outer). */
void hmo_mul_bary2d(double *u, const double *U, int64_t ldu, const double *F, int64_t ldf,
                    const double *V, int64_t ldv, int64_t m, int64_t n, int64_t r,
                    const double *v, int64_t i0, int64_t j0)
{
    double stack_tmp[128];
    double *temp1 = r <= 64 ? stack_tmp : (double *)malloc((size_t)(2 * r) * sizeof(double));
    double *temp2 = temp1 + (r <= 64 ? 64 : r);
    for (int64_t k = 0; k < r; k++) {
        double t = 0.0;
        const double *Vk = V + k * ldv;
        for (int64_t j = 0; j < n; j++) t += Vk[j] * v[j0 + j];
        temp1[k] = t;
        temp2[k] = 0.0;
    }
    for (int64_t l = 0; l < r; l++) {
        double t1 = temp1[l];
        for (int64_t k = 0; k < r; k++) temp2[k] += F[k + l * ldf] * t1;
    }
    for (int64_t k = 0; k < r; k++) {
        double t2 = temp2[k];
        const double *Uk = U + k * ldu;
        double *uu = u + i0;
        for (int64_t i = 0; i < m; i++) uu[i] += Uk[i] * t2;
    }
    if (temp1 != stack_tmp) free(temp1);
}

/* ------------------------------------------------------------------ */
/* EvenBarycentricMatrix                            */
/* ------------------------------------------------------------------ */

/* BarycentricMatrix.jl:18-37: w[i] = sum_k lambda_k * inv(2i-a-b-(b-a)x_k) (sequential
 * in k), W[k,i] = lambda_k * inv((2i-a-b-(b-a)x_k) * w[i]).  The integer part 2i-a-b is
 * exact; (b-a)*x_k is one rounded product.  barycentricmatrix (:61-89) builds the same
 * numbers transposed. */
void hmo_evenbary_weights(int64_t a, int64_t b, double *w, double *W)
{
    enum { RMAX = 64 };
    int r = hmo_blockrank_f64();
    double xk[RMAX], lam[RMAX];
    hmo_chebyshevpoints(r, 1, xk);
    hmo_chebyshevbarycentricweights(r, 1, lam);
    const double span = (double)(b - a);
    for (int64_t i = a; i <= b; i++) {
        const double two_i = (double)(2 * i - a - b);
        double acc = 0.0;
        for (int k = 0; k < r; k++) acc += lam[k] * (1.0 / (two_i - span * xk[k]));
        w[i - a] = acc;
        for (int k = 0; k < r; k++) W[k + (i - a) * r] = lam[k] * (1.0 / ((two_i - span * xk[k]) * acc));
    }
}

/* One half of algebra.jl:168-239: beta_k = sum over the columns j = jfirst, jfirst+2, ...
 * of v[j0+j] F[j,k], then u[i0+i] += sum_k beta_k W[k,i] over the rows i = ifirst, ifirst+2, ... */
static void evenbary_half(double *u, const double *W, int64_t ldw, const double *F, int64_t ldf,
                          int64_t m, int64_t n, int64_t r, const double *v, int64_t i0, int64_t j0,
                          int64_t ifirst, int64_t jfirst)
{
    double beta[64];
    for (int64_t k = 0; k < r; k++) {
        double bk = 0.0;
        for (int64_t j = jfirst; j < n; j += 2) bk += v[j0 + j] * F[j + k * ldf];
        beta[k] = bk;
    }
    for (int64_t i = ifirst; i < m; i += 2) {
        double ui = 0.0;
        for (int64_t k = 0; k < r; k++) ui += beta[k] * W[k + i * ldw];
        u[i0 + i] += ui;
    }
}

/* 0-based: the 1-based odd rows/columns are the even offsets here.  With an even total
 * shift, odd columns feed odd rows then even columns feed even rows; with an odd shift the
 * column classes swap (algebra.jl:172-236).  `shift_even` is passed separately so a walk
 * that relocates u (hmo_mul_omp) keeps the reference's parity. */
static void evenbary_apply(double *u, const double *W, int64_t ldw, const double *F, int64_t ldf,
                           int64_t m, int64_t n, int64_t r, const double *v, int64_t i0, int64_t j0,
                           int shift_even)
{
    if (shift_even) {
        evenbary_half(u, W, ldw, F, ldf, m, n, r, v, i0, j0, 0, 0);
        evenbary_half(u, W, ldw, F, ldf, m, n, r, v, i0, j0, 1, 1);
    } else {
        evenbary_half(u, W, ldw, F, ldf, m, n, r, v, i0, j0, 0, 1);
        evenbary_half(u, W, ldw, F, ldf, m, n, r, v, i0, j0, 1, 0);
    }
}

void hmo_mul_evenbary(double *u, const double *W, int64_t ldw, const double *F, int64_t ldf,
                      int64_t m, int64_t n, int64_t r, const double *v, int64_t i0, int64_t j0)
{
    evenbary_apply(u, W, ldw, F, ldf, m, n, r, v, i0, j0, ((i0 + j0) & 1) == 0);
}

/* BarycentricMatrix.jl:48-59: nonzero iff size(B,1)+size(B,2)+i+j is even (1-based i, j;
 * the +2 of the 0-based form does not change parity).  Note this is the matrix's own
 * parity rule and differs from mul!'s when m+n is odd; both are restated as written. */
double hmo_evenbary_getindex(const double *W, int64_t ldw, const double *F, int64_t ldf, int64_t m,
                             int64_t n, int64_t r, int64_t i, int64_t j)
{
    double ret = 0.0;
    if (((m + n + i + j) & 1) == 0)
        for (int64_t k = 0; k < r; k++) ret += F[j + k * ldf] * W[k + i * ldw];
    return ret;
}

/* ------------------------------------------------------------------ */
/* @hierarchical container: src/hierarchical.jl:49-69                  */
/* ------------------------------------------------------------------ */

typedef struct hmo_block {
    int kind;        /* HMO_NONE / NODE / LOWRANK / DENSE / BARY2D / EVENBARY */
    int64_t m, n, r; /* leaf extents */
    double *U;       /* dense: A (ld m); low-rank families: U (ld m); EVENBARY: W (r x m, ld r) */
    double *S;       /* LOWRANK: Sigma (r); BARY2D: F (r x r) */
    double *V;       /* n x r (ld n); EVENBARY: F */
    hmo_node *child;
} hmo_block;

struct hmo_node {
    int M, N;
    hmo_block *b; /* column-major M x N like the reference's Matrix{...}(undef,M,N) */
};

static hmo_block *blk(const hmo_node *h, int m, int n) { return &h->b[m + (size_t)n * h->M]; }

static double *dupmat(const double *A, int64_t rows, int64_t cols, int64_t ld)
{
    size_t cnt = (size_t)(rows > 0 ? rows : 0) * (size_t)(cols > 0 ? cols : 0);
    double *p = (double *)malloc((cnt ? cnt : 1) * sizeof(double));
    if (!p) return NULL;
    for (int64_t j = 0; j < cols; j++)
        if (rows > 0) memcpy(p + j * rows, A + j * ld, (size_t)rows * sizeof(double));
    return p;
}

hmo_node *hmo_node_create(int M, int N)
{
    hmo_node *h = (hmo_node *)calloc(1, sizeof(hmo_node));
    if (!h) return NULL;
    h->M = M;
    h->N = N;
    h->b = (hmo_block *)calloc(M > 0 && N > 0 ? (size_t)M * (size_t)N : 1, sizeof(hmo_block));
    return h;
}

static void block_clear(hmo_block *b)
{
    if (b->kind == HMO_NODE) hmo_node_free(b->child);
    free(b->U);
    free(b->S);
    free(b->V);
    memset(b, 0, sizeof(*b));
}

void hmo_node_free(hmo_node *h)
{
    if (!h) return;
    for (int i = 0; i < h->M * h->N; i++) block_clear(&h->b[i]);
    free(h->b);
    free(h);
}

static hmo_block *slot(hmo_node *h, int m, int n)
{
    if (!h || m < 0 || n < 0 || m >= h->M || n >= h->N) return NULL;
    hmo_block *b = blk(h, m, n);
    block_clear(b);
    return b;
}

int hmo_node_set_node(hmo_node *h, int m, int n, hmo_node *child)
{
    hmo_block *b = slot(h, m, n);
    if (!b) return -1;
    b->kind = HMO_NODE;
    b->child = child;
    return 0;
}

/* ownership-taking variants used by the assembler (no copy) */
static void set_dense_own(hmo_node *h, int m, int n, double *A, int64_t rows, int64_t cols)
{
    hmo_block *b = slot(h, m, n);
    b->kind = HMO_DENSE;
    b->m = rows;
    b->n = cols;
    b->U = A;
}

static void set_bary_own(hmo_node *h, int m, int n, double *U, double *F, double *V, int64_t rows,
                         int64_t cols, int64_t r)
{
    hmo_block *b = slot(h, m, n);
    b->kind = HMO_BARY2D;
    b->m = rows;
    b->n = cols;
    b->r = r;
    b->U = U;
    b->S = F;
    b->V = V;
}

int hmo_node_set_dense(hmo_node *h, int m, int n, const double *A, int64_t rows, int64_t cols,
                       int64_t lda)
{
    if (!slot(h, m, n)) return -1;
    set_dense_own(h, m, n, dupmat(A, rows, cols, lda), rows, cols);
    return 0;
}

int hmo_node_set_lowrank(hmo_node *h, int m, int n, const double *U, int64_t ldu, const double *S,
                         const double *V, int64_t ldv, int64_t rows, int64_t cols, int64_t r)
{
    hmo_block *b = slot(h, m, n);
    if (!b) return -1;
    b->kind = HMO_LOWRANK;
    b->m = rows;
    b->n = cols;
    b->r = r;
    b->U = dupmat(U, rows, r, ldu);
    b->S = dupmat(S, r, 1, r);
    b->V = dupmat(V, cols, r, ldv);
    return 0;
}

int hmo_node_set_bary2d(hmo_node *h, int m, int n, const double *U, int64_t ldu, const double *F,
                        int64_t ldf, const double *V, int64_t ldv, int64_t rows, int64_t cols,
                        int64_t r)
{
    if (!slot(h, m, n)) return -1;
    set_bary_own(h, m, n, dupmat(U, rows, r, ldu), dupmat(F, r, r, ldf), dupmat(V, cols, r, ldv),
                 rows, cols, r);
    return 0;
}

int hmo_node_set_evenbary(hmo_node *h, int m, int n, const double *W, int64_t ldw, const double *F,
                          int64_t ldf, int64_t rows, int64_t cols, int64_t r)
{
    hmo_block *b = slot(h, m, n);
    if (!b || r > 64) return -1;
    b->kind = HMO_EVENBARY;
    b->m = rows;
    b->n = cols;
    b->r = r;
    b->U = dupmat(W, r, rows, ldw);
    b->V = dupmat(F, cols, r, ldf);
    return 0;
}

/* the `assigned` code of hierarchical.jl:84-91 */
int hmo_node_assigned(const hmo_node *h, int m, int n)
{
    switch (blk(h, m, n)->kind) {
    case HMO_NODE: return 1;
    case HMO_LOWRANK:
    case HMO_BARY2D:
    case HMO_EVENBARY: return 2;
    case HMO_DENSE: return 3;
    default: return 0;
    }
}

/* blocksize(H, m, n, k) -- hierarchical.jl:78-97: size of the stored block, 0 if
 * unassigned; nested blocks recurse through size(). */
int64_t hmo_blocksize(const hmo_node *h, int m, int n, int k)
{
    const hmo_block *b = blk(h, m, n);
    switch (b->kind) {
    case HMO_NODE: return hmo_size(b->child, k);
    case HMO_NONE: return 0;
    default: return k == 1 ? b->m : b->n;
    }
}

/* size(H) -- hierarchical.jl:33-47: rows summed down the LAST block column,
 * columns summed along the FIRST block row. */
int64_t hmo_size(const hmo_node *h, int k)
{
    int64_t s = 0;
    if (h->M == 0 || h->N == 0) return 0;
    if (k == 1)
        for (int m = 0; m < h->M; m++) s += hmo_blocksize(h, m, h->N - 1, 1);
    else
        for (int n = 0; n < h->N; n++) s += hmo_blocksize(h, 0, n, 2);
    return s;
}

/* getindex(H, i, j) -- hierarchical.jl:120-147 + leaf getindex methods
 * (LowRankMatrix.jl:50-58 sums k = r..1; BarycentricMatrix.jl:222-234). */
double hmo_getindex(const hmo_node *h, int64_t i, int64_t j)
{
    int m = 0, n = 0;
    while (m < h->M) {
        int64_t r = hmo_blocksize(h, m, h->N - 1, 1);
        if (i >= r) {
            i -= r;
            m += 1;
        } else
            break;
    }
    while (n < h->N) {
        int64_t s = hmo_blocksize(h, 0, n, 2);
        if (j >= s) {
            j -= s;
            n += 1;
        } else
            break;
    }
    if (m >= h->M || n >= h->N) return NAN; /* BoundsError in the reference */
    const hmo_block *b = blk(h, m, n);
    switch (b->kind) {
    case HMO_NODE: return hmo_getindex(b->child, i, j);
    case HMO_DENSE: return b->U[i + j * b->m];
    case HMO_LOWRANK: {
        double ret = 0.0;
        for (int64_t k = b->r - 1; k >= 0; k--)
            ret += b->U[i + k * b->m] * b->S[k] * b->V[j + k * b->n];
        return ret;
    }
    case HMO_BARY2D: {
        double ret = 0.0;
        for (int64_t k = 0; k < b->r; k++) {
            double temp = 0.0;
            for (int64_t l = 0; l < b->r; l++) temp += b->S[k + l * b->r] * b->V[j + l * b->n];
            ret += b->U[i + k * b->m] * temp;
        }
        return ret;
    }
    case HMO_EVENBARY: return hmo_evenbary_getindex(b->U, b->r, b->V, b->n, b->m, b->n, b->r, i, j);
    default: return 0.0;
    }
}

/* ------------------------------------------------------------------ */
/* walks: KernelMatrix.jl:17-45, HierarchicalMatrix.jl:24-52           */
/* ------------------------------------------------------------------ */

static void leaf_apply(double *y, const hmo_block *b, const double *x, int64_t i0, int64_t j0,
                       int64_t incx, int64_t incy)
{
    switch (b->kind) {
    case HMO_DENSE: hmo_mul_dense(y, b->U, b->m, b->n, b->m, x, i0, j0, incx, incy); break;
    case HMO_LOWRANK:
        hmo_mul_lowrank(y, b->U, b->m, b->S, b->V, b->n, b->m, b->n, b->r, x, i0, j0, incx, incy);
        break;
    case HMO_BARY2D:
        /* the reference has only the unit-stride method (algebra.jl:243); the
         * strided form below is the obvious extension and is exercised only by
         * this repo's stride tests */
        if (incx == 1 && incy == 1) {
            hmo_mul_bary2d(y, b->U, b->m, b->S, b->r, b->V, b->n, b->m, b->n, b->r, x, i0, j0);
        } else {
            double t1[64], t2[64];
            int64_t r = b->r;
            for (int64_t k = 0; k < r; k++) {
                double t = 0.0;
                for (int64_t j = 0; j < b->n; j++) t += b->V[j + k * b->n] * x[j0 + j * incx];
                t1[k] = t;
                t2[k] = 0.0;
            }
            for (int64_t l = 0; l < r; l++)
                for (int64_t k = 0; k < r; k++) t2[k] += b->S[k + l * r] * t1[l];
            for (int64_t k = 0; k < r; k++)
                for (int64_t i = 0; i < b->m; i++) y[i0 + i * incy] += b->U[i + k * b->m] * t2[k];
        }
        break;
    case HMO_EVENBARY:
        /* only the unit-stride 5-argument method exists (algebra.jl:168) */
        if (incx == 1 && incy == 1)
            hmo_mul_evenbary(y, b->U, b->r, b->V, b->n, b->m, b->n, b->r, x, i0, j0);
        break;
    default: break;
    }
}

/* Double loop over `assigned`, m outer / n inner, running offsets p (rows, from
 * the last block column) and q (columns, from the first block row), both
 * scaled by the strides (HierarchicalMatrix.jl:45,47; KernelMatrix.jl:38,40
 * is the unit-stride case). */
void hmo_mul(double *y, const hmo_node *h, const double *x, int64_t i0, int64_t j0, int64_t incx,
             int64_t incy)
{
    int64_t p = 0;
    for (int m = 0; m < h->M; m++) {
        int64_t q = 0;
        for (int n = 0; n < h->N; n++) {
            const hmo_block *b = blk(h, m, n);
            if (b->kind == HMO_NODE)
                hmo_mul(y, b->child, x, i0 + p, j0 + q, incx, incy);
            else if (b->kind != HMO_NONE)
                leaf_apply(y, b, x, i0 + p, j0 + q, incx, incy);
            q += incx * hmo_blocksize(h, 0, n, 2);
        }
        p += incy * hmo_blocksize(h, m, h->N - 1, 1);
    }
}

/* ------------------------------------------------------------------ */
/* adjoint apply y += H' x                           */
/* ------------------------------------------------------------------ */

/* The reference defines no adjoint of its hierarchical types.  This walks the tree like
 * mul! (offsets as in KernelMatrix.jl:24-41) and applies to each leaf the reference's own
 * transposed leaf rule: dense -- algebra.jl:52-65 (y_i += sum_j A[j,i] x_j, inner sum
 * first); LowRankMatrix -- algebra.jl:138-159 (temp = Sigma .* (U' x); y += V temp);
 * BarycentricMatrix2D by analogy (temp1 = U' x; temp2 = F' temp1; y += V temp2).
 * x has size(H,1) entries starting at x[i0], y size(H,2) starting at y[j0]. */
void hmo_mul_adjoint(double *y, const hmo_node *h, const double *x, int64_t i0, int64_t j0)
{
    int64_t p = 0;
    for (int m = 0; m < h->M; m++) {
        int64_t q = 0;
        for (int n = 0; n < h->N; n++) {
            const hmo_block *b = blk(h, m, n);
            if (b->kind == HMO_NODE) {
                hmo_mul_adjoint(y, b->child, x, i0 + p, j0 + q);
            } else if (b->kind == HMO_DENSE) {
                hmo_mul_dense_t(y, b->U, b->m, b->n, b->m, x, j0 + q, i0 + p, 1, 1);
            } else if (b->kind == HMO_EVENBARY) {
                /* transpose of the masked interpolant: the row class that mul! pairs with a
                 * column class feeds it back (beta = W x over the class, y += F beta) */
                int even = ((i0 + p + j0 + q) & 1) == 0;
                for (int64_t cls = 0; cls < 2; cls++) {
                    int64_t jfirst = even ? cls : 1 - cls;
                    double beta[64];
                    for (int64_t k = 0; k < b->r; k++) {
                        double t = 0.0;
                        for (int64_t i = cls; i < b->m; i += 2) t += b->U[k + i * b->r] * x[i0 + p + i];
                        beta[k] = t;
                    }
                    for (int64_t j = jfirst; j < b->n; j += 2) {
                        double t = 0.0;
                        for (int64_t k = 0; k < b->r; k++) t += b->V[j + k * b->n] * beta[k];
                        y[j0 + q + j] += t;
                    }
                }
            } else if (b->kind != HMO_NONE) {
                double t1[64], t2[64];
                int64_t r = b->r;
                for (int64_t k = 0; k < r; k++) {
                    double t = 0.0;
                    for (int64_t i = 0; i < b->m; i++) t += b->U[i + k * b->m] * x[i0 + p + i];
                    t1[k] = t;
                }
                if (b->kind == HMO_LOWRANK) {
                    for (int64_t k = 0; k < r; k++) t2[k] = t1[k] * b->S[k];
                } else {
                    for (int64_t l = 0; l < r; l++) {
                        double t = 0.0;
                        for (int64_t k = 0; k < r; k++) t += b->S[k + l * r] * t1[k];
                        t2[l] = t;
                    }
                }
                for (int64_t k = 0; k < r; k++)
                    for (int64_t j = 0; j < b->n; j++) y[j0 + q + j] += b->V[j + k * b->n] * t2[k];
            }
            q += hmo_blocksize(h, 0, n, 2);
        }
        p += hmo_blocksize(h, m, h->N - 1, 1);
    }
}

/* ------------------------------------------------------------------ */
/* scale!: H*Diagonal(b) and Diagonal(b)*H, in place  */
/* ------------------------------------------------------------------ */

/* scale!(H, b, jstart) -- HierarchicalMatrix.jl:54-80: n outer, m inner, the column
 * offset q advances by blocksize(H,1,n,2); leaves by algebra.jl:280-297: dense
 * C[i,j] = A[i,j]*b[j], LowRankMatrix: V[j,k] = b[j]*V[j,k].  BarycentricMatrix2D has
 * no scale! in the reference; it is treated like LowRankMatrix (V rows).  j0 0-based. */
void hmo_scale_cols(hmo_node *h, const double *b, int64_t j0)
{
    int64_t q = 0;
    for (int n = 0; n < h->N; n++) {
        for (int m = 0; m < h->M; m++) {
            hmo_block *k = blk(h, m, n);
            if (k->kind == HMO_NODE) {
                hmo_scale_cols(k->child, b, j0 + q);
            } else if (k->kind == HMO_DENSE) {
                for (int64_t j = 0; j < k->n; j++)
                    for (int64_t i = 0; i < k->m; i++) k->U[i + j * k->m] = k->U[i + j * k->m] * b[j0 + q + j];
            } else if (k->kind != HMO_NONE) { /* EVENBARY (no reference method): rows of F, same loop */
                for (int64_t c = 0; c < k->r; c++)
                    for (int64_t j = 0; j < k->n; j++) k->V[j + c * k->n] = k->V[j + c * k->n] * b[j0 + q + j];
            }
        }
        q += hmo_blocksize(h, 0, n, 2);
    }
}

/* scale!(b, H, istart) -- HierarchicalMatrix.jl:82-108 with algebra.jl:299-315:
 * dense C[i,j] = A[i,j]*b[i], LowRankMatrix: U[i,k] = U[i,k]*b[i].  i0 0-based. */
void hmo_scale_rows(const double *b, hmo_node *h, int64_t i0)
{
    int64_t p = 0;
    for (int m = 0; m < h->M; m++) {
        for (int n = 0; n < h->N; n++) {
            hmo_block *k = blk(h, m, n);
            if (k->kind == HMO_NODE) {
                hmo_scale_rows(b, k->child, i0 + p);
            } else if (k->kind == HMO_DENSE) {
                for (int64_t j = 0; j < k->n; j++)
                    for (int64_t i = 0; i < k->m; i++) k->U[i + j * k->m] = k->U[i + j * k->m] * b[i0 + p + i];
            } else if (k->kind == HMO_EVENBARY) { /* no reference method: columns of W */
                for (int64_t i = 0; i < k->m; i++)
                    for (int64_t c = 0; c < k->r; c++) k->U[c + i * k->r] = k->U[c + i * k->r] * b[i0 + p + i];
            } else if (k->kind != HMO_NONE) {
                for (int64_t c = 0; c < k->r; c++)
                    for (int64_t i = 0; i < k->m; i++) k->U[i + c * k->m] = k->U[i + c * k->m] * b[i0 + p + i];
            }
        }
        p += hmo_blocksize(h, m, h->N - 1, 1);
    }
}

/* ------------------------------------------------------------------ */
/* leaf enumeration (same order and the same offset rule as the walk)  */
/* ------------------------------------------------------------------ */

typedef struct {
    hmo_leaf *out;
    int64_t cap, cnt;
} leaf_sink;

static void walk_leaves(const hmo_node *h, int64_t i0, int64_t j0, int depth, leaf_sink *s)
{
    int64_t p = 0;
    for (int m = 0; m < h->M; m++) {
        int64_t q = 0;
        for (int n = 0; n < h->N; n++) {
            const hmo_block *b = blk(h, m, n);
            if (b->kind == HMO_NODE) {
                walk_leaves(b->child, i0 + p, j0 + q, depth + 1, s);
            } else if (b->kind != HMO_NONE) {
                if (s->out && s->cnt < s->cap) {
                    hmo_leaf *l = &s->out[s->cnt];
                    l->kind = b->kind;
                    l->depth = depth;
                    l->row0 = i0 + p;
                    l->col0 = j0 + q;
                    l->m = b->m;
                    l->n = b->n;
                    l->r = b->r;
                    l->A = b->U;
                    l->S = b->S;
                    l->V = b->V;
                }
                s->cnt++;
            }
            q += hmo_blocksize(h, 0, n, 2);
        }
        p += hmo_blocksize(h, m, h->N - 1, 1);
    }
}

int64_t hmo_count_leaves(const hmo_node *h)
{
    leaf_sink s = {NULL, 0, 0};
    walk_leaves(h, 0, 0, 0, &s);
    return s.cnt;
}

int64_t hmo_list_leaves(const hmo_node *h, hmo_leaf *out, int64_t cap)
{
    leaf_sink s = {out, cap, 0};
    walk_leaves(h, 0, 0, 0, &s);
    return s.cnt;
}

int64_t hmo_stored_words(const hmo_node *h)
{
    int64_t w = 0;
    for (int i = 0; i < h->M * h->N; i++) {
        const hmo_block *b = &h->b[i];
        switch (b->kind) {
        case HMO_NODE: w += hmo_stored_words(b->child); break;
        case HMO_DENSE: w += b->m * b->n; break;
        case HMO_LOWRANK: w += (b->m + b->n) * b->r + b->r; break;
        case HMO_BARY2D: w += (b->m + b->n) * b->r + b->r * b->r; break;
        case HMO_EVENBARY: w += (b->m + b->n) * b->r; break; /* W and F; mul! reads nothing else */
        default: break;
        }
    }
    return w;
}

static void count_nodes(const hmo_node *h, int depth, int64_t *cnt, int *maxdepth)
{
    *cnt += 1;
    if (depth > *maxdepth) *maxdepth = depth;
    for (int i = 0; i < h->M * h->N; i++)
        if (h->b[i].kind == HMO_NODE) count_nodes(h->b[i].child, depth + 1, cnt, maxdepth);
}

int64_t hmo_count_nodes(const hmo_node *h, int *maxdepth)
{
    int64_t cnt = 0;
    int md = 0;
    count_nodes(h, 1, &cnt, &md);
    if (maxdepth) *maxdepth = md;
    return cnt;
}

/* All-cores variant for the CPU baseline: identical per-leaf arithmetic, leaves
 * split into contiguous runs of about equal stored words, one private y per
 * thread, summed in thread order at the end. */
void hmo_mul_omp(double *y, const hmo_node *h, const double *x, int64_t i0, int64_t j0,
                 int nthreads)
{
    int64_t nl = hmo_count_leaves(h);
    int64_t rows = hmo_size(h, 1);
    if (nthreads < 1) nthreads = 1;
    hmo_leaf *lv = (hmo_leaf *)malloc((size_t)(nl ? nl : 1) * sizeof(hmo_leaf));
    hmo_list_leaves(h, lv, nl);
    int64_t *cum = (int64_t *)malloc((size_t)(nl + 1) * sizeof(int64_t));
    cum[0] = 0;
    for (int64_t l = 0; l < nl; l++) {
        int64_t w = lv[l].kind == HMO_DENSE ? lv[l].m * lv[l].n
                                            : (lv[l].m + lv[l].n) * lv[l].r + lv[l].r * lv[l].r;
        cum[l + 1] = cum[l] + w;
    }
    double *priv = (double *)calloc((size_t)(rows ? rows : 1) * nthreads, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int t = 0, nt = 1;
#ifdef _OPENMP
        t = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        double *yt = priv + (size_t)t * rows;
        int64_t lo = cum[nl] / nt * t, hi = t == nt - 1 ? cum[nl] + 1 : cum[nl] / nt * (t + 1);
        for (int64_t l = 0; l < nl; l++) {
            if (cum[l] < lo || cum[l] >= hi) continue;
            const hmo_leaf *f = &lv[l];
            if (f->kind == HMO_DENSE)
                hmo_mul_dense(yt, f->A, f->m, f->n, f->m, x, f->row0, j0 + f->col0, 1, 1);
            else if (f->kind == HMO_LOWRANK)
                hmo_mul_lowrank(yt, f->A, f->m, f->S, f->V, f->n, f->m, f->n, f->r, x, f->row0,
                                j0 + f->col0, 1, 1);
            else if (f->kind == HMO_EVENBARY)
                evenbary_apply(yt, f->A, f->r, f->V, f->n, f->m, f->n, f->r, x, f->row0, j0 + f->col0,
                               ((i0 + f->row0 + j0 + f->col0) & 1) == 0);
            else
                hmo_mul_bary2d(yt, f->A, f->m, f->S, f->r, f->V, f->n, f->m, f->n, f->r, x,
                               f->row0, j0 + f->col0);
        }
#ifdef _OPENMP
#pragma omp barrier
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < rows; i++) {
            double s = 0.0;
            for (int tt = 0; tt < nt; tt++) s += priv[(size_t)tt * rows + i];
            y[i0 + i] += s;
        }
    }
    free(priv);
    free(cum);
    free(lv);
}

/* ------------------------------------------------------------------ */
/* assembly                                                            */
/* ------------------------------------------------------------------ */

/* BarycentricPoly2D (BarycentricMatrix.jl:147-178) followed by
 * BarycentricMatrix2D + update! (:236-297).  x rows [i0,i1), y rows [j0,j1);
 * U is (i1-i0) x r, F r x r, V (j1-j0) x r, all column-major, tight. */
void hmo_bary2d_build(int kernel, double a, double b, double c, double d, const double *x,
                      int64_t i0, int64_t i1, const double *y, int64_t j0, int64_t j1, double *U,
                      double *F, double *V)
{
    enum { RMAX = 64 };
    int r = hmo_blockrank_f64();
    double xn[RMAX], yn[RMAX], lx[RMAX], ly[RMAX];
    int64_t m = i1 > i0 ? i1 - i0 : 0, n = j1 > j0 ? j1 - j0 : 0;
    hmo_chebyshevpoints(r, 1, xn);
    hmo_chebyshevpoints(r, 1, yn);
    hmo_chebyshevbarycentricweights(r, 1, lx);
    hmo_chebyshevbarycentricweights(r, 1, ly);

    double ab2 = 0.5 * (a + b), ba2 = 0.5 * (b - a);
    for (int p = 0; p < r; p++) xn[p] = ab2 + ba2 * xn[p];
    double cd2 = 0.5 * (c + d), dc2 = 0.5 * (d - c);
    for (int q = 0; q < r; q++) yn[q] = cd2 + dc2 * yn[q];

    for (int nn = 0; nn < r; nn++) /* F[m,n] = f(x[m], y[n]) -- :170-175 */
        for (int mm = 0; mm < r; mm++) F[mm + nn * r] = hmo_kernel_eval(kernel, xn[mm], yn[nn]);

    /* update! -- :248-297: U[i,m] = lambda_m * inv(x_i - x_m), then each row is
     * divided by its sequential sum over m; likewise V. */
    for (int mm = 0; mm < r; mm++)
        for (int64_t i = 0; i < m; i++) U[i + mm * m] = lx[mm] * (1.0 / (x[i0 + i] - xn[mm]));
    for (int64_t i = 0; i < m; i++) {
        double t = 0.0;
        for (int mm = 0; mm < r; mm++) t += U[i + mm * m];
        for (int mm = 0; mm < r; mm++) U[i + mm * m] /= t;
    }
    for (int nn = 0; nn < r; nn++)
        for (int64_t j = 0; j < n; j++) V[j + nn * n] = ly[nn] * (1.0 / (y[j0 + j] - yn[nn]));
    for (int64_t j = 0; j < n; j++) {
        double t = 0.0;
        for (int nn = 0; nn < r; nn++) t += V[j + nn * n];
        for (int nn = 0; nn < r; nn++) V[j + nn * n] /= t;
    }
}

typedef struct {
    int kernel;
    const double *x, *y;
    int64_t nx, ny;
    int bs, r;
    int err;
    int depth; /* a cluster of >= BLOCKSIZE coincident points never splits: StackOverflowError upstream */
} asm_ctx;

static int64_t rlen(int64_t a, int64_t b) { return b > a ? b - a : 0; }
static size_t nz(int64_t cnt) { return cnt > 0 ? (size_t)cnt : 1; }

/* T[f(x[i], y[j]) for i in ir, j in jr] -- KernelMatrix.jl:57-60 */
static void put_dense(asm_ctx *c, hmo_node *h, int m, int n, int64_t i0, int64_t i1, int64_t j0,
                      int64_t j1)
{
    int64_t rows = rlen(i0, i1), cols = rlen(j0, j1);
    double *A = (double *)malloc(nz(rows * cols) * sizeof(double));
    for (int64_t j = 0; j < cols; j++)
        for (int64_t i = 0; i < rows; i++)
            A[i + j * rows] = hmo_kernel_eval(c->kernel, c->x[i0 + i], c->y[j0 + j]);
    set_dense_own(h, m, n, A, rows, cols);
}

static void put_bary(asm_ctx *c, hmo_node *h, int m, int n, double a, double b, double cc,
                     double d, int64_t i0, int64_t i1, int64_t j0, int64_t j1)
{
    int64_t rows = rlen(i0, i1), cols = rlen(j0, j1);
    int r = c->r;
    double *U = (double *)malloc(nz(rows * r) * sizeof(double));
    double *F = (double *)malloc((size_t)r * r * sizeof(double));
    double *V = (double *)malloc(nz(cols * r) * sizeof(double));
    hmo_bary2d_build(c->kernel, a, b, cc, d, c->x, i0, i1, c->y, j0, j1, U, F, V);
    set_bary_own(h, m, n, U, F, V, rows, cols, r);
}

/* variant 0: KernelMatrix  (KernelMatrix.jl:49-70)   diagonal node
 * variant 1: KernelMatrix1 (:72-93)  dense corner at block (2,1)
 * variant 2: KernelMatrix2 (:95-116) dense corner at block (1,2) */
static hmo_node *assemble(asm_ctx *c, int variant, int64_t i0, int64_t i1, int64_t j0, int64_t j1,
                          double a, double b, double cc, double d)
{
    int64_t im, jm;
    if (c->depth > 1200 || hmo_indsplit(c->x, c->nx, i0, i1, a, b, &im) ||
        hmo_indsplit(c->y, c->ny, j0, j1, cc, d, &jm)) {
        c->err = 1;
        return NULL;
    }
    c->depth++;
    double ab2 = 0.5 * (a + b), cd2 = 0.5 * (cc + d);
    int leaf = rlen(i0, im) < c->bs && rlen(im, i1) < c->bs && rlen(j0, jm) < c->bs &&
               rlen(jm, j1) < c->bs;
    hmo_node *h = hmo_node_create(2, 2), *ch;
    if (variant == 0) {
        if (leaf) {
            put_dense(c, h, 0, 0, i0, im, j0, jm);
            put_dense(c, h, 0, 1, i0, im, jm, j1);
            put_dense(c, h, 1, 0, im, i1, j0, jm);
            put_dense(c, h, 1, 1, im, i1, jm, j1);
        } else {
            if ((ch = assemble(c, 0, i0, im, j0, jm, a, ab2, cc, cd2))) hmo_node_set_node(h, 0, 0, ch);
            if ((ch = assemble(c, 1, i0, im, jm, j1, a, ab2, cd2, d))) hmo_node_set_node(h, 0, 1, ch);
            if ((ch = assemble(c, 2, im, i1, j0, jm, ab2, b, cc, cd2))) hmo_node_set_node(h, 1, 0, ch);
            if ((ch = assemble(c, 0, im, i1, jm, j1, ab2, b, cd2, d))) hmo_node_set_node(h, 1, 1, ch);
        }
    } else if (variant == 1) {
        put_bary(c, h, 0, 0, a, ab2, cc, cd2, i0, im, j0, jm);
        put_bary(c, h, 0, 1, a, ab2, cd2, d, i0, im, jm, j1);
        if (leaf)
            put_dense(c, h, 1, 0, im, i1, j0, jm);
        else if ((ch = assemble(c, 1, im, i1, j0, jm, ab2, b, cc, cd2)))
            hmo_node_set_node(h, 1, 0, ch);
        put_bary(c, h, 1, 1, ab2, b, cd2, d, im, i1, jm, j1);
    } else {
        put_bary(c, h, 0, 0, a, ab2, cc, cd2, i0, im, j0, jm);
        if (leaf)
            put_dense(c, h, 0, 1, i0, im, jm, j1);
        else if ((ch = assemble(c, 2, i0, im, jm, j1, a, ab2, cd2, d)))
            hmo_node_set_node(h, 0, 1, ch);
        put_bary(c, h, 1, 0, ab2, b, cc, cd2, im, i1, j0, jm);
        put_bary(c, h, 1, 1, ab2, b, cd2, d, im, i1, jm, j1);
    }
    c->depth--;
    return h;
}

/* KernelMatrix(f, x, y, a, b, c, d) -- KernelMatrix.jl:47 */
hmo_node *hmo_kernelmatrix(int kernel, const double *x, int64_t nx, const double *y, int64_t ny,
                           double a, double b, double c, double d)
{
    asm_ctx ctx = {kernel, x, y, nx, ny, hmo_blocksize_f64(), hmo_blockrank_f64(), 0, 0};
    hmo_node *h = assemble(&ctx, 0, 0, nx, 0, ny, a, b, c, d);
    if (ctx.err) {
        hmo_node_free(h);
        return NULL;
    }
    return h;
}

/* KF*b of examples/Kernel.jl:75-78 in long double: the independent yardstick
 * the example itself uses for the hierarchical product. */
void hmo_dense_kernel_matvec_ld(int kernel, const double *x, int64_t nx, const double *y,
                                int64_t ny, const double *b, double *out)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < nx; i++) {
        long double s = 0.0L;
        for (int64_t j = 0; j < ny; j++) {
            long double d = (long double)x[i] - (long double)y[j], k;
            switch (kernel) {
            case HMO_CAUCHY: k = 1.0L / d; break;
            case HMO_COULOMB: k = 1.0L / (d * d); break;
            case HMO_COULOMBPRIME: k = 1.0L / (d * d * d); break;
            default: k = logl(fabsl(d)); break;
            }
            s += k * (long double)b[j];
        }
        out[i] = (double)s;
    }
}
