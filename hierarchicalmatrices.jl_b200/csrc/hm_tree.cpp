// hm_tree.cpp -- geometric-bisection block tree of a KernelMatrix, host side.
//
// Follows /root/reference/src/KernelMatrix.jl:47-116 (the three mutually
// recursive constructors) and src/BarycentricMatrix.jl:299-307 (indsplit), but
// records only index ranges and interpolation boxes: U, V, F and the dense
// leaves are evaluated on the device straight into the packed streams
// (hm_kernels.cu: hm_fill1/3/core).  Offsets of the leaves are derived with the
// reference's own size rule (src/hierarchical.jl:33-47: rows from the last
// block column, columns from the first block row), not from the index ranges,
// so degenerate splits land where the reference's walk would put them.
#include "hm_tree.h"
#include "hm_layout.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>

int hm_blockrank_double()
{
    // 2round(Int, half(T)*log(3+sqrt(T(8)), inv(eps(T)))), round-half-even
    double v = 0.5 * (std::log(1.0 / DBL_EPSILON) / std::log(3.0 + std::sqrt(8.0)));
    return 2 * (int)std::nearbyint(v);
}

int hm_blocksize_double() { return 4 * hm_blockrank_double(); }

namespace {

// sin(pi q) for 0 <= q <= 1, via extended precision after folding to [0, 1/4]
double sinpi_unit(double q)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    double t = q > 0.5 ? 1.0 - q : q;
    long double v = t <= 0.25 ? sinl(pi * (long double)t) : cosl(pi * (long double)(0.5 - t));
    return (double)v;
}

enum SlotKind { SLOT_NONE = 0, SLOT_NODE = 1, SLOT_BARY = 2, SLOT_DENSE = 3 };

struct Slot {
    int kind = SLOT_NONE;
    int child = -1;
    int64_t i0 = 0, i1 = 0, j0 = 0, j1 = 0; // half-open point ranges
    double a = 0, b = 0, c = 0, d = 0;
};

struct Node {
    Slot s[2][2];
    int64_t rows = 0, cols = 0;
};

struct Builder {
    const double *x, *y;
    int64_t nx, ny;
    int bs, r;
    std::vector<Node> nodes;
    bool failed = false;
    bool too_deep = false;
    bool mono_x = false, mono_y = false; // point sets verified non-increasing
    int depth = 0;
    // Bisection halves the box every level; once it has collapsed to a point (about 1100
    // levels from [-1, 1]) a cluster of >= BLOCKSIZE coincident points can never be split
    // and the reference recurses until StackOverflowError.
    static constexpr int MAX_DEPTH = 1200;

    static int64_t len(int64_t lo, int64_t hi) { return hi > lo ? hi - lo : 0; }

    // descending points: first index whose value drops below the box midpoint.  The reference scans
    // linearly (BarycentricMatrix.jl:299-307); on a point set verified to be non-increasing (no NaN) a
    // bisection finds the same index, and the tree costs O(nodes log N) instead of O(N depth) reads
    // (0.5 s of the 2^22 assembly).  Anything else, and the empty-range quirk, takes the literal scan.
    bool split(const double *p, int64_t np, int64_t first, int64_t end, double lo, double hi,
               int64_t &mid)
    {
        const double pivot = 0.5 * (lo + hi);
        const bool mono = p == x ? mono_x : mono_y;
        if (mono && first >= 0 && first < np && end - 1 >= first) {
            const int64_t stop = end < np ? end : np;
            int64_t a_ = first, b_ = stop;
            while (a_ < b_) {
                const int64_t m = a_ + ((b_ - a_) >> 1);
                if (p[m] >= pivot)
                    a_ = m + 1;
                else
                    b_ = m;
            }
            if (a_ == stop && end > np) return false; // the scan would have run past the point set
            mid = a_;
            return true;
        }
        int64_t i = first;
        do {
            if (i < 0 || i >= np) return false; // BoundsError in the reference
            if (!(p[i] >= pivot)) break;
            ++i;
        } while (i <= end - 1);
        mid = i;
        return true;
    }

    Slot leaf(int kind, int64_t i0, int64_t i1, int64_t j0, int64_t j1, double a, double b, double c,
              double d)
    {
        Slot s;
        s.kind = kind;
        s.i0 = i0;
        s.i1 = i1;
        s.j0 = j0;
        s.j1 = j1;
        s.a = a;
        s.b = b;
        s.c = c;
        s.d = d;
        return s;
    }

    Slot nested(int variant, int64_t i0, int64_t i1, int64_t j0, int64_t j1, double a, double b,
                double c, double d)
    {
        Slot s;
        int idx = build(variant, i0, i1, j0, j1, a, b, c, d);
        if (idx < 0) return s;
        s.kind = SLOT_NODE;
        s.child = idx;
        return s;
    }

    int64_t extent(const Slot &s, int k) const
    {
        if (s.kind == SLOT_NODE) return k == 1 ? nodes[(size_t)s.child].rows : nodes[(size_t)s.child].cols;
        if (s.kind == SLOT_NONE) return 0;
        return k == 1 ? len(s.i0, s.i1) : len(s.j0, s.j1);
    }

    // variant 0: diagonal node; 1: dense corner bottom-left; 2: dense corner top-right
    int build(int variant, int64_t i0, int64_t i1, int64_t j0, int64_t j1, double a, double b,
              double c, double d)
    {
        int64_t im, jm;
        if (depth > MAX_DEPTH) {
            failed = too_deep = true;
            return -1;
        }
        if (!split(x, nx, i0, i1, a, b, im) || !split(y, ny, j0, j1, c, d, jm)) {
            failed = true;
            return -1;
        }
        struct DepthScope {
            int &d;
            explicit DepthScope(int &dd) : d(dd) { ++d; }
            ~DepthScope() { --d; }
        } scope(depth);
        const double xm = 0.5 * (a + b), ym = 0.5 * (c + d);
        const bool small = len(i0, im) < bs && len(im, i1) < bs && len(j0, jm) < bs && len(jm, j1) < bs;
        Node nd;
        if (variant == 0) {
            if (small) {
                nd.s[0][0] = leaf(SLOT_DENSE, i0, im, j0, jm, a, xm, c, ym);
                nd.s[0][1] = leaf(SLOT_DENSE, i0, im, jm, j1, a, xm, ym, d);
                nd.s[1][0] = leaf(SLOT_DENSE, im, i1, j0, jm, xm, b, c, ym);
                nd.s[1][1] = leaf(SLOT_DENSE, im, i1, jm, j1, xm, b, ym, d);
            } else {
                nd.s[0][0] = nested(0, i0, im, j0, jm, a, xm, c, ym);
                nd.s[0][1] = nested(1, i0, im, jm, j1, a, xm, ym, d);
                nd.s[1][0] = nested(2, im, i1, j0, jm, xm, b, c, ym);
                nd.s[1][1] = nested(0, im, i1, jm, j1, xm, b, ym, d);
            }
        } else {
            // the three well-separated quadrants are rank-r interpolants
            nd.s[0][0] = leaf(SLOT_BARY, i0, im, j0, jm, a, xm, c, ym);
            nd.s[1][1] = leaf(SLOT_BARY, im, i1, jm, j1, xm, b, ym, d);
            if (variant == 1) {
                nd.s[0][1] = leaf(SLOT_BARY, i0, im, jm, j1, a, xm, ym, d);
                nd.s[1][0] = small ? leaf(SLOT_DENSE, im, i1, j0, jm, xm, b, c, ym)
                                   : nested(1, im, i1, j0, jm, xm, b, c, ym);
            } else {
                nd.s[1][0] = leaf(SLOT_BARY, im, i1, j0, jm, xm, b, c, ym);
                nd.s[0][1] = small ? leaf(SLOT_DENSE, i0, im, jm, j1, a, xm, ym, d)
                                   : nested(2, i0, im, jm, j1, a, xm, ym, d);
            }
        }
        if (failed) return -1;
        nd.rows = extent(nd.s[0][1], 1) + extent(nd.s[1][1], 1); // last block column
        nd.cols = extent(nd.s[0][0], 2) + extent(nd.s[0][1], 2); // first block row
        nodes.push_back(nd);
        return (int)nodes.size() - 1;
    }

    void emit(int idx, int64_t r0, int64_t c0, std::vector<HmLeaf> &out) const
    {
        const Node &nd = nodes[(size_t)idx];
        int64_t p = 0;
        for (int m = 0; m < 2; m++) {
            int64_t q = 0;
            for (int n = 0; n < 2; n++) {
                const Slot &s = nd.s[m][n];
                if (s.kind == SLOT_NODE) {
                    emit(s.child, r0 + p, c0 + q, out);
                } else if (s.kind != SLOT_NONE) {
                    HmLeaf l{};
                    l.kind = s.kind == SLOT_DENSE ? HM_LEAF_DENSE : HM_LEAF_BARY2D;
                    l.source = HM_SRC_KERNEL;
                    l.row0 = r0 + p;
                    l.col0 = c0 + q;
                    l.m = len(s.i0, s.i1);
                    l.n = len(s.j0, s.j1);
                    l.ru = l.rv = s.kind == SLOT_DENSE ? 0 : r;
                    l.xi0 = s.i0;
                    l.yj0 = s.j0;
                    l.a = s.a;
                    l.b = s.b;
                    l.c = s.c;
                    l.d = s.d;
                    out.push_back(l);
                }
                q += extent(nd.s[0][n], 2);
            }
            p += extent(nd.s[m][1], 1);
        }
    }
};

} // namespace

void hm_cheb_nodes_weights(int n, double *nodes, double *weights)
{
    const int h = n / 2;
    for (int k = 0; k < n; k++) nodes[k] = weights[k] = 0.0;
    for (int k = 1; k <= h; k++) {
        nodes[k - 1] = sinpi_unit(((double)(n - 2 * k) + 1.0) / (double)(2 * n));
        nodes[n - k] = -nodes[k - 1];
    }
    for (int k = 1; k <= h + 1 && k <= n; k++) weights[k - 1] = sinpi_unit(((double)(2 * k) - 1.0) / (double)(2 * n));
    for (int k = 1; k <= h; k++) weights[n - k] = weights[k - 1];
    for (int k = 2; k <= n; k += 2) weights[k - 1] = -weights[k - 1];
}

std::string hm_kernel_tree(const double *x, int64_t nx, const double *y, int64_t ny, double a,
                           double b, double c, double d, std::vector<HmLeaf> &leaves,
                           int64_t &nrows, int64_t &ncols)
{
    hm_fault_checkpoint();
    Builder bld;
    bld.x = x;
    bld.y = y;
    bld.nx = nx;
    bld.ny = ny;
    bld.bs = hm_blocksize_double();
    bld.r = hm_blockrank_double();
    auto non_increasing = [](const double *p, int64_t n) {
        if (getenv("HMB200_TREE_LINEAR")) return false; // (test hook: the reference's literal scan everywhere)
        if (n > 0 && !(p[0] == p[0])) return false;
        for (int64_t i = 0; i + 1 < n; i++)
            if (!(p[i] >= p[i + 1])) return false;
        return true;
    };
    bld.mono_x = non_increasing(x, nx);
    bld.mono_y = x == y ? bld.mono_x : non_increasing(y, ny);
    int root = bld.build(0, 0, nx, 0, ny, a, b, c, d);
    if (bld.too_deep)
        return "KernelMatrix: a cluster of coincident points cannot be bisected (StackOverflowError in the reference)";
    if (root < 0 || bld.failed) return "KernelMatrix: index split ran past the end of the point set (BoundsError in the reference)";
    nrows = bld.nodes[(size_t)root].rows;
    ncols = bld.nodes[(size_t)root].cols;
    leaves.clear();
    hm_fault_checkpoint();
    bld.emit(root, 0, 0, leaves);
    // the leaves address points by their own ranges; they must exist
    for (const HmLeaf &l : leaves)
        if (l.xi0 < 0 || l.xi0 + l.m > nx || l.yj0 < 0 || l.yj0 + l.n > ny)
            return "KernelMatrix: leaf range outside the point set";
    return "";
}
