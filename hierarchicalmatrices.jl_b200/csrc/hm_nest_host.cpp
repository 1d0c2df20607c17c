// hm_nest_host.cpp -- host builder of the nested-basis form (hm_nest.h).  Pure C++.
#include "hm_nest.h"

#include "hm_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

namespace {

constexpr int R = HM_NEST_R;
constexpr int MAX_DEPTH = 60;

// Cluster tree over one point set: boxes by exact halving of the root box, points by the reference's
// indsplit rule (the first half takes the leading points with p >= midpoint; the sets are descending).
struct Trie {
    const double *pts = nullptr;
    int64_t npts = 0;
    std::vector<HmNestNode> nodes;
    std::vector<double> ba, bb; // box of every node as the reference computes it: (a, b), a may exceed b
    std::vector<int32_t> depth;
    std::vector<char> needed;   // the box of some low-rank leaf

    int add(double a, double b, int64_t p0, int64_t np, int parent, int which, int dep)
    {
        HmNestNode n{};
        n.mid = 0.5 * (a + b);
        const double half = 0.5 * (b - a);
        n.ih = half != 0.0 ? 1.0 / half : 0.0;
        n.p0 = (int32_t)p0;
        n.np = (int32_t)np;
        n.child0 = -1;
        n.parent = parent;
        n.which = which;
        nodes.push_back(n);
        ba.push_back(a);
        bb.push_back(b);
        depth.push_back(dep);
        needed.push_back(0);
        return (int)nodes.size() - 1;
    }

    bool splittable(int i) const
    {
        const double m = 0.5 * (ba[(size_t)i] + bb[(size_t)i]);
        return depth[(size_t)i] < MAX_DEPTH && m != ba[(size_t)i] && m != bb[(size_t)i];
    }

    void split(int i)
    {
        const double a = ba[(size_t)i], b = bb[(size_t)i], m = 0.5 * (a + b); // half(T)*(a+b)
        const int64_t p0 = nodes[(size_t)i].p0, np = nodes[(size_t)i].np;
        // indsplit: i advances while x[i] >= ab2 (src/BarycentricMatrix.jl:299-307)
        const double *lo = pts + p0, *hi = pts + p0 + np;
        const double *cut = std::partition_point(lo, hi, [m](double v) { return v >= m; });
        const int64_t n0 = cut - lo;
        const int dep = depth[(size_t)i] + 1;
        const int c0 = add(a, m, p0, n0, i, 0, dep);
        add(m, b, p0 + n0, np - n0, i, 1, dep);
        nodes[(size_t)i].child0 = c0;
    }

    // the node whose box is exactly (ta, tb); -1 when the box is not a dyadic descendant of the root
    int find(double ta, double tb)
    {
        int cur = 0;
        for (;;) {
            const double a = ba[(size_t)cur], b = bb[(size_t)cur];
            if (ta == a && tb == b) return cur;
            if (nodes[(size_t)cur].child0 < 0) {
                if (!splittable(cur)) return -1;
                split(cur);
            }
            const double tc = 0.5 * (ta + tb);
            const bool first = std::fabs(tc - a) < std::fabs(tc - b);
            cur = nodes[(size_t)cur].child0 + (first ? 0 : 1);
        }
    }
};

// Boxes are grouped into tiers by their point count (tier k: at most HM_NEST_TIER0 * HM_NEST_GROWTH^k points) and
// every maximal connected set of boxes of one tier is a subtree handled by one CTA.  A box's halves
// are in its own subtree (deeper, so earlier in the upward order) or in a lower tier (an earlier
// launch).  Cutting by points rather than depth keeps the subtrees balanced on graded point sets
// (Chebyshev points: a depth-10 box at the end of the interval holds 30 times the points of one in the middle).
void schedule(const Trie &T, HmNestTree &out)
{
    out.nodes = T.nodes;
    const size_t nn = T.nodes.size();
    std::vector<int32_t> tier(nn);
    int ntiers = 1;
    // (HMB200_NEST_TIERS="points of the finest tier,growth per tier": tuning experiments)
    int64_t tier0 = HM_NEST_TIER0, growth = HM_NEST_GROWTH;
    if (const char *e = getenv("HMB200_NEST_TIERS")) {
        long a = 0, b = 0;
        if (sscanf(e, "%ld,%ld", &a, &b) == 2 && a >= HM_NEST_BASE && b >= 2) {
            tier0 = a;
            growth = b;
        }
    }
    for (size_t i = 0; i < nn; i++) {
        int k = 0;
        int64_t cap = tier0;
        while (T.nodes[i].np > cap) {
            cap *= growth;
            k++;
        }
        tier[i] = k;
        ntiers = std::max(ntiers, k + 1);
    }
    out.order.clear();
    out.grp.clear();
    out.sub_g0.clear();
    out.tier_sub0.assign((size_t)ntiers + 1, 0);
    out.max_group = 0;
    std::vector<int32_t> stack, sub;
    auto emit_groups = [&](std::vector<int32_t> &ids) {
        // deepest first; stable inside a depth
        std::stable_sort(ids.begin(), ids.end(), [&](int32_t u, int32_t v) { return T.depth[(size_t)u] > T.depth[(size_t)v]; });
        out.sub_g0.push_back((int32_t)out.grp.size());
        size_t i = 0;
        while (i < ids.size()) {
            size_t j = i;
            while (j < ids.size() && T.depth[(size_t)ids[j]] == T.depth[(size_t)ids[i]]) j++;
            out.grp.push_back((int32_t)out.order.size());
            for (size_t k = i; k < j; k++) out.order.push_back(ids[k]);
            out.max_group = std::max(out.max_group, (int)(j - i));
            i = j;
        }
    };
    int nsub = 0;
    for (int k = 0; k < ntiers; k++) {
        out.tier_sub0[(size_t)k] = nsub;
        for (size_t r = 0; r < nn; r++) {
            if (tier[r] != k) continue;
            const int32_t par = T.nodes[r].parent;
            if (par >= 0 && tier[(size_t)par] == k) continue; // not the root of its subtree
            sub.clear();
            stack.assign(1, (int32_t)r);
            while (!stack.empty()) {
                const int32_t u = stack.back();
                stack.pop_back();
                sub.push_back(u);
                const int32_t c0 = T.nodes[(size_t)u].child0;
                if (c0 >= 0) {
                    if (tier[(size_t)c0] == k) stack.push_back(c0);
                    if (tier[(size_t)c0 + 1] == k) stack.push_back(c0 + 1);
                }
            }
            emit_groups(sub);
            nsub++;
        }
    }
    out.tier_sub0[(size_t)ntiers] = nsub;
    out.base.clear();
    for (size_t i = 0; i < nn; i++)
        if (T.nodes[i].child0 < 0) out.base.push_back((int32_t)i);
    out.sub_g0.push_back((int32_t)out.grp.size());
    out.grp.push_back((int32_t)out.order.size());
}

long double kernel_ld(int id, long double d)
{
    switch (id) {
    case 0: return 1.0L / d;
    case 1: return 1.0L / (d * d);
    case 2: return 1.0L / (d * d * d);
    default: return logl(fabsl(d));
    }
}

} // namespace

std::string hm_nest_build(const HmLayout &L, const double *x, int64_t nx, const double *y, int64_t ny, double a,
                          double b, double c, double d, int kernel_id, const std::vector<HmFreeRun> &frun3,
                          HmNest &out)
{
    hm_fault_checkpoint();
    out = HmNest();
    if (kernel_id < 0 || kernel_id > 3) return "kernel is not translation invariant";
    if (nx <= 0 || ny <= 0 || nx >= ((int64_t)1 << 31) || ny >= ((int64_t)1 << 31)) return "empty or oversized point set";
    for (int64_t i = 1; i < nx; i++)
        if (!(x[i] <= x[i - 1])) return "row points are not in descending order";
    for (int64_t j = 1; j < ny; j++)
        if (!(y[j] <= y[j - 1])) return "column points are not in descending order";
    Trie TR, TC;
    TR.pts = x;
    TR.npts = nx;
    TC.pts = y;
    TC.npts = ny;
    TR.add(a, b, 0, nx, -1, 0, 0);
    TC.add(c, d, 0, ny, -1, 0, 0);

    // ---- boxes of the low-rank leaves
    const size_t ncore = L.cores.size();
    std::vector<int32_t> rnode(ncore), cnode(ncore);
    for (size_t k = 0; k < ncore; k++) {
        const HmLeaf &l = L.leaves[(size_t)L.core_leaf[k]];
        if (l.kind != HM_LEAF_BARY2D || l.ru != R || l.rv != R) return "a low-rank leaf is not a rank-20 BarycentricMatrix2D";
        if (l.row0 != l.xi0 || l.col0 != l.yj0) return "rows / columns are not numbered like the points";
        const int ri = TR.find(l.a, l.b), ci = TC.find(l.c, l.d);
        if (ri < 0 || ci < 0) return "a leaf box is not a dyadic descendant of the root box";
        if (TR.nodes[(size_t)ri].p0 != l.xi0 || TR.nodes[(size_t)ri].np != l.m || TC.nodes[(size_t)ci].p0 != l.yj0 ||
            TC.nodes[(size_t)ci].np != l.n)
            return "a leaf's index range differs from its box's points";
        TR.needed[(size_t)ri] = 1;
        TC.needed[(size_t)ci] = 1;
        rnode[k] = ri;
        cnode[k] = ci;
    }
    // ---- finest boxes: halve until no box holds more than HM_NEST_BASE points
    for (Trie *T : {&TR, &TC})
        for (size_t i = 0; i < T->nodes.size(); i++) // grows while iterating
            if (T->nodes[i].child0 < 0 && T->nodes[i].np > HM_NEST_BASE && T->splittable((int)i)) T->split((int)i);
    hm_fault_checkpoint();
    schedule(TR, out.rows);
    schedule(TC, out.cols);

    // ---- distinct cores: G = C F C', F[k][l] = f((mid_I - mid_J) + half_I xi_k - half_J xi_l)
    long double xi[R], Cm[R][R];
    for (int k = 1; k <= R; k++) xi[k - 1] = cosl(3.14159265358979323846264338327950288L * (2 * k - 1) / (2.0L * R));
    for (int k = 0; k < R; k++) {
        long double tm2 = 1.0L, tm1 = xi[k];
        for (int q = 0; q < R; q++) {
            long double tq = q == 0 ? 1.0L : q == 1 ? xi[k] : 2.0L * xi[k] * tm1 - tm2;
            if (q >= 2) {
                tm2 = tm1;
                tm1 = tq;
            }
            Cm[q][k] = (q == 0 ? 1.0L : 2.0L) * tq / R;
        }
    }
    std::map<std::tuple<double, double, double>, int32_t> seen;
    std::vector<int32_t> core_of(ncore);
    std::vector<std::tuple<double, double, double>> keys;
    for (size_t k = 0; k < ncore; k++) {
        const HmLeaf &l = L.leaves[(size_t)L.core_leaf[k]];
        const double dm = 0.5 * (l.a + l.b) - 0.5 * (l.c + l.d), hi = 0.5 * (l.b - l.a), hj = 0.5 * (l.d - l.c);
        auto key = std::make_tuple(dm, hi, hj);
        auto it = seen.find(key);
        if (it == seen.end()) {
            it = seen.emplace(key, (int32_t)keys.size()).first;
            keys.push_back(key);
        }
        core_of[k] = it->second;
    }
    if (keys.size() > 20000) return "too many distinct cores (boxes are not on a common dyadic grid)";
    hm_fault_checkpoint();
    out.cores.assign(keys.size() * (size_t)(R * R), 0.0);
    for (size_t u = 0; u < keys.size(); u++) {
        const long double dm = std::get<0>(keys[u]), hi = std::get<1>(keys[u]), hj = std::get<2>(keys[u]);
        long double F[R][R], W[R][R];
        for (int k = 0; k < R; k++)
            for (int l = 0; l < R; l++) F[k][l] = kernel_ld(kernel_id, dm + (hi * xi[k] - hj * xi[l]));
        for (int q = 0; q < R; q++) // W = C F
            for (int l = 0; l < R; l++) {
                long double s = 0.0L;
                for (int k = 0; k < R; k++) s += Cm[q][k] * F[k][l];
                W[q][l] = s;
            }
        double *G = out.cores.data() + u * (size_t)(R * R);
        for (int q = 0; q < R; q++) // G = W C'
            for (int p = 0; p < R; p++) {
                long double s = 0.0L;
                for (int l = 0; l < R; l++) s += W[q][l] * Cm[p][l];
                G[q + p * R] = (double)s;
            }
    }

    // ---- leaves per row box
    const size_t nr = TR.nodes.size();
    out.rleaf_begin.assign(nr + 1, 0);
    for (size_t k = 0; k < ncore; k++) out.rleaf_begin[(size_t)rnode[k] + 1]++;
    for (size_t i = 0; i < nr; i++) out.rleaf_begin[i + 1] += out.rleaf_begin[i];
    out.rleaf.resize(ncore);
    {
        std::vector<int32_t> fill(out.rleaf_begin.begin(), out.rleaf_begin.end() - 1);
        for (size_t k = 0; k < ncore; k++) out.rleaf[(size_t)fill[(size_t)rnode[k]]++] = HmNestLeaf{core_of[k], cnode[k]};
    }

    // ---- transfer maps: T_q((eta -+ 1) / 2) = sum_p M[q][p] T_p(eta), by the discrete Chebyshev transform
    out.M.assign(4 * (size_t)(R * R), 0.0);
    for (int w = 0; w < 2; w++)
        for (int q = 0; q < R; q++)
            for (int p = 0; p <= q; p++) { // degree q: the map is lower triangular
                long double s = 0.0L;
                for (int k = 0; k < R; k++) {
                    const long double arg = 0.5L * (xi[k] + (w ? 1.0L : -1.0L));
                    long double t0 = 1.0L, t1 = arg, tq = q == 0 ? 1.0L : arg;
                    for (int j = 2; j <= q; j++) {
                        tq = 2.0L * arg * t1 - t0;
                        t0 = t1;
                        t1 = tq;
                    }
                    long double u0 = 1.0L, u1 = xi[k], tp = p == 0 ? 1.0L : xi[k];
                    for (int j = 2; j <= p; j++) {
                        tp = 2.0L * xi[k] * u1 - u0;
                        u0 = u1;
                        u1 = tp;
                    }
                    s += tq * tp;
                }
                const double m = (double)((p == 0 ? 1.0L : 2.0L) * s / R);
                out.M[(size_t)w * R * R + (size_t)q * R + p] = m;       // [q][p]
                out.M[(size_t)(2 + w) * R * R + (size_t)p * R + q] = m; // transposed: [p][q]
            }

    // ---- dense part: the stage-3 items without their low-rank runs
    out.round_begin.assign(L.round_begin.size(), 0);
    for (size_t r = 0; r + 1 < L.round_begin.size(); r++) {
        out.round_begin[r] = (int64_t)out.items3.size();
        for (int64_t i = L.round_begin[r]; i < L.round_begin[r + 1]; i++) {
            const HmItem &it = L.items3[(size_t)i];
            HmItem ni = it;
            ni.run0 = (int32_t)out.runs.size();
            int32_t pos = 0;
            for (int32_t k = it.run0; k < it.run0 + it.nrun; k++) {
                const HmRun &rr = L.runs[(size_t)k];
                if (rr.src < 0) continue;
                const HmFreeRun &fr = frun3[(size_t)k];
                if ((int32_t)out.runs.size() > ni.run0) {
                    // the next columns of the same rows: one longer run (KernelMatrix: the dense leaves of
                    // a row cluster are neighbours, so an item usually ends up with a single run)
                    HmRun &lr = out.runs.back();
                    HmFreeRun &lf = out.frun.back();
                    if (lr.src + lr.len == rr.src && lf.yoff + lf.kn == fr.yoff && lf.xoff == fr.xoff && rr.len == fr.kn &&
                        lr.len == lf.kn) {
                        lr.len += rr.len;
                        lf.kn += fr.kn;
                        pos += rr.len;
                        continue;
                    }
                }
                HmRun nr2 = rr;
                nr2.pos = pos;
                pos += rr.len;
                out.runs.push_back(nr2);
                out.frun.push_back(fr);
            }
            ni.nrun = (int32_t)out.runs.size() - ni.run0;
            ni.S = pos;
            if (ni.nrun == 0) continue;
            out.zcap = std::max(out.zcap, (int)pos);
            out.items3.push_back(ni);
        }
        // longest items first (the layout's order is by stored words, low-rank runs included): the hardware
        // hands out CTAs in order, so the launch ends on its cheapest items
        auto work = [](const HmItem &it) {
            const int nrh = (it.F + 15) >> 4;
            const int slots = (nrh == 3 || nrh == 5) ? 8 * nrh : 32 * ((it.F + 31) >> 5); // per column, as the kernel maps rows
            return (int64_t)slots * it.S;
        };
        std::stable_sort(out.items3.begin() + out.round_begin[r], out.items3.end(),
                         [&](const HmItem &a, const HmItem &b) { return work(a) > work(b); });
    }
    if (!out.round_begin.empty()) out.round_begin.back() = (int64_t)out.items3.size();

    // ---- can the dense pass evaluate the low-rank part of its rows as well?
    out.fused_eval = false;
    if (out.round_begin.size() == 2) {
        std::vector<int32_t> fin = out.rows.base; // finest row boxes by first row (they tile the rows)
        std::sort(fin.begin(), fin.end(), [&](int32_t u, int32_t v) { return TR.nodes[(size_t)u].p0 < TR.nodes[(size_t)v].p0; });
        out.item_box.assign(out.items3.size(), -1);
        int64_t covered = 0;
        bool ok = true;
        for (size_t i = 0; i < out.items3.size() && ok; i++) {
            const HmItem &it = out.items3[i];
            auto pos = std::upper_bound(fin.begin(), fin.end(), it.out, [&](int64_t row, int32_t u) {
                return row < TR.nodes[(size_t)u].p0;
            });
            while (pos != fin.begin() && TR.nodes[(size_t) * (pos - 1)].np == 0) --pos; // skip empty boxes
            if (pos == fin.begin()) {
                ok = false;
                break;
            }
            // the last box that starts at or before the item's first row and is not empty
            int32_t u = *(pos - 1);
            const HmNestNode &nd = TR.nodes[(size_t)u];
            if (it.out < nd.p0 || it.out + it.F > (int64_t)nd.p0 + nd.np) ok = false;
            out.item_box[i] = u;
            covered += it.F;
        }
        out.fused_eval = ok && covered == L.row_end - L.row_begin;
        if (!out.fused_eval) out.item_box.clear();
    }
    out.fin.assign(TR.nodes.size(), -1);
    for (size_t i = 0; i < out.rows.base.size(); i++) out.fin[(size_t)out.rows.base[i]] = (int32_t)i;
    if (out.fused_eval) {
        for (size_t i = 0; i < out.items3.size(); i++) {
            HmItem it = out.items3[i];
            const int32_t box = out.item_box[i];
            const int32_t r0 = it.run0;
            it.run0 = (int32_t)out.runsp.size();
            // (the merged dense run is cut into pieces of at most 32 columns again: the panel kernel deals
            // whole runs out to its warp groups)
            for (int32_t k = r0; k < r0 + it.nrun; k++) {
                const HmRun &rr = out.runs[(size_t)k];
                const HmFreeRun &fr = out.frun[(size_t)k];
                for (int32_t o = 0; o < rr.len; o += 32) {
                    const int32_t len = std::min<int32_t>(32, rr.len - o);
                    out.runsp.push_back(HmRun{rr.src + o, len, rr.pos + o});
                    out.frunp.push_back(HmFreeRun{fr.mid, fr.half, fr.xoff, fr.yoff + o, 0, len});
                }
            }
            it.nrun = (int32_t)out.runsp.size() - it.run0;
            out.runsp.push_back(HmRun{~(int32_t)(R * out.fin[(size_t)box]), R, it.S});
            out.frunp.push_back(HmFreeRun{0.5 * (TR.ba[(size_t)box] + TR.bb[(size_t)box]),
                                          0.5 * (TR.bb[(size_t)box] - TR.ba[(size_t)box]), it.out, 0, 0, R});
            it.nrun += 1;
            it.S += R;
            if (it.nrun > HM_MAXRUNS) { // outside the panel kernel's run table: no panel form
                out.items3p.clear();
                out.runsp.clear();
                out.frunp.clear();
                out.fused_eval = false;
                out.item_box.clear();
                break;
            }
            out.items3p.push_back(it);
        }
    }
    return "";
}
