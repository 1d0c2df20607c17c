// hm_layout.cpp -- host planner.  See DESIGN.md "Data layout in HBM".
//
// The reference applies the operator by a recursive walk over `assigned`
// (/root/reference/src/KernelMatrix.jl:17-45, src/HierarchicalMatrix.jl:24-52),
// one leaf kernel per block (src/algebra.jl:37-48, 110-131, 243-277).  Here the
// same leaves are regrouped so that every GPU thread block streams one
// contiguous slab and every output element has exactly one writer:
//
//   stage 1  t_b = V_b' x          V-stream, grouped by column segment
//   stage 2  s_b = F_b t_b | S_b.*t_b
//   stage 3  y  += U_b s_b + A x   U-stream (+ dense tiles), grouped by row segment
#include "hm_layout.h"

#include <algorithm>
#include <chrono>
void hm_trace_point(const char *name)
{
    static std::chrono::steady_clock::time_point t_prev = std::chrono::steady_clock::now();
    if (!getenv("HMB200_PLAN_TRACE")) return;
    auto t_now = std::chrono::steady_clock::now();
    fprintf(stderr, "plan %s: +%.3f s\n", name, std::chrono::duration<double>(t_now - t_prev).count());
    t_prev = t_now;
}
#define HM_TRACE_POINT(name) hm_trace_point(name)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <new>
#include <numeric>

namespace {

inline bool leaf_live(const HmLeaf &l)
{
    if (l.m <= 0 || l.n <= 0) return false;
    if (l.kind != HM_LEAF_DENSE && (l.ru <= 0 || l.rv <= 0)) return false;
    return true;
}

inline int64_t zlen(const HmLeaf &l) { return l.kind == HM_LEAF_DENSE ? l.n : l.ru; }

inline int64_t core_words_of(const HmLeaf &l)
{
    if (l.kind == HM_LEAF_BARY2D) return (int64_t)l.ru * l.rv;
    if (l.kind == HM_LEAF_LOWRANK) return l.ru;
    return 0;
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Split the sorted boundary list into pieces of at most `cap` elements; pieces
// never cross a boundary.  Returns piece starts plus a final sentinel.
std::vector<int64_t> make_pieces(const std::vector<int64_t> &bounds, int64_t cap, bool even)
{
    std::vector<int64_t> starts;
    for (size_t i = 0; i + 1 < bounds.size(); i++) {
        int64_t s = bounds[i], e = bounds[i + 1], len = e - s;
        if (len <= 0) continue;
        int64_t np = (len + cap - 1) / cap;
        int64_t sz = (len + np - 1) / np;
        if (even && (sz & 1) && sz + 1 <= cap) sz += 1;
        for (int64_t p = s; p < e; p += sz) starts.push_back(p);
    }
    starts.push_back(bounds.empty() ? 0 : bounds.back());
    return starts;
}

// piece -> covering leaves (CSR), leaves in increasing index order.
struct Cover {
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
};

template <class RangeFn>
Cover cover_pieces(const std::vector<int64_t> &starts, size_t nleaves, RangeFn range)
{
    size_t np = starts.size() - 1;
    Cover c;
    c.ptr.assign(np + 1, 0);
    for (int pass = 0; pass < 2; pass++) {
        std::vector<int64_t> cur;
        if (pass == 1) {
            std::partial_sum(c.ptr.begin(), c.ptr.end(), c.ptr.begin());
            c.idx.resize((size_t)c.ptr[np]);
            cur.assign(c.ptr.begin(), c.ptr.end() - 1);
        }
        for (size_t li = 0; li < nleaves; li++) {
            int64_t lo, hi;
            if (!range(li, lo, hi) || hi <= lo) continue;
            size_t p = (size_t)(std::upper_bound(starts.begin(), starts.end() - 1, lo) - starts.begin());
            p = p ? p - 1 : 0;
            for (; p < np && starts[p] < hi; p++) {
                if (starts[p + 1] <= lo) continue;
                if (pass == 0)
                    c.ptr[p + 1]++;
                else
                    c.idx[(size_t)cur[p]++] = (int32_t)li;
            }
        }
    }
    return c;
}

} // namespace

namespace {
thread_local int64_t g_fault_countdown = -1;
}

void hm_fault_arm(int64_t nth) { g_fault_countdown = nth; }

void hm_fault_checkpoint()
{
    if (g_fault_countdown < 0) return;
    if (g_fault_countdown-- == 0) throw std::bad_alloc();
}

std::vector<int64_t> hm_partition_rows(const std::vector<HmLeaf> &leaves, int64_t nrows,
                                       int nparts)
{
    std::vector<int64_t> cuts((size_t)nparts + 1, nrows);
    cuts[0] = 0;
    if (nparts <= 1 || nrows <= 0) return cuts;
    hm_fault_checkpoint();
    // candidate cut points: block-row boundaries, refined every 128 rows
    std::vector<int64_t> b{0, nrows};
    for (const HmLeaf &l : leaves) {
        if (!leaf_live(l)) continue;
        b.push_back(std::min(std::max<int64_t>(l.row0, 0), nrows));
        b.push_back(std::min(std::max<int64_t>(l.row0 + l.m, 0), nrows));
    }
    std::sort(b.begin(), b.end());
    b.erase(std::unique(b.begin(), b.end()), b.end());
    std::vector<int64_t> cand;
    for (size_t i = 0; i + 1 < b.size(); i++)
        for (int64_t p = b[i]; p < b[i + 1]; p += 128) cand.push_back(p);
    cand.push_back(nrows);
    const size_t nc = cand.size();
    // cost of the part [cand[i], cand[j]) = rowc[j] - rowc[i]      (U rows, dense tile rows, y)
    //                                     + vbeg[j] - vend[i]      (V + core of every low-rank leaf
    //                                                               with row0 < cand[j] and row end > cand[i])
    //                                     + ncols                  (x)
    std::vector<double> diff(nc + 1, 0.0), vbeg(nc, 0.0), vend(nc, 0.0);
    double ncols_seen = 0.0;
    for (const HmLeaf &l : leaves) {
        if (!leaf_live(l)) continue;
        int64_t lo = std::max<int64_t>(l.row0, 0), hi = std::min(l.row0 + l.m, nrows);
        if (hi <= lo) continue;
        ncols_seen = std::max(ncols_seen, (double)(l.col0 + l.n));
        size_t ilo = (size_t)(std::lower_bound(cand.begin(), cand.end(), lo) - cand.begin());
        size_t ihi = (size_t)(std::lower_bound(cand.begin(), cand.end(), hi) - cand.begin());
        diff[ilo] += (double)zlen(l);
        diff[ihi] -= (double)zlen(l);
        if (l.kind != HM_LEAF_DENSE) {
            const double vw = (double)l.n * l.rv + (double)core_words_of(l);
            if (ilo + 1 < nc) vbeg[ilo + 1] += vw; // counted by every cut j > ilo (row0 < cand[j])
            vend[ihi] += vw;                       // dropped by every cut i >= ihi (row end <= cand[i])
        }
    }
    std::vector<double> rowc(nc, 0.0);
    double w = 0.0;
    for (size_t i = 0; i + 1 < nc; i++) {
        w += diff[i];
        rowc[i + 1] = rowc[i] + (w + 1.0) * (double)(cand[i + 1] - cand[i]); // +1: the y word
    }
    for (size_t i = 1; i < nc; i++) {
        vbeg[i] += vbeg[i - 1];
        vend[i] += vend[i - 1];
    }
    auto cost = [&](size_t i, size_t j) { return rowc[j] - rowc[i] + vbeg[j] - vend[i] + ncols_seen; };
    // the farthest j > i with cost(i, j) <= T (cost grows with j); at least i + 1
    auto reach = [&](size_t i, double T) {
        size_t lo = i + 1, hi = nc - 1;
        if (cost(i, lo) > T) return lo;
        while (lo < hi) {
            size_t mid = (lo + hi + 1) / 2;
            if (cost(i, mid) <= T)
                lo = mid;
            else
                hi = mid - 1;
        }
        return lo;
    };
    auto parts_needed = [&](double T) {
        size_t i = 0;
        int k = 0;
        while (i < nc - 1) {
            if (cost(i, i + 1) > T) return nparts + 1; // a single candidate interval exceeds T
            i = reach(i, T);
            if (++k > nparts) return k;
        }
        return k;
    };
    double tlo = 0.0, thi = cost(0, nc - 1);
    for (int it = 0; it < 60; it++) {
        double mid = 0.5 * (tlo + thi);
        if (parts_needed(mid) <= nparts)
            thi = mid;
        else
            tlo = mid;
    }
    // greedy cuts for the bound thi: every part but the last is filled up to the bound
    std::vector<size_t> ci((size_t)nparts + 1, nc - 1);
    ci[0] = 0;
    for (int p = 1; p < nparts; p++) ci[(size_t)p] = ci[(size_t)p - 1] < nc - 1 ? reach(ci[(size_t)p - 1], thi) : nc - 1;
    for (int p = 1; p < nparts; p++) cuts[(size_t)p] = cand[ci[(size_t)p]];
    return cuts;
}

std::string hm_build_layout(const std::vector<HmLeaf> &all, int64_t nrows, int64_t ncols, int part,
                            int nparts, const HmLayoutParams &prm, HmLayout &L)
{
    if (nparts < 1 || part < 0 || part >= nparts) return "part index out of range";
    if (nrows < 0 || ncols < 0) return "negative matrix extent";
    if (nrows >= (int64_t)1 << 31 || ncols >= (int64_t)1 << 31) return "matrix extent >= 2^31";
    if (prm.cmax1 > prm.smax || prm.cmax0 > prm.smax) return "layout parameters: cmax > smax";
    hm_fault_checkpoint();
    L = HmLayout();
    L.nrows = nrows;
    L.ncols = ncols;
    L.part = part;
    L.nparts = nparts;

    HM_TRACE_POINT("0");
    // ---- accounting over the whole operator (SURVEY 8d formula) ----
    for (const HmLeaf &l : all) {
        if (l.kind == HM_LEAF_DENSE) {
            L.n_dense++;
            L.dense_words += std::max<int64_t>(l.m, 0) * std::max<int64_t>(l.n, 0);
        } else {
            (l.kind == HM_LEAF_BARY2D ? L.n_bary2d : L.n_lowrank)++;
            int64_t cw = core_words_of(l);
            L.lowrank_words += std::max<int64_t>(l.m, 0) * l.ru + std::max<int64_t>(l.n, 0) * l.rv + cw - l.extra_words;
            L.core_words_all += cw;
        }
    }

    std::vector<int64_t> cuts = hm_partition_rows(all, nrows, nparts);
    const int64_t rlo = cuts[(size_t)part], rhi = cuts[(size_t)part + 1];
    L.row_begin = rlo;
    L.row_end = rhi;

    HM_TRACE_POINT("1");
    // ---- leaves of this part ----
    for (size_t i = 0; i < all.size(); i++) {
        const HmLeaf &l = all[i];
        if (!leaf_live(l)) continue;
        if (l.row0 + l.m <= rlo || l.row0 >= rhi) continue;
        L.leaves.push_back(l);
        L.leaf_global.push_back((int64_t)i);
    }
    const std::vector<HmLeaf> &lv = L.leaves;
    const size_t nl = lv.size();
    if (nl >= ((size_t)1 << 31)) return "too many leaves";

    HM_TRACE_POINT("2");
    // ---- stage 2 tables ----
    std::vector<int32_t> core_of(nl, -1);
    for (size_t i = 0; i < nl; i++) {
        const HmLeaf &l = lv[i];
        int64_t rows = std::min(l.row0 + l.m, rhi) - std::max(l.row0, rlo);
        if (l.kind == HM_LEAF_DENSE) {
            L.part_words += rows * l.n;
            L.part_dense_words += rows * l.n;
            continue;
        }
        if (l.kind == HM_LEAF_LOWRANK && l.ru != l.rv) return "LowRankMatrix leaf with ru != rv";
        HmCoreBlock cb{};
        cb.kind = l.kind;
        cb.ru = l.ru;
        cb.rv = l.rv;
        cb.core = L.core_words;
        cb.soff = (int32_t)L.s_words;
        core_of[i] = (int32_t)L.cores.size();
        L.cores.push_back(cb);
        L.core_leaf.push_back((int32_t)i);
        L.core_words += align_up(core_words_of(l), 2);
        L.s_words += l.ru;
        L.max_r = std::max(L.max_r, std::max(l.ru, l.rv));
        L.part_words += rows * l.ru + l.n * l.rv + core_words_of(l);
        L.part_u_words += rows * l.ru;
        L.part_v_words += l.n * l.rv;
        L.part_core_words += core_words_of(l);
        if (L.s_words >= ((int64_t)1 << 31) - 64) return "stage-2 vector too long";
    }

    // collectors for the adjoint apply
    struct AdjCol {
        int64_t c0, c1, base; // y[j] += PQ[base + j] for c0 <= j < c1
    };
    struct AdjQ {
        int32_t core, k0;
        int64_t off; // PQ offset of the piece's first column (columns k0.. of the leaf)
    };
    std::vector<AdjCol> adj_cols;
    std::vector<AdjQ> adj_q;

    HM_TRACE_POINT("3");
    // ---- stage 3: row segments ----
    {
        std::vector<int64_t> b{rlo, rhi};
        for (const HmLeaf &l : lv) {
            b.push_back(std::min(std::max(l.row0, rlo), rhi));
            b.push_back(std::min(std::max(l.row0 + l.m, rlo), rhi));
        }
        std::sort(b.begin(), b.end());
        b.erase(std::unique(b.begin(), b.end()), b.end());
        std::vector<int64_t> starts = make_pieces(b, prm.rmax, true);
        size_t np = starts.size() - 1;
        Cover cov = cover_pieces(starts, nl, [&](size_t li, int64_t &lo, int64_t &hi) {
            lo = std::max(lv[li].row0, rlo);
            hi = std::min(lv[li].row0 + lv[li].m, rhi);
            return true;
        });

        struct Tmp {
            HmItem it;
            int round;
            int64_t words;
        };
        std::vector<Tmp> tmp;
        tmp.reserve(np);
        // one run / fill entry (and one adjoint entry) per (piece, covering leaf) pair, a few more where a
        // leaf's columns are cut: reserved up front, these vectors hold millions of records at N = 2^22
        L.runs.reserve(L.runs.size() + cov.idx.size() + np);
        L.fill3.reserve(L.fill3.size() + cov.idx.size() + np);
        adj_q.reserve(adj_q.size() + cov.idx.size());
        int64_t slab = 0;
        int maxround = 0;
        int64_t qwords = 0;
        for (size_t p = 0; p < np; p++) {
            int64_t ps = starts[p], pe = starts[p + 1];
            int32_t F = (int32_t)(pe - ps), Fp = F + (F & 1);
            int round = 0;
            HmItem cur{};
            auto open = [&]() {
                cur = HmItem{};
                cur.slab = slab;
                cur.out = ps;
                cur.F = F;
                cur.Fp = Fp;
                cur.S = 0;
                cur.run0 = (int32_t)L.runs.size();
                cur.nrun = 0;
            };
            auto close = [&]() {
                int64_t words = (int64_t)cur.Fp * cur.S;
                cur.aux = qwords; // adjoint: this item's S row-dot results
                qwords += cur.S;
                tmp.push_back(Tmp{cur, round, words});
                slab = align_up(slab + words, 16);
                maxround = std::max(maxround, round);
            };
            open();
            for (int64_t e = cov.ptr[p]; e < cov.ptr[p + 1]; e++) {
                int32_t li = cov.idx[(size_t)e];
                const HmLeaf &l = lv[(size_t)li];
                int64_t total = zlen(l), k = 0;
                while (k < total) {
                    int64_t room = prm.smax - cur.S;
                    // a low-rank leaf's columns stay together (the adjoint sums them as one piece)
                    const bool lr_whole = l.kind != HM_LEAF_DENSE && room < total && total <= prm.smax;
                    if (room <= 0 || cur.nrun >= prm.maxruns || (lr_whole && cur.S > 0)) {
                        close();
                        round++;
                        open();
                        continue;
                    }
                    int64_t take = std::min(total - k, room);
                    HmRun r;
                    r.src = l.kind == HM_LEAF_DENSE ? (int32_t)(l.col0 + k)
                                                    : ~(int32_t)(L.cores[(size_t)core_of[(size_t)li]].soff + k);
                    r.len = (int32_t)take;
                    r.pos = cur.S;
                    L.runs.push_back(r);
                    HmFill f{};
                    f.dst = cur.slab + (int64_t)cur.S * Fp;
                    f.leaf = li;
                    f.off = (int32_t)(ps - l.row0);
                    f.k0 = (int32_t)k;
                    f.kn = (int32_t)take;
                    f.F = F;
                    f.Fp = Fp;
                    f.S = 0;
                    L.fill3.push_back(f);
                    if (l.kind == HM_LEAF_DENSE)
                        adj_cols.push_back(AdjCol{l.col0 + k, l.col0 + k + take, qwords + cur.S - (l.col0 + k)});
                    else
                        adj_q.push_back(AdjQ{core_of[(size_t)li], (int32_t)k, qwords + cur.S});
                    cur.S += (int32_t)take;
                    cur.nrun++;
                    k += take;
                }
            }
            close();
        }
        if (L.runs.size() >= ((size_t)1 << 31)) return "too many stage-3 runs";
        L.ustream_words = slab;
        L.pq_words = qwords;
        std::stable_sort(tmp.begin(), tmp.end(), [](const Tmp &a, const Tmp &b) {
            if (a.round != b.round) return a.round < b.round;
            return a.words > b.words;
        });
        L.items3.reserve(tmp.size());
        L.round_begin.assign((size_t)maxround + 2, 0);
        for (const Tmp &t : tmp) {
            L.items3.push_back(t.it);
            L.round_begin[(size_t)t.round + 1]++;
        }
        for (size_t r = 1; r < L.round_begin.size(); r++) L.round_begin[r] += L.round_begin[r - 1];
    }

    HM_TRACE_POINT("4");
    // ---- stage 1: column segments ----
    hm_fault_checkpoint();
    {
        std::vector<int32_t> npl(L.cores.size(), 0);
        struct Pending {
            int32_t core;
            int32_t off;
        };
        std::vector<Pending> pend; // (core, partial offset) in column order per core
        std::vector<std::pair<int64_t, HmItem>> tmp;
        int64_t slab = 0, pout = 0;

        // class 0: joint slabs over the small leaves
        std::vector<int64_t> b;
        for (const HmLeaf &l : lv)
            if (l.kind != HM_LEAF_DENSE && l.n < prm.nbig) {
                b.push_back(l.col0);
                b.push_back(l.col0 + l.n);
            }
        std::sort(b.begin(), b.end());
        b.erase(std::unique(b.begin(), b.end()), b.end());
        if (b.size() >= 2) {
            std::vector<int64_t> starts = make_pieces(b, prm.cmax0, false);
            size_t np = starts.size() - 1;
            Cover cov = cover_pieces(starts, nl, [&](size_t li, int64_t &lo, int64_t &hi) {
                const HmLeaf &l = lv[li];
                if (l.kind == HM_LEAF_DENSE || l.n >= prm.nbig) return false;
                lo = l.col0;
                hi = l.col0 + l.n;
                return true;
            });
            L.fill1.reserve(L.fill1.size() + cov.idx.size());
            L.s1ent.reserve(L.s1ent.size() + cov.idx.size());
            pend.reserve(pend.size() + cov.idx.size());
            for (size_t p = 0; p < np; p++) {
                if (cov.ptr[p + 1] == cov.ptr[p]) continue;
                int64_t ps = starts[p], pe = starts[p + 1];
                int64_t F = 0;
                for (int64_t e = cov.ptr[p]; e < cov.ptr[p + 1]; e++) F += lv[(size_t)cov.idx[(size_t)e]].rv;
                if (F >= ((int64_t)1 << 24)) return "stage-1 item too wide";
                HmItem it{};
                it.slab = slab;
                it.out = pout;
                it.F = (int32_t)F;
                it.Fp = (int32_t)(F + (F & 1));
                it.S = (int32_t)(pe - ps);
                it.zoff = (int32_t)ps;
                it.aux = L.pq_words;
                adj_cols.push_back(AdjCol{ps, pe, L.pq_words - ps});
                L.pq_words += pe - ps;
                L.adj_max_f = std::max<int>(L.adj_max_f, (int)(F + (F & 1)));
                it.run0 = (int32_t)L.s1ent.size();
                it.nrun = (int32_t)(cov.ptr[p + 1] - cov.ptr[p]);
                int64_t fofs = 0;
                for (int64_t e = cov.ptr[p]; e < cov.ptr[p + 1]; e++) {
                    int32_t li = cov.idx[(size_t)e];
                    const HmLeaf &l = lv[(size_t)li];
                    HmFill f{};
                    f.dst = slab + fofs;
                    f.leaf = li;
                    f.off = (int32_t)(ps - l.col0);
                    f.k0 = 0;
                    f.kn = l.rv;
                    f.F = 0;
                    f.Fp = it.Fp;
                    f.S = it.S;
                    L.fill1.push_back(f);
                    int32_t c = core_of[(size_t)li];
                    pend.push_back(Pending{c, (int32_t)(pout + fofs)});
                    L.s1ent.push_back(c);
                    npl[(size_t)c]++;
                    fofs += l.rv;
                }
                int64_t words = (int64_t)it.Fp * it.S;
                tmp.emplace_back(words, it);
                slab = align_up(slab + words, 16);
                pout += it.Fp;
            }
        }
        // class 1: one leaf per item
        for (size_t li = 0; li < nl; li++) {
            const HmLeaf &l = lv[li];
            if (l.kind == HM_LEAF_DENSE || l.n < prm.nbig) continue;
            int64_t np = (l.n + prm.cmax1 - 1) / prm.cmax1;
            int64_t sz = (l.n + np - 1) / np;
            const int64_t rleaf = L.pq_words; // the leaf's n row-dot results are contiguous
            adj_cols.push_back(AdjCol{l.col0, l.col0 + l.n, rleaf - l.col0});
            L.pq_words += l.n;
            L.adj_max_f = std::max<int>(L.adj_max_f, l.rv + (l.rv & 1));
            for (int64_t c0 = 0; c0 < l.n; c0 += sz) {
                int64_t c1 = std::min(c0 + sz, l.n);
                HmItem it{};
                it.slab = slab;
                it.out = pout;
                it.F = l.rv;
                it.Fp = l.rv + (l.rv & 1);
                it.S = (int32_t)(c1 - c0);
                it.zoff = (int32_t)(l.col0 + c0);
                it.aux = rleaf + c0;
                HmFill f{};
                f.dst = slab;
                f.leaf = (int32_t)li;
                f.off = (int32_t)c0;
                f.k0 = 0;
                f.kn = l.rv;
                f.Fp = it.Fp;
                f.S = it.S;
                L.fill1.push_back(f);
                int32_t c = core_of[li];
                pend.push_back(Pending{c, (int32_t)pout});
                it.run0 = (int32_t)L.s1ent.size();
                it.nrun = 1;
                L.s1ent.push_back(c);
                npl[(size_t)c]++;
                int64_t words = (int64_t)it.Fp * it.S;
                tmp.emplace_back(words, it);
                slab = align_up(slab + words, 16);
                pout += it.Fp;
            }
        }
        if (pout >= ((int64_t)1 << 31) - 64) return "stage-1 partial array too long";
        L.vstream_words = slab;
        L.partial_words = pout;
        std::stable_sort(tmp.begin(), tmp.end(),
                         [](const std::pair<int64_t, HmItem> &a, const std::pair<int64_t, HmItem> &b) {
                             return a.first > b.first;
                         });
        L.items1.reserve(tmp.size());
        for (auto &t : tmp) L.items1.push_back(t.second);
        // partial lists, column order per core (class-0 pieces were emitted in
        // increasing column order, class-1 pieces likewise; a leaf is in one class)
        int32_t acc = 0;
        for (size_t c = 0; c < L.cores.size(); c++) {
            L.cores[c].pl0 = acc;
            L.cores[c].npl = 0;
            acc += npl[c];
        }
        L.plist.assign((size_t)acc, 0);
        for (const Pending &pd : pend) {
            HmCoreBlock &cb = L.cores[(size_t)pd.core];
            L.plist[(size_t)(cb.pl0 + cb.npl++)] = pd.off;
        }
    }

    HM_TRACE_POINT("5");
    // ---- chunked orderings for the host-pointer path ----
    {
        const int NC = HM_NCHUNK;
        L.xchunk.assign((size_t)NC + 1, ncols);
        L.ychunk.assign((size_t)NC + 1, rhi);
        // Unequal chunks: only the first upload and the last download are exposed (nothing computes beside
        // them), so x goes up as 1/16, 3/16, 4/16, 8/16 of the columns and y comes down as 8/16, 4/16, 3/16,
        // 1/16 of the rows; every later copy is shorter than the kernels it hides behind.
        // (HMB200_CHUNKS=equal: four quarters, the round-1 form.)
        static_assert(HM_NCHUNK == 4, "chunk fractions are written for four chunks");
        const char *ce = getenv("HMB200_CHUNKS");
        const bool equal = ce && ce[0] == 'e';
        int64_t xcum[5] = {0, equal ? 16 : 4, equal ? 32 : 16, equal ? 48 : 32, 64};
        int64_t ycum[5] = {0, equal ? 16 : 32, equal ? 32 : 48, equal ? 48 : 60, 64};
        if (ce && ce[0] >= '0' && ce[0] <= '9') { // experiment: "a,b,c" = cumulative 64ths of x; y mirrored
            int a = 0, b = 0, c = 0;
            if (sscanf(ce, "%d,%d,%d", &a, &b, &c) == 3 && 0 < a && a < b && b < c && c < 64) {
                xcum[1] = a, xcum[2] = b, xcum[3] = c;
                ycum[1] = 64 - c, ycum[2] = 64 - b, ycum[3] = 64 - a;
            }
        }
        for (int k = 0; k < NC; k++) {
            L.xchunk[(size_t)k] = (ncols * xcum[k] / 64) & ~(int64_t)511;
            L.ychunk[(size_t)k] = rlo;
        }
        auto chunk_of = [&](const std::vector<int64_t> &bnd, int64_t pos) {
            int k = (int)(std::upper_bound(bnd.begin(), bnd.end() - 1, pos) - bnd.begin()) - 1;
            return std::min(std::max(k, 0), NC - 1);
        };
        // stage 1: by the chunk that holds the last column the item reads
        std::vector<std::pair<int, size_t>> key;
        for (size_t i = 0; i < L.items1.size(); i++)
            key.emplace_back(chunk_of(L.xchunk, (int64_t)L.items1[i].zoff + std::max(L.items1[i].S, 1) - 1), i);
        std::stable_sort(key.begin(), key.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        L.c1_begin.assign((size_t)NC + 1, 0);
        for (auto &kv : key) {
            L.items1c.push_back(L.items1[kv.second]);
            L.c1_begin[(size_t)kv.first + 1]++;
        }
        for (int k = 0; k < NC; k++) L.c1_begin[(size_t)k + 1] += L.c1_begin[(size_t)k];
        // stage 3 (only when everything fits one round): by row chunk; chunk boundaries sit on
        // item boundaries so that each chunk's rows are one contiguous range of y
        if (L.round_begin.size() == 2) {
            std::vector<size_t> order(L.items3.size());
            std::iota(order.begin(), order.end(), (size_t)0);
            std::stable_sort(order.begin(), order.end(),
                             [&](size_t a, size_t b) { return L.items3[a].out < L.items3[b].out; });
            const int64_t rows = rhi - rlo;
            size_t pos = 0;
            std::vector<std::vector<size_t>> groups((size_t)NC);
            for (int k = 0; k < NC; k++) {
                const int64_t target = rlo + rows * ycum[k + 1] / 64;
                L.ychunk[(size_t)k] = pos < order.size() ? L.items3[order[pos]].out : rhi;
                while (pos < order.size() && (k == NC - 1 || L.items3[order[pos]].out < target)) groups[(size_t)k].push_back(order[pos++]);
            }
            L.ychunk[0] = rlo;
            L.c3_begin.assign((size_t)NC + 1, 0);
            for (int k = 0; k < NC; k++) {
                auto &g = groups[(size_t)k];
                std::stable_sort(g.begin(), g.end(), [&](size_t a, size_t b) {
                    return (int64_t)L.items3[a].Fp * L.items3[a].S > (int64_t)L.items3[b].Fp * L.items3[b].S;
                });
                for (size_t i : g) L.items3c.push_back(L.items3[i]);
                L.c3_begin[(size_t)k + 1] = (int64_t)L.items3c.size();
            }
        }
    }

    HM_TRACE_POINT("6");
    // ---- adjoint tables ----
    {
        if (L.pq_words >= ((int64_t)1 << 31) - 64) return "adjoint work buffer too long";
        L.core_q0.assign(L.cores.size(), 0);
        L.core_qn.assign(L.cores.size(), 0);
        for (const AdjQ &q : adj_q) {
            if (q.k0 != 0) return "internal: split low-rank entry";
            L.core_qn[(size_t)q.core]++;
        }
        int32_t acc = 0;
        for (size_t c = 0; c < L.cores.size(); c++) {
            L.core_q0[c] = acc;
            acc += L.core_qn[c];
            L.core_qn[c] = 0;
        }
        L.qlist.assign((size_t)acc, 0);
        for (const AdjQ &q : adj_q) // emitted in increasing row order per leaf
            L.qlist[(size_t)(L.core_q0[(size_t)q.core] + L.core_qn[(size_t)q.core]++)] = (int32_t)q.off;
        // final gather: column intervals on which the set of contributors is constant
        std::vector<int64_t> b{0, ncols};
        for (const AdjCol &a : adj_cols) {
            b.push_back(a.c0);
            b.push_back(a.c1);
        }
        std::sort(b.begin(), b.end());
        b.erase(std::unique(b.begin(), b.end()), b.end());
        std::vector<int64_t> starts = make_pieces(b, 256, false);
        size_t np = starts.size() - 1;
        Cover cov = cover_pieces(starts, adj_cols.size(), [&](size_t i, int64_t &lo, int64_t &hi) {
            lo = adj_cols[i].c0;
            hi = adj_cols[i].c1;
            return true;
        });
        L.colsegs.reserve(np);
        L.colbases.reserve(cov.idx.size());
        for (size_t p = 0; p < np; p++) {
            HmColSeg cs;
            cs.c0 = (int32_t)starts[p];
            cs.c1 = (int32_t)starts[p + 1];
            cs.b0 = (int32_t)L.colbases.size();
            cs.nb = (int32_t)(cov.ptr[p + 1] - cov.ptr[p]);
            for (int64_t e = cov.ptr[p]; e < cov.ptr[p + 1]; e++) L.colbases.push_back(adj_cols[(size_t)cov.idx[(size_t)e]].base);
            L.colsegs.push_back(cs);
        }
        if (L.colbases.size() >= ((size_t)1 << 31)) return "too many adjoint gather entries";
    }

    for (const HmItem &it : L.items1)
        if ((int64_t)it.Fp * it.S >= ((int64_t)1 << 31)) return "stage-1 item too large";
    for (const HmItem &it : L.items3)
        if ((int64_t)it.Fp * it.S >= ((int64_t)1 << 31)) return "stage-3 item too large";
    HM_TRACE_POINT("end");
    return "";
}
