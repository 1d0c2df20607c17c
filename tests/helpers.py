"""Test helpers: feed oracle-built block trees to the C ABI, error norms."""
import ctypes as C

import numpy as np

_dp = C.POINTER(C.c_double)


def relinf(a, b):
    """‖a − b‖∞ / ‖b‖∞ -- the parity metric of BASELINE.json (tolerance 1e-12)."""
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


TOL = 1e-12  # BASELINE.json north_star: relative ∞-norm error in Float64


class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, device=0):
    """torch view (no copy) of n Float64 words at a raw device address the library handed out."""
    import torch
    return torch.as_tensor(_DevArr(ptr, n), device=f"cuda:{device}")


def push_oracle_leaves(hm, O, tree, builder, parity=0):
    """hm_builder_add_* for every leaf of an oracle tree (pointers into the oracle's arrays)."""
    L = hm.lib()
    arr, n = tree.leaves()
    for i in range(n):
        f = arr[i]
        if f.kind == O.EVENBARY:
            st = L.hm_builder_add_evenbary(builder, f.A, max(f.r, 1), f.V, max(f.n, 1), f.m, f.n, f.r,
                                           f.row0, f.col0, parity)
        elif f.kind == O.DENSE:
            st = L.hm_builder_add_dense(builder, f.A, f.m, f.n, max(f.m, 1), f.row0, f.col0)
        elif f.kind == O.LOWRANK:
            st = L.hm_builder_add_lowrank(builder, f.A, max(f.m, 1), f.S, f.V, max(f.n, 1), f.m, f.n, f.r,
                                          f.row0, f.col0)
        else:
            st = L.hm_builder_add_bary2d(builder, f.A, max(f.m, 1), f.S, max(f.r, 1), f.V, max(f.n, 1), f.m,
                                         f.n, f.r, f.row0, f.col0)
        hm._lib.check(st)
    return n


def plan_from_oracle_tree(hm, O, tree, device=0, part=0, nparts=1, parity=0):
    L = hm.lib()
    nrows, ncols = tree.shape
    b = C.c_void_p()
    hm._lib.check(L.hm_builder_create(C.byref(b), nrows, ncols, 0, device))
    try:
        push_oracle_leaves(hm, O, tree, b, parity)
        h = C.c_void_p()
        if nparts == 1:
            hm._lib.check(L.hm_plan_finalize(b, (C.c_int32 * 1)(device), 1, C.byref(h)))
        else:
            hm._lib.check(L.hm_plan_finalize_part(b, part, nparts, C.byref(h)))
    finally:
        L.hm_builder_destroy(b)
    return hm.Plan(h.value, device)


def stats_from_oracle_tree(hm, O, tree, part=0, nparts=1):
    L = hm.lib()
    nrows, ncols = tree.shape
    b = C.c_void_p()
    hm._lib.check(L.hm_builder_create(C.byref(b), nrows, ncols, 0, -1))
    try:
        push_oracle_leaves(hm, O, tree, b)
        s = hm._lib.Stats()
        hm._lib.check(L.hm_builder_layout_stats(b, part, nparts, C.byref(s)))
    finally:
        L.hm_builder_destroy(b)
    return s.asdict()


def random_lowrank_tree(hm, rng, n, leaf=48, r=7, depth=0, maxdepth=6):
    """HODLR-style HierarchicalMatrix (LowRankMatrix off-diagonal, dense diagonal) of
    size n x n as product-side objects, with ragged splits."""
    if n <= leaf or depth >= maxdepth:
        H = hm.HierarchicalMatrix(np.float64, 1, 1)
        H[hm.Block(1), hm.Block(1)] = np.asfortranarray(rng.standard_normal((n, n)))
        return H
    n1 = int(n * rng.uniform(0.35, 0.65))
    n2 = n - n1
    H = hm.HierarchicalMatrix(np.float64, 2, 2)
    H[hm.Block(1), hm.Block(1)] = random_lowrank_tree(hm, rng, n1, leaf, r, depth + 1, maxdepth)
    H[hm.Block(2), hm.Block(2)] = random_lowrank_tree(hm, rng, n2, leaf, r, depth + 1, maxdepth)
    for (bm, bn, mm, nn) in ((1, 2, n1, n2), (2, 1, n2, n1)):
        U = np.asfortranarray(rng.standard_normal((mm, r)))
        V = np.asfortranarray(rng.standard_normal((nn, r)))
        S = np.abs(rng.standard_normal(r)) + 0.1
        H[hm.Block(bm), hm.Block(bn)] = hm.LowRankMatrix(U, S, V)
    return H


def oracle_tree_from_mirror(O, H):
    """Product-side block tree -> oracle tree (same structure, same data)."""
    T = O.Tree.create(H.M, H.N)
    for m in range(H.M):
        for n in range(H.N):
            A = H._block(m, n)
            if A is None:
                continue
            if hasattr(A, "assigned"):
                T.set_node(m, n, oracle_tree_from_mirror(O, A))
            elif isinstance(A, np.ndarray):
                T.set_dense(m, n, A)
            elif hasattr(A, "W"):
                T.set_evenbary(m, n, A.W, A.F)
            elif hasattr(A, "U") and hasattr(A, "F"):
                T.set_bary2d(m, n, A.U, A.F, A.V)
            else:
                T.set_lowrank(m, n, A.U, A.S, A.V)
    return T
