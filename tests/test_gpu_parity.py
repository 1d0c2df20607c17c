"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerance: 1e-12 relative ∞-norm in Float64 (BASELINE.json north_star; summation
order differs from the reference).  The oracle is unpinned against Julia (none in
this image) -- see oracle/hm_oracle.h; it is itself checked against the dense kernel
product and the golden fixtures in tests/test_oracle.py.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import (TOL, oracle_tree_from_mirror, plan_from_oracle_tree, random_lowrank_tree, relinf)

pytestmark = pytest.mark.gpu


def _vec(n, seed=0):
    return np.random.default_rng(seed).standard_normal(n)


# ------------------------------------------------------------------ KernelMatrix through the builder
@pytest.mark.parametrize("dist,N", [("cheb", 1000), ("cheb", 4096), ("unif", 4096), ("quad", 1000),
                                    ("cheb", 300), ("unif", 77), ("cheb", 10000)])
def test_kernelmatrix_builder_matches_oracle(hm, O, dist, N):
    x, y, (a, b, c, d) = O.example_points(N, dist)
    K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    plan = plan_from_oracle_tree(hm, O, K)
    v = _vec(N)
    ref = K.matvec(v)
    out = np.zeros(N)
    plan.matvec(v, out, accumulate=False)
    assert relinf(out, ref) <= TOL
    # mul! accumulates: y += H x  (algebra.jl:43,126,272)
    y0 = _vec(N, 1)
    out2 = y0.copy()
    plan.matvec(v, out2, accumulate=True)
    assert relinf(out2, y0 + ref) <= TOL
    st = plan.stats()
    assert st["algorithmic_bytes"] == 8 * K.stored_words() + 16 * N


def test_deterministic_run_to_run(hm, O):
    x, y, (a, b, c, d) = O.example_points(4096, "cheb")
    plan = plan_from_oracle_tree(hm, O, O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d))
    v = _vec(4096)
    outs = []
    for _ in range(3):
        o = np.zeros(4096)
        plan.matvec(v, o, accumulate=False)
        outs.append(o)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


# ------------------------------------------------------------------ on-device assembly
@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
@pytest.mark.parametrize("dist,N", [("cheb", 4096), ("unif", 1000), ("quad", 1000)])
def test_device_assembly_matches_oracle(hm, O, kernel, dist, N):
    x, y, (a, b, c, d) = O.example_points(N, dist)
    Kref = O.kernelmatrix(kernel, x, y, a, b, c, d)
    f = [hm.cauchykernel, hm.coulombkernel, hm.coulombprimekernel, hm.logkernel][kernel]
    K = hm.KernelMatrix(f, x, y, a, b, c, d, device=0)
    assert K.size() == Kref.shape
    v = _vec(N)
    ref = Kref.matvec(v)
    out = K * v
    assert relinf(out, ref) <= TOL
    # factors: bit-identical to the oracle's for the rational kernels (IEEE div/mul/add,
    # same operation order); log() differs by an ulp between libm and CUDA
    plan = K.plan()
    arr, n = Kref.leaves()
    assert plan.num_leaves() == n
    for i in list(range(0, n, max(1, n // 40))) + [n - 1]:
        info = plan.leaf_info(i)
        lf = arr[i]
        assert (info["row0"], info["col0"], info["m"], info["n"]) == (lf.row0, lf.col0, lf.m, lf.n)
        if lf.kind == O.DENSE:
            A = np.ctypeslib.as_array(lf.A, shape=(lf.n, lf.m)).T
            got = plan.read_leaf(i, 3)
            if kernel == 3:
                assert np.allclose(got, A, rtol=1e-15, atol=1e-15)
            else:
                assert np.array_equal(got, A)
        else:
            U = np.ctypeslib.as_array(lf.A, shape=(lf.r, lf.m)).T
            V = np.ctypeslib.as_array(lf.V, shape=(lf.r, lf.n)).T
            F = np.ctypeslib.as_array(lf.S, shape=(lf.r, lf.r)).T
            assert np.array_equal(plan.read_leaf(i, 0), U)
            assert np.array_equal(plan.read_leaf(i, 2), V)
            if kernel == 3:
                assert np.allclose(plan.read_leaf(i, 1), F, rtol=1e-15, atol=1e-15)
            else:
                assert np.array_equal(plan.read_leaf(i, 1), F)


def test_device_assembly_against_dense_golden(hm):
    """examples/Kernel.jl:78 with an actual assertion: K*b against the dense kernel
    product evaluated at 60 digits (tests/golden/cauchy_dense_*.json)."""
    import json
    import os
    for n in (300, 1000):
        g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"cauchy_dense_{n}.json")))
        b = np.array([float.fromhex(h) for h in g["b"]])
        Kb = np.array([float.fromhex(h) for h in g["Kb"]])
        x, y = hm.chebyshevpoints(n), hm.chebyshevpoints(n, 2)
        K = hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0, device=0)
        out = K * b
        assert np.linalg.norm(out - Kb) / np.linalg.norm(Kb) < 1e-13


# ------------------------------------------------------------------ reference test cases (runtests.jl:15-33)
def test_dense_offset_stride_cases(hm, O):
    """The four dense offset/stride cases of test/runtests.jl:15-33, through mul_ with a
    one-block HierarchicalMatrix (the transpose cases use the transposed matrix as the block)."""
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.random((10, 5)))
    x = rng.random(40)
    eps = np.finfo(np.float64).eps

    def one_block(M):
        H = hm.HierarchicalMatrix(np.float64, 1, 1)
        H[hm.Block(1), hm.Block(1)] = np.asfortranarray(M)
        return H

    H, Ht = one_block(A), one_block(A.T)
    y = np.zeros(40)
    hm.mul_(y, H, x, 1, 1)
    assert np.linalg.norm(y[0:10] - A @ x[0:5]) <= 4 * eps * np.linalg.norm(A @ x[0:5])
    assert not y[10:].any()

    y[:] = 0
    hm.mul_(y, H, x, 5, 5, 2, 2)
    assert np.linalg.norm(y[4:23:2] - A @ x[4:13:2]) <= 4 * eps * np.linalg.norm(A @ x[4:13:2])
    assert not y[5:23:2].any() and not y[:4].any() and not y[23:].any()

    y[:] = 0
    hm.mul_(y, Ht, x, 1, 5, 2, 1)
    assert np.linalg.norm(y[0:5] - A.T @ x[4:23:2]) <= 4 * eps * np.linalg.norm(A.T @ x[4:23:2])

    y[:] = 0
    hm.mul_(y, Ht, x, 6, 3, 1, 3)
    assert np.linalg.norm(y[5:18:3] - A.T @ x[2:12]) <= 4 * eps * np.linalg.norm(A.T @ x[2:12])

    # integer data: exact, like the BigFloat cases of runtests.jl:35-52
    Ai = np.asfortranarray(rng.integers(-8, 9, (10, 5)).astype(np.float64))
    xi = rng.integers(-8, 9, 40).astype(np.float64)
    yi = np.zeros(40)
    hm.mul_(yi, one_block(Ai), xi, 5, 5, 2, 2)
    assert np.array_equal(yi[4:23:2], Ai @ xi[4:13:2])


# ------------------------------------------------------------------ HierarchicalMatrix (LowRankMatrix + Matrix)
@pytest.mark.parametrize("n,seed", [(500, 1), (3000, 2), (129, 3)])
def test_hierarchicalmatrix_lowrank_matches_oracle(hm, O, n, seed):
    rng = np.random.default_rng(seed)
    H = random_lowrank_tree(hm, rng, n)
    T = oracle_tree_from_mirror(O, H)
    assert H.size() == T.shape == (n, n)
    v = rng.standard_normal(n)
    ref = T.matvec(v)
    out = H * v
    assert relinf(out, ref) <= TOL
    # strided mul!: k interleaved right-hand sides, column c via (c, c, k, k) -- runtests.jl:23-25
    k = 3
    X = np.asfortranarray(rng.standard_normal((k, n)))
    Y = np.zeros((k, n), order="F")
    Yref = np.zeros((k, n), order="F")
    for c in range(1, k + 1):
        hm.mul_(Y, H, X, c, c, k, k)
        T.mul(Yref.reshape(-1, order="F"), X.reshape(-1, order="F"), c - 1, c - 1, k, k)
    assert relinf(Y, Yref) <= TOL
    # H * X
    Xc = np.asfortranarray(rng.standard_normal((n, 4)))
    Z = H * Xc
    for c in range(4):
        assert relinf(Z[:, c], T.matvec(np.ascontiguousarray(Xc[:, c]))) <= TOL


def test_kernelmatrix_as_lowrankmatrix(hm, O):
    """the same tree with every BarycentricMatrix2D rewritten as a
    LowRankMatrix (U·Uf, Σ, V·Vf from svd(F)) exercises the U Σ V' path."""
    N = 2000
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    arr, n = K.leaves()
    L = hm.lib()
    bld = C.c_void_p()
    hm._lib.check(L.hm_builder_create(C.byref(bld), N, N, 0, 0))
    T = O.Tree.create(1, 1)  # oracle side: apply leaf by leaf
    ref = np.zeros(N)
    v = _vec(N)
    dp = C.POINTER(C.c_double)
    for i in range(n):
        lf = arr[i]
        if lf.kind == O.DENSE:
            hm._lib.check(L.hm_builder_add_dense(bld, lf.A, lf.m, lf.n, max(lf.m, 1), lf.row0, lf.col0))
            A = np.ctypeslib.as_array(lf.A, shape=(lf.n, lf.m)).T
            O.mul_dense(ref, A, v, lf.row0, lf.col0)
        else:
            U = np.ctypeslib.as_array(lf.A, shape=(lf.r, lf.m)).T
            V = np.ctypeslib.as_array(lf.V, shape=(lf.r, lf.n)).T
            F = np.ctypeslib.as_array(lf.S, shape=(lf.r, lf.r)).T
            Uf, S, Vft = np.linalg.svd(F)
            U2 = np.asfortranarray(U @ Uf)
            V2 = np.asfortranarray(V @ Vft.T)
            S = np.ascontiguousarray(S)
            hm._lib.check(L.hm_builder_add_lowrank(bld, U2.ctypes.data_as(dp), max(lf.m, 1), S.ctypes.data_as(dp),
                                                   V2.ctypes.data_as(dp), max(lf.n, 1), lf.m, lf.n, lf.r,
                                                   lf.row0, lf.col0))
            O.mul_lowrank(ref, U2, S, V2, v, lf.row0, lf.col0)
    h = C.c_void_p()
    hm._lib.check(L.hm_plan_finalize(bld, (C.c_int32 * 1)(0), 1, C.byref(h)))
    L.hm_builder_destroy(bld)
    plan = hm.Plan(h.value, 0)
    out = np.zeros(N)
    plan.matvec(v, out, accumulate=False)
    assert relinf(out, ref) <= TOL
    st = plan.stats()
    assert st["n_lowrank"] == n - st["n_dense"] and st["n_bary2d"] == 0


# ------------------------------------------------------------------ edge cases
def test_edge_cases(hm, O):
    rng = np.random.default_rng(5)
    # unassigned blocks are zero (hierarchical.jl:94,115); empty and ragged blocks
    H = hm.HierarchicalMatrix(np.float64, 3, 3)
    sizes_r, sizes_c = [5, 0, 131], [7, 3, 90]
    for m in range(3):
        for n in range(3):
            if (m, n) in ((0, 1), (2, 0)):
                continue  # left unassigned
            if (m + n) % 2 == 0:
                H[hm.Block(m + 1), hm.Block(n + 1)] = np.asfortranarray(rng.standard_normal((sizes_r[m], sizes_c[n])))
            else:
                r = 4
                H[hm.Block(m + 1), hm.Block(n + 1)] = hm.LowRankMatrix(
                    rng.standard_normal((sizes_r[m], r)), rng.standard_normal(r), rng.standard_normal((sizes_c[n], r)))
    T = oracle_tree_from_mirror(O, H)
    assert H.size() == T.shape
    v = rng.standard_normal(H.size(2))
    assert relinf(H * v, T.matvec(v)) <= TOL

    # overlapping leaves add up (legal at the ABI, as in the reference walk)
    L = hm.lib()
    b = C.c_void_p()
    hm._lib.check(L.hm_builder_create(C.byref(b), 300, 200, 0, 0))
    dp = C.POINTER(C.c_double)
    A1 = np.asfortranarray(rng.standard_normal((300, 200)))
    A2 = np.asfortranarray(rng.standard_normal((150, 100)))
    hm._lib.check(L.hm_builder_add_dense(b, A1.ctypes.data_as(dp), 300, 200, 300, 0, 0))
    hm._lib.check(L.hm_builder_add_dense(b, A2.ctypes.data_as(dp), 150, 100, 150, 75, 33))
    U = np.asfortranarray(rng.standard_normal((260, 9)))
    V = np.asfortranarray(rng.standard_normal((199, 9)))
    S = rng.standard_normal(9)
    hm._lib.check(L.hm_builder_add_lowrank(b, U.ctypes.data_as(dp), 260, S.ctypes.data_as(dp), V.ctypes.data_as(dp),
                                           199, 260, 199, 9, 40, 1))
    h = C.c_void_p()
    hm._lib.check(L.hm_plan_finalize(b, None, 1, C.byref(h)))
    L.hm_builder_destroy(b)
    plan = hm.Plan(h.value, 0)
    v = rng.standard_normal(200)
    ref = A1 @ v
    ref[75:225] += A2 @ v[33:133]
    ref[40:300] += U @ (S * (V.T @ v[1:200]))
    out = np.zeros(300)
    plan.matvec(v, out, accumulate=False)
    assert relinf(out, ref) <= 1e-13

    # empty operator and operator with no leaves
    H0 = hm.HierarchicalMatrix(np.float64, 2, 2)
    assert H0.size() == (0, 0)
    assert (H0 * np.zeros(0)).shape == (0,)


def _one_col(hm, A):
    H = hm.HierarchicalMatrix(np.float64, 1, 1)
    H[hm.Block(1), hm.Block(1)] = A
    return H


def test_large_single_blocks(hm, O):
    """Blocks far larger than one work item: a 5000 x 4700 dense block (z longer than the
    shared-memory staging -> several stage-3 rounds) and a tall low-rank block."""
    rng = np.random.default_rng(7)
    H = hm.HierarchicalMatrix(np.float64, 2, 1)
    A = np.asfortranarray(rng.standard_normal((1500, 4700)))
    H[hm.Block(1), hm.Block(1)] = A
    r = 33
    Lr = hm.LowRankMatrix(rng.standard_normal((9000, r)), rng.standard_normal(r), rng.standard_normal((4700, r)))
    H[hm.Block(2), hm.Block(1)] = Lr
    v = rng.standard_normal(4700)
    ref = np.concatenate([A @ v, Lr.U @ (Lr.S * (Lr.V.T @ v))])
    out = H * v
    assert relinf(out, ref) <= TOL
    assert H.plan().stats()["n_stage3_rounds"] >= 2
    # a leaf with more stage-1 partial sums than one warp handles (n > 128 * 4096 columns)
    H3 = hm.HierarchicalMatrix(np.float64, 1, 2)
    L3 = hm.LowRankMatrix(rng.standard_normal((300, 6)), rng.standard_normal(6), rng.standard_normal((540001, 6)))
    B3 = hm.BarycentricMatrix2D(rng.standard_normal((300, 20)), rng.standard_normal((20, 20)),
                                rng.standard_normal((530000, 20)))
    H3[hm.Block(1), hm.Block(1)] = L3
    K3 = hm.KernelMatrix(np.float64, 1, 1)
    K3[hm.Block(1), hm.Block(1)] = B3
    v3 = rng.standard_normal(540001)
    assert relinf(H3.__class__.__mul__(_one_col(hm, L3), v3), L3.U @ (L3.S * (L3.V.T @ v3))) <= TOL
    v4 = rng.standard_normal(530000)
    assert relinf(K3 * v4, B3.U @ (B3.F @ (B3.V.T @ v4))) <= TOL


# ------------------------------------------------------------------ row parts (multi-GPU layout on one device)
@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_row_parts_tile_the_product(hm, O, nparts):
    N = 4096
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    v = _vec(N)
    ref = K.matvec(v)
    out = np.full(N, np.nan)
    covered = np.zeros(N, dtype=int)
    for p in range(nparts):
        plan = plan_from_oracle_tree(hm, O, K, part=p, nparts=nparts)
        st = plan.stats()
        plan.matvec(v, out, accumulate=False)  # writes only the owned rows
        covered[st["row_begin"]:st["row_end"]] += 1
    assert (covered == 1).all()
    assert relinf(out, ref) <= TOL
    # device-assembled parts
    out2 = np.full(N, np.nan)
    for p in range(nparts):
        Kp = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=p, nparts=nparts)
        Kp.plan().matvec(v, out2, accumulate=False)
    assert relinf(out2, ref) <= TOL


# ------------------------------------------------------------------ error behaviour
def test_errors(hm):
    L = hm.lib()
    b = C.c_void_p()
    hm._lib.check(L.hm_builder_create(C.byref(b), 10, 10, 0, 0))
    dp = C.POINTER(C.c_double)
    A = np.zeros((4, 4), order="F")
    assert L.hm_builder_add_dense(b, A.ctypes.data_as(dp), 4, 4, 4, 8, 0) == 4      # HM_ERR_RANGE
    assert L.hm_builder_add_dense(b, A.ctypes.data_as(dp), 4, 4, 3, 0, 0) == 3      # HM_ERR_SHAPE (ld < m)
    assert L.hm_builder_add_dense(b, None, 4, 4, 4, 0, 0) == 2                      # HM_ERR_NULL
    assert L.hm_builder_add_dense(b, A.ctypes.data_as(dp), -1, 4, 4, 0, 0) == 3
    assert b"NULL" in L.hm_last_error() or b"negative" in L.hm_last_error()
    L.hm_builder_destroy(b)
    H = hm.HierarchicalMatrix(np.float64, 1, 1)
    H[hm.Block(1), hm.Block(1)] = np.zeros((3, 3), order="F")
    with pytest.raises(IndexError):
        hm.mul_(np.zeros(2), H, np.zeros(3))
    with pytest.raises(TypeError):
        hm.mul_(np.zeros(3, dtype=np.float32), H, np.zeros(3))
    xs, ys = np.zeros(3), np.zeros(3)
    assert L.hm_matvec(H.plan().handle, xs.ctypes.data_as(dp), 0, ys.ctypes.data_as(dp), 1, 1) == 1   # HM_ERR_INVALID
    assert L.hm_matvec(H.plan().handle, xs.ctypes.data_as(dp), 1, ys.ctypes.data_as(dp), -1, 1) == 1
    assert L.hm_matvec(H.plan().handle, None, 1, ys.ctypes.data_as(dp), 1, 1) == 2                      # HM_ERR_NULL
    assert L.hm_matmat(H.plan().handle, xs.ctypes.data_as(dp), 2, ys.ctypes.data_as(dp), 3, 1, 0) == 3  # ld < n


# ------------------------------------------------------------------ many right-hand sides (DMMA panel kernels)
@pytest.mark.parametrize("nrhs", [2, 16, 17, 33, 64, 65, 130])
def test_matmat_kernelmatrix(hm, O, nrhs):
    N = 3000
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    rng = np.random.default_rng(nrhs)
    X = np.asfortranarray(rng.standard_normal((N, nrhs)))
    Y = K * X
    assert Y.shape == (N, nrhs)
    for c_ in sorted({0, 1, nrhs // 2, nrhs - 1}):
        assert relinf(Y[:, c_], Kref.matvec(np.ascontiguousarray(X[:, c_]))) <= TOL
    # accumulate form with leading dimensions larger than the extents
    Xb = np.zeros((N + 5, nrhs), order="F")
    Xb[:N] = X
    Y0 = np.asfortranarray(rng.standard_normal((N + 3, nrhs)))
    Yb = Y0.copy(order="F")
    dp = C.POINTER(C.c_double)
    hm._lib.check(hm.lib().hm_matmat(K.plan().handle, Xb.ctypes.data_as(dp), N + 5, Yb.ctypes.data_as(dp), N + 3,
                                     nrhs, 1))
    assert relinf(Yb[:N], Y0[:N] + Y) <= TOL
    assert np.array_equal(Yb[N:], Y0[N:])
    # the panel path agrees with the single-vector path to rounding
    y1 = K * np.ascontiguousarray(X[:, 0])
    assert relinf(Y[:, 0], y1) <= 1e-13


def test_matmat_lowrank_wide_and_multiround(hm, O):
    rng = np.random.default_rng(11)
    # generic tree: ragged LowRankMatrix/dense blocks
    H = random_lowrank_tree(hm, rng, 1500)
    T = oracle_tree_from_mirror(O, H)
    X = np.asfortranarray(rng.standard_normal((1500, 24)))
    Y = H * X
    for c_ in (0, 7, 23):
        assert relinf(Y[:, c_], T.matvec(np.ascontiguousarray(X[:, c_]))) <= TOL
    # one block far larger than a work item: z longer than the staging (several stage-3
    # rounds), a fast dimension wider than one pass (rank 150 -> stage-1 F = 150 > 128)
    H2 = hm.HierarchicalMatrix(np.float64, 2, 1)
    A = np.asfortranarray(rng.standard_normal((700, 4700)))
    H2[hm.Block(1), hm.Block(1)] = A
    r = 150
    Lr = hm.LowRankMatrix(rng.standard_normal((2500, r)), rng.standard_normal(r), rng.standard_normal((4700, r)))
    H2[hm.Block(2), hm.Block(1)] = Lr
    X2 = np.asfortranarray(rng.standard_normal((4700, 20)))
    ref = np.vstack([A @ X2, Lr.U @ (Lr.S[:, None] * (Lr.V.T @ X2))])
    Y2 = H2 * X2
    assert relinf(Y2, ref) <= TOL


def test_matmat_row_parts(hm, O):
    N = 4096
    x, y, (a, b, c, d) = O.example_points(N, "unif")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    X = np.asfortranarray(np.random.default_rng(3).standard_normal((N, 16)))
    Y = np.full((N, 16), np.nan, order="F")
    for p in range(3):
        Kp = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=p, nparts=3)
        Kp.plan().matmat(X, Y, accumulate=False)
    for c_ in (0, 15):
        assert relinf(Y[:, c_], Kref.matvec(np.ascontiguousarray(X[:, c_]))) <= TOL


def test_pinned_host_buffers(hm, O):
    """hm_matvec with pinned (page-locked) host vectors, as bench.py's e2e leg uses them."""
    import torch
    N = 4096
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    plan = K.plan()
    v = _vec(N)
    ref = np.zeros(N)
    plan.matvec(v, ref, accumulate=False)
    yp = torch.zeros(N, dtype=torch.float64).pin_memory().numpy()
    xp = torch.from_numpy(v.copy()).pin_memory().numpy()
    plan.matvec(xp, yp, accumulate=False)
    assert np.array_equal(yp, ref)
    plan.matvec(xp, yp, accumulate=True)
    assert relinf(yp, 2 * ref) <= 1e-15


# ------------------------------------------------------------------ rmul!/lmul!/scale! on the packed operator
def test_scale_updates_the_device_operator(hm, O):
    rng = np.random.default_rng(21)
    n = 1800
    H = random_lowrank_tree(hm, rng, n)
    T = oracle_tree_from_mirror(O, H)
    v = rng.standard_normal(n)
    bc, br = rng.standard_normal(n + 3), rng.standard_normal(n + 2)
    _ = H * v                                   # plan exists before the update: updated in place, not rebuilt
    plan_before = H.plan()
    hm.scale_(H, bc, 4)                         # scale!(H, b, jstart = 4): columns
    T.scale_cols(bc, 3)
    assert H.plan() is plan_before
    assert relinf(H * v, T.matvec(v)) <= TOL
    hm.scale_(br, H, 3)                         # scale!(b, H, istart = 3): rows
    T.scale_rows(br, 2)
    assert relinf(H * v, T.matvec(v)) <= TOL
    H.invalidate()                              # the host mirror was kept consistent: a fresh plan agrees
    assert relinf(H * v, T.matvec(v)) <= TOL
    # rmul!/lmul! on a device-assembled KernelMatrix (Diagonally scaled Cauchy operator)
    N = 3000
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    dcol, drow = rng.standard_normal(N), rng.standard_normal(N)
    hm.rmul_(K, dcol)
    hm.lmul_(drow, K)
    Kref.scale_cols(dcol).scale_rows(drow)
    w = rng.standard_normal(N)
    assert relinf(K * w, Kref.matvec(w)) <= TOL
    X = np.asfortranarray(rng.standard_normal((N, 16)))
    Y = K * X
    assert relinf(Y[:, 5], Kref.matvec(np.ascontiguousarray(X[:, 5]))) <= TOL
    # large single blocks: several stage-3 rounds, own stage-1 items
    H2 = hm.HierarchicalMatrix(np.float64, 1, 1)
    A = np.asfortranarray(rng.standard_normal((300, 4500)))
    H2[hm.Block(1), hm.Block(1)] = A.copy(order="F")
    v2 = rng.standard_normal(4500)
    _ = H2 * v2
    b2 = rng.standard_normal(4500)
    hm.rmul_(H2, b2)
    assert relinf(H2 * v2, A @ (b2 * v2)) <= TOL


def test_fused_stage2_option(hm, O):
    """HMB200_FUSE_STAGE2=1 (core apply in the tail of stage 1) gives the same result."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import hmb200_loader; hm = hmb200_loader.load();"
        "x, y = hm.chebyshevpoints(5000), hm.chebyshevpoints(5000, 2);"
        "K = hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0, device=0);"
        "v = np.random.default_rng(0).standard_normal(5000); u = K * v; u2 = K * v;"
        "assert np.array_equal(u, u2); np.save(sys.argv[1], u)"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("0", "1"):
        path = f"/tmp/hm_fuse_{flag}.npy"
        env = dict(os.environ, HMB200_FUSE_STAGE2=flag)
        subprocess.check_call([sys.executable, "-c", code, path], env=env)
        outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1])  # same arithmetic, same order


# ------------------------------------------------------------------ adjoint apply H'x
@pytest.mark.parametrize("dist,N", [("cheb", 4096), ("unif", 1000), ("quad", 3000), ("unif", 77)])
def test_adjoint_kernelmatrix(hm, O, dist, N):
    x, y, (a, b, c, d) = O.example_points(N, dist)
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    w = _vec(N, 5)
    ref = Kref.rmatvec(w)
    out = hm.adjoint(K) * w
    assert relinf(out, ref) <= TOL
    # <w, K v> == <K' w, v>
    v = _vec(N, 6)
    assert abs(w @ (K * v) - out @ v) <= 1e-11 * np.linalg.norm(w) * np.linalg.norm(K * v)
    # mul!(y, K', x): accumulates
    y0 = _vec(N, 7)
    y1 = y0.copy()
    hm.mul_(y1, hm.adjoint(K), w)
    assert relinf(y1, y0 + ref) <= TOL
    assert np.array_equal(hm.adjoint(K) * w, out)  # deterministic


def test_adjoint_generic_trees(hm, O):
    rng = np.random.default_rng(31)
    H = random_lowrank_tree(hm, rng, 2100)
    T = oracle_tree_from_mirror(O, H)
    w = rng.standard_normal(2100)
    assert relinf(hm.adjoint(H) * w, T.rmatvec(w)) <= TOL
    # offsets and strides on both vectors
    k = 3
    X = np.asfortranarray(rng.standard_normal((k, 2100)))
    Y = np.zeros((k, 2100), order="F")
    hm.mul_(Y, hm.adjoint(H), X, 2, 3, k, k)
    assert relinf(Y[1], T.rmatvec(np.ascontiguousarray(X[2]))) <= TOL
    assert not Y[0].any() and not Y[2].any()
    # rectangular, unassigned and empty blocks
    G = hm.HierarchicalMatrix(np.float64, 3, 3)
    sr, sc = [5, 0, 131], [7, 3, 90]
    for m in range(3):
        for n in range(3):
            if (m, n) in ((0, 1), (2, 0)):
                continue
            if (m + n) % 2 == 0:
                G[hm.Block(m + 1), hm.Block(n + 1)] = np.asfortranarray(rng.standard_normal((sr[m], sc[n])))
            else:
                G[hm.Block(m + 1), hm.Block(n + 1)] = hm.LowRankMatrix(
                    rng.standard_normal((sr[m], 4)), rng.standard_normal(4), rng.standard_normal((sc[n], 4)))
    TG = oracle_tree_from_mirror(O, G)
    wg = rng.standard_normal(G.size(1))
    assert relinf(hm.adjoint(G) * wg, TG.rmatvec(wg)) <= TOL
    # blocks far larger than one work item: wide dense block (several rounds), rank-150 leaf
    H2 = hm.HierarchicalMatrix(np.float64, 2, 1)
    A = np.asfortranarray(rng.standard_normal((700, 4700)))
    H2[hm.Block(1), hm.Block(1)] = A
    Lr = hm.LowRankMatrix(rng.standard_normal((2500, 150)), rng.standard_normal(150), rng.standard_normal((4700, 150)))
    H2[hm.Block(2), hm.Block(1)] = Lr
    w2 = rng.standard_normal(3200)
    ref = A.T @ w2[:700] + Lr.V @ (Lr.S * (Lr.U.T @ w2[700:]))
    assert relinf(hm.adjoint(H2) * w2, ref) <= TOL
    # after rmul!/lmul! the adjoint sees the scaled operator
    bc = rng.standard_normal(2100)
    hm.rmul_(H, bc)
    T.scale_cols(bc)
    assert relinf(hm.adjoint(H) * w, T.rmatvec(w)) <= TOL


def test_adjoint_row_parts_sum(hm, O):
    N = 4096
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    w = _vec(N, 8)
    acc = np.zeros(N)
    for p in range(3):
        Kp = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=p, nparts=3)
        acc += hm.adjoint(Kp) * w        # contribution of the part's rows
    assert relinf(acc, Kref.rmatvec(w)) <= TOL


def test_panel_kernel_variants_agree(hm, O):
    """The two implementations of the multi-RHS kernels (register-streaming default,
    HMB200_PANEL=tma bulk-copy/mbarrier pipeline) compute the same products."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import hmb200_loader; hm = hmb200_loader.load();"
        "x, y = hm.chebyshevpoints(6000), hm.chebyshevpoints(6000, 2);"
        "K = hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0, device=0);"
        "X = np.asfortranarray(np.random.default_rng(0).standard_normal((6000, 40)));"
        "np.save(sys.argv[1], np.hstack([K * X, K * X[:, :10], K * X[:, 10:37]]))"   # 64-, 16- and 32-wide panels
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("stream", "tma", "mm"):
        path = f"/tmp/hm_panel_{flag}.npy"
        subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, HMB200_PANEL=flag))
        outs.append(np.load(path))
    for o in outs[1:]:
        assert o.shape == (6000, 77) and relinf(outs[0], o) <= 1e-13
    assert relinf(outs[2][:, 40:50], outs[2][:, :10]) <= 1e-13 and relinf(outs[2][:, 50:], outs[2][:, 10:37]) <= 1e-13
    x, y, (a, b, c, d) = O.example_points(6000, "cheb")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    X = np.asfortranarray(np.random.default_rng(0).standard_normal((6000, 40)))
    assert relinf(outs[1][:, 39], Kref.matvec(np.ascontiguousarray(X[:, 39]))) <= TOL


def test_matvec_device_allgather_abi(hm, O):
    """hm_matvec_device_allgather stores the owned rows into every peer buffer (here three
    buffers on the same device stand in for the ranks' symmetric buffers); row parts written
    by different plans tile each buffer exactly like the all-gather of y would."""
    import torch
    N = 4096
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    v = _vec(N, 9)
    ref = Kref.matvec(v)
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(v).to(dev)
    bufs = [torch.full((N,), float("nan"), dtype=torch.float64, device=dev) for _ in range(3)]
    ptrs = [t.data_ptr() for t in bufs]
    for p in range(3):  # rank p owns part p and writes its rows into all three buffers
        Kp = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=p, nparts=3)
        Kp.plan().matvec_device_allgather(xd.data_ptr(), ptrs, p, accumulate=False,
                                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    for t in bufs:
        assert relinf(t.cpu().numpy(), ref) <= TOL
    # accumulate reads this rank's own buffer
    K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    y0 = _vec(N, 10)
    bufs[1].copy_(torch.from_numpy(y0))
    K.plan().matvec_device_allgather(xd.data_ptr(), ptrs, 1, accumulate=True,
                                     stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for t in bufs:
        assert relinf(t.cpu().numpy(), y0 + ref) <= TOL
    # the same through matrix-free plans (their stage 3 stores into the peer buffers as well)
    for t in bufs:
        t.fill_(float("nan"))
    for p in range(3):
        Kp = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=p, nparts=3, matrix_free=True)
        Kp.plan().matvec_device_allgather(xd.data_ptr(), ptrs, p, accumulate=False,
                                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    for t in bufs:
        assert relinf(t.cpu().numpy(), ref) <= TOL
    L = hm.lib()
    arr = (C.c_uint64 * 1)(0)
    assert L.hm_matvec_device_allgather(K.plan().handle, xd.data_ptr(), arr, 1, 0, 0, None) == 2   # NULL peer
    assert L.hm_matvec_device_allgather(K.plan().handle, xd.data_ptr(), arr, 1, 3, 0, None) == 1   # self out of range


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_leaf_soup(hm, seed):
    """Arbitrary leaves at arbitrary (overlapping) positions -- legal at the ABI, contributions
    add -- against a dense numpy reference: matvec, accumulate, adjoint, panel, scaling, parts."""
    rng = np.random.default_rng(100 + seed)
    nr, ncol = int(rng.integers(50, 900)), int(rng.integers(50, 900))
    D = np.zeros((nr, ncol))
    L = hm.lib()
    dp = C.POINTER(C.c_double)
    leaves = []
    for _ in range(int(rng.integers(1, 40))):
        m, n = int(rng.integers(0, min(nr, 300) + 1)), int(rng.integers(0, min(ncol, 300) + 1))
        r0, c0 = int(rng.integers(0, nr - m + 1)), int(rng.integers(0, ncol - n + 1))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            A = np.asfortranarray(rng.standard_normal((m, n)))
            D[r0:r0 + m, c0:c0 + n] += A
            leaves.append((3, r0, c0, A))
        else:
            r = int(rng.integers(1, 41))
            U, V = np.asfortranarray(rng.standard_normal((m, r))), np.asfortranarray(rng.standard_normal((n, r)))
            if kind == 1:
                S = rng.standard_normal(r)
                D[r0:r0 + m, c0:c0 + n] += (U * S) @ V.T
                leaves.append((2, r0, c0, U, S, V))
            else:
                F = np.asfortranarray(rng.standard_normal((r, r)))
                D[r0:r0 + m, c0:c0 + n] += U @ F @ V.T
                leaves.append((4, r0, c0, U, F, V))

    def build(part=0, nparts=1):
        b = C.c_void_p()
        hm._lib.check(L.hm_builder_create(C.byref(b), nr, ncol, 0, 0))
        for lf in leaves:
            if lf[0] == 3:
                A = lf[3]
                hm._lib.check(L.hm_builder_add_dense(b, A.ctypes.data_as(dp), A.shape[0], A.shape[1],
                                                     max(A.shape[0], 1), lf[1], lf[2]))
            elif lf[0] == 2:
                U, S, V = lf[3:]
                hm._lib.check(L.hm_builder_add_lowrank(b, U.ctypes.data_as(dp), max(U.shape[0], 1), S.ctypes.data_as(dp),
                                                       V.ctypes.data_as(dp), max(V.shape[0], 1), U.shape[0],
                                                       V.shape[0], U.shape[1], lf[1], lf[2]))
            else:
                U, F, V = lf[3:]
                hm._lib.check(L.hm_builder_add_bary2d(b, U.ctypes.data_as(dp), max(U.shape[0], 1), F.ctypes.data_as(dp),
                                                      F.shape[0], V.ctypes.data_as(dp), max(V.shape[0], 1),
                                                      U.shape[0], V.shape[0], U.shape[1], lf[1], lf[2]))
        h = C.c_void_p()
        hm._lib.check(L.hm_plan_finalize_part(b, part, nparts, C.byref(h)))
        L.hm_builder_destroy(b)
        return hm.Plan(h.value, 0)

    scale = max(np.abs(D).max(), 1.0)
    tol = 1e-12 * scale * ncol  # entries of D are sums of products of O(1) numbers
    plan = build()
    v, w = rng.standard_normal(ncol), rng.standard_normal(nr)
    out = np.zeros(nr)
    plan.matvec(v, out, accumulate=False)
    assert np.max(np.abs(out - D @ v)) <= tol
    y0 = rng.standard_normal(nr)
    out2 = y0.copy()
    plan.matvec(v, out2, accumulate=True)
    assert np.max(np.abs(out2 - (y0 + D @ v))) <= tol
    adj = np.zeros(ncol)
    plan.rmatvec(w, adj, accumulate=False)
    assert np.max(np.abs(adj - D.T @ w)) <= tol * nr / ncol + tol
    X = np.asfortranarray(rng.standard_normal((ncol, 19)))
    Y = np.zeros((nr, 19), order="F")
    plan.matmat(X, Y, accumulate=False)
    assert np.max(np.abs(Y - D @ X)) <= tol
    bc, br = rng.standard_normal(ncol), rng.standard_normal(nr)
    plan.scale(bc, 0)
    plan.scale(br, 1)
    plan.matvec(v, out, accumulate=False)
    assert np.max(np.abs(out - br * (D @ (bc * v)))) <= tol * 10
    # row parts tile the product
    res = np.full(nr, np.nan)
    for p in range(3):
        build(p, 3).matvec(v, res, accumulate=False)
    assert np.max(np.abs(res - D @ v)) <= tol


# ---------------------------------------------------------------- EvenBarycentricMatrix
def _cauchy_int(x, j):
    return 1.0 / (x - j)


def test_evenbary_leaf_mul_matches_oracle(hm, O):
    """mul!(u, B::EvenBarycentricMatrix, v, istart, jstart) (algebra.jl:168-239) on the GPU: both
    offset parities, ragged sizes, accumulation into u; also the golden 60-digit product."""
    rng = np.random.default_rng(31)
    for (a, b, c, d) in ((1, 100, 300, 420), (1, 333, 1000, 2501), (5, 6, 40, 40)):
        B = hm.EvenBarycentricMatrix(np.float64, _cauchy_int, a, b, c, d)
        m, n = B.shape
        for (istart, jstart) in ((1, 1), (1, 2), (4, 2), (3, 3)):
            v = rng.standard_normal(jstart - 1 + n)
            u0 = rng.standard_normal(istart - 1 + m + 2)
            ref = O.mul_evenbary(u0.copy(), B.W, B.F, v, istart - 1, jstart - 1)
            got = hm.mul_(u0.copy(), B, v, istart, jstart)
            assert relinf(got, ref) <= TOL
            assert np.array_equal(got[:istart - 1], u0[:istart - 1]) and np.array_equal(got[-2:], u0[-2:])
        assert relinf(B * v[-n:], O.mul_evenbary(np.zeros(m), B.W, B.F, v[-n:])) <= TOL
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "evenbary_cauchy.json")))
    B = hm.EvenBarycentricMatrix(_cauchy_int, g["a"], g["b"], g["c"], g["d"])
    v = np.array([float.fromhex(h) for h in g["v"]])
    for shift in (0, 1):
        ref = np.array([float.fromhex(h) for h in g["u"][str(shift)]])
        vv = np.concatenate([np.zeros(shift), v])
        assert relinf(hm.mul_(np.zeros(B.shape[0]), B, vv, 1, 1 + shift), ref) <= TOL


def test_evenbary_in_block_tree(hm, O):
    """A custom `@hierarchical` type with EvenBarycentricMatrix and Matrix leaves: the walk's
    absolute offsets select each leaf's parity class; adjoint, panels and row parts follow."""
    rng = np.random.default_rng(32)
    T = hm.hierarchical("ParityMatrix", hm.EvenBarycentricMatrix, hm.Matrix)
    H = T(np.float64, 2, 2)
    H[hm.Block(1), hm.Block(1)] = hm.EvenBarycentricMatrix(_cauchy_int, 1, 161, 400, 690)     # 161 x 291
    H[hm.Block(1), hm.Block(2)] = np.asfortranarray(rng.standard_normal((161, 33)))
    H[hm.Block(2), hm.Block(1)] = hm.EvenBarycentricMatrix(_cauchy_int, 1, 245, -700, -410)   # 245 x 291
    inner = T(np.float64, 1, 1)
    inner[hm.Block(1), hm.Block(1)] = hm.EvenBarycentricMatrix(_cauchy_int, 1, 245, 300, 332)  # 245 x 33
    H[hm.Block(2), hm.Block(2)] = inner
    assert H.size() == (406, 324)
    Tr = oracle_tree_from_mirror(O, H)
    for (istart, jstart) in ((1, 1), (2, 1), (3, 6), (2, 2)):
        x = rng.standard_normal(jstart - 1 + 324)
        y0 = rng.standard_normal(istart - 1 + 406)
        ref = Tr.mul(y0.copy(), x, istart - 1, jstart - 1)
        got = hm.mul_(y0.copy(), H, x, istart, jstart)
        assert relinf(got, ref) <= TOL
    x = rng.standard_normal(324)
    assert relinf(H * x, Tr.matvec(x)) <= TOL
    w = rng.standard_normal(406)
    assert relinf(hm.adjoint(H) * w, Tr.rmatvec(w)) <= TOL
    X = np.asfortranarray(rng.standard_normal((324, 5)))
    Y = H * X
    for k in range(5):
        assert relinf(Y[:, k], Tr.matvec(np.ascontiguousarray(X[:, k]))) <= TOL
    res = np.full(406, np.nan)
    for p in range(3):
        hm.flatten(H, 0, p, 3).matvec(x, res, accumulate=False)
    assert relinf(res, Tr.matvec(x)) <= TOL
    # the plan counts the leaf at the reference's size; the zero-interleaved packing is 2x
    st = H.plan().stats()
    assert st["lowrank_words"] == ((161 + 291) + (245 + 291) + (245 + 33)) * 20


def test_barycentricmatrix_lowrank_family(hm, O):
    """barycentricmatrix (BarycentricMatrix.jl:61-89) gives a LowRankMatrix; in a
    HierarchicalMatrix it applies like any other (algebra.jl:110-131)."""
    rng = np.random.default_rng(33)
    H = hm.HierarchicalMatrix(np.float64, 1, 2)
    L1 = hm.barycentricmatrix(np.float64, _cauchy_int, 1, 300, 500, 900)
    H[hm.Block(1), hm.Block(1)] = L1
    H[hm.Block(1), hm.Block(2)] = np.asfortranarray(rng.standard_normal((300, 17)))
    Tr = oracle_tree_from_mirror(O, H)
    x = rng.standard_normal(401 + 17)
    assert relinf(H * x, Tr.matvec(x)) <= TOL
    i = np.arange(1, 301)[:, None]
    j = np.arange(500, 901)[None, :]
    assert relinf((H * x)[:300] - H[hm.Block(1), hm.Block(2)] @ x[401:], (1.0 / (i - j)) @ x[:401]) <= 1e-11


def test_leaf_level_mul_reads_like_runtests(hm, O):
    """test/runtests.jl:15-33 verbatim: mul!(y, A, x, ...) and mul!(y, transpose(A), x, ...) on a
    bare Matrix; then the same call forms for LowRankMatrix (algebra.jl:110-159) and
    BarycentricMatrix2D (algebra.jl:243-277) against the oracle's leaf applies."""
    rng = np.random.default_rng(0)
    eps = np.finfo(np.float64).eps
    A = np.asfortranarray(rng.random((10, 5)))
    x = rng.random(40)
    y = np.zeros_like(x)
    hm.mul_(y, A, x, 1, 1)
    assert np.linalg.norm(y[0:10] - A @ x[0:5]) <= 4 * eps * np.linalg.norm(A @ x[0:5])
    y[:] = 0
    hm.mul_(y, A, x, 5, 5, 2, 2)
    assert np.linalg.norm(y[4:23:2] - A @ x[4:13:2]) <= 4 * eps * np.linalg.norm(A @ x[4:13:2])
    y[:] = 0
    hm.mul_(y, hm.transpose(A), x, 1, 5, 2, 1)
    assert np.linalg.norm(y[0:5] - A.T @ x[4:23:2]) <= 4 * eps * np.linalg.norm(A.T @ x[4:23:2])
    y[:] = 0
    hm.mul_(y, hm.transpose(A), x, 6, 3, 1, 3)
    assert np.linalg.norm(y[5:18:3] - A.T @ x[2:12]) <= 4 * eps * np.linalg.norm(A.T @ x[2:12])
    assert not y[:5].any() and not y[6:18:3].any() and not y[18:].any()
    # exact on integer data, like the BigFloat block runtests.jl:35-52
    Ai = np.asfortranarray(rng.integers(-8, 9, (10, 5)).astype(np.float64))
    xi = rng.integers(-8, 9, 40).astype(np.float64)
    yi = np.zeros(40)
    hm.mul_(yi, hm.transpose(Ai), xi, 6, 3, 1, 3)
    assert np.array_equal(yi[5:18:3], Ai.T @ xi[2:12])

    U, V, S = rng.standard_normal((37, 6)), rng.standard_normal((23, 6)), rng.standard_normal(6)
    Lr = hm.LowRankMatrix(np.asfortranarray(U), S, np.asfortranarray(V))
    x = rng.standard_normal(200)
    y0 = rng.standard_normal(200)
    ref = O.mul_lowrank(y0.copy(), U, S, V, x, 3, 7, 2, 3)
    assert relinf(hm.mul_(y0.copy(), Lr, x, 4, 8, 2, 3), ref) <= TOL
    got = hm.mul_(y0.copy(), hm.adjoint(Lr), x, 2, 5, 3, 2)          # y[2 + 2k] += (V S U') x[5 + 3i]
    ref = y0.copy()
    ref[1:1 + 2 * 23:2] += (V * S) @ (U.T @ x[4:4 + 3 * 37:3])
    assert relinf(got, ref) <= TOL

    Ub, Fb, Vb = rng.standard_normal((50, 20)), rng.standard_normal((20, 20)), rng.standard_normal((31, 20))
    B2 = hm.BarycentricMatrix2D(np.asfortranarray(Ub), np.asfortranarray(Fb), np.asfortranarray(Vb))
    ref = O.mul_bary2d(y0.copy(), Ub, Fb, Vb, x, 9, 4)
    assert relinf(hm.mul_(y0.copy(), B2, x, 10, 5), ref) <= TOL
    with pytest.raises(TypeError):
        hm.mul_(y0.copy(), B2, x, 1, 1, 2, 1)
    with pytest.raises(IndexError):
        hm.mul_(np.zeros(9), A, x, 1, 1)


def test_reference_example_runs(hm, capsys):
    """test/runtests.jl:59 includes examples/Kernel.jl (and asserts nothing); the same script on
    this engine, at its smaller size, with the error it prints bounded."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "kernel.py")
    spec = importlib.util.spec_from_file_location("hm_example_kernel", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    worst = mod.main([1000])
    out = capsys.readouterr().out
    assert out.count("2-norm relative error") == 8     # 4 kernels x 2 point families
    assert worst <= 1e-10


def test_host_copy_pipeline_is_bit_identical(hm, O):
    """hm_matvec pipelines its host copies against the kernels above ~2 MB of vector data (chunked
    x upload ahead of chunk-ordered stage 1, chunked y download behind chunk-ordered stage 3).
    Every output still has one writer and the same summation order: bit-identical to the plain
    path, with accumulation, and for a row part."""
    N = 1 << 19
    x, y, (a, b, c, d) = O.example_points(N, "unif")
    rng = np.random.default_rng(77)
    v, y0 = rng.standard_normal(N), rng.standard_normal(N)
    for part, nparts, mfree in ((0, 1, False), (1, 3, False), (0, 1, True)):
        K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=part, nparts=nparts, matrix_free=mfree)
        P = K.plan()
        outs = []
        for flag in (None, "1"):
            if flag is None:
                os.environ.pop("HMB200_NO_COPY_PIPELINE", None)
            else:
                os.environ["HMB200_NO_COPY_PIPELINE"] = flag
            try:
                u = y0.copy()
                P.matvec(v, u, accumulate=True)
                w = np.full(N, np.nan)
                P.matvec(v, w, accumulate=False)
            finally:
                os.environ.pop("HMB200_NO_COPY_PIPELINE", None)
            outs.append((u, w))
        st = P.stats()
        r0, r1 = st["row_begin"], st["row_end"]
        assert np.array_equal(outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1][r0:r1], outs[1][1][r0:r1])
        assert np.isnan(outs[0][1][:r0]).all() and np.isnan(outs[0][1][r1:]).all()   # only owned rows are written
        assert np.array_equal(outs[0][0][:r0], y0[:r0]) and np.array_equal(outs[0][0][r1:], y0[r1:])
        assert relinf(outs[0][0][r0:r1] - y0[r0:r1], outs[0][1][r0:r1]) <= 1e-12


# ---------------------------------------------------------------- matrix-free apply
@pytest.mark.parametrize("kernel", ["cauchykernel", "coulombkernel", "coulombprimekernel", "logkernel"])
@pytest.mark.parametrize("dist,N", [("cheb", 4096), ("unif", 3000), ("quad", 1000), ("cheb", 77 * 2), ("unif", 20011)])
def test_matrix_free_matches_oracle_and_stored(hm, O, kernel, dist, N):
    """hm_assemble_kernel_free: nothing but the cores is stored, every U, V and dense entry is
    evaluated inside mul!.  Same operator as the stored plan (and as the oracle's KernelMatrix)."""
    x, y, (a, b, c, d) = O.example_points(N, dist)
    f = getattr(hm, kernel)
    Kf = hm.KernelMatrix(f, x, y, a, b, c, d, device=0, matrix_free=True)
    Ks = hm.KernelMatrix(f, x, y, a, b, c, d, device=0)
    Kref = O.kernelmatrix(getattr(O, kernel[:-6].upper()), x, y, a, b, c, d)
    v = _vec(N, 5)
    u = Kf * v
    assert relinf(u, Kref.matvec(v)) <= TOL
    assert relinf(u, Ks * v) <= 1e-13
    y0 = _vec(N + 3, 6)
    got = hm.mul_(y0.copy(), Kf, np.concatenate([np.zeros(2), v]), 4, 3)      # offsets, accumulate
    ref = Kref.mul(y0.copy(), np.concatenate([np.zeros(2), v]), 3, 2)
    assert relinf(got, ref) <= TOL
    st = Kf.plan().stats()
    assert st["u_stream_bytes"] == 0 and st["v_stream_bytes"] == 0 and st["stored_bytes"] == 8 * st["core_words"]
    assert st["algorithmic_bytes"] == Ks.plan().stats()["algorithmic_bytes"]


def test_matrix_free_parts_determinism_and_unsupported(hm, O):
    N = 6000
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    v = _vec(N, 8)
    full = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, matrix_free=True)
    u1, u2 = full * v, full * v
    assert np.array_equal(u1, u2)                                   # run-to-run deterministic
    res = np.full(N, np.nan)
    for p in range(3):
        part = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=p, nparts=3, matrix_free=True)
        part.plan().matvec(v, res, accumulate=False)
    assert relinf(res, u1) <= 1e-13
    P = full.plan()
    with pytest.raises(hm.HmError):
        part.plan().rmatvec(v, np.zeros(N))                         # the adjoint needs the whole operator
    with pytest.raises(hm.HmError):
        P.scale(np.ones(N), 0)
    with pytest.raises(hm.HmError):
        P.read_leaf(0, 3 if P.leaf_info(0)["kind"] == 3 else 0)


@pytest.mark.parametrize("kernel,dist,N", [("cauchykernel", "cheb", 4096), ("cauchykernel", "unif", 3000),
                                           ("coulombkernel", "quad", 1000), ("logkernel", "cheb", 77 * 2),
                                           ("coulombprimekernel", "unif", 20011), ("logkernel", "unif", 9000)])
def test_matrix_free_adjoint(hm, O, kernel, dist, N):
    """y (+)= K'w on a matrix-free plan: the matrix-free plan of the transposed leaves with the point sets
    exchanged (odd kernels through -w).  Against the oracle's adjoint restatement, the stored plan's adjoint,
    and <w, K v> = <K'w, v>; both forms of the matrix-free apply."""
    x, y, (a, b, c, d) = O.example_points(N, dist)
    ref = O.kernelmatrix(getattr(O, kernel[:-6].upper()), x, y, a, b, c, d)
    w, v = _vec(N, 21), _vec(N, 22)
    want = ref.rmatvec(w)
    Ks = hm.KernelMatrix(getattr(hm, kernel), x, y, a, b, c, d, device=0)
    ys = np.zeros(N)
    Ks.plan().rmatvec(w, ys, accumulate=False)
    for env in (None, "cheb"):
        if env:
            os.environ["HMB200_FREE_FORM"] = env
        try:
            Kf = hm.KernelMatrix(getattr(hm, kernel), x, y, a, b, c, d, device=0, matrix_free=True)
            P = Kf.plan()
            got = np.zeros(N)
            P.rmatvec(w, got, accumulate=False)
        finally:
            os.environ.pop("HMB200_FREE_FORM", None)
        assert relinf(got, want) <= TOL
        assert relinf(got, ys) <= 1e-12
        y0 = _vec(N, 23)
        acc = y0.copy()
        P.rmatvec(w, acc, accumulate=True)                           # accumulate, twice the same answer
        assert relinf(acc - y0, got) <= 1e-12
        Kv = Kf * v
        assert abs(np.dot(w, Kv) - np.dot(got, v)) <= 1e-11 * np.linalg.norm(w) * np.linalg.norm(Kv)


@pytest.mark.parametrize("N", [1, 2, 39, 80, 81, 161, 257, 1281])
def test_matrix_free_small_and_ragged(hm, O, N):
    """Matrix-free plans on the sizes where the tree degenerates (one dense leaf, one split, ragged halves):
    matvec, accumulate, 3 and 17 right-hand sides and the adjoint against the oracle; rectangular operators
    (more rows than columns and the other way round)."""
    for nx, ny in ((N, N), (N, max(1, (2 * N) // 3)), (max(1, N // 2), N)):
        i = np.arange(1, nx + 1, dtype=np.float64)
        j = np.arange(1, ny + 1, dtype=np.float64)
        x, y = 1.0 - 2.0 * (i - 0.5) / nx, 1.0 - 2.0 * (j - 0.25) / ny - 1e-3
        ref = O.kernelmatrix(O.CAUCHY, x, y, 1.0, -1.0, 1.0, -1.0)
        K = hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0, device=0, matrix_free=True)
        P = K.plan()
        v, w, y0 = _vec(ny, 41), _vec(nx, 42), _vec(nx, 43)
        want = ref.matvec(v)
        scale = max(np.max(np.abs(want)), 1e-300)
        got = np.zeros(nx)
        P.matvec(v, got, accumulate=False)
        assert np.max(np.abs(got - want)) <= TOL * scale
        acc = y0.copy()
        P.matvec(v, acc, accumulate=True)
        assert np.max(np.abs(acc - y0 - want)) <= 1e-12 * max(scale, np.max(np.abs(y0)))
        for nrhs in (3, 17):
            X = np.asfortranarray(np.random.default_rng(nrhs).standard_normal((ny, nrhs)))
            Y = K * X
            for c in range(nrhs):
                wc = ref.matvec(np.ascontiguousarray(X[:, c]))
                assert np.max(np.abs(Y[:, c] - wc)) <= TOL * max(np.max(np.abs(wc)), 1e-300)
        wt = ref.rmatvec(w)
        gt = np.zeros(ny)
        P.rmatvec(w, gt, accumulate=False)
        assert np.max(np.abs(gt - wt)) <= TOL * max(np.max(np.abs(wt)), 1e-300)


def test_matrix_free_nested_overlap_and_graph(hm, O):
    """Nested-basis matvec: the dense leaves run on the plan's second stream beside the tree passes
    (fork / join by events).  Same result as the single-stream form (HMB200_NEST_OVERLAP=0), with
    accumulate, back to back on one stream, and captured into a CUDA graph and replayed."""
    import torch
    N = 12000   # (even: no first-kind point coincides with a second-kind one)
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    v, y0 = _vec(N, 31), _vec(N, 32)
    ref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d).matvec(v)
    outs = []
    for env in ("0", None):
        if env:
            os.environ["HMB200_NEST_OVERLAP"] = env
        try:
            K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, matrix_free=True)
        finally:
            os.environ.pop("HMB200_NEST_OVERLAP", None)
        assert K.plan().form == 3
        u = K * v
        acc = y0.copy()
        K.plan().matvec(v, acc, accumulate=True)
        assert relinf(u, ref) <= TOL and relinf(acc - y0, ref) <= 1e-12
        outs.append(u)
    assert relinf(outs[0], outs[1]) <= 1e-13
    # device pointers: dependent chain on one stream, eager and as a replayed graph
    P = K.plan()
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(v).to(dev)
    bufs = [torch.zeros(N, dtype=torch.float64, device=dev) for _ in range(3)]
    st = torch.cuda.Stream(device=dev)

    def chain():
        src = xd
        for bb in bufs:
            P.matvec_device(src.data_ptr(), bb.data_ptr(), accumulate=False, stream=st.cuda_stream)
            src = bb

    with torch.cuda.stream(st):
        chain()
    st.synchronize()
    eager = [bb.clone() for bb in bufs]
    assert relinf(eager[0].cpu().numpy(), ref) <= TOL
    for bb in bufs:
        bb.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        chain()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    for e, bb in zip(eager, bufs):
        assert torch.equal(e, bb)


@pytest.mark.parametrize("nrhs", [2, 16, 17, 33, 64, 70])
@pytest.mark.parametrize("kernel,dist,N", [("cauchykernel", "cheb", 4096), ("cauchykernel", "unif", 3000),
                                           ("coulombkernel", "quad", 1000), ("logkernel", "cheb", 77 * 2),
                                           ("coulombprimekernel", "unif", 20011)])
def test_matrix_free_matmat_matches_oracle(hm, O, kernel, dist, N, nrhs):
    """H*X on a matrix-free plan (hm_free_panel.cu): U, V and dense entries are generated in MMA
    fragment layout and consumed by FP64 tensor-core MMAs, for every panel width (16 / 32 / 64 and
    ragged last panels).  Same columns as the oracle's mul! (src/KernelMatrix.jl:9-12 per column)."""
    x, y, (a, b, c, d) = O.example_points(N, dist)
    f = getattr(hm, kernel)
    Kf = hm.KernelMatrix(f, x, y, a, b, c, d, device=0, matrix_free=True)
    Kref = O.kernelmatrix(getattr(O, kernel[:-6].upper()), x, y, a, b, c, d)
    rng = np.random.default_rng(100 + nrhs)
    X = np.asfortranarray(rng.standard_normal((N, nrhs)))
    Y = Kf * X
    assert Y.shape == (N, nrhs)
    for c_ in sorted({0, 1, nrhs // 2, nrhs - 1}):
        assert relinf(Y[:, c_], Kref.matvec(np.ascontiguousarray(X[:, c_]))) <= TOL
    # every column agrees with the plan's own single-vector path to rounding, and the product is deterministic
    for c_ in sorted({0, nrhs - 1}):
        assert relinf(Y[:, c_], Kf * np.ascontiguousarray(X[:, c_])) <= 1e-13
    assert np.array_equal(Y, Kf * X)
    # accumulate form with leading dimensions larger than the extents
    Xb = np.zeros((N + 5, nrhs), order="F")
    Xb[:N] = X
    Y0 = np.asfortranarray(rng.standard_normal((N + 3, nrhs)))
    Yb = Y0.copy(order="F")
    dp = C.POINTER(C.c_double)
    hm._lib.check(hm.lib().hm_matmat(Kf.plan().handle, Xb.ctypes.data_as(dp), N + 5, Yb.ctypes.data_as(dp), N + 3,
                                     nrhs, 1))
    assert relinf(Yb[:N], Y0[:N] + Y) <= TOL
    assert np.array_equal(Yb[N:], Y0[N:])


def test_matrix_free_matmat_row_parts(hm, O):
    """Row parts of a matrix-free plan tile the panel product."""
    N, nrhs = 6000, 24
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    X = np.asfortranarray(np.random.default_rng(3).standard_normal((N, nrhs)))
    whole = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, matrix_free=True) * X
    res = np.full((N, nrhs), np.nan, order="F")
    for p in range(3):
        part = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0, part=p, nparts=3, matrix_free=True)
        part.plan().matmat(X, res, accumulate=False)
    assert relinf(res, whole) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("dist,N", [("cheb", 3000), ("unif", 2048), ("quad", 1000)])
@pytest.mark.parametrize("kname", ["exp", "lorentz"])
def test_arbitrary_kernel_function_matches_oracle(hm, O, dist, N, kname):
    """KernelMatrix(f, x, y, a, b, c, d) takes any f::Function (KernelMatrix.jl:47): a fifth and a
    sixth kernel, exp(-|x - y|) and 1/(1 + (x - y)^2), through hm_assemble_kernel_fn (cores and dense
    leaves evaluated by the host callback, U and V on the device) against the oracle given the same
    function."""
    import math
    x, y, (a, b, c, d) = O.example_points(N, dist)
    if kname == "exp":
        fv, fs = (lambda p, q: np.exp(-np.abs(p - q))), (lambda p, q: math.exp(-abs(p - q)))
    else:
        fv = fs = lambda p, q: 1.0 / (1.0 + (p - q) * (p - q))
    K = hm.KernelMatrix(fv, x, y, a, b, c, d, device=0)
    O.set_user_kernel(fs)
    Kref = O.kernelmatrix(O.USER, x, y, a, b, c, d)
    v = np.random.default_rng(11).standard_normal(N)
    assert relinf(K * v, Kref.matvec(v)) <= TOL
    # the factors themselves: U and V bit-identical as for the built-in kernels; cores and dense leaves
    # bit-identical for the rational kernel (IEEE arithmetic on both sides), to an ulp or two for exp
    # (numpy's vectorised exp vs libm's)
    plan = K.plan()
    arr, n = Kref.leaves()
    assert plan.num_leaves() == n
    same = np.array_equal if kname == "lorentz" else (lambda p, q: np.allclose(p, q, rtol=4e-16, atol=0))
    for i in range(0, n, max(1, n // 40)):
        o = arr[i]
        if o.kind == O.DENSE:
            assert same(plan.read_leaf(i, 3), np.ctypeslib.as_array(o.A, shape=(o.n, o.m)).T)
        else:
            assert same(plan.read_leaf(i, 1), np.ctypeslib.as_array(o.S, shape=(o.r, o.r)).T)
            assert np.array_equal(plan.read_leaf(i, 0), np.ctypeslib.as_array(o.A, shape=(o.r, o.m)).T)
            assert np.array_equal(plan.read_leaf(i, 2), np.ctypeslib.as_array(o.V, shape=(o.r, o.n)).T)
    if kname == "exp":
        return
    # a failing callback surfaces as its Python exception, not as a crash
    with pytest.raises(ZeroDivisionError):
        hm.KernelMatrix(lambda p, q: 1 / 0, x, y, a, b, c, d, device=0)
    # the built-in Cauchy kernel given as a plain function reproduces the device-evaluated operator
    K0 = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    K1 = hm.KernelMatrix(lambda p, q: 1.0 / (p - q), x, y, a, b, c, d, device=0)
    assert np.array_equal(K0 * v, K1 * v)


@pytest.mark.gpu
def test_hierarchical_plus_lowrank_on_device(hm, O):
    """SURVEY 8f row f4, first step: G = H +- L (algebra.jl:394-524, recompressed on the host as in the
    reference) applied on the device against the oracle walk of the same tree and against H*x +- L*x."""
    rng = np.random.default_rng(9)
    n = 1500
    H = random_lowrank_tree(hm, rng, n, leaf=64, r=6)
    L = hm.LowRankMatrix(rng.standard_normal((n, 4)), np.array([3.0, 2.0, 1.0, 0.5]), rng.standard_normal((n, 4)))
    v = rng.standard_normal(n)
    Hx = oracle_tree_from_mirror(O, H).matvec(v)
    Lx = (L.U * L.S) @ (L.V.T @ v)
    for G, sign in ((H + L, 1.0), (H - L, -1.0)):
        out = G * v
        assert relinf(out, oracle_tree_from_mirror(O, G).matvec(v)) <= TOL
        assert relinf(out, Hx + sign * Lx) <= 1e-11


def test_hierarchicalcholesky_factor_on_device(hm, O):
    """R = hierarchicalcholesky(A) (cholesky.jl:12-94, host recursion as in the
    reference) is an ordinary upper-stored HierarchicalMatrix; R*x and R'*y run on the device and
    R'(R x) reproduces A x; the oracle walk of the same factor agrees with the device product."""
    rng = np.random.default_rng(21)
    n = 1600
    x = np.sort(rng.uniform(0.0, 80.0, n))

    def hodlr(p):
        m = len(p)
        if m <= 100:
            return np.asfortranarray(np.exp(-np.abs(p[:, None] - p[None, :])) + 0.5 * np.eye(m))
        h = m // 2
        H = hm.HierarchicalMatrix(2, 2)
        H[hm.Block(1), hm.Block(1)] = hodlr(p[:h])
        H[hm.Block(1), hm.Block(2)] = hm.svdtrunc(np.exp(-np.abs(p[:h, None] - p[None, h:])))
        H[hm.Block(2), hm.Block(2)] = hodlr(p[h:])
        return H

    A = hodlr(x)
    Ad = np.exp(-np.abs(x[:, None] - x[None, :])) + 0.5 * np.eye(n)
    R = hm.hierarchicalcholesky(A)
    v = rng.standard_normal(n)
    Rv = R * v                                                     # device, forward streams
    assert relinf(Rv, oracle_tree_from_mirror(O, R).matvec(v)) <= TOL
    RtRv = hm.adjoint(R) * Rv                                      # device, adjoint apply
    assert relinf(RtRv, Ad @ v) <= 1e-11
    # a solve through the factor: z = R \ (R' \ b) with the host triangular solves of cholesky.jl:150-228
    import scipy.linalg
    w = hm.solvetransposed(R, v)
    z = scipy.linalg.solve_triangular(R.todense(), w, lower=False)
    assert relinf(Ad @ z, v) <= 1e-9
