"""The CPU oracle against everything that can pin it without Julia:
  - the assertions of the reference's own tests (test/runtests.jl:5-7, :15-52),
  - golden fixtures computed independently with mpmath (tests/golden/make_golden.py),
  - the example's own yardstick, the dense kernel product (examples/Kernel.jl:78).
The oracle stays "parity unpinned" against the Julia implementation itself (no julia in
the image, no golden vectors in the reference) -- see oracle/hm_oracle.h.
"""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hex(v):
    return np.array([float.fromhex(h) for h in v])


def test_blockrank_even_and_values(O):
    # runtests.jl:5-7: BLOCKRANK(T) is even
    assert O.blockrank(np.float64) % 2 == 0 and O.blockrank(np.float32) % 2 == 0
    assert O.blockrank(np.float64) == 20 and O.blockrank(np.float32) == 10
    assert O.blocksize() == 80  # BLOCKSIZE = 4 BLOCKRANK


def test_cheb_nodes_weights_golden(O):
    g = json.load(open(os.path.join(GOLD, "cheb20.json")))
    nodes, weights = _hex(g["nodes"]), _hex(g["weights"])
    x, lam = O.chebyshevpoints(20), O.chebyshevbarycentricweights(20)
    # correctly rounded sin(pi q) from mpmath vs the long-double evaluation: <= 1 ulp
    assert np.max(np.abs(x - nodes)) <= np.finfo(float).eps
    assert np.max(np.abs(lam - weights)) <= np.finfo(float).eps
    assert np.array_equal(x, -x[::-1]) and np.all(np.diff(x) < 0)
    assert lam[0] > 0 and lam[1] < 0


def test_point_samples_golden(O):
    g = json.load(open(os.path.join(GOLD, "points_samples.json")))
    for key, rec in g.items():
        n, kind = (int(a) for a in key.split("_"))
        x = O.chebyshevpoints(n, kind)
        got = x[np.array(rec["k"]) - 1]
        assert np.max(np.abs(got - _hex(rec["x"])) / np.abs(_hex(rec["x"]))) <= 2 * np.finfo(float).eps
        assert np.all(np.diff(x) < 0)  # descending, as the assembler requires


def test_dense_leaf_reference_cases(O):
    """test/runtests.jl:15-33 -- the only numerical assertions the reference has on this path."""
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.random((10, 5)))
    x = rng.random(40)
    eps = np.finfo(float).eps
    y = np.zeros(40)
    O.mul_dense(y, A, x, 0, 0)                       # mul!(y, A, x, 1, 1)
    assert np.linalg.norm(y[:10] - A @ x[:5]) <= eps * np.linalg.norm(A @ x[:5]) * 4
    y[:] = 0
    O.mul_dense(y, A, x, 4, 4, 2, 2)                 # mul!(y, A, x, 5, 5, 2, 2)
    assert np.linalg.norm(y[4:23:2] - A @ x[4:13:2]) <= eps * np.linalg.norm(A @ x[4:13:2]) * 4
    y[:] = 0
    O.mul_dense(y, A, x, 0, 4, 2, 1, transpose=True)  # mul!(y, transpose(A), x, 1, 5, 2, 1)
    assert np.linalg.norm(y[:5] - A.T @ x[4:23:2]) <= eps * np.linalg.norm(A.T @ x[4:23:2]) * 4
    y[:] = 0
    O.mul_dense(y, A, x, 5, 2, 1, 3, transpose=True)  # mul!(y, transpose(A), x, 6, 3, 1, 3)
    assert np.linalg.norm(y[5:18:3] - A.T @ x[2:12]) <= eps * np.linalg.norm(A.T @ x[2:12]) * 4


def test_dense_leaf_exact_cases(O):
    """runtests.jl:35-52 uses BigFloat and `==`; integer data is exact in Float64."""
    rng = np.random.default_rng(1)
    A = np.asfortranarray(rng.integers(-9, 10, (10, 5)).astype(float))
    x = rng.integers(-9, 10, 40).astype(float)
    y = np.zeros(40)
    O.mul_dense(y, A, x, 0, 0)
    assert np.array_equal(y[:10], A @ x[:5])
    y[:] = 0
    O.mul_dense(y, A, x, 4, 4, 2, 2)
    assert np.array_equal(y[4:23:2], A @ x[4:13:2])
    y[:] = 0
    O.mul_dense(y, A, x, 0, 4, 2, 1, transpose=True)
    assert np.array_equal(y[:5], A.T @ x[4:23:2])
    y[:] = 0
    O.mul_dense(y, A, x, 5, 2, 1, 3, transpose=True)
    assert np.array_equal(y[5:18:3], A.T @ x[2:12])
    # accumulate semantics: mul! adds into y
    y0 = rng.integers(-9, 10, 40).astype(float)
    y = y0.copy()
    O.mul_dense(y, A, x, 0, 0)
    assert np.array_equal(y[:10], y0[:10] + A @ x[:5]) and np.array_equal(y[10:], y0[10:])


def test_lowrank_and_bary_leaves(O):
    rng = np.random.default_rng(2)
    m, n, r = 37, 53, 6
    U, V, S = rng.standard_normal((m, r)), rng.standard_normal((n, r)), rng.standard_normal(r)
    F = rng.standard_normal((r, r))
    x = rng.standard_normal(3 * n + 5)
    y = np.zeros(2 * m + 7)
    O.mul_lowrank(y, U, S, V, x, 3, 2, 3, 2)
    ref = U @ (S * (V.T @ x[2:2 + 3 * n:3]))  # no conjugation, V[j,k]*x (algebra.jl:118)
    assert np.allclose(y[3:3 + 2 * m:2], ref, rtol=1e-13, atol=1e-13)
    u = np.zeros(m + 4)
    O.mul_bary2d(u, U, F, V, x, 4, 1)
    assert np.allclose(u[4:], U @ (F @ (V.T @ x[1:1 + n])), rtol=1e-13, atol=1e-13)
    assert not u[:4].any()


def test_size_rule_and_walk(O):
    """size(): rows from the LAST block column, columns from the FIRST block row
    (hierarchical.jl:36-44); unassigned blocks are zero; walk == dense product."""
    rng = np.random.default_rng(3)
    T = O.Tree.create(2, 2)
    A11, A12 = rng.standard_normal((4, 6)), rng.standard_normal((4, 3))
    A22 = rng.standard_normal((5, 3))
    T.set_dense(0, 0, A11)
    T.set_dense(0, 1, A12)
    T.set_dense(1, 1, A22)  # (1,0) unassigned
    assert T.shape == (9, 9)
    assert T.assigned(1, 0) == 0 and T.assigned(0, 0) == 3 and T.blocksize(1, 0, 1) == 0
    D = np.zeros((9, 9))
    D[:4, :6], D[:4, 6:], D[4:, 6:] = A11, A12, A22
    G = np.array([[T.getindex(i, j) for j in range(9)] for i in range(9)])
    assert np.array_equal(G, D)
    x = rng.standard_normal(9)
    assert np.allclose(T.matvec(x), D @ x, rtol=1e-14, atol=1e-14)
    # nested + low-rank + codes
    P = O.Tree.create(1, 2)
    P.set_node(0, 0, T)
    U, S, V = rng.standard_normal((9, 3)), rng.standard_normal(3), rng.standard_normal((7, 3))
    P.set_lowrank(0, 1, U, S, V)
    assert P.assigned(0, 0) == 1 and P.assigned(0, 1) == 2 and P.shape == (9, 16)
    Dp = np.hstack([D, (U * S) @ V.T])
    xx = rng.standard_normal(16)
    assert np.allclose(P.matvec(xx), Dp @ xx, rtol=1e-13, atol=1e-13)
    assert abs(P.getindex(2, 12) - Dp[2, 12]) < 1e-13
    # strided walk (HierarchicalMatrix.jl:24-52): offsets advance by INCX*cols, INCY*rows
    k = 3
    X = rng.standard_normal(16 * k)
    Y = np.zeros(9 * k)
    P.mul(Y, X, 1, 2, k, k)
    assert np.allclose(Y[1::k], Dp @ X[2::k], rtol=1e-13, atol=1e-13)
    assert not Y[0::k].any() and not Y[2::k].any()


def test_indsplit(O):
    x = np.array([0.9, 0.7, 0.5, 0.2, -0.1, -0.6])
    assert O.indsplit(x, 0, 6, 1.0, -1.0) == ((0, 4), (4, 6))       # midpoint 0: x >= 0 first
    assert O.indsplit(x, 0, 6, 1.0, 0.0) == ((0, 3), (3, 6))        # x[i] >= 0.5 inclusive
    assert O.indsplit(x, 2, 6, 0.0, -1.0) == ((2, 5), (5, 6))
    assert O.indsplit(x, 0, 3, 0.0, -1.0) == ((0, 3), (3, 3))       # everything in the first half
    assert O.indsplit(x, 3, 6, 1.0, 0.9) == ((3, 3), (3, 6))        # nothing in the first half
    # reference quirk: the first element is read before the range is tested
    assert O.indsplit(x, 2, 2, 1.0, -1.0) == ((2, 3), (3, 2))
    with pytest.raises(IndexError):
        O.indsplit(x, 6, 6, 1.0, -1.0)                              # BoundsError


@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
@pytest.mark.parametrize("dist,N", [("cheb", 1000), ("quad", 1000), ("unif", 2500)])
def test_kernelmatrix_vs_dense_product(O, kernel, dist, N):
    """examples/Kernel.jl:61-107 with an assertion: 2-norm relative error of K*b against
    the dense kernel product (long double)."""
    x, y, (a, b, c, d) = O.example_points(N, dist)
    K = O.kernelmatrix(kernel, x, y, a, b, c, d)
    assert K.shape == (N, N)
    v = np.random.default_rng(4).standard_normal(N)
    u = K.matvec(v)
    ref = O.dense_kernel_matvec_ld(kernel, x, y, v)
    assert np.linalg.norm(u - ref) / np.linalg.norm(ref) < 5e-14
    # all-cores variant performs the same per-leaf arithmetic
    u2 = np.zeros(N)
    K.mul_omp(u2, v, 4)
    assert np.max(np.abs(u2 - u)) <= 1e-13 * np.max(np.abs(u))


def test_kernelmatrix_vs_mpmath_golden(O):
    for n in (300, 1000):
        g = json.load(open(os.path.join(GOLD, f"cauchy_dense_{n}.json")))
        b, Kb = _hex(g["b"]), _hex(g["Kb"])
        x, y, (a, bb, c, d) = O.example_points(n, "cheb")
        u = O.kernelmatrix(O.CAUCHY, x, y, a, bb, c, d).matvec(b)
        assert np.linalg.norm(u - Kb) / np.linalg.norm(Kb) < 1e-13
        assert np.linalg.norm(O.dense_kernel_matvec_ld(O.CAUCHY, x, y, b) - Kb) / np.linalg.norm(Kb) < 1e-15


def test_tree_shape_matches_survey(O):
    """274 dense + 510 low-rank leaves at N = 4096 (Chebyshev), 190 + 342 and
    23 237 376 algorithmic bytes for the uniform set."""
    for dist, nd, nl in (("cheb", 274, 510), ("unif", 190, 342)):
        x, y, (a, b, c, d) = O.example_points(4096, dist)
        K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
        arr, n = K.leaves()
        dense = sum(1 for i in range(n) if arr[i].kind == O.DENSE)
        assert (dense, n - dense) == (nd, nl)
        for i in range(n):
            lf = arr[i]
            if lf.kind == O.DENSE:
                assert lf.m < 80 and lf.n < 80  # below BLOCKSIZE
            else:
                assert lf.r == 20
        if dist == "unif":
            assert 8 * K.stored_words() + 16 * 4096 == 23237376
    # bary factors: rows of U and V sum to one (row normalisation, BarycentricMatrix.jl:263-270)
    U, F, V = O.bary2d_build(O.CAUCHY, 1.0, 0.5, -0.5, -1.0, x, 0, 50, y, 3000, 3100)
    assert np.allclose(U.sum(axis=1), 1.0, atol=1e-13) and np.allclose(V.sum(axis=1), 1.0, atol=1e-13)
    assert np.allclose(U @ F @ V.T, 1.0 / (x[:50, None] - y[None, 3000:3100]), rtol=1e-13)


def test_scale_walks(O):
    """scale!(H, b, jstart) / scale!(b, H, istart) (HierarchicalMatrix.jl:54-108, leaf rules
    algebra.jl:280-315): H*Diagonal(b) and Diagonal(b)*H."""
    rng = np.random.default_rng(8)
    T = O.Tree.create(2, 2)
    A11, A22 = rng.standard_normal((4, 6)), rng.standard_normal((5, 3))
    U, S, V = rng.standard_normal((4, 2)), rng.standard_normal(2), rng.standard_normal((3, 2))
    T.set_dense(0, 0, A11)
    T.set_lowrank(0, 1, U, S, V)
    T.set_dense(1, 1, A22)
    D = np.zeros((9, 9))
    D[:4, :6], D[:4, 6:], D[4:, 6:] = A11, (U * S) @ V.T, A22
    bc, br = rng.standard_normal(12), rng.standard_normal(11)
    T.scale_cols(bc, 2)
    T.scale_rows(br, 1)
    Dn = br[1:10, None] * D * bc[None, 2:11]
    G = np.array([[T.getindex(i, j) for j in range(9)] for i in range(9)])
    assert np.allclose(G, Dn, rtol=1e-14, atol=1e-14)


def test_adjoint_walk(O):
    """y += H'x with the reference's transposed leaf rules (algebra.jl:52-65, 138-159)."""
    rng = np.random.default_rng(12)
    for dist, N in (("cheb", 900), ("quad", 500)):
        x, y, (a, b, c, d) = O.example_points(N, dist)
        K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
        w, v = rng.standard_normal(N), rng.standard_normal(N)
        D = 1.0 / (x[:, None] - y[None, :])
        assert np.linalg.norm(K.rmatvec(w) - D.T @ w) / np.linalg.norm(D.T @ w) < 1e-13
        assert abs(w @ K.matvec(v) - K.rmatvec(w) @ v) <= 1e-12 * np.linalg.norm(w) * np.linalg.norm(K.matvec(v))


# ---------------------------------------------------------------- EvenBarycentricMatrix
def _cauchy_int(x, j):
    return 1.0 / (x - j)


def test_evenbary_golden(O):
    """w, W and the masked product against the 60-digit evaluation of BarycentricMatrix.jl:18-45
    and algebra.jl:168-239 (tests/golden/make_golden.py: evenbary)."""
    g = json.load(open(os.path.join(GOLD, "evenbary_cauchy.json")))
    w, W, F = O.evenbary_factors(_cauchy_int, g["a"], g["b"], g["c"], g["d"])
    idx = np.array(g["w_idx"])
    assert np.max(np.abs(w[idx] / _hex(g["w"]) - 1)) <= 1e-14
    for i in idx:
        ref = _hex(g["W_cols"][str(i)])
        assert np.max(np.abs(W[:, i] - ref)) <= 1e-14 * np.max(np.abs(ref))
    v = _hex(g["v"])
    m, n = W.shape[1], F.shape[0]
    for shift in (0, 1):
        u = np.zeros(m + 3)
        vv = np.concatenate([np.zeros(2 + shift), v])
        O.mul_evenbary(u, W, F, vv, 2, 2 + shift)   # mul!(u, B, v, 3, 3 + shift)
        ref = _hex(g["u"][str(shift)])
        assert np.all(u[:2] == 0) and u[-1] == 0
        assert np.max(np.abs(u[2:2 + m] - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_evenbary_interpolates_and_masks(O):
    """The factors reproduce f on the grid (well-separated ranges) and mul! keeps exactly the
    entries whose absolute i+j is even, for either parity of the offsets and ragged sizes."""
    rng = np.random.default_rng(5)
    for (a, b, c, d) in ((1, 100, 300, 420), (7, 93, -250, -120), (1, 2, 50, 52), (10, 40, 100, 100)):
        w, W, F = O.evenbary_factors(_cauchy_int, a, b, c, d)
        M = (F @ W).T
        i = np.arange(a, b + 1)[:, None]
        j = np.arange(c, d + 1)[None, :]
        assert np.max(np.abs(M - 1.0 / (i - j))) <= 1e-13 * np.max(np.abs(M))
        m, n = M.shape
        for (i0, j0) in ((0, 0), (3, 4), (5, 5), (0, 1)):
            v = rng.standard_normal(j0 + n)
            u0 = rng.standard_normal(i0 + m)
            u = u0.copy()
            O.mul_evenbary(u, W, F, v, i0, j0)
            ii = np.arange(m)[:, None] + i0
            jj = np.arange(n)[None, :] + j0
            ref = (M * ((ii + jj) % 2 == 0)) @ v[j0:]
            assert np.max(np.abs(u[i0:] - u0[i0:] - ref)) <= 1e-13 * max(np.max(np.abs(ref)), 1e-300)
            assert np.array_equal(u[:i0], u0[:i0])
        # getindex keeps its own rule: size-parity, not offset-parity (BarycentricMatrix.jl:51)
        for (p, q) in ((0, 0), (0, 1), (m - 1, n - 1)):
            val = O.evenbary_getindex(W, F, p, q)
            assert val == (0.0 if (m + n + p + q) % 2 else float(np.dot(F[q, :], W[:, p]))) or \
                abs(val - float(np.dot(F[q, :], W[:, p]))) <= 1e-15 * abs(val)


def test_evenbary_tree_walks(O):
    """EvenBarycentricMatrix leaves inside a block tree: the walk hands each leaf its absolute
    offsets (KernelMatrix.jl:24-41), so the active parity class follows row0+col0; adjoint and the
    threaded walk agree with the dense form."""
    rng = np.random.default_rng(6)
    _, W1, F1 = O.evenbary_factors(_cauchy_int, 1, 61, 200, 290)      # 61 x 91
    _, W2, F2 = O.evenbary_factors(_cauchy_int, 1, 45, -300, -210)    # 45 x 91
    D1 = rng.standard_normal((61, 33))
    D2 = rng.standard_normal((45, 33))
    T = O.Tree.create(2, 2)
    T.set_evenbary(0, 0, W1, F1)
    T.set_dense(0, 1, D1)
    T.set_evenbary(1, 0, W2, F2)
    T.set_dense(1, 1, D2)
    assert T.shape == (106, 124) and T.assigned(0, 0) == 2 and T.assigned(0, 1) == 3
    assert T.stored_words() == (61 + 91) * 20 + (45 + 91) * 20 + 61 * 33 + 45 * 33
    M1, M2 = (F1 @ W1).T, (F2 @ W2).T

    def dense(shift):
        A = np.zeros((106, 124))
        ii, jj = np.arange(106)[:, None], np.arange(124)[None, :]
        A[:61, :91] = M1
        A[61:, :91] = M2
        A[:, :91] *= ((ii + jj[:, :91] + shift) % 2 == 0)
        A[:61, 91:] = D1
        A[61:, 91:] = D2
        return A

    for (i0, j0) in ((0, 0), (2, 3)):
        A = dense((i0 + j0) % 2)
        x = rng.standard_normal(j0 + 124)
        y = np.zeros(i0 + 106)
        T.mul(y, x, i0, j0)
        ref = A @ x[j0:]
        assert np.max(np.abs(y[i0:] - ref)) <= 1e-13 * np.max(np.abs(ref))
        y2 = np.zeros(i0 + 106)
        T.mul_omp(y2, x, 3, i0, j0)
        assert np.max(np.abs(y2 - y)) <= 1e-13 * np.max(np.abs(ref))
    xr = rng.standard_normal(106)
    ref = dense(0).T @ xr
    assert np.max(np.abs(T.rmatvec(xr) - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_bary2d_block_golden(O):
    """U, F, V of one BarycentricMatrix2D block against the 60-digit evaluation of
    BarycentricMatrix.jl:147-178 / 248-297 (tests/golden/make_golden.py: bary2d_block)."""
    g = json.load(open(os.path.join(GOLD, "bary2d_block.json")))
    x, y = _hex(g["x"]), _hex(g["y"])
    U, F, V = O.bary2d_build(O.CAUCHY, g["a"], g["b"], g["c"], g["d"], x, 0, len(x), y, 0, len(y))
    Ug = np.array([_hex(r) for r in g["U"]])
    Vg = np.array([_hex(r) for r in g["V"]])
    Fg = np.array([_hex(r) for r in g["F"]])
    assert np.max(np.abs(F - Fg) / np.abs(Fg)) <= 4 * np.finfo(float).eps
    # lambda/(x - node) loses digits only where x nearly hits a node; rows are normalised to sum 1
    assert np.max(np.abs(U - Ug)) <= 1e-13 * np.max(np.abs(Ug))
    assert np.max(np.abs(V - Vg)) <= 1e-13 * np.max(np.abs(Vg))
    assert np.max(np.abs(U.sum(axis=1) - 1)) <= 1e-14 and np.max(np.abs(V.sum(axis=1) - 1)) <= 1e-14
    # and the block it represents is the kernel: U F V' = 1/(x - y) to interpolation accuracy
    K = 1.0 / (x[:, None] - y[None, :])
    assert np.max(np.abs(U @ F @ V.T - K)) <= 1e-13 * np.max(np.abs(K))
