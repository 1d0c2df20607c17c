"""CPU-side tests of the product: the C-ABI library loads and exports every symbol of
include/hmb200.h, the host planner (tree of index ranges, stream layout, row partition)
and the Python mirror of the reference API.  No compute calls (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import stats_from_oracle_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_abi_exports_every_declared_symbol(hm):
    header = open(os.path.join(ROOT, "include", "hmb200.h")).read()
    declared = set(re.findall(r"HM_API\s+[\w\s\*]+?\b(hm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    assert declared == set(hm._lib.SIGNATURES), declared ^ set(hm._lib.SIGNATURES)
    L = hm.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.hm_version() >= 100
    assert L.hm_blockrank_f64() == 20 and L.hm_blocksize_f64() == 80


def test_constants_mirror(hm):
    # BLOCKRANK is even for every float type (runtests.jl:5-7); values of SURVEY §2
    assert [hm.BLOCKRANK(t) for t in (np.float64, np.float32, np.float16)] == [20, 10, 4]
    assert hm.BLOCKRANK(np.complex128) == 20
    assert hm.BLOCKSIZE(np.float64) == 80
    assert all(hm.BLOCKRANK(t) % 2 == 0 for t in (np.float16, np.float32, np.float64))


def test_product_points_match_oracle(hm, O):
    for n in (20, 1000, 4097):
        for kind in (1, 2):
            assert np.array_equal(hm.chebyshevpoints(n, kind), O.chebyshevpoints(n, kind))


@pytest.mark.parametrize("dist,N", [("cheb", 4096), ("unif", 4096), ("quad", 1000), ("cheb", 300), ("unif", 50)])
def test_kernel_tree_leaves_match_oracle(hm, O, dist, N):
    """The product's own range-tree builder against the oracle's assembled tree: same
    leaves, same walk order, same offsets (KernelMatrix.jl:47-116, :17-45)."""
    x, y, (a, b, c, d) = O.example_points(N, dist)
    K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    arr, n = K.leaves()
    dp = C.POINTER(C.c_double)
    cnt = C.c_int64()
    buf = (hm._lib.TreeLeaf * n)()
    hm._lib.check(hm.lib().hm_kernel_tree_leaves(x.ctypes.data_as(dp), N, y.ctypes.data_as(dp), N, a, b, c, d,
                                                 buf, n, C.byref(cnt)))
    assert cnt.value == n
    for i in range(n):
        o, p = arr[i], buf[i]
        assert p.kind == (3 if o.kind == O.DENSE else 4)
        assert (p.row0, p.col0, p.m, p.n) == (o.row0, o.col0, o.m, o.n)
        if o.kind != O.DENSE:
            assert p.rank == 20
            # the box of the block reproduces the oracle's factors
            if i % 37 == 0:
                U, F, V = O.bary2d_build(O.CAUCHY, p.a, p.b, p.c, p.d, x, p.xi0, p.xi0 + p.m, y, p.yj0, p.yj0 + p.n)
                assert np.array_equal(U, np.ctypeslib.as_array(o.A, shape=(o.r, o.m)).T)
                assert np.array_equal(F, np.ctypeslib.as_array(o.S, shape=(o.r, o.r)).T)
                assert np.array_equal(V, np.ctypeslib.as_array(o.V, shape=(o.r, o.n)).T)


def test_kernel_tree_reference_error(hm, O):
    # all points in the upper half of the box: the reference recurses into the empty
    # trailing range N+1:N, whose indsplit reads x[N+1] -> BoundsError
    x = np.linspace(1.0, 0.5, 200)
    dp = C.POINTER(C.c_double)
    s = hm._lib.Stats()
    rc = hm.lib().hm_assemble_kernel_stats(x.ctypes.data_as(dp), 200, x.ctypes.data_as(dp), 200, 1.0, -1.0, 1.0, -1.0,
                                           0, 1, C.byref(s))
    assert rc == 9  # HM_ERR_REFERENCE
    assert b"BoundsError" in hm.lib().hm_last_error()
    with pytest.raises(RuntimeError):
        O.kernelmatrix(O.CAUCHY, x, x, 1.0, -1.0, 1.0, -1.0)
    # a cluster of >= BLOCKSIZE coincident points can never be bisected: the reference recurses
    # until StackOverflowError; the planner reports it instead of overflowing its own stack
    xc = np.concatenate([np.linspace(1.0, 0.31, 50), np.full(200, 0.3), np.linspace(0.29, -1.0, 50)])
    rc = hm.lib().hm_assemble_kernel_stats(xc.ctypes.data_as(dp), 300, xc.ctypes.data_as(dp), 300, 1.0, -1.0, 1.0, -1.0,
                                           0, 1, C.byref(s))
    assert rc == 9 and (b"StackOverflow" in hm.lib().hm_last_error() or b"BoundsError" in hm.lib().hm_last_error())
    # ... while a box that fits the points is fine
    rc = hm.lib().hm_assemble_kernel_stats(x.ctypes.data_as(dp), 200, x.ctypes.data_as(dp), 200, 1.0, 0.5, 1.0, 0.5,
                                           0, 1, C.byref(s))
    assert rc == 0 and s.nrows == 200


def test_kernel_tree_bisection_equals_linear_scan(hm, O):
    """indsplit (BarycentricMatrix.jl:299-307) is a linear scan; on point sets verified to be non-increasing
    the planner bisects instead.  Same leaves -- ranges, offsets, boxes, error returns -- as the literal scan
    (HMB200_TREE_LINEAR=1) on graded, uniform, tied, rectangular and ill-fitting inputs and on the sizes where
    ranges become empty."""
    dp = C.POINTER(C.c_double)

    def leaves(x, y, a, b, c, d):
        cnt = C.c_int64()
        rc = hm.lib().hm_kernel_tree_leaves(x.ctypes.data_as(dp), len(x), y.ctypes.data_as(dp), len(y), a, b, c, d,
                                            None, 0, C.byref(cnt))
        if rc:
            return ("error", rc)
        arr = (hm._lib.TreeLeaf * max(cnt.value, 1))()
        hm._lib.check(hm.lib().hm_kernel_tree_leaves(x.ctypes.data_as(dp), len(x), y.ctypes.data_as(dp), len(y),
                                                     a, b, c, d, arr, cnt.value, C.byref(cnt)))
        return [(l.kind, l.rank, l.row0, l.col0, l.m, l.n, l.xi0, l.yj0, l.a, l.b, l.c, l.d) for l in arr[:cnt.value]]

    cases = [O.example_points(N, dist) for N in (1, 2, 3, 39, 79, 80, 81, 160, 161, 257, 1000, 4096, 20011)
             for dist in ("cheb", "unif", "quad")]
    rng = np.random.default_rng(1)
    for N in (500, 5000):
        x = np.sort(np.round(rng.uniform(-1, 1, N), 3))[::-1].copy()      # ties
        y = np.sort(rng.uniform(-1, 1, N // 2))[::-1].copy()              # rectangular
        cases.append((x, y, (1.0, -1.0, 1.0, -1.0)))
        cases.append((x, y, (0.5, -0.5, 2.0, -2.0)))                      # boxes that do not fit the points
    xs = np.linspace(1.0, -1.0, 3000)
    xs[10], xs[11] = xs[11], xs[10]                                       # not monotone: literal scan either way
    cases.append((xs, xs.copy(), (1.0, -1.0, 1.0, -1.0)))
    try:
        for x, y, (a, b, c, d) in cases:
            os.environ.pop("HMB200_TREE_LINEAR", None)
            fast = leaves(x, y, a, b, c, d)
            os.environ["HMB200_TREE_LINEAR"] = "1"
            assert leaves(x, y, a, b, c, d) == fast
    finally:
        os.environ.pop("HMB200_TREE_LINEAR", None)


def test_layout_stats_match_survey(hm, O):
    """leaf counts and algorithmic bytes (the roofline numerator)."""
    for dist, nd, nl in (("cheb", 274, 510), ("unif", 190, 342)):
        x, y, (a, b, c, d) = O.example_points(4096, dist)
        st = hm.KernelMatrix.layout_stats(x, y, a, b, c, d)
        assert (st["n_dense"], st["n_bary2d"], st["n_lowrank"]) == (nd, nl, 0)
        K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
        assert st["algorithmic_bytes"] == 8 * K.stored_words() + 16 * 4096
        assert st["dense_words"] + st["lowrank_words"] == K.stored_words()
        assert st["core_words"] == 400 * nl
        # builder path (leaves pushed one by one) lays out identically
        st2 = stats_from_oracle_tree(hm, O, K)
        for k in ("algorithmic_bytes", "stored_bytes", "n_stage1_items", "n_stage3_items", "partial_bytes"):
            assert st[k] == st2[k], k
        # padding overhead of the packed streams stays small
        assert st["stored_bytes"] <= 1.02 * 8 * K.stored_words()
        assert st["part_v_words"] + st["part_u_words"] + st["part_core_words"] + st["part_dense_words"] == K.stored_words()
    assert st["algorithmic_bytes"] == 23237376  # uniform set, SURVEY 8(d)


def test_layout_stats_2pow20(hm):
    """62 794 / 125 460 leaves (Chebyshev), 49 150 / 98 214 and
    14 021 327 616 B (uniform) at N = 2^20."""
    n = 1 << 20
    x, y = hm.chebyshevpoints(n), hm.chebyshevpoints(n, 2)
    st = hm.KernelMatrix.layout_stats(x, y, 1.0, -1.0, 1.0, -1.0)
    assert (st["n_dense"], st["n_bary2d"]) == (62794, 125460)
    assert abs(st["algorithmic_bytes"] - 14078717200) <= 4096  # survey's figure, to its sinpi rounding
    i = np.arange(1, n + 1, dtype=np.float64)
    st = hm.KernelMatrix.layout_stats(1.0 - 2.0 * (i - 0.5) / n, 1.0 - 2.0 * (i - 0.25) / n, 1.0, -1.0, 1.0, -1.0)
    assert (st["n_dense"], st["n_bary2d"]) == (49150, 98214)
    assert st["algorithmic_bytes"] == 14021327616
    assert st["stored_bytes"] < 1.01 * st["algorithmic_bytes"]


@pytest.mark.parametrize("nparts", [2, 3, 4, 8])
def test_row_partition(hm, O, nparts):
    x, y, (a, b, c, d) = O.example_points(20000, "cheb")
    whole = hm.KernelMatrix.layout_stats(x, y, a, b, c, d)
    parts = [hm.KernelMatrix.layout_stats(x, y, a, b, c, d, p, nparts) for p in range(nparts)]
    assert parts[0]["row_begin"] == 0 and parts[-1]["row_end"] == 20000
    for p, q in zip(parts, parts[1:]):
        assert p["row_end"] == q["row_begin"]
    words = [p["part_words"] for p in parts]
    # balanced by stored words; V and F of blocks that straddle a cut are replicated
    assert max(words) <= 1.15 * (sum(words) / nparts)
    total = whole["dense_words"] + whole["lowrank_words"]
    assert total <= sum(words) <= 1.25 * total
    assert sum(p["part_u_words"] + p["part_dense_words"] for p in parts) == whole["part_u_words"] + whole["part_dense_words"]


@pytest.mark.parametrize("dist", ["cheb", "unif"])
def test_row_partition_balances_true_bytes_2pow20(hm, dist):
    """hm_partition_rows minimises the largest per-part byte count *including* the V / F
    replicas of leaves that straddle a cut."""
    n = 1 << 20
    if dist == "cheb":
        x, y = hm.chebyshevpoints(n), hm.chebyshevpoints(n, 2)
    else:
        i = np.arange(1, n + 1, dtype=np.float64)
        x, y = 1.0 - 2.0 * (i - 0.5) / n, 1.0 - 2.0 * (i - 0.25) / n
    whole = hm.KernelMatrix.layout_stats(x, y, 1.0, -1.0, 1.0, -1.0)["algorithmic_bytes"]
    for nparts in (2, 4, 8):
        b = [hm.KernelMatrix.layout_stats(x, y, 1.0, -1.0, 1.0, -1.0, p, nparts)["part_algorithmic_bytes"]
             for p in range(nparts)]
        # (the uniform tree on 4 parts has its optimum at the clean quarter cuts, 2.4 % apart:
        # moving a cut into a block replicates more V words than it moves)
        assert max(b) <= 1.03 * (sum(b) / nparts), (nparts, b)
        # against the unreplicated ideal: the round-1 cuts were at 1.131 (cheb) / 1.122 (unif) for 8
        assert max(b) <= (1.0 + 0.01 * (nparts + 1)) * whole / nparts, (nparts, b)


def test_no_exception_crosses_the_abi(hm, O):
    """every entry point returns a status; a host allocation failure inside the
    planner or the tree builder becomes HM_ERR_NOMEM, not std::terminate."""
    L = hm.lib()
    dp = C.POINTER(C.c_double)
    x, y, (a, b, c, d) = O.example_points(3000, "cheb")
    s = hm._lib.Stats()
    args = (x.ctypes.data_as(dp), 3000, y.ctypes.data_as(dp), 3000, a, b, c, d)
    assert L.hm_assemble_kernel_stats(*args, 0, 2, C.byref(s)) == 0
    hit = 0
    for nth in range(8):
        L.hm_debug_fail_alloc(nth)
        rc = L.hm_assemble_kernel_stats(*args, 0, 2, C.byref(s))
        L.hm_debug_fail_alloc(-1)
        if rc != 0:
            assert rc == 6 and b"memory" in L.hm_last_error()  # HM_ERR_NOMEM
            hit += 1
    assert hit >= 4  # tree entry, leaf emission, layout entry, partition, stage-1 tables
    cnt = C.c_int64()
    L.hm_debug_fail_alloc(0)
    assert L.hm_kernel_tree_leaves(*args, None, 0, C.byref(cnt)) == 6
    L.hm_debug_fail_alloc(-1)
    # structure-only builder: the layout of pushed leaves
    bld = C.c_void_p()
    hm._lib.check(L.hm_builder_create(C.byref(bld), 100, 100, 0, -1))
    hm._lib.check(L.hm_builder_add_dense(bld, None, 100, 100, 100, 0, 0))
    L.hm_debug_fail_alloc(0)
    assert L.hm_builder_layout_stats(bld, 0, 1, C.byref(s)) == 6
    L.hm_debug_fail_alloc(-1)
    assert L.hm_builder_layout_stats(bld, 0, 1, C.byref(s)) == 0 and s.n_dense == 1
    L.hm_builder_destroy(bld)
    # the library is still usable afterwards
    assert L.hm_assemble_kernel_stats(*args, 0, 1, C.byref(s)) == 0 and s.nrows == 3000


def test_builder_validation_and_structure_only_mode(hm):
    L = hm.lib()
    b = C.c_void_p()
    assert L.hm_builder_create(C.byref(b), 10, 10, 1, -1) == 8      # dtype: only Float64
    assert L.hm_builder_create(C.byref(b), -1, 10, 0, -1) == 3
    hm._lib.check(L.hm_builder_create(C.byref(b), 100, 80, 0, -1))  # structure only, pointers unused
    assert L.hm_builder_add_dense(b, None, 10, 10, 10, 95, 0) == 4  # HM_ERR_RANGE
    assert b"outside" in L.hm_last_error()
    assert L.hm_builder_add_dense(b, None, -1, 10, 10, 0, 0) == 3   # HM_ERR_SHAPE
    assert L.hm_builder_add_lowrank(b, None, 1, None, None, 1, 5, 5, -2, 0, 0) == 3
    hm._lib.check(L.hm_builder_add_dense(b, None, 60, 80, 60, 0, 0))
    hm._lib.check(L.hm_builder_add_lowrank(b, None, 40, None, None, 80, 40, 80, 5, 60, 0))
    hm._lib.check(L.hm_builder_add_bary2d(b, None, 40, None, 5, None, 30, 40, 30, 5, 60, 50))  # overlaps: legal
    hm._lib.check(L.hm_builder_add_dense(b, None, 0, 7, 1, 3, 3))   # empty block
    s = hm._lib.Stats()
    hm._lib.check(L.hm_builder_layout_stats(b, 0, 1, C.byref(s)))
    assert (s.n_dense, s.n_lowrank, s.n_bary2d) == (2, 1, 1)
    assert s.dense_words == 60 * 80 and s.lowrank_words == (40 + 80) * 5 + 5 + (40 + 30) * 5 + 25
    assert s.algorithmic_bytes == 8 * (s.dense_words + s.lowrank_words) + 8 * 180
    assert L.hm_builder_layout_stats(b, 2, 2, C.byref(s)) == 1      # part out of range
    h = C.c_void_p()
    assert L.hm_plan_finalize(b, None, 1, C.byref(h)) == 5          # HM_ERR_STATE: nothing staged
    assert L.hm_plan_finalize(b, None, 2, C.byref(h)) == 8          # one plan drives one GPU
    L.hm_builder_destroy(b)


def test_no_cpu_fallback(hm):
    """Without a CUDA device every compute entry point fails loudly (no silent fallback)."""
    if _has_gpu():
        pytest.skip("a GPU is present")
    L = hm.lib()
    b = C.c_void_p()
    assert L.hm_builder_create(C.byref(b), 10, 10, 0, 0) == 7       # HM_ERR_CUDA
    x = hm.chebyshevpoints(300)
    with pytest.raises(hm.HmError) as e:
        hm.KernelMatrix(hm.cauchykernel, x, x + 1e-3, 1.0, -1.0, 1.0, -1.0, device=0)
    assert e.value.status == 7
    with pytest.raises(hm.HmError) as e:
        hm.KernelMatrix(hm.cauchykernel, x, x + 1e-3, 1.0, -1.0, 1.0, -1.0, device=0, matrix_free=True)
    assert e.value.status == 7
    with pytest.raises(hm.HmError) as e:   # any-f constructor (host callback): no CPU path either
        hm.KernelMatrix(lambda p, q: np.exp(-np.abs(p - q)), x, x + 1e-3, 1.0, -1.0, 1.0, -1.0, device=0)
    assert e.value.status == 7
    with pytest.raises(hm.HmError):     # matrix_free belongs to the assembling constructor only
        hm.KernelMatrix(np.float64, 2, 2, matrix_free=True)
    H = hm.HierarchicalMatrix(np.float64, 1, 1)
    H[hm.Block(1), hm.Block(1)] = np.zeros((3, 3), order="F")
    with pytest.raises(hm.HmError):
        H * np.zeros(3)


def test_python_mirror_container_semantics(hm):
    """@hierarchical container behaviour (hierarchical.jl:31-172) of the host mirror."""
    rng = np.random.default_rng(0)
    H = hm.HierarchicalMatrix(np.float64, 2, 2)
    assert H.blocksize() == (2, 2) and H.size() == (0, 0)
    assert hasattr(H, "HierarchicalMatrixblocks") and hasattr(H, "LowRankMatrixblocks") and hasattr(H, "Matrixblocks")
    A = np.asfortranarray(rng.standard_normal((4, 6)))
    Lr = hm.LowRankMatrix(rng.standard_normal((4, 2)), rng.standard_normal(2), rng.standard_normal((3, 2)))
    H[hm.Block(1), hm.Block(1)] = A
    H[hm.Block(1), hm.Block(2)] = Lr
    G = hm.HierarchicalMatrix(np.float64, 1, 1)
    G[hm.Block(1), hm.Block(1)] = np.asfortranarray(rng.standard_normal((5, 3)))
    H[hm.Block(2), hm.Block(2)] = G
    assert H.assigned.tolist() == [[3, 2], [0, 1]]           # codes: 1 nested, 2 LowRank, 3 Matrix, 0 none
    assert H.size() == (9, 9)                                 # rows: last block column; cols: first block row
    assert H.blocksize(2, 1, 1) == 0 and H.blocksize(1, 2) == (4, 3) and hm.blocksize(H, 2, 2, 1) == 5
    assert H[1, 1] == A[0, 0] and H[5, 2] == 0.0              # unassigned block reads as zero
    assert abs(H[2, 8] - (Lr.U[1] * Lr.S) @ Lr.V[1]) < 1e-14
    assert H[9, 9] == G.Matrixblocks[0, 0][4, 2]
    # setindex! silently ignores values whose type matches no field (hierarchical.jl:155-169)
    H[hm.Block(2), hm.Block(1)] = hm.BarycentricMatrix2D(np.zeros((5, 2)), np.zeros((2, 2)), np.zeros((6, 2)))
    H[hm.Block(2), hm.Block(1)] = np.zeros((5, 6), dtype=np.float32)
    assert H.assigned[1, 0] == 0
    # leaves in walk order with the walk's offsets
    lv = H.leaves()
    assert [(k, r, c) for k, r, c, _ in lv] == [(3, 0, 0), (2, 0, 6), (3, 4, 6)]
    st = H.stats()
    assert st["n_dense"] == 2 and st["n_lowrank"] == 1 and st["nrows"] == 9
    K = hm.KernelMatrix(np.float64, 1, 2)
    assert hasattr(K, "KernelMatrixblocks") and hasattr(K, "BarycentricMatrix2Dblocks")
    K[hm.Block(1), hm.Block(1)] = hm.BarycentricMatrix2D(np.zeros((5, 2)), np.zeros((2, 2)), np.zeros((6, 2)))
    K[hm.Block(1), hm.Block(2)] = Lr                          # not a KernelMatrix leaf type
    assert K.assigned.tolist() == [[2, 0]]
    assert K.size() == (0, 6)  # rows come from the LAST block column, which is unassigned here
    with pytest.raises(hm.HmError):
        hm.hierarchical("Odd", dict)
    # mul_ argument checking happens before any device work
    with pytest.raises(TypeError):
        hm.mul_(np.zeros(5), K, np.zeros(6), 1, 1, 2, 2)      # KernelMatrix has no strided mul!
    with pytest.raises(IndexError):
        hm.mul_(np.zeros(8), H, np.zeros(9))                  # y too short
    with pytest.raises(ValueError):
        hm.mul_(np.zeros((3, 9)), H, np.zeros((3, 9)))        # C-order 2-D: not Julia's linear indexing


def test_python_mirror_scale_host_side(hm, O):
    """scale_ / rmul_ / lmul_ keep the host mirror's blocks consistent with the reference's
    scale! walks (no plan -> no device work)."""
    from helpers import oracle_tree_from_mirror, random_lowrank_tree
    rng = np.random.default_rng(9)
    H = random_lowrank_tree(hm, rng, 400)
    T = oracle_tree_from_mirror(O, H)
    bc, br = rng.standard_normal(405), rng.standard_normal(401)
    hm.scale_(H, bc, 3)
    hm.scale_(br, H, 2)
    T.scale_cols(bc, 2)
    T.scale_rows(br, 1)
    T2 = oracle_tree_from_mirror(O, H)
    v = rng.standard_normal(400)
    assert np.allclose(T2.matvec(v), T.matvec(v), rtol=1e-13, atol=1e-13)
    hm.rmul_(H, bc[:400])
    hm.lmul_(br[:400], H)
    with pytest.raises(IndexError):
        hm.scale_(H, bc[:100], 1)
    with pytest.raises(TypeError):
        hm.scale_(bc, br)


# ---------------------------------------------------------------- EvenBarycentricMatrix
def test_evenbary_host_factors_bit_identical_to_oracle(hm, O):
    """The host mirror builds w, W, F with the reference's operation order
    (BarycentricMatrix.jl:18-45) -- bit for bit what the oracle restatement builds."""
    assert np.array_equal(hm.chebyshevbarycentricweights(20), O.chebyshevbarycentricweights(20))
    assert np.array_equal(hm.chebyshevbarycentricweights(7), O.chebyshevbarycentricweights(7))
    assert np.array_equal(hm.chebyshevbarycentricweights(6, kind=2), O.chebyshevbarycentricweights(6, 2))
    f = lambda x, j: 1.0 / (x - j)  # noqa: E731
    for (a, b, c, d) in ((1, 100, 300, 420), (-40, 37, 90, 91)):
        B = hm.EvenBarycentricMatrix(np.float64, f, a, b, c, d)
        w, W, F = O.evenbary_factors(f, a, b, c, d)
        assert B.shape == (b - a + 1, d - c + 1) and B.size(2) == d - c + 1
        assert np.array_equal(B.w, w) and np.array_equal(B.W, W) and np.array_equal(B.F, F)
        for (i, j) in ((1, 1), (1, 2), (b - a + 1, d - c + 1)):
            assert B[i, j] == O.evenbary_getindex(W, F, i - 1, j - 1)
        L = hm.barycentricmatrix(np.float64, f, a, b, c, d)   # BarycentricMatrix.jl:61-89
        assert isinstance(L, hm.LowRankMatrix) and np.array_equal(L.U, W.T) and np.array_equal(L.V, F)
        assert np.array_equal(L.S, np.ones(20))
    with pytest.raises(TypeError):
        hm.EvenBarycentricMatrix(f, 1.0, 5, 1, 5)


def test_evenbary_layout_accounting(hm):
    """An EvenBarycentricMatrix leaf is counted at the reference's (m+n) r words although it is
    packed zero-interleaved at rank 2r; parity and shapes are validated at the ABI."""
    L = hm.lib()
    b = C.c_void_p()
    hm._lib.check(L.hm_builder_create(C.byref(b), 500, 400, 0, -1))
    hm._lib.check(L.hm_builder_add_evenbary(b, None, 20, None, 300, 200, 300, 20, 10, 50, 1))
    s = hm._lib.Stats()
    hm._lib.check(L.hm_builder_layout_stats(b, 0, 1, C.byref(s)))
    st = s.asdict()
    assert st["lowrank_words"] == (200 + 300) * 20 and st["n_lowrank"] == 1
    assert st["algorithmic_bytes"] == 8 * ((200 + 300) * 20 + 500 + 400)
    assert st["v_stream_bytes"] + st["u_stream_bytes"] >= 2 * 8 * (200 + 300) * 20
    assert L.hm_builder_add_evenbary(b, None, 20, None, 300, 200, 300, 20, 10, 50, 2) == 1     # HM_ERR_INVALID
    assert L.hm_builder_add_evenbary(b, None, 20, None, 300, 200, 300, 20, 400, 50, 0) == 4    # HM_ERR_RANGE
    assert L.hm_builder_add_evenbary(b, None, 20, None, 300, 200, 300, -1, 0, 0, 0) == 3     # HM_ERR_SHAPE
    L.hm_builder_destroy(b)

    T = hm.hierarchical("ParityMatrix", hm.EvenBarycentricMatrix, hm.Matrix)
    H = T(np.float64, 1, 2)
    f = lambda x, j: 1.0 / (x - j)  # noqa: E731
    H[hm.Block(1), hm.Block(1)] = hm.EvenBarycentricMatrix(f, 1, 50, 100, 160)
    H[hm.Block(1), hm.Block(2)] = np.ones((50, 9), order="F")
    assert H.size() == (50, 70) and H.assigned.tolist() == [[2, 3]] and H.has_parity_leaves()
    assert H[3, 5] == H[hm.Block(1), hm.Block(1)][3, 5] and H[1, 62] == 1.0
    assert H.stats()["lowrank_words"] == (50 + 61) * 20
    with pytest.raises(TypeError):
        hm.rmul_(H, np.ones(70))       # the reference has no scale! for this leaf type
    with pytest.raises(TypeError):
        hm.mul_(np.zeros(100), H, np.ones(140), 1, 1, 2, 2)   # nor a strided mul!


def _dense_of(H):
    A = np.zeros(H.size())
    for kind, r0, c0, B in H.leaves():
        A[r0:r0 + B.shape[0], c0:c0 + B.shape[1]] += B if isinstance(B, np.ndarray) else B.todense()
    return A


def test_lowrank_algebra_host(hm, O):
    """SURVEY 8f row f4, first step: LowRankMatrix +, -, *, svdtrunc, getrank, lrzeros
    (LowRankMatrix.jl:60-171) in the host mirror against the oracle restatement and dense arithmetic."""
    rng = np.random.default_rng(0)
    L1 = hm.LowRankMatrix(rng.standard_normal((40, 3)), np.array([3.0, 2.0, 1.0]), rng.standard_normal((25, 3)))
    L2 = hm.LowRankMatrix(rng.standard_normal((40, 2)), np.array([1.5, 0.5]), rng.standard_normal((25, 2)))
    for sign, G in ((1.0, L1 + L2), (-1.0, L1 - L2)):
        U, S, V = O.lowrank_combine(L1.U, L1.S, L1.V, L2.U, L2.S, L2.V, sign)
        assert G.rank() == len(S) == 5 and np.all(np.diff(G.S) <= 0)
        assert np.allclose(G.S, S, rtol=1e-14) and np.allclose(np.abs(G.U), np.abs(U), atol=1e-12)
        assert np.allclose(G.todense(), L1.todense() + sign * L2.todense(), atol=1e-13)
        assert np.allclose(G.U.T @ G.U, np.eye(5), atol=1e-13) and np.allclose(G.V.T @ G.V, np.eye(5), atol=1e-13)
    # a sum whose exact rank is lower than r1 + r2 is truncated by getrank (tol = r*eps(sigma_1))
    L3 = hm.LowRankMatrix(L1.U[:, :2], np.array([1.0, 1.0]), L1.V[:, :2])
    assert (L1 + L3).rank() == 3
    assert hm.getrank(np.array([1.0, 1e-3, 1e-17])) == 2 and hm.getrank(np.array([])) == 0
    assert O.getrank(np.array([1.0, 1e-3, 1e-17])) == 2
    A = rng.standard_normal((30, 4)) @ rng.standard_normal((4, 18))
    T = hm.svdtrunc(A)
    assert T.rank() == 4 and np.allclose(T.todense(), A, atol=1e-13)
    Z = hm.lrzeros(np.float64, 7, 9)
    assert Z.shape == (7, 9) and Z.rank() == 0 and np.all(Z.todense() == 0)
    W = hm.LowRankMatrix(rng.standard_normal((25, 2)), np.array([2.0, 1.0]), rng.standard_normal((11, 2)))
    P = L1 * W                                                   # LowRankMatrix.jl:113-118
    assert P.shape == (40, 11) and P.rank() == 2 and np.allclose(P.todense(), L1.todense() @ W.todense(), atol=1e-12)
    assert np.allclose((2.0 * L1).todense(), 2 * L1.todense()) and np.allclose((L1 / 4.0).todense(), L1.todense() / 4)
    sub = L1[3:20, 5:9]                                          # getindex(L, ir, jr), :60-62
    assert np.allclose(sub.todense(), L1.todense()[3:20, 5:9])
    assert L1[2, 3] == L1.U[1, 2] * L1.S[2] * L1.V[2, 2] + L1.U[1, 1] * L1.S[1] * L1.V[2, 1] + L1.U[1, 0] * L1.S[0] * L1.V[2, 0]


def test_hierarchical_plus_minus_lowrank_host(hm):
    """H + L, L + H, H - L, L - H (algebra.jl:394-524): same block structure, LowRankMatrix blocks
    recompressed, Matrix blocks dense; equal to the dense sum."""
    from helpers import random_lowrank_tree
    rng = np.random.default_rng(5)
    n = 700
    H = random_lowrank_tree(hm, rng, n, leaf=60, r=5)
    L = hm.LowRankMatrix(rng.standard_normal((n, 3)), np.array([2.0, 1.0, 0.5]), rng.standard_normal((n, 3)))
    Hd, Ld = _dense_of(H), L.todense()
    for G, ref in ((H + L, Hd + Ld), (L + H, Ld + Hd), (H - L, Hd - Ld), (L - H, Ld - Hd)):
        assert type(G) is type(H) and G.size() == H.size()
        assert np.array_equal(G.assigned, H.assigned)
        assert np.allclose(_dense_of(G), ref, atol=1e-12 * np.abs(ref).max())
        kinds = [k for k, _, _, _ in G.leaves()]
        assert kinds == [k for k, _, _, _ in H.leaves()]
        assert max(B.rank() for k, _, _, B in G.leaves() if k == 2) <= 8
    with pytest.raises(TypeError):
        hm.KernelMatrix(np.float64, 1, 1) + L


# ---------------------------------------------------------------- hierarchicalcholesky
def _spd_hodlr(hm, x, leaf, shift=0.5):
    """Upper-stored HODLR form of the SPD kernel matrix exp(-|x_i - x_j|) + shift*I on sorted points
    (off-diagonal blocks are numerically low rank): the input cholesky.jl:1-10 describes."""
    n = len(x)
    if n <= leaf:
        return np.asfortranarray(np.exp(-np.abs(x[:, None] - x[None, :])) + shift * np.eye(n))
    h = n // 2
    H = hm.HierarchicalMatrix(2, 2)
    H[hm.Block(1), hm.Block(1)] = _spd_hodlr(hm, x[:h], leaf, shift)
    H[hm.Block(1), hm.Block(2)] = hm.svdtrunc(np.exp(-np.abs(x[:h, None] - x[None, h:])))
    H[hm.Block(2), hm.Block(2)] = _spd_hodlr(hm, x[h:], leaf, shift)
    return H


def _upper_dense(R):
    return R if isinstance(R, np.ndarray) else R.todense()


def test_hierarchicalcholesky_host(hm, O):
    """hierarchicalcholesky / solvetransposed (cholesky.jl:12-228) in the host mirror: R'R = A, the factor
    keeps A's block structure with the (2,1) blocks unassigned, agrees with the unique dense upper factor
    (oracle restatement of the all-dense method), and solvetransposed(R, b) = R' \\ b for vectors,
    matrices and LowRankMatrix right-hand sides."""
    import scipy.linalg
    rng = np.random.default_rng(7)
    n = 600
    x = np.sort(rng.uniform(0.0, 40.0, n))
    A = _spd_hodlr(hm, x, leaf=80)
    Ad = np.exp(-np.abs(x[:, None] - x[None, :])) + 0.5 * np.eye(n)
    assert np.allclose(np.triu(A.todense()), np.triu(Ad), atol=1e-13)   # only the upper blocks are stored
    assert np.all(A.todense()[n // 2:, :n // 2] == 0)
    R = hm.hierarchicalcholesky(A)
    assert type(R) is hm.HierarchicalMatrix and R.size() == (n, n)
    assert R.assigned.tolist() == [[1, 2], [0, 1]]       # nested, LowRankMatrix, unassigned, nested
    Rd = R.todense()
    assert np.allclose(np.tril(Rd, -1), 0.0)
    assert np.max(np.abs(Rd.T @ Rd - Ad)) <= 1e-12 * np.abs(Ad).max() * n
    # the upper factor of an SPD matrix is unique: compare with the dense block restatement
    h = n // 2
    Rref = O.block_cholesky_dense(Ad[:h, :h], Ad[:h, h:], Ad[h:, h:])
    assert np.max(np.abs(Rd - Rref)) <= 1e-11
    # solvetransposed
    b = rng.standard_normal(n)
    xs = hm.solvetransposed(R, b)
    assert np.max(np.abs(xs - scipy.linalg.solve_triangular(Rref, b, trans="T"))) <= 1e-10 * np.abs(xs).max()
    B = np.asfortranarray(rng.standard_normal((n, 3)))
    XB = hm.solvetransposed(R, B)
    assert XB.shape == (n, 3) and np.max(np.abs(Rd.T @ XB - B)) <= 1e-10
    L = hm.LowRankMatrix(rng.standard_normal((n, 2)), np.array([2.0, 1.0]), rng.standard_normal((17, 2)))
    XL = hm.solvetransposed(R, L)
    assert isinstance(XL, hm.LowRankMatrix) and XL.shape == (n, 17) and XL.V is L.V
    assert np.max(np.abs(Rd.T @ XL.todense() - L.todense())) <= 1e-9
    # a full solve A z = b through the factor: z = R \ (R' \ b)
    z = scipy.linalg.solve_triangular(Rd, xs, lower=False)
    assert np.max(np.abs(Ad @ z - b)) <= 1e-9 * np.abs(b).max()
    # L' * L2 (LowRankMatrix.jl:120-126) against the oracle restatement and dense arithmetic
    L2 = hm.LowRankMatrix(rng.standard_normal((n, 3)), np.array([3.0, 1.0, 0.5]), rng.standard_normal((9, 3)))
    P = L.adjoint_mul(L2)
    U, S, V = O.lowrank_adjoint_times(L.U, L.S, L.V, L2.U, L2.S, L2.V)
    assert P.shape == (17, 9) and np.allclose(P.S, S, rtol=1e-13)
    assert np.allclose(P.todense(), L.todense().T @ L2.todense(), atol=1e-10)
    assert np.allclose(L.adjoint_mul(b), L.todense().T @ b, atol=1e-10)


@pytest.mark.parametrize("c11,c12,c22", [(1, 2, 1), (1, 2, 3), (1, 3, 1), (1, 3, 3), (3, 2, 1), (3, 2, 3), (3, 3, 1), (3, 3, 3)])
def test_hierarchicalcholesky_eight_cases(hm, c11, c12, c22):
    """The eight Val{M11}, Val{M12}, Val{M22} methods of cholesky.jl:22-94 and :163-228: nested or dense
    diagonal blocks, LowRankMatrix or dense (1,2) block."""
    rng = np.random.default_rng(c11 * 100 + c12 * 10 + c22)
    n = 240
    x = np.sort(rng.uniform(0.0, 20.0, n))
    Ad = np.exp(-np.abs(x[:, None] - x[None, :])) + 0.5 * np.eye(n)
    h = n // 2
    A = hm.HierarchicalMatrix(2, 2)
    A[hm.Block(1), hm.Block(1)] = _spd_hodlr(hm, x[:h], leaf=40 if c11 == 1 else n)
    A[hm.Block(2), hm.Block(2)] = _spd_hodlr(hm, x[h:], leaf=40 if c22 == 1 else n)
    off = np.asfortranarray(Ad[:h, h:])
    A[hm.Block(1), hm.Block(2)] = hm.svdtrunc(off) if c12 == 2 else off
    assert (A.assigned[0, 0], A.assigned[0, 1], A.assigned[1, 1]) == (c11, c12, c22)
    R = hm.hierarchicalcholesky(A)
    Rd = R.todense()
    assert np.allclose(np.tril(Rd, -1), 0.0)
    assert np.max(np.abs(Rd.T @ Rd - Ad)) <= 1e-11
    assert R.assigned[0, 1] == c12 and R.assigned[1, 0] == 0 and R.assigned[0, 0] == c11
    # a dense R12 makes A22 - R12'R12 dense (generic AbstractMatrix arithmetic), as in the reference
    assert R.assigned[1, 1] == (3 if c12 == 3 else c22)
    b = rng.standard_normal(n)
    assert np.max(np.abs(Rd.T @ hm.solvetransposed(R, b) - b)) <= 1e-10


def test_hierarchicalcholesky_rejects_what_the_reference_rejects(hm):
    rng = np.random.default_rng(1)
    D = np.asfortranarray(np.eye(6) * 2.0)
    A = hm.HierarchicalMatrix(2, 2)
    A[hm.Block(1), hm.Block(1)] = hm.svdtrunc(D)          # low-rank diagonal block: no method (cholesky.jl:5-9)
    A[hm.Block(1), hm.Block(2)] = D
    A[hm.Block(2), hm.Block(2)] = D
    with pytest.raises(TypeError):
        hm.hierarchicalcholesky(A)
    B = hm.HierarchicalMatrix(3, 3)                        # @assert blocksize(A) == (2, 2)
    with pytest.raises(AssertionError):
        hm.hierarchicalcholesky(B)
    with pytest.raises(np.linalg.LinAlgError):             # PosDefException
        hm.hierarchicalcholesky(np.asfortranarray(-np.eye(4)))
    with pytest.raises(TypeError):
        hm.hierarchicalcholesky(hm.svdtrunc(D))
