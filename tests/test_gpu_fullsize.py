"""Full-size checks at BASELINE.json's N = 2^20 (configs[1], [2]): parity against the oracle
and size-independent properties (linearity, additivity over row parts, panel == vector
path).  Tolerance 1e-12 relative ∞-norm (north_star)."""
import numpy as np
import pytest

from helpers import TOL, relinf

pytestmark = pytest.mark.gpu

N = 1 << 20


@pytest.fixture(scope="module")
def K(hm):
    x, y = hm.chebyshevpoints(N), hm.chebyshevpoints(N, 2)
    return hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0, device=0)


def test_fullsize_layout(K):
    st = K.plan().stats()
    assert (st["n_dense"], st["n_bary2d"]) == (62794, 125460)
    assert abs(st["algorithmic_bytes"] - 14078717200) <= 4096
    assert st["stored_bytes"] <= 1.01 * st["algorithmic_bytes"]         # padding of the packed streams
    assert st["partial_bytes"] <= 0.005 * st["algorithmic_bytes"]
    assert st["n_stage3_rounds"] == 1


def test_fullsize_parity_with_oracle(hm, O, K):
    x, y, (a, b, c, d) = O.example_points(N, "cheb")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    v = np.random.default_rng(0).standard_normal(N)
    ref = Kref.matvec(v)
    out = K * v
    assert relinf(out, ref) <= TOL
    # mul! accumulates into y
    y0 = np.random.default_rng(1).standard_normal(N)
    y1 = y0.copy()
    hm.mul_(y1, K, v)
    assert relinf(y1, y0 + ref) <= TOL
    # assembled factors of a few leaves, bit for bit
    arr, n = Kref.leaves()
    plan = K.plan()
    for i in (0, 1, n // 3, n // 2, n - 2):
        lf = arr[i]
        if lf.kind == O.DENSE:
            assert np.array_equal(plan.read_leaf(i, 3), np.ctypeslib.as_array(lf.A, shape=(lf.n, lf.m)).T)
        else:
            assert np.array_equal(plan.read_leaf(i, 0), np.ctypeslib.as_array(lf.A, shape=(lf.r, lf.m)).T)
            assert np.array_equal(plan.read_leaf(i, 2), np.ctypeslib.as_array(lf.V, shape=(lf.r, lf.n)).T)


def test_fullsize_linearity_and_determinism(K):
    rng = np.random.default_rng(2)
    u, v = rng.standard_normal(N), rng.standard_normal(N)
    Ku, Kv = K * u, K * v
    comb = K * (0.75 * u - 2.5 * v)
    assert relinf(comb, 0.75 * Ku - 2.5 * Kv) <= TOL
    assert np.array_equal(K * u, Ku)                                    # run-to-run identical
    e = np.zeros(N)
    e[N // 3] = 1.0                                                     # a column of K: 1/(x_i - y_j)
    col = K * e
    import hmb200_loader
    hm = hmb200_loader.load()
    x, y = hm.chebyshevpoints(N), hm.chebyshevpoints(N, 2)
    exact = 1.0 / (x - y[N // 3])
    assert np.max(np.abs(col - exact) / np.abs(exact)) < 1e-11          # entrywise interpolation accuracy


def test_fullsize_row_parts_add_up(hm, K):
    x, y = hm.chebyshevpoints(N), hm.chebyshevpoints(N, 2)
    v = np.random.default_rng(3).standard_normal(N)
    whole = K * v
    out = np.full(N, np.nan)
    rows = 0
    for p in range(4):
        Kp = hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0, device=0, part=p, nparts=4)
        st = Kp.plan().stats()
        Kp.plan().matvec(v, out, accumulate=False)
        rows += st["row_end"] - st["row_begin"]
        del Kp
    assert rows == N
    assert relinf(out, whole) <= TOL


def test_fullsize_panel_matches_vector_path(K):
    rng = np.random.default_rng(4)
    X = np.asfortranarray(rng.standard_normal((N, 16)))
    Y = K * X
    for c in (0, 9, 15):
        assert relinf(Y[:, c], K * np.ascontiguousarray(X[:, c])) <= TOL


def test_fullsize_adjoint(K):
    rng = np.random.default_rng(5)
    v, w = rng.standard_normal(N), rng.standard_normal(N)
    import hmb200_loader
    hm = hmb200_loader.load()
    Ktw = hm.adjoint(K) * w
    Kv = K * v
    assert abs(w @ Kv - Ktw @ v) <= 1e-11 * np.linalg.norm(w) * np.linalg.norm(Kv)
    e = np.zeros(N)
    e[N // 5] = 1.0                                                     # a row of K: 1/(x_i - y_j)
    x, y = hm.chebyshevpoints(N), hm.chebyshevpoints(N, 2)
    row = hm.adjoint(K) * e
    exact = 1.0 / (x[N // 5] - y)
    assert np.max(np.abs(row - exact) / np.abs(exact)) < 1e-11


def test_builder_path_at_scale_matches_device_assembly(hm, O):
    """3.5 GB of leaves pushed one by one through hm_builder_add_* (pinned staging window,
    several device arena chunks, device-side repacking) give the same packed operator as
    the on-device assembly: identical factors, identical layout, bit-identical product."""
    from helpers import plan_from_oracle_tree
    n = 1 << 18
    x, y, (a, b, c, d) = O.example_points(n, "cheb")
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    plan = plan_from_oracle_tree(hm, O, Kref)
    Kdev = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    v = np.random.default_rng(6).standard_normal(n)
    out_b = np.zeros(n)
    plan.matvec(v, out_b, accumulate=False)
    out_d = Kdev * v
    assert np.array_equal(out_b, out_d)
    assert relinf(out_b, Kref.matvec(v)) <= TOL
    sb, sd = plan.stats(), Kdev.plan().stats()
    for k in ("algorithmic_bytes", "stored_bytes", "n_stage1_items", "n_stage3_items", "partial_bytes"):
        assert sb[k] == sd[k]


def test_fullsize_matrix_free(hm, K):
    """The matrix-free plan (hm_assemble_kernel_free) is the same operator as the stored one at
    full size: products agree to rounding, a unit vector reproduces a column of the kernel, and
    nothing but the cores is resident."""
    x, y = hm.chebyshevpoints(N), hm.chebyshevpoints(N, 2)
    Kf = hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0, device=0, matrix_free=True)
    v = np.random.default_rng(9).standard_normal(N)
    # (the Chebyshev-series form of the matrix-free kernels interpolates at the same nodes but sums a
    # different, equally valid expansion: measured 4e-13 from the stored operator at this size, against
    # 2e-16 for the barycentric form -- both inside the 1e-12 parity bound)
    dev = relinf(Kf * v, K * v)
    print(f"matrix-free vs stored at N={N}: {dev:.3e}")
    assert dev <= TOL
    assert np.array_equal(Kf * v, Kf * v)
    j = 777_777
    e = np.zeros(N)
    e[j] = 1.0
    col = Kf * e
    rows = np.array([0, 1, 5000, j - 1, j, j + 1, N // 2, N - 1])
    assert relinf(col[rows], 1.0 / (x[rows] - y[j])) <= 1e-12
    st = Kf.plan().stats()
    assert st["stored_bytes"] == 8 * st["core_words"] < 0.04 * st["algorithmic_bytes"]
