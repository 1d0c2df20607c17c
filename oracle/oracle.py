"""ctypes binding of the CPU oracle (oracle/hm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/hm_oracle.h.  PARITY UNPINNED (no Julia in
this image, no golden vectors in the reference).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhm_oracle.so")

CAUCHY, COULOMB, COULOMBPRIME, LOG, USER = 0, 1, 2, 3, 4
NONE, NODE, LOWRANK, DENSE, BARY2D, EVENBARY = 0, 1, 2, 3, 4, 5

_dp = C.POINTER(C.c_double)
_i64 = C.c_int64


class Leaf(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("depth", C.c_int32),
        ("row0", _i64),
        ("col0", _i64),
        ("m", _i64),
        ("n", _i64),
        ("r", _i64),
        ("A", _dp),
        ("S", _dp),
        ("V", _dp),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile) if the .so is missing/stale."""
    src = os.path.join(_HERE, "hm_oracle.c")
    stale = (not os.path.exists(_SO)) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)
    )
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp = C.c_void_p
    sig = {
        "hmo_blockrank_f64": (C.c_int, []),
        "hmo_blockrank_f32": (C.c_int, []),
        "hmo_blocksize_f64": (C.c_int, []),
        "hmo_sinpi": (C.c_double, [C.c_double]),
        "hmo_chebyshevpoints": (None, [_i64, C.c_int, _dp]),
        "hmo_chebyshevbarycentricweights": (None, [_i64, C.c_int, _dp]),
        "hmo_indsplit": (C.c_int, [_dp, _i64, _i64, _i64, C.c_double, C.c_double, C.POINTER(_i64)]),
        "hmo_kernel_eval": (C.c_double, [C.c_int, C.c_double, C.c_double]),
        "hmo_mul_dense": (None, [_dp, _dp, _i64, _i64, _i64, _dp, _i64, _i64, _i64, _i64]),
        "hmo_mul_dense_t": (None, [_dp, _dp, _i64, _i64, _i64, _dp, _i64, _i64, _i64, _i64]),
        "hmo_mul_lowrank": (None, [_dp, _dp, _i64, _dp, _dp, _i64, _i64, _i64, _i64, _dp, _i64, _i64, _i64, _i64]),
        "hmo_mul_bary2d": (None, [_dp, _dp, _i64, _dp, _i64, _dp, _i64, _i64, _i64, _i64, _dp, _i64, _i64]),
        "hmo_evenbary_weights": (None, [_i64, _i64, _dp, _dp]),
        "hmo_mul_evenbary": (None, [_dp, _dp, _i64, _dp, _i64, _i64, _i64, _i64, _dp, _i64, _i64]),
        "hmo_evenbary_getindex": (C.c_double, [_dp, _i64, _dp, _i64, _i64, _i64, _i64, _i64, _i64]),
        "hmo_node_set_evenbary": (C.c_int, [vp, C.c_int, C.c_int, _dp, _i64, _dp, _i64, _i64, _i64, _i64]),
        "hmo_node_create": (vp, [C.c_int, C.c_int]),
        "hmo_node_free": (None, [vp]),
        "hmo_node_set_node": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "hmo_node_set_dense": (C.c_int, [vp, C.c_int, C.c_int, _dp, _i64, _i64, _i64]),
        "hmo_node_set_lowrank": (C.c_int, [vp, C.c_int, C.c_int, _dp, _i64, _dp, _dp, _i64, _i64, _i64, _i64]),
        "hmo_node_set_bary2d": (C.c_int, [vp, C.c_int, C.c_int, _dp, _i64, _dp, _i64, _dp, _i64, _i64, _i64, _i64]),
        "hmo_node_assigned": (C.c_int, [vp, C.c_int, C.c_int]),
        "hmo_blocksize": (_i64, [vp, C.c_int, C.c_int, C.c_int]),
        "hmo_size": (_i64, [vp, C.c_int]),
        "hmo_getindex": (C.c_double, [vp, _i64, _i64]),
        "hmo_mul": (None, [_dp, vp, _dp, _i64, _i64, _i64, _i64]),
        "hmo_mul_omp": (None, [_dp, vp, _dp, _i64, _i64, C.c_int]),
        "hmo_mul_adjoint": (None, [_dp, vp, _dp, _i64, _i64]),
        "hmo_scale_cols": (None, [vp, _dp, _i64]),
        "hmo_scale_rows": (None, [_dp, vp, _i64]),
        "hmo_kernelmatrix": (vp, [C.c_int, _dp, _i64, _dp, _i64, C.c_double, C.c_double, C.c_double, C.c_double]),
        "hmo_bary2d_build": (None, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _i64, _i64, _dp, _i64, _i64, _dp, _dp, _dp]),
        "hmo_count_leaves": (_i64, [vp]),
        "hmo_list_leaves": (_i64, [vp, C.POINTER(Leaf), _i64]),
        "hmo_stored_words": (_i64, [vp]),
        "hmo_count_nodes": (_i64, [vp, C.POINTER(C.c_int)]),
        "hmo_dense_kernel_matvec_ld": (None, [C.c_int, _dp, _i64, _dp, _i64, _dp, _dp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _p(a: np.ndarray):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_dp)


def _f(a) -> np.ndarray:
    """Column-major float64 copy-if-needed."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


# ---------------------------------------------------------------- constants
def blockrank(dtype=np.float64) -> int:
    return lib().hmo_blockrank_f32() if np.dtype(dtype) == np.float32 else lib().hmo_blockrank_f64()


def blocksize() -> int:
    return lib().hmo_blocksize_f64()


def chebyshevpoints(n: int, kind: int = 1) -> np.ndarray:
    out = np.empty(n)
    lib().hmo_chebyshevpoints(n, kind, _p(out))
    return out


def chebyshevbarycentricweights(n: int, kind: int = 1) -> np.ndarray:
    out = np.empty(n)
    lib().hmo_chebyshevbarycentricweights(n, kind, _p(out))
    return out


def indsplit(x: np.ndarray, i0: int, i1: int, a: float, b: float):
    """0-based half-open; returns ((i0, mid), (mid, i1)) or raises IndexError."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    mid = _i64()
    if lib().hmo_indsplit(_p(x), len(x), i0, i1, a, b, C.byref(mid)):
        raise IndexError("indsplit: BoundsError in the reference")
    return (i0, mid.value), (mid.value, i1)


# ---------------------------------------------------------------- leaf applies
def mul_dense(y, A, x, i0=0, j0=0, incx=1, incy=1, transpose=False):
    A = _f(A)
    m, n = A.shape
    fn = lib().hmo_mul_dense_t if transpose else lib().hmo_mul_dense
    fn(_p(y), _p(A), m, n, max(m, 1), _p(x), i0, j0, incx, incy)
    return y


def mul_lowrank(y, U, S, V, x, i0=0, j0=0, incx=1, incy=1):
    U, V, S = _f(U), _f(V), np.ascontiguousarray(S, dtype=np.float64)
    m, r = U.shape
    n = V.shape[0]
    lib().hmo_mul_lowrank(_p(y), _p(U), max(m, 1), _p(S), _p(V), max(n, 1), m, n, r, _p(x), i0, j0, incx, incy)
    return y


def mul_bary2d(u, U, F, V, v, i0=0, j0=0):
    U, V, F = _f(U), _f(V), _f(F)
    m, r = U.shape
    n = V.shape[0]
    lib().hmo_mul_bary2d(_p(u), _p(U), max(m, 1), _p(F), max(r, 1), _p(V), max(n, 1), m, n, r, _p(v), i0, j0)
    return u


def bary2d_build(kernel, a, b, c, d, x, i0, i1, y, j0, j1):
    r = blockrank()
    m, n = max(i1 - i0, 0), max(j1 - j0, 0)
    U = np.zeros((m, r), order="F")
    F = np.zeros((r, r), order="F")
    V = np.zeros((n, r), order="F")
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    lib().hmo_bary2d_build(kernel, a, b, c, d, _p(x), i0, i1, _p(y), j0, j1, _p(U), _p(F), _p(V))
    return U, F, V


# ---------------------------------------------------------------- EvenBarycentricMatrix
def evenbary_factors(f, a: int, b: int, c: int, d: int):
    """EvenBarycentricMatrix(Float64, f, a, b, c, d) -- BarycentricMatrix.jl:18-45: returns
    (w, W, F) with W r x (b-a+1) and F (d-c+1) x r; f(x, j) is the user kernel of a real
    abscissa and an integer index (the reference passes the element type first)."""
    r = blockrank()
    m, n = b - a + 1, d - c + 1
    w = np.zeros(max(m, 0))
    W = np.zeros((r, max(m, 0)), order="F")
    lib().hmo_evenbary_weights(a, b, _p(w), _p(W))
    xk = chebyshevpoints(r, 1)
    F = np.zeros((max(n, 0), r), order="F")
    for k in range(r):
        node = (a + b) / 2 + (b - a) * xk[k] / 2
        for j in range(c, d + 1):
            F[j - c, k] = f(node, j)
    return w, W, F


def mul_evenbary(u, W, F, v, i0=0, j0=0):
    W, F = _f(W), _f(F)
    r, m = W.shape
    n = F.shape[0]
    lib().hmo_mul_evenbary(_p(u), _p(W), max(r, 1), _p(F), max(n, 1), m, n, r, _p(v), i0, j0)
    return u


def evenbary_getindex(W, F, i, j) -> float:
    W, F = _f(W), _f(F)
    r, m = W.shape
    n = F.shape[0]
    return lib().hmo_evenbary_getindex(_p(W), max(r, 1), _p(F), max(n, 1), m, n, r, i, j)


# ---------------------------------------------------------------- trees
class Tree:
    """Owning handle on an oracle block tree (an `@hierarchical` instance)."""

    def __init__(self, handle, owned=True):
        if not handle:
            raise RuntimeError("oracle: the reference would throw here (NULL tree)")
        self.h = handle
        self.owned = owned

    def __del__(self):
        if getattr(self, "owned", False) and self.h and _lib is not None:
            _lib.hmo_node_free(self.h)
            self.h = None

    # -- construction (setindex! with Block indices, 0-based m,n)
    @staticmethod
    def create(M: int, N: int) -> "Tree":
        return Tree(lib().hmo_node_create(M, N))

    def set_node(self, m, n, child: "Tree"):
        assert child.owned
        lib().hmo_node_set_node(self.h, m, n, child.h)
        child.owned = False  # ownership moves into the parent

    def set_dense(self, m, n, A):
        A = _f(A)
        lib().hmo_node_set_dense(self.h, m, n, _p(A), A.shape[0], A.shape[1], max(A.shape[0], 1))

    def set_lowrank(self, m, n, U, S, V):
        U, V, S = _f(U), _f(V), np.ascontiguousarray(S, dtype=np.float64)
        lib().hmo_node_set_lowrank(
            self.h, m, n, _p(U), max(U.shape[0], 1), _p(S), _p(V), max(V.shape[0], 1),
            U.shape[0], V.shape[0], U.shape[1])

    def set_bary2d(self, m, n, U, F, V):
        U, V, F = _f(U), _f(V), _f(F)
        r = U.shape[1]
        lib().hmo_node_set_bary2d(
            self.h, m, n, _p(U), max(U.shape[0], 1), _p(F), max(r, 1), _p(V), max(V.shape[0], 1),
            U.shape[0], V.shape[0], r)

    def set_evenbary(self, m, n, W, F):
        W, F = _f(W), _f(F)
        r = W.shape[0]
        rc = lib().hmo_node_set_evenbary(self.h, m, n, _p(W), max(r, 1), _p(F), max(F.shape[0], 1),
                                         W.shape[1], F.shape[0], r)
        assert rc == 0

    # -- queries
    @property
    def shape(self):
        return (lib().hmo_size(self.h, 1), lib().hmo_size(self.h, 2))

    def assigned(self, m, n) -> int:
        return lib().hmo_node_assigned(self.h, m, n)

    def blocksize(self, m, n, k) -> int:
        return lib().hmo_blocksize(self.h, m, n, k)

    def getindex(self, i, j) -> float:
        return lib().hmo_getindex(self.h, i, j)

    def stored_words(self) -> int:
        return lib().hmo_stored_words(self.h)

    def count_nodes(self):
        d = C.c_int()
        n = lib().hmo_count_nodes(self.h, C.byref(d))
        return n, d.value

    def leaves(self):
        n = lib().hmo_count_leaves(self.h)
        arr = (Leaf * max(n, 1))()
        lib().hmo_list_leaves(self.h, arr, n)
        return arr, n

    # -- mul!(y, H, x, istart, jstart, INCX, INCY); 0-based offsets; accumulates
    def mul(self, y, x, i0=0, j0=0, incx=1, incy=1):
        lib().hmo_mul(_p(y), self.h, _p(x), i0, j0, incx, incy)
        return y

    def mul_omp(self, y, x, nthreads, i0=0, j0=0):
        lib().hmo_mul_omp(_p(y), self.h, _p(x), i0, j0, nthreads)
        return y

    # rmul!(H, Diagonal(b)) / lmul!(Diagonal(b), H) -- HierarchicalMatrix.jl:15-16
    def scale_cols(self, b, j0=0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        lib().hmo_scale_cols(self.h, _p(b), j0)
        return self

    def scale_rows(self, b, i0=0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        lib().hmo_scale_rows(_p(b), self.h, i0)
        return self

    # H' * x
    def rmatvec(self, x):
        y = np.zeros(self.shape[1])
        lib().hmo_mul_adjoint(_p(y), self.h, _p(np.ascontiguousarray(x, dtype=np.float64)), 0, 0)
        return y

    # H * x
    def matvec(self, x):
        y = np.zeros(self.shape[0])
        return self.mul(y, np.ascontiguousarray(x, dtype=np.float64))


_USER_KERNEL_T = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)
_user_kernel_keep = None


def set_user_kernel(f):
    """Register a scalar Python function f(x, y) as kernel id USER (any f::Function, KernelMatrix.jl:47)."""
    global _user_kernel_keep
    _user_kernel_keep = _USER_KERNEL_T(lambda a, b: float(f(a, b)))
    lib().hmo_set_user_kernel.argtypes = [_USER_KERNEL_T]
    lib().hmo_set_user_kernel.restype = None
    lib().hmo_set_user_kernel(_user_kernel_keep)


def kernelmatrix(kernel, x, y, a, b, c, d) -> Tree:
    """KernelMatrix(f, x, y, a, b, c, d) -- KernelMatrix.jl:47."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return Tree(lib().hmo_kernelmatrix(kernel, _p(x), len(x), _p(y), len(y), a, b, c, d))


def dense_kernel_matvec_ld(kernel, x, y, b) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.empty(len(x))
    lib().hmo_dense_kernel_matvec_ld(kernel, _p(x), len(x), _p(y), len(y), _p(b), _p(out))
    return out


# ---------------------------------------------------------------- low-rank algebra
def getrank(sigma) -> int:
    """LowRankMatrix.jl:70-80: tol = r*eps(first(σ)), trailing values <= tol dropped."""
    r = len(sigma)
    if r == 0:
        return 0
    tol = r * np.spacing(sigma[0])
    while r >= 1 and not sigma[r - 1] > tol:
        r -= 1
    return r


def lowrank_combine(U1, S1, V1, U2, S2, V2, sign=1.0):
    """(+)(L1, L2) / (-)(L1, L2) -- LowRankMatrix.jl:95-111, step by step:
    QRU = qr!(hcat(U1, U2)); QRV = qr!(hcat(V1, V2));
    SVD = svd!(QRU.R * Diagonal(vcat(S1, +-S2)) * QRV.R'); r = getrank(SVD.S);
    U = (QRU.Q*SVD.U)[:, 1:r], S = SVD.S[1:r], V = (QRV.Q*SVD.V)[:, 1:r]."""
    QU, RU = np.linalg.qr(np.concatenate([U1, U2], axis=1))
    QV, RV = np.linalg.qr(np.concatenate([V1, V2], axis=1))
    core = RU @ np.diag(np.concatenate([S1, sign * np.asarray(S2)])) @ RV.T
    Us, sv, Vh = np.linalg.svd(core)
    r = getrank(sv)
    return (QU @ Us)[:, :r], sv[:r], (QV @ Vh.T)[:, :r]


def lowrank_adjoint_times(U1, S1, V1, U2, S2, V2):
    """(*)(L1', L2) -- LowRankMatrix.jl:120-126: SVD = svd!(S1 * U1'U2 * S2); r = getrank(SVD.S);
    LowRankMatrix((V1*SVD.U)[:, 1:r], SVD.S[1:r], (V2*SVD.V)[:, 1:r])."""
    Us, sv, Vh = np.linalg.svd(np.diag(S1) @ (U1.T @ U2) @ np.diag(S2))
    r = getrank(sv)
    return (V1 @ Us)[:, :r], sv[:r], (V2 @ Vh.T)[:, :r]


def block_cholesky_dense(A11, A12, A22):
    """cholesky.jl:18-94 with every block dense (the Val{3}, Val{3}, Val{3} method, :86-94):
    R11 = chol(A11); R12 = R11' \\ A12 column by column (:128-135, :150-152); R22 = chol(A22 - R12'R12).
    For a symmetric positive definite A the upper factor is unique, so the hierarchical factor of the
    same A -- whatever its block structure -- has to agree with this one up to its truncation error."""
    import scipy.linalg
    R11 = np.linalg.cholesky(np.triu(A11) + np.triu(A11, 1).T).T
    R12 = np.zeros_like(A12)
    for j in range(A12.shape[1]):
        R12[:, j] = scipy.linalg.solve_triangular(R11, A12[:, j], trans="T", lower=False)
    S = A22 - R12.T @ R12
    R22 = np.linalg.cholesky(np.triu(S) + np.triu(S, 1).T).T
    n1 = A11.shape[0]
    R = np.zeros((n1 + A22.shape[0],) * 2)
    R[:n1, :n1], R[:n1, n1:], R[n1:, n1:] = R11, R12, R22
    return R


# ---------------------------------------------------------------- synthetic inputs
def example_points(N: int, dist: str = "cheb"):
    """Point sets of the benchmark configs: "cheb" = examples/Kernel.jl:61-62,
    "unif" = the uniform interlaced set of SURVEY 8(d).  Returns x, y, (a,b,c,d)."""
    if dist == "cheb":
        return chebyshevpoints(N, 1), chebyshevpoints(N, 2), (1.0, -1.0, 1.0, -1.0)
    if dist == "unif":
        i = np.arange(1, N + 1, dtype=np.float64)
        return 1.0 - 2.0 * (i - 0.5) / N, 1.0 - 2.0 * (i - 0.25) / N, (1.0, -1.0, 1.0, -1.0)
    if dist == "quad":  # examples/Kernel.jl:88-92
        i = np.arange(N, 0, -1, dtype=np.float64)
        x = i * (i + 1.0)
        y = (i + 0.5) * (i + 1.5)
        return x, y, (x.max(), x.min(), y.max(), y.min())
    raise ValueError(dist)
