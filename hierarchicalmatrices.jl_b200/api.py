"""Host-side mirror of the HierarchicalMatrices.jl API for the `mul!` hot path.

Julia is not available in the build image, so the host side that the reference
keeps in Julia is mirrored here in Python with the reference's names, argument
meaning (1-based `Block`, `istart`/`jstart`, `INCX`/`INCY`) and error behaviour.
The Julia shim that binds the same C ABI with `ccall` is in julia/.

    reference (file:line under /root/reference)              here
    -------------------------------------------------------  -------------------------
    BLOCKRANK, BLOCKSIZE      src/HierarchicalMatrices.jl:5-7   BLOCKRANK, BLOCKSIZE
    Block                     src/block.jl:8-10                 Block
    LowRankMatrix             src/LowRankMatrix.jl:28-38        LowRankMatrix
    BarycentricMatrix2D       src/BarycentricMatrix.jl:208-218  BarycentricMatrix2D
    EvenBarycentricMatrix     src/BarycentricMatrix.jl:5-59     EvenBarycentricMatrix
    barycentricmatrix         src/BarycentricMatrix.jl:61-89    barycentricmatrix
    @hierarchical             src/hierarchical.jl:4-232         hierarchical()
    HierarchicalMatrix        src/HierarchicalMatrix.jl:1       HierarchicalMatrix
    KernelMatrix              src/KernelMatrix.jl:1,47          KernelMatrix
    blocksize, size           src/hierarchical.jl:31-47,76-97   blocksize, size
    mul!, *                   src/KernelMatrix.jl:5-45,         mul_, H * x, H @ x
                              src/HierarchicalMatrix.jl:5-52

All arithmetic of `mul_` / `*` runs in the CUDA library (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import HmError, Stats

_dp = C.POINTER(C.c_double)

__all__ = [
    "BLOCKRANK", "BLOCKSIZE", "Block", "LowRankMatrix", "BarycentricMatrix2D", "Matrix",
    "hierarchical", "HierarchicalMatrix", "KernelMatrix", "blocksize", "size", "mul_",
    "cauchykernel", "coulombkernel", "coulombprimekernel", "logkernel", "Plan", "flatten",
    "chebyshevpoints", "HmError", "rmul_", "lmul_", "scale_", "adjoint", "Adjoint",
    "chebyshevbarycentricweights", "EvenBarycentricMatrix", "barycentricmatrix", "Transpose", "transpose",
    "dist_unique_id", "getrank", "svdtrunc", "lrzeros", "hierarchicalcholesky", "solvetransposed",
]

Matrix = np.ndarray  # the dense leaf type of the reference


# --------------------------------------------------------------------------- constants
def BLOCKRANK(T=np.float64) -> int:
    """2round(Int, half(T)*log(3+sqrt(T(8)), inv(eps(T)))) -- HierarchicalMatrices.jl:5."""
    T = np.dtype(T)
    if T.kind == "c":
        T = np.dtype(f"f{T.itemsize // 2}")
    if T == np.float64:
        return int(_lib.lib().hm_blockrank_f64())
    eps = T.type(np.finfo(T).eps)
    v = T.type(0.5) * (np.log(T.type(1) / eps) / np.log(T.type(3) + np.sqrt(T.type(8))))
    return 2 * int(np.rint(v))


def BLOCKSIZE(T=np.float64) -> int:
    """4BLOCKRANK(T) -- HierarchicalMatrices.jl:7."""
    return 4 * BLOCKRANK(T)


def _sinpi(q: np.ndarray) -> np.ndarray:
    """sinpi for 0 <= q <= 1: folded to [0, 1/4] and evaluated in long double."""
    q = np.where(q > 0.5, 1.0 - q, q)
    ql = q.astype(np.longdouble)
    pi = np.longdouble("3.14159265358979323846264338327950288")
    return np.where(q <= 0.25, np.sin(pi * ql), np.cos(pi * (np.longdouble(0.5) - ql))).astype(np.float64)


def chebyshevpoints(n: int, kind: int = 1) -> np.ndarray:
    """chebyshevpoints(Float64, n; kind) -- BarycentricMatrix.jl:92-111 (sinpi in long double)."""
    k = np.arange(1, n // 2 + 1, dtype=np.float64)
    q = (n - 2 * k + 1.0) / (2.0 * n) if kind == 1 else (n - 2 * k + 1.0) / (2.0 * (n - 1))
    v = _sinpi(q)
    x = np.zeros(n)
    x[: n // 2] = v
    x[n - (n // 2):] = -v[::-1]
    return x


def chebyshevbarycentricweights(n: int, kind: int = 1) -> np.ndarray:
    """chebyshevbarycentricweights(Float64, n; kind) -- BarycentricMatrix.jl:114-136."""
    lam = np.zeros(n)
    if kind == 1:
        h = n // 2
        if n < 1:
            raise IndexError("BoundsError: chebyshevbarycentricweights needs n >= 1")
        k = np.arange(1, h + 2, dtype=np.float64)
        lam[: h + 1] = _sinpi((2 * k - 1.0) / (2.0 * n))
        lam[n - h:] = lam[:h][::-1]
    else:
        lam[:] = 1.0
    lam[1::2] *= -1.0
    if kind != 1:
        lam[0] *= 0.5
        lam[n - 1] *= 0.5
    return lam


class Block:
    """`Block(K)`, 1-based block index -- block.jl:8-10."""

    __slots__ = ("K",)

    def __init__(self, K: int):
        self.K = int(K)

    def __int__(self):
        return self.K

    def __eq__(self, o):
        return isinstance(o, Block) and o.K == self.K

    def __hash__(self):
        return hash(("Block", self.K))

    def __repr__(self):
        return f"Block({self.K})"


# --------------------------------------------------------------------------- kernels
class _Kernel:
    """One of the four kernels of examples/Kernel.jl:34-37; callable on the host
    (scalars or broadcastable arrays) and known to the device assembler by id."""

    def __init__(self, name, kid, fn):
        self.__name__ = name
        self.id = kid
        self._fn = fn

    def __call__(self, x, y):
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        if x.ndim == 1 and y.ndim == 1:  # f(X::Vector, Y::Vector) = [f(x, y) for x in X, y in Y] -- Kernel.jl:39-41
            return self._fn(x[:, None], y[None, :])
        return self._fn(x, y)

    def __repr__(self):
        return self.__name__


cauchykernel = _Kernel("cauchykernel", 0, lambda x, y: 1.0 / (x - y))
coulombkernel = _Kernel("coulombkernel", 1, lambda x, y: 1.0 / ((x - y) * (x - y)))
coulombprimekernel = _Kernel("coulombprimekernel", 2, lambda x, y: 1.0 / ((x - y) * (x - y) * (x - y)))
logkernel = _Kernel("logkernel", 3, lambda x, y: np.log(np.abs(x - y)))


# --------------------------------------------------------------------------- leaf types
def _fmat(a, name):
    a = np.asarray(a)
    if a.ndim != 2:
        raise TypeError(f"{name} must be a matrix")
    return a


class LowRankMatrix:
    """A = U Σ V' (no conjugation in mul!, algebra.jl:118) -- LowRankMatrix.jl:28-38."""

    def __init__(self, U, S, V):
        self.U = _fmat(U, "U")
        self.V = _fmat(V, "V")
        S = np.asarray(S)
        self.S = np.diag(S).copy() if S.ndim == 2 else S  # accepts Diagonal-as-matrix
        if not (self.U.shape[1] == self.V.shape[1] == self.S.shape[0]):
            raise ValueError("LowRankMatrix: rank mismatch between U, Σ and V")
        if not (self.U.dtype == self.V.dtype == self.S.dtype):
            raise TypeError("LowRankMatrix: U, Σ, V must share one element type")

    Σ = property(lambda self: self.S)
    dtype = property(lambda self: self.U.dtype)
    shape = property(lambda self: (self.U.shape[0], self.V.shape[0]))

    def rank(self):
        return self.S.shape[0]

    # ---- low-rank algebra (first step): LowRankMatrix.jl:60-171.  Host side, as in
    # the reference (LAPACK QR / SVD through numpy); the results are ordinary leaves for the device path.
    def __getitem__(self, key):
        """`L[ir, jr]` with 0-based Python slices -- getindex(L, ir::UnitRange, jr::UnitRange), :60-62."""
        ir, jr = key
        if isinstance(ir, slice) and isinstance(jr, slice):
            return LowRankMatrix(self.U[ir, :], self.S, self.V[jr, :])
        ret = self.dtype.type(0)  # scalar getindex, 1-based, k = r..1 (:50-58)
        for k in range(self.rank() - 1, -1, -1):
            ret += self.U[ir - 1, k] * self.S[k] * self.V[jr - 1, k]
        return ret

    def todense(self):
        return (self.U * self.S) @ self.V.T

    def _combine(self, other, sign):
        # (+)/(-)(L1, L2), :95-111: QR of [U1 U2] and [V1 V2], SVD of Ru diag(S1, +-S2) Rv', truncation
        if not isinstance(other, LowRankMatrix):
            return NotImplemented
        if self.shape != other.shape:
            raise ValueError("DimensionMismatch")
        Qu, Ru = np.linalg.qr(np.hstack([self.U, other.U]))
        Qv, Rv = np.linalg.qr(np.hstack([self.V, other.V]))
        Us, sv, Vt = np.linalg.svd((Ru * np.concatenate([self.S, sign * other.S])) @ Rv.T)
        r = getrank(sv)
        return LowRankMatrix((Qu @ Us)[:, :r], sv[:r], (Qv @ Vt.T)[:, :r])

    def __add__(self, other):
        if isinstance(other, _HierarchicalBase):
            return other.__radd__(self)
        if isinstance(other, np.ndarray):  # generic AbstractMatrix + : elementwise, a dense Matrix
            return self.todense() + other
        return self._combine(other, 1.0)

    def __sub__(self, other):
        if isinstance(other, _HierarchicalBase):
            return other.__rsub__(self)
        if isinstance(other, np.ndarray):
            return self.todense() - other
        return self._combine(other, -1.0)

    def __radd__(self, other):
        return other + self.todense() if isinstance(other, np.ndarray) else NotImplemented

    def __rsub__(self, other):
        return other - self.todense() if isinstance(other, np.ndarray) else NotImplemented

    def __mul__(self, other):
        if isinstance(other, LowRankMatrix):  # :113-118
            Us, sv, Vt = np.linalg.svd((self.S[:, None] * (self.V.T @ other.U)) * other.S[None, :])
            r = getrank(sv)
            return LowRankMatrix((self.U @ Us)[:, :r], sv[:r], (other.V @ Vt.T)[:, :r])
        if np.isscalar(other):  # :160-161
            return LowRankMatrix(self.U, self.S * other, self.V)
        x = np.asarray(other)  # L*x, algebra.jl:88-95
        return (self.U * self.S) @ (self.V.T @ x)

    def __rmul__(self, other):
        return LowRankMatrix(self.U, other * self.S, self.V) if np.isscalar(other) else NotImplemented

    def adjoint_mul(self, other):
        """`L1' * L2` (LowRankMatrix.jl:120-126; the step R12'R12 of cholesky.jl:24) and `L' * x`
        (algebra.jl:97-104, real element type)."""
        if isinstance(other, LowRankMatrix):
            Us, sv, Vt = np.linalg.svd((self.S[:, None] * (self.U.T @ other.U)) * other.S[None, :])
            r = getrank(sv)
            return LowRankMatrix((self.V @ Us)[:, :r], sv[:r], (other.V @ Vt.T)[:, :r])
        return (self.V * self.S) @ (self.U.T @ np.asarray(other))

    def __truediv__(self, other):  # :162
        return LowRankMatrix(self.U, self.S / other, self.V)


def getrank(sigma) -> int:
    """`getrank(σ)` -- LowRankMatrix.jl:70-80: trailing singular values <= r*eps(σ[1]) are dropped."""
    sigma = np.asarray(sigma)
    r = len(sigma)
    if r == 0:
        return 0
    tol = r * np.spacing(sigma[0])
    while r >= 1:
        if sigma[r - 1] > tol:
            return r
        r -= 1
    return r


def svdtrunc(A) -> "LowRankMatrix":
    """`svdtrunc(A)` -- LowRankMatrix.jl:82-86 (also `convert(LowRankMatrix, A)`, :68)."""
    U, sv, Vt = np.linalg.svd(np.asarray(A), full_matrices=False)
    r = getrank(sv)
    return LowRankMatrix(U[:, :r], sv[:r], Vt.T[:, :r])


def lrzeros(T, m: int, n: int) -> "LowRankMatrix":
    """`lrzeros(T, m, n)` -- LowRankMatrix.jl:88-93."""
    return LowRankMatrix(np.zeros((m, 0), dtype=T), np.zeros(0, dtype=T), np.zeros((n, 0), dtype=T))


class BarycentricMatrix2D:
    """U F V' with U m×r, F r×r, V n×r -- BarycentricMatrix.jl:208-218 (factors as
    produced by update!, :248-297; B.B.F is `F` here)."""

    def __init__(self, U, F, V):
        self.U = _fmat(U, "U")
        self.F = _fmat(F, "F")
        self.V = _fmat(V, "V")
        r = self.F.shape[0]
        if not (self.F.shape[1] == r == self.U.shape[1] == self.V.shape[1]):
            raise ValueError("BarycentricMatrix2D: rank mismatch between U, F and V")
        if not (self.U.dtype == self.V.dtype == self.F.dtype):
            raise TypeError("BarycentricMatrix2D: U, F, V must share one element type")

    dtype = property(lambda self: self.U.dtype)
    shape = property(lambda self: (self.U.shape[0], self.V.shape[0]))


# --------------------------------------------------------------------------- plan handle
def dist_unique_id() -> bytes:
    """128-byte id for `Plan.dist_init` (ncclGetUniqueId); create on one rank, send to all."""
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().hm_dist_get_id(buf))
    return buf.raw


class Plan:
    """Owning handle on an `hm_plan` (immutable packed operator on one GPU)."""

    def __init__(self, handle, device):
        self._h = C.c_void_p(handle)
        self.device = device
        self._stats = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        lib = getattr(_lib, "_lib", None) if _lib is not None else None  # module globals vanish at interpreter exit
        if h and lib is not None:
            lib.hm_plan_destroy(h)

    close = __del__

    @property
    def handle(self):
        return self._h

    def stats(self) -> dict:
        if self._stats is None:
            s = Stats()
            _lib.check(_lib.lib().hm_plan_stats(self._h, C.byref(s)))
            self._stats = s.asdict()
        return self._stats

    @property
    def shape(self):
        s = self.stats()
        return (s["nrows"], s["ncols"])

    @property
    def form(self) -> int:
        """0 stored streams, 1 matrix-free (barycentric), 2 matrix-free (Chebyshev), 3 matrix-free nested-basis."""
        f = C.c_int32()
        _lib.check(_lib.lib().hm_plan_form(self._h, C.byref(f)))
        return f.value

    @property
    def launches_per_matvec(self) -> int:
        return int(_lib.lib().hm_plan_launches_per_matvec(self._h))

    @staticmethod
    def _check_vec(a, nm, n, inc, off):
        """The raw pointer handed to the C ABI must cover off + (n-1)*inc (cudaMemcpy has no bounds)."""
        if not isinstance(a, np.ndarray) or a.dtype != np.float64:
            raise TypeError(f"{nm} must be a Float64 array (MethodError in the reference: all eltypes equal)")
        if a.ndim != 1 or (a.size > 1 and a.strides[0] != 8):
            raise ValueError(f"{nm} must be a contiguous vector (use incx/incy for strides)")
        if inc < 1 or off < 0:
            raise ValueError(f"{nm}: stride must be >= 1 and offset >= 0")
        if n > 0 and off + (n - 1) * inc >= a.size:
            raise ValueError(f"{nm} has {a.size} elements; offset {off} + ({n}-1)*{inc} is out of range "
                             f"(BoundsError in the reference)")

    # y[i*incy] (+)= (H x)[i]; host arrays, strides in elements
    def matvec(self, x: np.ndarray, y: np.ndarray, incx=1, incy=1, accumulate=True, xoff=0, yoff=0):
        nr, nc = self.shape
        self._check_vec(x, "x", nc, incx, xoff)
        self._check_vec(y, "y", nr, incy, yoff)
        px = C.cast(x.ctypes.data + xoff * 8, _dp)
        py = C.cast(y.ctypes.data + yoff * 8, _dp)
        _lib.check(_lib.lib().hm_matvec(self._h, px, incx, py, incy, 1 if accumulate else 0))
        return y

    # y[j*incy] (+)= (H' x)[j]
    def rmatvec(self, x: np.ndarray, y: np.ndarray, incx=1, incy=1, accumulate=True, xoff=0, yoff=0):
        nr, nc = self.shape
        self._check_vec(x, "x", nr, incx, xoff)
        self._check_vec(y, "y", nc, incy, yoff)
        px = C.cast(x.ctypes.data + xoff * 8, _dp)
        py = C.cast(y.ctypes.data + yoff * 8, _dp)
        _lib.check(_lib.lib().hm_matvec_adjoint(self._h, px, incx, py, incy, 1 if accumulate else 0))
        return y

    def matmat(self, X: np.ndarray, Y: np.ndarray, accumulate=True):
        nr, nc = self.shape
        for a, nm, rows in ((X, "X", nc), (Y, "Y", nr)):
            if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.ndim != 2:
                raise TypeError(f"{nm} must be a Float64 matrix")
            if not a.flags.f_contiguous:
                raise ValueError(f"{nm} must be column-major (Fortran order)")
            if a.shape[0] != rows:
                raise ValueError(f"{nm} has {a.shape[0]} rows, the operator needs {rows}")
        if X.shape[1] != Y.shape[1]:
            raise ValueError("X and Y must have the same number of columns")
        nrhs = X.shape[1]
        _lib.check(_lib.lib().hm_matmat(
            self._h, X.ctypes.data_as(_dp), max(X.shape[0], 1), Y.ctypes.data_as(_dp), max(Y.shape[0], 1),
            nrhs, 1 if accumulate else 0))
        return Y

    # device pointers (ints), e.g. torch tensors' data_ptr(); enqueued on `stream`
    def matvec_device(self, dx: int, dy: int, accumulate=False, stream: int = 0):
        _lib.check(_lib.lib().hm_matvec_device(self._h, dx, dy, 1 if accumulate else 0, stream))

    # multi-GPU: owned rows of y are stored into every rank's y buffer (peer-mapped pointers)
    def matvec_device_allgather(self, dx: int, ypeers, self_rank: int, accumulate=False, stream: int = 0):
        arr = (C.c_uint64 * len(ypeers))(*[int(a) for a in ypeers])
        _lib.check(_lib.lib().hm_matvec_device_allgather(self._h, dx, arr, len(ypeers), self_rank,
                                                         1 if accumulate else 0, stream))

    # ---- multi-GPU: this plan is block-row part `rank` of `nranks`, one process per GPU ----
    def dist_init(self, uid: bytes, nranks: int, rank: int):
        """Collective.  `uid` = `dist_unique_id()` of one rank, distributed by the caller."""
        if len(uid) != 128:
            raise ValueError("uid must be the 128 bytes of dist_unique_id()")
        buf = C.create_string_buffer(bytes(uid), 128)
        _lib.check(_lib.lib().hm_dist_init(self._h, buf, nranks, rank))
        self.nranks, self.rank = nranks, rank

    def dist_init_torch(self, group=None):
        """`dist_init` with the id broadcast over an initialised `torch.distributed` group."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.dist_init(box[0], world, rank)

    def dist_buffers(self):
        """((x slot 0, x slot 1), (y slot 0, y slot 1)) device addresses; y slots are full length."""
        x2, y2 = (C.c_void_p * 2)(), (C.c_void_p * 2)()
        _lib.check(_lib.lib().hm_dist_buffers(self._h, x2, y2))
        return (int(x2[0]), int(x2[1])), (int(y2[0]), int(y2[1]))

    def dist_bcast_x(self, dx_root: int, root: int = 0, slot: int = 0, stream: int = 0):
        _lib.check(_lib.lib().hm_dist_bcast_x(self._h, dx_root or None, root, slot, stream))

    def dist_push_x(self, dx_root: int, root: int = 0, slot: int = 0, stream: int = 0):
        """Copy-engine replication of x from the root into every rank's x slot; complete on a rank once
        the next barrier (the next `dist_matvec_device`) has completed there."""
        _lib.check(_lib.lib().hm_dist_push_x(self._h, dx_root or None, root, slot, stream))

    def dist_matvec_device(self, dx: int, yslot: int = 0, accumulate=False, stream: int = 0):
        _lib.check(_lib.lib().hm_dist_matvec_device(self._h, dx, yslot, 1 if accumulate else 0, stream))

    def dist_barrier(self, stream: int = 0):
        _lib.check(_lib.lib().hm_dist_barrier(self._h, stream))

    def dist_check(self):
        _lib.check(_lib.lib().hm_dist_check(self._h))

    def dist_matvec(self, x, y, root: int = 0, accumulate=False, incx=1, incy=1):
        """Host arrays: x is read on `root` only (pass None elsewhere), y (or None) gets the whole
        result on every rank that passes one."""
        nr, nc = self.shape
        if x is not None:
            self._check_vec(x, "x", nc, incx, 0)
        if y is not None:
            self._check_vec(y, "y", nr, incy, 0)
        px = x.ctypes.data_as(_dp) if x is not None else None
        py = y.ctypes.data_as(_dp) if y is not None else None
        _lib.check(_lib.lib().hm_dist_matvec(self._h, px, incx, py, incy, root, 1 if accumulate else 0))
        return y

    def rmatvec_device(self, dx: int, dy: int, accumulate=False, stream: int = 0):
        _lib.check(_lib.lib().hm_matvec_adjoint_device(self._h, dx, dy, 1 if accumulate else 0, stream))

    def matmat_device(self, dX: int, ldx: int, dY: int, ldy: int, nrhs: int, accumulate=False, stream: int = 0):
        _lib.check(_lib.lib().hm_matmat_device(self._h, dX, ldx, dY, ldy, nrhs, 1 if accumulate else 0, stream))

    # H <- H*Diagonal(b) (side 0) or Diagonal(b)*H (side 1), on the device, no re-planning
    def scale(self, b: np.ndarray, side: int, off: int = 0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        pb = C.cast(b.ctypes.data + off * 8, _dp)
        _lib.check(_lib.lib().hm_plan_scale(self._h, pb, 1, side))

    def timing_begin(self, max_calls: int):
        _lib.check(_lib.lib().hm_plan_timing_begin(self._h, max_calls))

    def timing_end(self):
        """(ms of stage 1, 2, 3 summed over the timed matvecs, number of matvecs)."""
        ms = (C.c_double * 3)()
        n = C.c_int64()
        _lib.check(_lib.lib().hm_plan_timing_end(self._h, ms, C.byref(n)))
        return [ms[0], ms[1], ms[2]], n.value

    # test hooks
    def num_leaves(self) -> int:
        n = C.c_int64()
        _lib.check(_lib.lib().hm_plan_num_leaves(self._h, C.byref(n)))
        return n.value

    def leaf_info(self, i: int) -> dict:
        kind = C.c_int32()
        v = [C.c_int64() for _ in range(5)]
        _lib.check(_lib.lib().hm_plan_leaf_info(self._h, i, C.byref(kind), *[C.byref(a) for a in v]))
        return dict(kind=kind.value, row0=v[0].value, col0=v[1].value, m=v[2].value, n=v[3].value, r=v[4].value)

    def read_leaf(self, i: int, which: int) -> np.ndarray:
        info = self.leaf_info(i)
        m, n, r = info["m"], info["n"], info["r"]
        dense = info["kind"] == 3
        shape = {0: (m, r), 1: (r, r) if info["kind"] == 4 else (r, 1), 2: (n, r), 3: (m, n)}[which]
        if dense != (which == 3):
            raise ValueError("leaf kind has no such factor")
        out = np.zeros(shape, order="F")
        _lib.check(_lib.lib().hm_plan_read_leaf(self._h, i, which, out.ctypes.data_as(_dp), out.size))
        return out


class EvenBarycentricMatrix:
    """`EvenBarycentricMatrix(T, f, a, b, c, d)` -- BarycentricMatrix.jl:5-45: the rank-BLOCKRANK
    barycentric interpolant in i of f(i, j) on the integer grid i = a..b, j = c..d, whose `mul!`
    (algebra.jl:168-239) keeps only the entries with even absolute i+j.  `f(x, j)` takes the real
    abscissa and the integer column (the reference also passes the element type first)."""

    def __init__(self, *args):
        if len(args) == 6:
            T, f, a, b, c, d = args
            if np.dtype(T) != np.float64:
                raise HmError(8, "only Float64 operators are supported")
        elif len(args) == 5:
            f, a, b, c, d = args
        else:
            raise TypeError("EvenBarycentricMatrix(T, f, a, b, c, d)")
        for v in (a, b, c, d):
            if not isinstance(v, (int, np.integer)):
                raise TypeError("MethodError: a, b, c, d must be Int")
        self.a, self.b, self.c, self.d = int(a), int(b), int(c), int(d)
        n = BLOCKRANK()
        self.x = chebyshevpoints(n)
        self.λ = chebyshevbarycentricweights(n)
        self.w, self.W = _bary1d_weights(self.a, self.b, self.x, self.λ)
        self.β = np.zeros(n)
        self.F = _bary1d_samples(f, self.a, self.b, self.c, self.d, self.x)
        self._plans = {}

    dtype = property(lambda self: self.W.dtype)
    shape = property(lambda self: (self.b - self.a + 1, self.d - self.c + 1))

    def size(self, k=None):
        return self.shape if k is None else self.shape[k - 1]

    # getindex -- BarycentricMatrix.jl:48-59 (1-based); its parity rule is the matrix's own
    def __getitem__(self, key):
        i, j = key
        m, n = self.shape
        if not (1 <= i <= m and 1 <= j <= n):
            raise IndexError("BoundsError")
        ret = 0.0
        if (m + n + i + j) % 2 == 0:
            for k in range(self.x.size):
                ret += self.F[j - 1, k] * self.W[k, i - 1]
        return ret

    def plan(self, parity: int = 0, device=None) -> "Plan":
        """Single-leaf plan serving mul! calls whose (istart-1)+(jstart-1) has this parity."""
        key = (parity & 1, device)
        if key not in self._plans:
            self._plans[key] = _single_leaf_plan(self, parity & 1, _current_device() if device is None else device)
        return self._plans[key]

    def invalidate(self):
        self._plans = {}

    def __mul__(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        if v.ndim != 1 or v.size != self.shape[1]:
            raise ValueError("DimensionMismatch")
        return mul_(np.zeros(self.shape[0]), self, v)

    __matmul__ = __mul__


def _bary1d_weights(a: int, b: int, x: np.ndarray, lam: np.ndarray):
    """w[i] = Σ_k λ_k inv(2i-a-b-(b-a)x_k) in k order; W[k,i] = λ_k inv((2i-a-b-(b-a)x_k) w[i])
    -- BarycentricMatrix.jl:23-36, same operation order."""
    m = max(b - a + 1, 0)
    two_i = (2 * np.arange(a, b + 1, dtype=np.int64) - a - b).astype(np.float64)
    den = two_i[None, :] - (float(b - a) * x)[:, None]  # n x m
    w = np.zeros(m)
    for k in range(x.size):
        w += lam[k] * (1.0 / den[k])
    W = np.asfortranarray(lam[:, None] * (1.0 / (den * w[None, :])))
    return w, W


def _bary1d_samples(f, a: int, b: int, c: int, d: int, x: np.ndarray) -> np.ndarray:
    """F[j-c+1, k] = f((a+b)/2 + (b-a)x_k/2, j) -- BarycentricMatrix.jl:38-43."""
    F = np.zeros((max(d - c + 1, 0), x.size), order="F")
    for k in range(x.size):
        node = (a + b) / 2 + (b - a) * x[k] / 2
        for j in range(c, d + 1):
            F[j - c, k] = f(node, j)
    return F


def barycentricmatrix(*args) -> LowRankMatrix:
    """`barycentricmatrix(T, f, a, b, c, d)` -- BarycentricMatrix.jl:61-89: the unmasked
    interpolant as LowRankMatrix(W', I, F)."""
    if len(args) == 6:
        args = args[1:]
    f, a, b, c, d = args
    x = chebyshevpoints(BLOCKRANK())
    _, W = _bary1d_weights(int(a), int(b), x, chebyshevbarycentricweights(BLOCKRANK()))
    return LowRankMatrix(np.asfortranarray(W.T), np.ones(x.size), _bary1d_samples(f, int(a), int(b), int(c), int(d), x))


def _current_device() -> int:
    """The CUDA device of this process: LOCAL_RANK under torchrun, else 0."""
    import os
    return int(os.environ.get("HMB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


# --------------------------------------------------------------------------- @hierarchical
def _leaf_kind(A):
    if isinstance(A, LowRankMatrix):
        return 2
    if isinstance(A, BarycentricMatrix2D):
        return 4
    if isinstance(A, EvenBarycentricMatrix):
        return 5
    if isinstance(A, np.ndarray):
        return 3
    return None


class _HierarchicalBase:
    """Common behaviour of the types `hierarchical()` generates (hierarchical.jl:18-232)."""

    _name = ""
    _types: tuple = ()

    # $HierarchicalType(::Type{T}, M, N) / (M, N) -- hierarchical.jl:54-69
    def __init__(self, *args):
        if len(args) == 2:
            T, (M, N) = np.float64, args
        elif len(args) == 3:
            T, M, N = args
        else:
            raise TypeError(f"{self._name}(T, M, N) or {self._name}(M, N)")
        self.T = np.dtype(T)
        self.M, self.N = int(M), int(N)
        self._fields = [f"{self._name}blocks"] + [f"{_type_name(t)}blocks" for t in self._types]
        for f in self._fields:
            setattr(self, f, np.empty((self.M, self.N), dtype=object))  # "#undef" slots
        self.assigned = np.zeros((self.M, self.N), dtype=np.int64)
        self._plan = None
        self._has_parity = None

    dtype = property(lambda self: self.T)

    # setindex!(H, A, Block(m), Block(n)) -- hierarchical.jl:149-172: the field is
    # picked by the exact type of A; no match is silently ignored, as in the reference.
    def __setitem__(self, key, A):
        B1, B2 = key
        if not (isinstance(B1, Block) and isinstance(B2, Block)):
            raise TypeError("only H[Block(m), Block(n)] = A is defined")
        m, n = B1.K - 1, B2.K - 1
        if not (0 <= m < self.M and 0 <= n < self.N):
            raise IndexError("BoundsError: block index out of range")
        code = self._code_for(A)
        if code is None:
            return
        getattr(self, self._fields[code - 1])[m, n] = A
        self.assigned[m, n] = code
        self._plan = None
        self._has_parity = None

    def _code_for(self, A):
        if type(A) is type(self):
            return 1 if A.T == self.T else None
        for l, t in enumerate(self._types):
            if t is Matrix:
                if isinstance(A, np.ndarray) and A.ndim == 2 and A.dtype == self.T:
                    return l + 2
            elif type(A) is t and A.dtype == self.T:
                return l + 2
        return None

    def _block(self, m, n):
        code = self.assigned[m, n]
        return None if code == 0 else getattr(self, self._fields[code - 1])[m, n]

    # blocksize(H, m, n, k) -- hierarchical.jl:78-97 (1-based m, n; k = 1 rows, 2 columns)
    def blocksize(self, m=None, n=None, k=None):
        if m is None:
            return (self.M, self.N)
        if k is None:
            return (self.blocksize(m, n, 1), self.blocksize(m, n, 2))
        A = self._block(m - 1, n - 1)
        if A is None:
            return 0
        return A.size(k) if isinstance(A, _HierarchicalBase) else A.shape[k - 1]

    # size(H) -- hierarchical.jl:33-47: rows down the LAST block column, columns
    # along the FIRST block row
    def size(self, k=None):
        if self.M == 0 or self.N == 0:
            p = q = 0
        else:
            p = sum(self.blocksize(m, self.N, 1) for m in range(1, self.M + 1))
            q = sum(self.blocksize(1, n, 2) for n in range(1, self.N + 1))
        return (p, q) if k is None else (p, q)[k - 1]

    shape = property(lambda self: self.size())

    # getindex(H, i, j) -- hierarchical.jl:120-147 (1-based)
    def __getitem__(self, key):
        i, j = key
        if isinstance(i, Block):
            return self._block(i.K - 1, j.K - 1)
        m = 1
        while m <= self.M:
            r = self.blocksize(m, self.N, 1)
            if i > r:
                i -= r
                m += 1
            else:
                break
        n = 1
        while n <= self.N:
            s = self.blocksize(1, n, 2)
            if j > s:
                j -= s
                n += 1
            else:
                break
        if m > self.M or n > self.N:
            raise IndexError("BoundsError")
        A = self._block(m - 1, n - 1)
        if A is None:
            return self.T.type(0)
        if isinstance(A, _HierarchicalBase):
            return A[i, j]
        if isinstance(A, np.ndarray):
            return A[i - 1, j - 1]
        if isinstance(A, EvenBarycentricMatrix):
            return A[i, j]
        if isinstance(A, LowRankMatrix):  # LowRankMatrix.jl:50-58, k = r..1
            ret = self.T.type(0)
            for k in range(A.rank() - 1, -1, -1):
                ret += A.U[i - 1, k] * A.S[k] * A.V[j - 1, k]
            return ret
        ret = self.T.type(0)  # BarycentricMatrix.jl:222-234
        for k in range(A.F.shape[0]):
            ret += A.U[i - 1, k] * np.dot(A.F[k, :], A.V[j - 1, :])
        return ret

    # ---- the planner's front end: walk `assigned` exactly as mul! does ----
    def leaves(self, i0=0, j0=0, out=None):
        """(kind, row0, col0, block) of every leaf in walk order with the offsets
        KernelMatrix.jl:24-41 / HierarchicalMatrix.jl:30-48 pass to it (0-based)."""
        out = [] if out is None else out
        p = 0
        for m in range(self.M):
            q = 0
            for n in range(self.N):
                A = self._block(m, n)
                if isinstance(A, _HierarchicalBase):
                    A.leaves(i0 + p, j0 + q, out)
                elif A is not None:
                    out.append((_leaf_kind(A), i0 + p, j0 + q, A))
                q += self.blocksize(1, n + 1, 2)
            p += self.blocksize(m + 1, self.N, 1)
        return out

    def has_parity_leaves(self) -> bool:
        """True when some leaf is an EvenBarycentricMatrix, whose active entries depend on the
        parity of the offsets mul! is called with (algebra.jl:172)."""
        if self._has_parity is None:
            self._has_parity = any(kind == 5 for kind, _, _, _ in self.leaves())
        return self._has_parity

    def plan(self, device=None, part=0, nparts=1, parity=0) -> Plan:
        """Flatten the tree and pack it on the device.  The plan is a snapshot cached on
        this object: `H[Block(m), Block(n)] = A` on it drops the cache; after mutating a
        nested block or a leaf's arrays in place call `invalidate()`.  `parity` is that of
        (istart-1)+(jstart-1) of the mul! calls to serve; it only matters for operators with
        EvenBarycentricMatrix leaves."""
        key = (device, part, nparts, parity & 1)
        if self._plan is not None and self._plan[0] == key:
            return self._plan[1]
        if self.T != np.float64:
            raise HmError(8, "only Float64 operators are supported")
        P = flatten(self, _current_device() if device is None else device, part, nparts, parity & 1)
        self._plan = (key, P)
        return P

    def invalidate(self):
        self._plan = None
        self._has_parity = None

    def stats(self, part=0, nparts=1) -> dict:
        """Planner only (no GPU needed): sizes of the packed layout."""
        return _structure_stats(self, part, nparts)

    # ---- *  (HierarchicalMatrix.jl:5-12, KernelMatrix.jl:5-12) ----
    def __mul__(self, x):
        x = np.asarray(x)
        TS = np.result_type(self.T, x.dtype)
        if TS != np.float64:
            raise HmError(8, "only Float64 operators are supported")
        if x.ndim == 1:
            y = np.zeros(self.size(1), dtype=TS)
            return mul_(y, self, np.ascontiguousarray(x, dtype=TS))
        if x.ndim == 2:
            # The reference has no usable matrix product here (HierarchicalMatrix fills
            # column 1 only, KernelMatrix falls back to O(N^2) getindex); this is the
            # column-wise product those methods are meant to compute.
            X = np.asfortranarray(x, dtype=TS)
            Y = np.zeros((self.size(1), X.shape[1]), dtype=TS, order="F")
            if X.shape[0] != self.size(2):
                raise ValueError("DimensionMismatch")
            return self.plan().matmat(X, Y, accumulate=True)
        raise TypeError("x must be a vector or a matrix")

    __matmul__ = __mul__

    # ---- H + L, L + H, H - L, L - H  (algebra.jl:394-524): the generated walk adds the matching
    # window L[pr, qr] to every assigned block -- nested blocks recursively, LowRankMatrix blocks
    # by recompression (LowRankMatrix.jl:95-111), Matrix blocks elementwise.  Host side; the result
    # is an ordinary HierarchicalMatrix for the device path.
    def _add_lowrank(self, L, hsign, lsign, h_first):
        if not isinstance(L, LowRankMatrix):
            return NotImplemented
        if LowRankMatrix not in self._types:
            raise TypeError("MethodError: +(::%s, ::LowRankMatrix) is defined for HierarchicalMatrix" % self._name)
        G = type(self)(np.result_type(self.T, L.dtype), self.M, self.N)
        p = 0
        for m in range(self.M):
            q = 0
            for n in range(self.N):
                A = self._block(m, n)
                if A is not None:
                    rows, cols = self.blocksize(m + 1, n + 1, 1), self.blocksize(m + 1, n + 1, 2)
                    Lw = L[p:p + rows, q:q + cols]
                    if isinstance(A, _HierarchicalBase):
                        B = A._add_lowrank(Lw, hsign, lsign, h_first)
                    elif isinstance(A, LowRankMatrix):
                        if h_first:   # H_mn (+-) L
                            B = A._combine(Lw, lsign)
                        else:         # L (+-) H_mn
                            B = Lw._combine(A, hsign)
                    else:
                        B = np.asfortranarray(hsign * A + lsign * Lw.todense())
                    G[Block(m + 1), Block(n + 1)] = B
                q += self.blocksize(1, n + 1, 2)
            p += self.blocksize(m + 1, self.N, 1)
        return G

    def todense(self) -> np.ndarray:
        """`Matrix(H)`: every assigned block written out (what the generic AbstractMatrix fallbacks of
        the reference compute through getindex, hierarchical.jl:120-147)."""
        D = np.zeros(self.size(), dtype=self.T, order="F")
        for kind, i0, j0, A in self.leaves():
            blk = A if isinstance(A, np.ndarray) else A.todense()
            D[i0:i0 + blk.shape[0], j0:j0 + blk.shape[1]] += blk
        return D

    def __add__(self, L):
        if isinstance(L, np.ndarray):  # generic AbstractMatrix arithmetic: a dense Matrix
            return self.todense() + L
        return self._add_lowrank(L, 1.0, 1.0, True)

    def __radd__(self, L):
        if isinstance(L, np.ndarray):
            return L + self.todense()
        return self._add_lowrank(L, 1.0, 1.0, False)

    def __sub__(self, L):
        if isinstance(L, np.ndarray):
            return self.todense() - L
        return self._add_lowrank(L, 1.0, -1.0, True)

    def __rsub__(self, L):
        if isinstance(L, np.ndarray):
            return L - self.todense()
        return self._add_lowrank(L, -1.0, 1.0, False)


def _type_name(t):
    return "Matrix" if t is Matrix else t.__name__


def hierarchical(name: str, *types):
    """`@hierarchical Name T1 T2 ...` -- hierarchical.jl:4: a block-matrix type whose
    blocks are `Name` itself (code 1) or one of the listed leaf types (codes 2, 3, ...)."""
    for t in types:
        if t is not Matrix and t not in (LowRankMatrix, BarycentricMatrix2D, EvenBarycentricMatrix):
            raise HmError(8, f"leaf type {t!r} is not on the accelerated path")
    return type(name, (_HierarchicalBase,), {"_name": name, "_types": tuple(types)})


HierarchicalMatrix = hierarchical("HierarchicalMatrix", LowRankMatrix, Matrix)
_KernelMatrixBlocks = hierarchical("KernelMatrix", BarycentricMatrix2D, Matrix)


# ---- hierarchicalcholesky / solvetransposed (/root/reference/src/cholesky.jl:12-228).
# The reference spells the recursion out as eight Val-dispatched methods per function, one per
# combination of the `assigned` codes of the (1,1), (1,2), (2,2) blocks; here the codes select the
# blocks and one body serves all eight.  Host side: dense leaves go through LAPACK exactly where the
# reference calls it (cholesky(Symmetric(A)).U, UpperTriangular(R)'\b), the low-rank algebra is the
# mirror above.  The factor is an ordinary HierarchicalMatrix: R*x and R'*x run on the device.
def _chol_blocks(A, what):
    if not isinstance(A, HierarchicalMatrix):
        raise TypeError(f"MethodError: {what}(::{type(A).__name__})")
    if A.blocksize() != (2, 2):
        raise AssertionError("blocksize(A) == (2, 2)")           # cholesky.jl:13, :158
    c11, c12, c22 = (int(A.assigned[0, 0]), int(A.assigned[0, 1]), int(A.assigned[1, 1]))
    if c11 not in (1, 3) or c22 not in (1, 3) or c12 not in (2, 3):
        # no method for Val{c11}, Val{c12}, Val{c22}: low-rank diagonal blocks and nested or
        # unassigned off-diagonal blocks are outside the eight cases (cholesky.jl:5-9)
        raise TypeError(f"MethodError: {what}(::HierarchicalMatrix, Val({c11}), Val({c12}), Val({c22}))")
    return A._block(0, 0), A._block(0, 1), A._block(1, 1)


def hierarchicalcholesky(A):
    """Upper Cholesky factor R (A = R'R) of a symmetric positive definite HierarchicalMatrix, of which
    only the upper blocks are read -- cholesky.jl:12-94.  `hierarchicalcholesky(A::Matrix)` is the
    dense factor (:18-20)."""
    if isinstance(A, np.ndarray):
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("DimensionMismatch: matrix is not square")
        # Symmetric(A) reads the upper triangle; LAPACK potrf
        Au = np.triu(A)
        return np.asfortranarray(np.linalg.cholesky(Au + np.triu(A, 1).T).T)
    A11, A12, A22 = _chol_blocks(A, "hierarchicalcholesky")
    R = HierarchicalMatrix(A.T, 2, 2)
    R11 = hierarchicalcholesky(A11)
    R[Block(1), Block(1)] = R11
    R12 = solvetransposed(R11, A12)
    R[Block(1), Block(2)] = R12
    if isinstance(R12, LowRankMatrix):
        LL = R12.adjoint_mul(R12)                 # R12'R12, LowRankMatrix.jl:120-126
        # H - L (algebra.jl:460-491) or Matrix - L (generic elementwise, a dense Matrix)
        S = A22 - LL.todense() if isinstance(A22, np.ndarray) else A22 - LL
    else:
        S = A22 - R12.T @ R12                     # generic AbstractMatrix arithmetic: dense
    if isinstance(S, np.ndarray):
        S = np.asfortranarray(S)
    R[Block(2), Block(2)] = hierarchicalcholesky(S)
    return R


def solvetransposed(R, B):
    """`R' \\ B` for the upper factor R: B a vector (cholesky.jl:154-228), a Matrix (column by column,
    :109-116, :128-135) or a LowRankMatrix (its U factor column by column, :99-107, :118-126)."""
    if isinstance(B, LowRankMatrix):
        Uh = np.zeros_like(B.U)
        for j in range(B.U.shape[1]):
            Uh[:, j] = solvetransposed(R, np.ascontiguousarray(B.U[:, j]))
        return LowRankMatrix(Uh, B.S, B.V)
    B = np.asarray(B)
    if B.ndim == 2:
        Mh = np.zeros_like(B, order="F")
        for j in range(B.shape[1]):
            Mh[:, j] = solvetransposed(R, np.ascontiguousarray(B[:, j]))
        return Mh
    if B.ndim != 1:
        raise TypeError("MethodError: solvetransposed")
    if isinstance(R, np.ndarray):                 # UpperTriangular(R)' \ b, :150-152
        import scipy.linalg
        if R.shape[0] != B.shape[0]:
            raise ValueError("DimensionMismatch")
        return scipy.linalg.solve_triangular(R, B, trans="T", lower=False, check_finite=False)
    R11, R12, R22 = _chol_blocks(R, "solvetransposed")
    n = R.size(2)
    s = R11.shape[1] if isinstance(R11, np.ndarray) else R11.size(2)
    b1, b2 = B[:s], B[s:n]
    x1 = solvetransposed(R11, b1)
    r12x = R12.adjoint_mul(x1) if isinstance(R12, LowRankMatrix) else R12.T @ x1
    x2 = solvetransposed(R22, b2 - r12x)
    return np.concatenate([x1, x2])


class KernelMatrix(_KernelMatrixBlocks):
    """`@hierarchical KernelMatrix BarycentricMatrix2D Matrix` (KernelMatrix.jl:1).

    KernelMatrix(T, M, N)            empty container, blocks set with H[Block(m), Block(n)] = A
    KernelMatrix(f, x, y, a, b, c, d) the assembling constructor (KernelMatrix.jl:47-116):
        the tree of index ranges is built on the host, U, V, F and the dense leaves
        are evaluated on the GPU straight into the packed streams when `f` is one of the
        four kernels of examples/Kernel.jl; for any other function f(x, y) (called with
        arrays) the cores and dense leaves are evaluated by f on the host, U and V on the GPU.
    """

    _name = "KernelMatrix"

    def __init__(self, *args, device=None, part=0, nparts=1, matrix_free=False):
        """`KernelMatrix(f, x, y, a, b, c, d)` assembles on the GPU.  `matrix_free=True` keeps no
        U, V or dense tiles: every `mul!` evaluates the entries from the point sets while applying
        them (FP64-bound instead of HBM-bound; the operator needs no memory beyond its r x r
        cores, so sizes whose packed form exceeds the GPU still fit)."""
        self.matrix_free = bool(matrix_free)
        if len(args) == 7:
            f, x, y, a, b, c, d = args
            if not callable(f):
                raise TypeError("MethodError: f must be a function f(x, y)")
            super().__init__(np.float64, 0, 0)
            x = np.ascontiguousarray(x, dtype=np.float64)
            y = np.ascontiguousarray(y, dtype=np.float64)
            dev = _current_device() if device is None else device
            h = C.c_void_p()
            if isinstance(f, _Kernel):
                build = _lib.lib().hm_assemble_kernel_free if matrix_free else _lib.lib().hm_assemble_kernel
                _lib.check(build(x.ctypes.data_as(_dp), len(x), y.ctypes.data_as(_dp), len(y), a, b, c, d, f.id, dev,
                                 part, nparts, C.byref(h)))
            else:
                # any f::Function (KernelMatrix.jl:47): the cores and the dense leaves are evaluated by f on
                # the host in large batches (f is called with two equally long arrays), U and V on the device
                if matrix_free:
                    raise HmError(8, "matrix_free needs one of the kernels the device can evaluate")
                err = []

                def batch(px, py, n, pout, _user):
                    try:
                        xa = np.ctypeslib.as_array(px, shape=(n,))
                        ya = np.ctypeslib.as_array(py, shape=(n,))
                        np.ctypeslib.as_array(pout, shape=(n,))[:] = np.asarray(f(xa, ya), dtype=np.float64)
                    except BaseException as exc:  # must not propagate through the C frames
                        err.append(exc)
                        np.ctypeslib.as_array(pout, shape=(n,))[:] = np.nan

                cb = _lib.KERNEL_FN(batch)
                _lib.check(_lib.lib().hm_assemble_kernel_fn(x.ctypes.data_as(_dp), len(x), y.ctypes.data_as(_dp), len(y),
                                                            a, b, c, d, cb, None, dev, part, nparts, C.byref(h)))
                if err:
                    Plan(h.value, dev).close()
                    raise err[0]
            self._assembled = Plan(h.value, dev)
            self.kernel = f
        else:
            if matrix_free:
                raise HmError(8, "matrix_free applies to KernelMatrix(f, x, y, a, b, c, d) only")
            super().__init__(*args)
            self._assembled = None

    def _code_for(self, A):
        if isinstance(A, KernelMatrix):
            return 1 if A.T == self.T and A._assembled is None else None
        return super()._code_for(A)

    def size(self, k=None):
        if self._assembled is not None:
            s = self._assembled.shape
            return s if k is None else s[k - 1]
        return super().size(k)

    shape = property(lambda self: self.size())

    def plan(self, device=None, part=0, nparts=1, parity=0) -> Plan:
        if self._assembled is not None:
            return self._assembled
        return super().plan(device, part, nparts, parity)

    def stats(self, part=0, nparts=1):
        if self._assembled is not None:
            return self._assembled.stats()
        return super().stats(part, nparts)

    @staticmethod
    def layout_stats(x, y, a, b, c, d, part=0, nparts=1) -> dict:
        """Planner only (no GPU): leaf counts and byte sizes of KernelMatrix(f,x,y,a,b,c,d)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        s = Stats()
        _lib.check(_lib.lib().hm_assemble_kernel_stats(
            x.ctypes.data_as(_dp), len(x), y.ctypes.data_as(_dp), len(y), a, b, c, d, part, nparts, C.byref(s)))
        return s.asdict()


# --------------------------------------------------------------------------- planner front end
def _push_leaves(H, b, parity=0):
    L = _lib.lib()
    keep = []  # keep converted arrays alive until the builder has copied them
    for kind, row0, col0, A in H.leaves():
        if kind == 5:
            W = np.asfortranarray(A.W, dtype=np.float64)
            F = np.asfortranarray(A.F, dtype=np.float64)
            keep += [W, F]
            r, m = W.shape
            n = F.shape[0]
            _lib.check(L.hm_builder_add_evenbary(b, W.ctypes.data_as(_dp), max(r, 1), F.ctypes.data_as(_dp),
                                                 max(n, 1), m, n, r, row0, col0, parity))
            continue
        if kind == 3:
            M = np.asfortranarray(A, dtype=np.float64)
            keep.append(M)
            _lib.check(L.hm_builder_add_dense(b, M.ctypes.data_as(_dp), M.shape[0], M.shape[1],
                                              max(M.shape[0], 1), row0, col0))
            continue
        U = np.asfortranarray(A.U, dtype=np.float64)
        V = np.asfortranarray(A.V, dtype=np.float64)
        m, r = U.shape
        n = V.shape[0]
        if kind == 2:
            S = np.ascontiguousarray(A.S, dtype=np.float64)
            keep += [U, V, S]
            _lib.check(L.hm_builder_add_lowrank(b, U.ctypes.data_as(_dp), max(m, 1), S.ctypes.data_as(_dp),
                                                V.ctypes.data_as(_dp), max(n, 1), m, n, r, row0, col0))
        else:
            F = np.asfortranarray(A.F, dtype=np.float64)
            keep += [U, V, F]
            _lib.check(L.hm_builder_add_bary2d(b, U.ctypes.data_as(_dp), max(m, 1), F.ctypes.data_as(_dp),
                                               max(r, 1), V.ctypes.data_as(_dp), max(n, 1), m, n, r, row0, col0))
    return keep


class _SingleLeaf:
    """A leaf standing alone as an operator (the reference's leaf-level mul! methods)."""

    def __init__(self, A):
        self.A = A

    def size(self):
        return tuple(self.A.shape)

    def leaves(self):
        return [(_leaf_kind(self.A), 0, 0, self.A)]


def _single_leaf_plan(A, parity: int, device: int) -> Plan:
    return flatten(_SingleLeaf(A), device, 0, 1, parity)


def flatten(H, device: int, part=0, nparts=1, parity=0) -> Plan:
    """Walk the block tree and pack it into device arrays (the planner)."""
    L = _lib.lib()
    nrows, ncols = H.size()
    b = C.c_void_p()
    _lib.check(L.hm_builder_create(C.byref(b), nrows, ncols, 0, device))
    try:
        _push_leaves(H, b, parity)
        h = C.c_void_p()
        if nparts == 1:
            dev = (C.c_int32 * 1)(device)
            _lib.check(L.hm_plan_finalize(b, dev, 1, C.byref(h)))
        else:
            _lib.check(L.hm_plan_finalize_part(b, part, nparts, C.byref(h)))
    finally:
        L.hm_builder_destroy(b)
    return Plan(h.value, device)


def _structure_stats(H, part=0, nparts=1) -> dict:
    L = _lib.lib()
    nrows, ncols = H.size()
    b = C.c_void_p()
    _lib.check(L.hm_builder_create(C.byref(b), nrows, ncols, 0, -1))
    try:
        _push_leaves(H, b)
        s = Stats()
        _lib.check(L.hm_builder_layout_stats(b, part, nparts, C.byref(s)))
    finally:
        L.hm_builder_destroy(b)
    return s.asdict()


# --------------------------------------------------------------------------- free functions
def blocksize(H, *a):
    return H.blocksize(*a)


def size(H, k=None):
    return H.size(k)


def _scale_host(H, b, off, side):
    """The reference's scale! walks on the host mirror's own blocks (kept consistent with the
    device plan): HierarchicalMatrix.jl:54-108, leaf rules algebra.jl:280-315."""
    p = 0
    outer, inner = (range(H.N), range(H.M)) if side == 0 else (range(H.M), range(H.N))
    for o in outer:
        for i in inner:
            m, n = (i, o) if side == 0 else (o, i)
            A = H._block(m, n)
            if A is None:
                continue
            if isinstance(A, _HierarchicalBase):
                _scale_host(A, b, off + p, side)
            elif isinstance(A, np.ndarray):
                if side == 0:
                    A *= b[off + p: off + p + A.shape[1]][None, :]
                else:
                    A *= b[off + p: off + p + A.shape[0]][:, None]
            elif side == 0:
                A.V *= b[off + p: off + p + A.V.shape[0]][:, None]
            else:
                A.U *= b[off + p: off + p + A.U.shape[0]][:, None]
        p += H.blocksize(1, o + 1, 2) if side == 0 else H.blocksize(o + 1, H.N, 1)


def scale_(a, b, start: int = 1):
    """`scale!(H, b, jstart)`: H <- H*Diagonal(b[jstart:...]) or `scale!(b, H, istart)`:
    H <- Diagonal(b[istart:...])*H (HierarchicalMatrix.jl:54-108), 1-based start.  The packed
    operator on the GPU is updated in place by streaming kernels (no re-planning)."""
    if isinstance(a, _HierarchicalBase):
        H, vec, side = a, np.asarray(b, dtype=np.float64), 0
    elif isinstance(b, _HierarchicalBase):
        H, vec, side = b, np.asarray(a, dtype=np.float64), 1
    else:
        raise TypeError("MethodError: scale!(H, b, jstart) or scale!(b, H, istart)")
    n = H.size(2 if side == 0 else 1)
    if start < 1 or start - 1 + n > vec.size:
        raise IndexError("BoundsError: b is too short")
    if getattr(H, "_assembled", None) is None and H.has_parity_leaves():
        raise TypeError("MethodError: the reference defines no scale! for EvenBarycentricMatrix")
    assembled = getattr(H, "_assembled", None)
    if assembled is not None:
        assembled.scale(vec, side, start - 1)
        return H
    _scale_host(H, vec, start - 1, side)
    if H._plan is not None:
        H._plan[1].scale(vec, side, start - 1)
    return H


def rmul_(H, b):
    """`rmul!(H, Diagonal(b))` -- HierarchicalMatrix.jl:15."""
    return scale_(H, b, 1)


def lmul_(b, H):
    """`lmul!(Diagonal(b), H)` -- HierarchicalMatrix.jl:16."""
    return scale_(b, H, 1)


class Adjoint:
    """`adjoint(H)` / `H'` of a hierarchical matrix (real case: the transpose).  The reference
    wraps only its leaves this way (algebra.jl:50-82, 133-159); here the whole operator
    applies through the same packed streams, reduced over the fast index."""

    def __init__(self, parent):
        self.parent = parent

    def size(self, k=None):
        p, q = self.parent.size()
        return (q, p) if k is None else (q, p)[k - 1]

    shape = property(lambda self: self.size())

    def __mul__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 1:
            raise TypeError("adjoint(H) * x is defined for vectors")
        if x.size != self.parent.size(1):
            raise ValueError("DimensionMismatch")
        y = np.zeros(self.parent.size(2))
        return self.parent.plan().rmatvec(x, y, accumulate=False)

    __matmul__ = __mul__


class Transpose:
    """`transpose(A)` / `A'` of a leaf (Matrix or LowRankMatrix): the wrappers the reference's
    leaf-level mul! methods take (algebra.jl:50-82, 133-159; test/runtests.jl:27-33)."""

    def __init__(self, parent):
        self.parent = parent

    shape = property(lambda self: tuple(self.parent.shape[::-1]))


def adjoint(H):
    """`adjoint(H)`; all operators here are real, so this is the transpose."""
    if isinstance(H, _HierarchicalBase):
        return Adjoint(H)
    if _leaf_kind(H) in (2, 3):
        return Transpose(H)
    raise TypeError("MethodError: adjoint of a hierarchical matrix, a Matrix or a LowRankMatrix")


transpose = adjoint


def _linear(a: np.ndarray, name: str) -> np.ndarray:
    """Julia linear indexing = column-major order; must be a view so y is updated in place."""
    if a.ndim == 1:
        if a.strides[0] != a.itemsize and a.size > 1:
            raise ValueError(f"{name} must be contiguous")
        return a
    if not a.flags.f_contiguous:
        raise ValueError(f"{name} must be column-major (Fortran order), as Julia arrays are")
    return a.reshape(-1, order="F")


def _offset_parity(H, istart, jstart, INCX, INCY) -> int:
    """Parity of (istart-1)+(jstart-1) when it matters (EvenBarycentricMatrix leaves), else 0."""
    if getattr(H, "_assembled", None) is not None or not H.has_parity_leaves():
        return 0
    if INCX != 1 or INCY != 1:
        raise TypeError("MethodError: EvenBarycentricMatrix has no strided mul! (algebra.jl:168)")
    return (istart + jstart) & 1


def mul_(y, H, x, istart: int = 1, jstart: int = 1, INCX: int = 1, INCY: int = 1):
    """`mul!(y, H, x, istart, jstart, INCX, INCY)`: y[istart+(i-1)INCY] += Σ_j H[i,j] x[jstart+(j-1)INCX]
    with 1-based linear indices (HierarchicalMatrix.jl:14-52, KernelMatrix.jl:14-45); returns y."""
    if isinstance(H, Adjoint):  # mul!(y, H', x, ...): y += H' x with the same offset/stride rules
        Hp = H.parent
        if not (y.dtype == x.dtype == Hp.T == np.float64):
            raise TypeError("MethodError: y, H and x must all be Float64")
        if INCX < 1 or INCY < 1 or istart < 1 or jstart < 1:
            raise IndexError("BoundsError: offsets and strides are 1-based positive integers")
        yl, xl = _linear(y, "y"), _linear(x, "x")
        nr, nc = Hp.size()
        if nc and istart - 1 + (nc - 1) * INCY >= yl.size:
            raise IndexError("BoundsError: y is too short")
        if nr and jstart - 1 + (nr - 1) * INCX >= xl.size:
            raise IndexError("BoundsError: x is too short")
        par = _offset_parity(Hp, istart, jstart, INCX, INCY)
        Hp.plan(parity=par).rmatvec(xl, yl, INCX, INCY, accumulate=True, xoff=jstart - 1, yoff=istart - 1)
        return y
    if isinstance(H, EvenBarycentricMatrix):  # leaf-level mul!(u, B, v, istart, jstart) -- algebra.jl:166-239
        if INCX != 1 or INCY != 1:
            raise TypeError("MethodError: mul!(u, ::EvenBarycentricMatrix, v, istart, jstart) has no strided form")
        if not (isinstance(y, np.ndarray) and isinstance(x, np.ndarray) and y.dtype == x.dtype == np.float64):
            raise TypeError("MethodError: u and v must be Float64 arrays")
        if istart < 1 or jstart < 1:
            raise IndexError("BoundsError: offsets are 1-based positive integers")
        yl, xl = _linear(y, "u"), _linear(x, "v")
        nr, nc = H.shape
        if istart - 1 + nr > yl.size or jstart - 1 + nc > xl.size:
            raise IndexError("BoundsError: u or v is too short")
        H.plan((istart + jstart) & 1).matvec(xl, yl, 1, 1, accumulate=True, xoff=jstart - 1, yoff=istart - 1)
        return y
    leaf, transposed = (H.parent, True) if isinstance(H, Transpose) else (H, False)
    if _leaf_kind(leaf) in (2, 3, 4):
        # leaf-level mul!(y, A, x, istart, jstart, INCX, INCY): Matrix algebra.jl:37-48, its
        # transpose :52-82, LowRankMatrix :110-131 / :138-159, BarycentricMatrix2D :243-277.
        # A one-leaf plan is packed for the call (leaves are plain arrays the caller may mutate).
        if _leaf_kind(leaf) == 4 and (INCX != 1 or INCY != 1 or transposed):
            raise TypeError("MethodError: BarycentricMatrix2D has only mul!(u, B, v, istart, jstart)")
        if not (isinstance(y, np.ndarray) and isinstance(x, np.ndarray)
                and y.dtype == x.dtype == leaf.dtype == np.float64):
            raise TypeError("MethodError: y, A and x must all be Float64")
        if isinstance(leaf, np.ndarray) and leaf.ndim != 2:
            raise TypeError("MethodError: A must be a matrix")
        if INCX < 1 or INCY < 1 or istart < 1 or jstart < 1:
            raise IndexError("BoundsError: offsets and strides are 1-based positive integers")
        yl, xl = _linear(y, "y"), _linear(x, "x")
        nr, nc = leaf.shape[::-1] if transposed else leaf.shape
        if nr and istart - 1 + (nr - 1) * INCY >= yl.size:
            raise IndexError("BoundsError: y is too short")
        if nc and jstart - 1 + (nc - 1) * INCX >= xl.size:
            raise IndexError("BoundsError: x is too short")
        P = _single_leaf_plan(leaf, 0, _current_device())
        apply = P.rmatvec if transposed else P.matvec
        apply(xl, yl, INCX, INCY, accumulate=True, xoff=jstart - 1, yoff=istart - 1)
        return y
    if not isinstance(H, _HierarchicalBase):
        raise TypeError("MethodError: H is not a hierarchical matrix")
    if isinstance(H, KernelMatrix) and (INCX != 1 or INCY != 1):
        raise TypeError("MethodError: mul!(u, ::KernelMatrix, v, istart, jstart) has no strided form")
    if not isinstance(y, np.ndarray) or not isinstance(x, np.ndarray):
        raise TypeError("y and x must be numpy arrays (y is updated in place)")
    if not (y.dtype == x.dtype == H.T == np.float64):
        raise TypeError("MethodError: y, H and x must all be Float64")
    if INCX < 1 or INCY < 1 or istart < 1 or jstart < 1:
        raise IndexError("BoundsError: offsets and strides are 1-based positive integers")
    yl, xl = _linear(y, "y"), _linear(x, "x")
    nr, nc = H.size()
    if nr and istart - 1 + (nr - 1) * INCY >= yl.size:
        raise IndexError("BoundsError: y is too short")
    if nc and jstart - 1 + (nc - 1) * INCX >= xl.size:
        raise IndexError("BoundsError: x is too short")
    par = _offset_parity(H, istart, jstart, INCX, INCY)
    H.plan(parity=par).matvec(xl, yl, INCX, INCY, accumulate=True, xoff=jstart - 1, yoff=istart - 1)
    return y
