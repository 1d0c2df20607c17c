"""The reference's example (examples/Kernel.jl) on the B200 engine.

Same experiment, same output lines: for each of the four asymptotically smooth kernels and
two families of point sets, build the hierarchical KernelMatrix, apply it to a random vector
and print the 2-norm relative error against the dense product (Kernel.jl:56-80, 84-110).
Here the operator is assembled and applied on the GPU through the C ABI.

    python examples/kernel.py            # N = 1000 and 10 000 as in the reference
    python examples/kernel.py 1000000    # one size of your choice (skips the dense check above 20 000)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmb200_loader  # noqa: E402

hm = hmb200_loader.load()


def point_sets(N):
    # Kernel.jl:61-62: Chebyshev points of the first and second kind
    yield "Chebyshev points", hm.chebyshevpoints(N), hm.chebyshevpoints(N, kind=2), (1.0, -1.0, 1.0, -1.0)
    # Kernel.jl:88-92: quadratic spacing on [0, inf)
    i = np.arange(N, 0, -1, dtype=np.float64)
    x, y = i * (i + 1), (i + 1 / 2) * (i + 3 / 2)
    yield "quadratic spacing", x, y, (x.max(), x.min(), y.max(), y.min())


def timed(label, fn):
    t0 = time.perf_counter()
    out = fn()
    print(f"  {time.perf_counter() - t0:10.6f} seconds  {label}")
    return out


def main(sizes):
    rng = np.random.default_rng(0)
    worst = 0.0
    for f in (hm.cauchykernel, hm.coulombkernel, hm.coulombprimekernel, hm.logkernel):
        for N in sizes:
            for name, x, y, (a, b, c, d) in point_sets(N):
                v = rng.standard_normal(N)
                print(f"\n{f} matrix construction at N = {N} ({name})\n")
                K = timed("KernelMatrix (GPU assembly)", lambda: hm.KernelMatrix(f, x, y, a, b, c, d))
                K = timed("KernelMatrix (GPU assembly)", lambda: hm.KernelMatrix(f, x, y, a, b, c, d))
                dense = N <= 20000
                if dense:
                    KF = timed("dense f(x, y)", lambda: f(x, y))
                print(f"\n{f} matrix-vector multiplication at N = {N}\n")
                u = timed("K*b", lambda: K * v)
                u = timed("K*b", lambda: K * v)
                if dense:
                    uf = timed("KF*b", lambda: KF @ v)
                    err = np.linalg.norm(u - uf) / np.linalg.norm(u)
                    worst = max(worst, err)
                    print(f"\n2-norm relative error: {err}\n")
    return worst


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [1000, 10_000])
