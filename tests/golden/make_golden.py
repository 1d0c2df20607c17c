"""Generates the committed golden fixtures of tests/golden/ (run once, in the build
container).  Everything here is computed with mpmath at 60 digits, independently of
oracle/ and of the CUDA library:

  cheb20.json        the 20 Chebyshev nodes and barycentric weights that
                     BLOCKRANK(Float64) = 20 selects
                     (/root/reference/src/BarycentricMatrix.jl:92-136): sin(pi q)
                     of the Float64-rounded argument q, correctly rounded
  points_*.json      sample entries of chebyshevpoints(Float64, N; kind) for large N
  cauchy_dense_*.json  K*b for the example's setup (examples/Kernel.jl:61-78) with the
                     kernel matrix applied densely at 60 digits -- the example's
                     own yardstick for the hierarchical product

  evenbary_cauchy.json  EvenBarycentricMatrix(Float64, (x,j) -> 1/(x-j), 1, 96, 300, 420)
                     (/root/reference/src/BarycentricMatrix.jl:18-45): sample entries of w
                     and W, and the masked product of algebra.jl:168-239 for an even and an
                     odd offset shift, all evaluated at 60 digits from the Float64 nodes/weights

  bary2d_block.json  one BarycentricMatrix2D block (/root/reference/src/BarycentricMatrix.jl:147-178,
                     236-297) for the Cauchy kernel on the box x in [0.5, 1], y in [-1, -0.5]:
                     U, F, V from the formulas at 60 digits (row-normalised lambda/(x - node))

The reference ships no golden vectors (test/runtests.jl draws from Julia's RNG), and
Julia is not installed here, so these are the strongest reference-independent pins
available.
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))


def sinpi_rounded(q: float) -> float:
    return float(mp.sin(mp.pi * mp.mpf(q)))


def chebpts(n, kind=1):
    x = [0.0] * n
    for k in range(1, n // 2 + 1):
        q = (n - 2 * k + 1.0) / (2.0 * n) if kind == 1 else (n - 2 * k + 1.0) / (2.0 * (n - 1))
        x[k - 1] = sinpi_rounded(q)
    for k in range(1, n // 2 + 1):
        x[n - k] = -x[k - 1]
    return x


def chebweights(n):
    lam = [0.0] * n
    for k in range(1, n // 2 + 2):
        lam[k - 1] = sinpi_rounded((2.0 * k - 1.0) / (2.0 * n))
    for k in range(1, n // 2 + 1):
        lam[n - k] = lam[k - 1]
    for k in range(2, n + 1, 2):
        lam[k - 1] = -lam[k - 1]
    return lam


def hexlist(v):
    return [float(a).hex() for a in v]


def main():
    json.dump({"n": 20, "nodes": hexlist(chebpts(20)), "weights": hexlist(chebweights(20))},
              open(os.path.join(HERE, "cheb20.json"), "w"), indent=1)

    samples = {}
    for n in (4096, 1 << 20):
        for kind in (1, 2):
            idx = sorted(set([1, 2, 3, n // 7, n // 3, n // 2 - 1, n // 2]))
            vals = []
            for k in idx:
                q = (n - 2 * k + 1.0) / (2.0 * n) if kind == 1 else (n - 2 * k + 1.0) / (2.0 * (n - 1))
                vals.append(sinpi_rounded(q))
            samples[f"{n}_{kind}"] = {"k": idx, "x": hexlist(vals)}
    json.dump(samples, open(os.path.join(HERE, "points_samples.json"), "w"), indent=1)

    # dense kernel product at 60 digits, N = 300 (ragged vs the 80-leaf size) and 1000
    for n in (300, 1000):
        x = chebpts(n, 1)
        y = chebpts(n, 2)
        b = np.random.default_rng(20261017 + n).standard_normal(n)
        xm = [mp.mpf(v) for v in x]
        ym = [mp.mpf(v) for v in y]
        bm = [mp.mpf(float(v)) for v in b]
        out = []
        for i in range(n):
            s = mp.mpf(0)
            for j in range(n):
                s += bm[j] / (xm[i] - ym[j])
            out.append(float(s))
        json.dump({"n": n, "seed": 20261017 + n, "b": hexlist(b), "Kb": hexlist(out)},
                  open(os.path.join(HERE, f"cauchy_dense_{n}.json"), "w"))


def evenbary():
    a, b, c, d, r = 1, 96, 300, 420, 20
    xk = [mp.mpf(v) for v in chebpts(r)]
    lam = [mp.mpf(v) for v in chebweights(r)]
    m, n = b - a + 1, d - c + 1
    den = [[mp.mpf(2 * i - a - b) - mp.mpf(b - a) * xk[k] for k in range(r)] for i in range(a, b + 1)]
    w = [sum(lam[k] / den[i][k] for k in range(r)) for i in range(m)]
    W = [[lam[k] / (den[i][k] * w[i]) for i in range(m)] for k in range(r)]  # W[k][i]
    nodes = [float(mp.mpf(a + b) / 2 + mp.mpf(b - a) * xk[k] / 2) for k in range(r)]  # as Float64 ops would give
    F = [[mp.mpf(1) / (mp.mpf(nodes[k]) - j) for k in range(r)] for j in range(c, d + 1)]  # F[j][k]
    v = np.random.default_rng(20261018).standard_normal(n)
    vm = [mp.mpf(float(t)) for t in v]
    prods = {}
    for shift in (0, 1):  # parity of (istart-1)+(jstart-1)
        out = []
        for i in range(m):
            s = mp.mpf(0)
            for j in range(n):
                if (shift + i + j) % 2 == 0:
                    s += sum(F[j][k] * W[k][i] for k in range(r)) * vm[j]
            out.append(float(s))
        prods[str(shift)] = hexlist(out)
    idx = [0, 1, 17, 48, 95]
    json.dump({"a": a, "b": b, "c": c, "d": d, "seed": 20261018, "v": hexlist(v),
               "w_idx": idx, "w": hexlist([w[i] for i in idx]),
               "W_cols": {str(i): hexlist([W[k][i] for k in range(r)]) for i in idx},
               "u": prods},
              open(os.path.join(HERE, "evenbary_cauchy.json"), "w"))


def bary2d_block():
    r = 20
    a, b, c, d = 1.0, 0.5, -0.5, -1.0          # descending boxes, as KernelMatrix passes them
    nodes = [mp.mpf(v) for v in chebpts(r)]
    lam = [mp.mpf(v) for v in chebweights(r)]
    rng = np.random.default_rng(20261019)
    x = np.sort(rng.uniform(0.5, 1.0, 37))[::-1].copy()
    y = np.sort(rng.uniform(-1.0, -0.5, 29))[::-1].copy()
    # mapped nodes as Float64 arithmetic gives them: (a+b)/2 + (b-a)/2 * node
    xn = [float(np.float64(0.5 * (a + b)) + np.float64(0.5 * (b - a)) * np.float64(float(t))) for t in nodes]
    yn = [float(np.float64(0.5 * (c + d)) + np.float64(0.5 * (d - c)) * np.float64(float(t))) for t in nodes]

    def factor(pts, nd):
        out = []
        for p in pts:
            row = [lam[k] / (mp.mpf(float(p)) - mp.mpf(nd[k])) for k in range(r)]
            tot = sum(row)
            out.append([float(v / tot) for v in row])
        return out
    U = factor(x, xn)
    V = factor(y, yn)
    F = [[float(mp.mpf(1) / (mp.mpf(xn[p]) - mp.mpf(yn[q]))) for q in range(r)] for p in range(r)]
    json.dump({"a": a, "b": b, "c": c, "d": d, "x": hexlist(x), "y": hexlist(y),
               "U": [hexlist(row) for row in U], "V": [hexlist(row) for row in V],
               "F": [hexlist(row) for row in F]},
              open(os.path.join(HERE, "bary2d_block.json"), "w"))


if __name__ == "__main__":
    evenbary()
    bary2d_block()
    main()
