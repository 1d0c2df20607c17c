"""World-size-2 `gloo` test of the N>1 host logic (runs on CPU): every rank plans its own
block-row part of the same operator; the parts must tile the rows, agree on the whole
operator's accounting, and the rank-0 broadcast / all-gather plumbing used by bench.py
must move the right slices."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import hmb200_loader
        hm = hmb200_loader.load()
        n = 6000
        x, y = hm.chebyshevpoints(n), hm.chebyshevpoints(n, 2)
        st = hm.KernelMatrix.layout_stats(x, y, 1.0, -1.0, 1.0, -1.0, rank, world)
        cuts = [None] * world
        dist.all_gather_object(cuts, (st["row_begin"], st["row_end"], st["part_words"], st["algorithmic_bytes"]))
        # x replicated from rank 0 (bench.py: dist.broadcast), y slices gathered with padding
        v = torch.from_numpy(np.random.default_rng(0).standard_normal(n)) if rank == 0 else torch.zeros(n, dtype=torch.float64)
        dist.broadcast(v, src=0)
        r0, r1 = st["row_begin"], st["row_end"]
        yloc = torch.zeros(n, dtype=torch.float64)
        yloc[r0:r1] = v[r0:r1] * (rank + 1)  # stand-in for the owned rows of y
        # the product itself, partitioned the way the GPUs partition it: this rank applies the rows
        # [r0, r1) of every leaf of the planner's own tree that meets them (a leaf whose rows straddle
        # the cut is applied in part by both ranks: U rows stay with their owners, V and F are used by
        # both), leaf arithmetic from the oracle's factor builder; gathered, it must be the oracle's mul!
        import ctypes as C
        from oracle import oracle as O
        dp = C.POINTER(C.c_double)
        cnt = C.c_int64()
        hm._lib.check(hm.lib().hm_kernel_tree_leaves(x.ctypes.data_as(dp), n, y.ctypes.data_as(dp), n, 1.0, -1.0, 1.0, -1.0,
                                                     None, 0, C.byref(cnt)))
        leaves = (hm._lib.TreeLeaf * cnt.value)()
        hm._lib.check(hm.lib().hm_kernel_tree_leaves(x.ctypes.data_as(dp), n, y.ctypes.data_as(dp), n, 1.0, -1.0, 1.0, -1.0,
                                                     leaves, cnt.value, C.byref(cnt)))
        vn = v.numpy()
        part = np.zeros(n)
        for l in leaves:
            a, b = max(l.row0, r0), min(l.row0 + l.m, r1)
            if b <= a:
                continue
            vv = vn[l.col0:l.col0 + l.n]
            if l.kind == 3:   # Matrix leaf: T[f(x[i], y[j])]
                part[a:b] += (1.0 / (x[l.xi0 + (a - l.row0):l.xi0 + (b - l.row0), None] - y[None, l.yj0:l.yj0 + l.n])) @ vv
            else:             # BarycentricMatrix2D leaf: U (F (V' v))
                U, F, V = O.bary2d_build(O.CAUCHY, l.a, l.b, l.c, l.d, x, l.xi0, l.xi0 + l.m, y, l.yj0, l.yj0 + l.n)
                part[a:b] += U[a - l.row0:b - l.row0] @ (F @ (V.T @ vv))
        prod = [None] * world
        dist.all_gather_object(prod, (r0, r1, part[r0:r1]))
        full = np.zeros(n)
        for a, b, seg in prod:
            full[a:b] = seg
        ref = O.kernelmatrix(O.CAUCHY, x, y, 1.0, -1.0, 1.0, -1.0).matvec(vn)
        prod_err = float(np.max(np.abs(full - ref)) / np.max(np.abs(ref)))
        maxrows = max(b - a for a, b, _, _ in cuts)
        pad = torch.zeros(maxrows, dtype=torch.float64)
        pad[: r1 - r0] = yloc[r0:r1]
        gathered = [torch.zeros(maxrows, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, pad)
        for qk, (a, b, _, _) in enumerate(cuts):
            yloc[a:b] = gathered[qk][: b - a]
        expect = v.clone()
        for qk, (a, b, _, _) in enumerate(cuts):
            expect[a:b] *= qk + 1
        q.put((rank, cuts, bool(torch.equal(yloc, expect)), float(v.sum()), prod_err))
    finally:
        dist.destroy_process_group()


def test_two_rank_row_partition_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cuts = res[0][1]
    assert res[1][1] == cuts                       # every rank computed the same partition
    assert cuts[0][0] == 0 and cuts[-1][1] == 6000
    assert cuts[0][1] == cuts[1][0]                # parts tile the rows
    assert cuts[0][3] == cuts[1][3]                # same whole-operator byte count
    w = [c[2] for c in cuts]
    assert max(w) <= 1.2 * sum(w) / 2              # balanced by stored words
    assert res[0][2] and res[1][2]                 # gathered y is identical and complete on both
    assert res[0][3] == res[1][3]                  # x was replicated
    assert res[0][4] <= 1e-12 and res[1][4] == res[0][4]   # the partitioned product is the oracle's mul! on both ranks
