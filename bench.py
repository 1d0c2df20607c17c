#!/usr/bin/env python
"""Headline benchmark: H-matrix matvec, Cauchy kernel, N = 2^20, Float64 (BASELINE.json
configs[1]).  One "step" = one y = K x through the three-stage CUDA path.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the restated
                                                            reference loops, all host cores)

N > 1: launched under torchrun, one process per GPU; the operator is partitioned by
block rows (each rank owns a row range of y), x is replicated with an NCCL broadcast
from rank 0 and the owned y slices are all-gathered, every step.  The whole operator is
fixed as N grows -> "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npoints", "--size", dest="n", type=int, default=1 << 20,
                    help="points per side (default 2^20)")
    ap.add_argument("--no-p2p", action="store_true",
                    help="N > 1: gather y with an NCCL all-gather instead of the peer-memory stores fused into stage 3")
    ap.add_argument("--no-graph", action="store_true",
                    help="N > 1: launch every step from Python instead of one CUDA graph of the whole loop")
    ap.add_argument("--dist", default="cheb", choices=["cheb", "unif"],
                    help="cheb = examples/Kernel.jl:61-62 point sets; unif = uniform interlaced")
    ap.add_argument("--no-gather", action="store_true", help="skip the all-gather of y (N > 1)")
    ap.add_argument("--no-matrix-free", action="store_true", help="skip the secondary matrix-free measurement")
    ap.add_argument("--matrix-free", action="store_true",
                    help="hm_assemble_kernel_free: store no U/V/dense tiles, evaluate the entries inside every matvec "
                         "(FP64-bound; not the headline configuration)")
    ap.add_argument("--nrhs", type=int, default=1,
                    help="> 1: time the multi-right-hand-side product (BASELINE configs[2]/[4]) instead")
    ap.add_argument("--adjoint", action="store_true",
                    help="single GPU: time the adjoint apply y = K'x (SURVEY 8f f2) instead of K x")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def points(hm, n, dist):
    if dist == "cheb":
        return hm.chebyshevpoints(n, 1), hm.chebyshevpoints(n, 2)
    i = np.arange(1, n + 1, dtype=np.float64)
    return 1.0 - 2.0 * (i - 0.5) / n, 1.0 - 2.0 * (i - 0.25) / n


def workload_name(n, dist):
    p = "Chebyshev 1st/2nd-kind points (examples/Kernel.jl:61-62)" if dist == "cheb" else "uniform interlaced points"
    return f"Cauchy 1/(x-y) KernelMatrix, N={n}, Float64, single-vector mul!, {p}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self, t_begin=None):
        """Samples taken after wall-clock time t_begin (the timed region); if the region was
        too short to catch three, all samples since start (warm-up + timed, same load)."""
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        recs = []
        for ln in self.lines:
            f = [a.strip() for a in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                recs.append((ts, float(f[1]), float(f[2]), [nm for nm, val in zip(names, f[5:9])
                                                            if val.lower().startswith("active")]))
            except ValueError:
                continue
        timed = [r for r in recs if t_begin is not None and r[0] >= t_begin]
        window = "timed region"
        if len(timed) < 3:
            timed, window = recs, "warm-up + timed region (timed region shorter than 3 samples)"
        sm = [r[1] for r in timed]
        reasons = sorted({nm for r in timed for nm in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(r[2] for r in timed) if timed else None,
                "samples": len(sm), "window": window, "reasons": reasons}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------- CPU legs (oracle = port)
def cpu_oracle_tree(n, dist):
    from oracle import oracle as O
    x, y, (a, b, c, d) = O.example_points(n, dist)
    t0 = time.perf_counter()
    K = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    return O, K, time.perf_counter() - t0


def host_mem_ok(n):
    need = 14.2e9 * (n / float(1 << 20)) * 1.15
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return float(ln.split()[1]) * 1024 > need
    except OSError:
        pass
    return True


def cpu_baseline(n, dist, v, y_gpu=None):
    """The restated reference loops on the host: one faithful single-thread matvec and
    two all-core matvecs of the same operator (bounded sample: ~5-10 s of CPU work
    after a ~20 s single-thread assembly)."""
    n_cpu = n
    while n_cpu > 4096 and not host_mem_ok(n_cpu):
        n_cpu //= 2
    O, K, t_build = cpu_oracle_tree(n_cpu, dist)
    vv = v[:n_cpu].copy()
    t0 = time.perf_counter()
    ref = K.matvec(vv)
    t1 = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    tt = []
    for _ in range(2):
        o = np.zeros(n_cpu)
        t0 = time.perf_counter()
        K.mul_omp(o, vv, cores)
        tt.append(time.perf_counter() - t0)
    out = {
        "value": 1.0 / min(tt), "unit": "matvecs/s", "cores": cores, "kind": "port",
        "sample": (f"full N={n_cpu} operator: 2 all-core matvecs (best {min(tt):.3f} s) and 1 single-thread "
                   f"reference-order matvec ({t1:.3f} s = {1.0 / t1:.3f} matvecs/s); restated reference "
                   f"(C port of the Julia loops, oracle/hm_oracle.c), not Julia; assembly {t_build:.1f} s"),
        "single_thread_value": 1.0 / t1,
    }
    parity = None
    if y_gpu is not None and n_cpu == n:
        parity = float(np.max(np.abs(y_gpu - ref)) / np.max(np.abs(ref)))
    return out, parity


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, dist = args.n, args.dist
    while n > 4096 and not host_mem_ok(n):
        n //= 2
    O, K, t_build = cpu_oracle_tree(n, dist)
    v = np.random.default_rng(0).standard_normal(n)
    cores = os.cpu_count() or 1
    o = np.zeros(n)
    t0 = time.perf_counter()
    K.mul_omp(o, v, cores)
    t_first = time.perf_counter() - t0
    for _ in range(max(args.warmup - 1, 0)):
        if t_first * args.warmup > 60:
            break
        K.mul_omp(o, v, cores)
    reps = max(1, min(args.steps, int(120.0 / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        K.mul_omp(o, v, cores)
    dt = (time.perf_counter() - t0) / reps
    val = 1.0 / dt
    words = K.stored_words()
    line = {
        "impl": "reference", "metric": "H-matvec matvecs/s", "value": val, "unit": "matvecs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(n, dist), "n": n, "dist": dist},
        "effective_gbs": (8 * words + 16 * n) * val / 1e9,
        "cpu_baseline": {"value": val, "unit": "matvecs/s", "cores": cores, "kind": "port",
                         "sample": f"full N={n} operator per step, {reps} of {args.steps} steps timed; restated "
                                   f"reference (C port, OpenMP over leaves), not Julia (julia is not installed); "
                                   f"assembly {t_build:.1f} s"},
        "e2e": {"value": val, "unit": "matvecs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- multi-RHS mode (single GPU)
def run_matmat(args, hm, torch, plan, st, px, py, dev, t_asm):
    """Y = K X with X of `nrhs` columns (BASELINE configs[2]: N = 2^20, 64 right-hand sides; FP64
    tensor-core panel kernels).  A step is one product; value = columns per second."""
    n, nrhs = args.n, args.nrhs
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.standard_normal((nrhs, n))).to(dev)   # column-major n x nrhs
    Y = torch.zeros((nrhs, n), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def run(k):
        for _ in range(k):
            plan.matmat_device(X.data_ptr(), n, Y.data_ptr(), n, nrhs, accumulate=False, stream=stream.cuda_stream)

    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    run(max(args.warmup, 3))
    torch.cuda.synchronize()
    plan.timing_begin(args.steps * ((nrhs + 63) // 64))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    run(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    stage_ms, ncalls = plan.timing_end()
    clocks = sampler.stop(t_wall0)
    # measured FP64 GEMM peak on this box (cuBLAS DGEMM 8192^3), the FP64-pipe denominator
    A = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    B = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(A, B)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        torch.matmul(A, B)
        a1.record()
        torch.cuda.synchronize()
        best = min(best, a0.elapsed_time(a1))
    dgemm_tflops = 2 * 8192 ** 3 / (best / 1e3) / 1e12
    del A, B
    words = st["dense_words"] + st["lowrank_words"]
    flops = 2.0 * words * nrhs
    cs = 16 if nrhs <= 16 else 32 if nrhs <= 32 else 64
    bytes_alg = 8 * words + 16 * n * nrhs
    peak, peak_src = measured_peak()
    # spot check of a few entries against dense kernel rows in long double
    Yh = Y.cpu().numpy()
    Xh = X.cpu().numpy()
    rows = np.unique(rng.integers(0, n, 12))
    cols = [0, nrhs - 1]
    xl, yl = px.astype(np.longdouble), py.astype(np.longdouble)
    err = 0.0
    for c in cols:
        dense = np.array([np.sum(Xh[c].astype(np.longdouble) / (xl[i] - yl)) for i in rows], dtype=np.float64)
        err = max(err, float(np.max(np.abs(Yh[c, rows] - dense)) / np.max(np.abs(dense))))
    tf = flops / (ms / 1e3) / 1e12
    gbs = bytes_alg / (ms / 1e3) / 1e9
    line = {
        "metric": "H-matmat columns/s", "value": nrhs / (ms / 1e3), "unit": "columns/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, args.dist).replace("single-vector mul!", f"{nrhs} right-hand sides"),
                   "n": n, "dist": args.dist, "nrhs": nrhs, "panel_width": cs, "assembly_s": round(t_asm, 3),
                   "l2": "inputs larger than L2"},
        "tflops": tf, "effective_gbs": gbs,
        "roofline": {"bound": "tensor" if tf / dgemm_tflops > gbs / peak else "hbm",
                     "kernel": "hm_panel_kernel (DMMA m8n8k4, stage 1 + stage 3)",
                     "achieved_tflops": tf, "peak_tflops": dgemm_tflops,
                     "peak_tflops_source": "cuBLAS DGEMM 8192^3 measured in this run (best of 5)",
                     "frac_fp64": tf / dgemm_tflops, "achieved_gbs": gbs, "peak_gbs": peak, "peak_gbs_source": peak_src,
                     "frac_hbm": gbs / peak, "traffic": None,
                     "ms_per_launch": {"stage1+in": stage_ms[0] / max(ncalls, 1), "stage2": stage_ms[1] / max(ncalls, 1),
                                       "stage3+out": stage_ms[2] / max(ncalls, 1)}},
        "gpu_launches": args.steps * ((nrhs + 63) // 64) * (plan.launches_per_matvec + 2), "clocks": clocks,
        "check_sampled_dense_entries_relerr": err,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import hmb200_loader
    hm = hmb200_loader.load()
    if not os.path.exists(hm._lib.LIB_PATH):
        # a fresh checkout carries no binaries: compile the CUDA library (nvcc, sm_100a) once;
        # under torchrun only local rank 0 builds, the others wait for the file
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            hm.build()
        else:
            for _ in range(600):
                if os.path.exists(hm._lib.LIB_PATH):
                    break
                time.sleep(0.5)
            time.sleep(1.0)
    hm.lib()  # fails loudly if the CUDA library is missing: there is no CPU fallback

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the matvec has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    px, py = points(hm, n, args.dist)
    t0 = time.perf_counter()
    K = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=local, part=rank, nparts=world,
                        matrix_free=args.matrix_free)
    torch.cuda.synchronize()
    t_asm = time.perf_counter() - t0
    plan = K.plan()
    st = plan.stats()
    r0, r1 = st["row_begin"], st["row_end"]

    if args.nrhs > 1:
        return run_matmat(args, hm, torch, plan, st, px, py, dev, t_asm)
    if args.adjoint:
        xa = torch.from_numpy(np.random.default_rng(0).standard_normal(n)).to(dev)
        ya = torch.zeros(n, dtype=torch.float64, device=dev)
        sm = torch.cuda.current_stream()
        for _ in range(max(args.warmup, 3)):
            plan.rmatvec_device(xa.data_ptr(), ya.data_ptr(), False, sm.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sm)
        for _ in range(args.steps):
            plan.rmatvec_device(xa.data_ptr(), ya.data_ptr(), False, sm.cuda_stream)
        e1.record(sm)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        rows = np.unique(np.random.default_rng(1).integers(0, n, 24))
        xl, yl, vl = px.astype(np.longdouble), py.astype(np.longdouble), xa.cpu().numpy().astype(np.longdouble)
        dense = np.array([np.sum(vl / (xl - yl[j])) for j in rows], dtype=np.float64)
        err = float(np.max(np.abs(ya.cpu().numpy()[rows] - dense)) / np.max(np.abs(dense)))
        peak, _ = measured_peak()
        print(json.dumps({"metric": "H-matvec adjoint matvecs/s", "value": 1e3 / ms, "unit": "matvecs/s", "n_gpus": 1,
                          "steps": args.steps, "ms_per_step": ms, "dtype": "f64",
                          "config": {"workload": workload_name(n, args.dist).replace("mul!", "adjoint mul!"), "n": n},
                          "effective_gbs": st["algorithmic_bytes"] / ms / 1e6,
                          "frac_of_measured_hbm": st["algorithmic_bytes"] / ms / 1e6 / peak,
                          "check_sampled_dense_columns_relerr": err}), flush=True)
        return

    v = np.random.default_rng(0).standard_normal(n)
    gather = dist_on and not args.no_gather
    stream = torch.cuda.current_stream()
    # two x / y buffers: the NCCL broadcast of x for step k+1 and the all-gather of y from
    # step k-1 run on a second stream while step k computes
    nbuf = 2 if dist_on else 1
    x_bufs = [torch.from_numpy(v).to(dev) if rank == 0 else torch.zeros(n, dtype=torch.float64, device=dev)
              for _ in range(nbuf)]
    y_bufs = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(nbuf)]
    p2p = None
    if dist_on:
        cuts = [None] * world
        dist.all_gather_object(cuts, (r0, r1))
        maxrows = max(b - a for a, b in cuts)
        pad = torch.zeros(maxrows, dtype=torch.float64, device=dev)
        gathered = torch.zeros(world * maxrows, dtype=torch.float64, device=dev)
        cs = torch.cuda.Stream(device=dev)
        x_ready = [torch.cuda.Event() for _ in range(nbuf)]
        xy_free = [torch.cuda.Event() for _ in range(nbuf)]
        y_ready = [torch.cuda.Event() for _ in range(nbuf)]
        y_done = [torch.cuda.Event() for _ in range(nbuf)]
        if gather and not args.no_p2p:
            # y lives in symmetric memory: every rank can store into every rank's buffer over
            # NVLink, so stage 3 writes its rows to all of them (all-gather fused into the kernel)
            # and only a device barrier remains on the second stream.
            ok = torch.ones(1, device=dev)
            try:
                import torch.distributed._symmetric_memory as symm_mem
                ysym = [symm_mem.empty(n, dtype=torch.float64, device=dev) for _ in range(nbuf)]
                hdls = [symm_mem.rendezvous(t, dist.group.WORLD) for t in ysym]
                for t in ysym:
                    t.zero_()
                p2p = {"bufs": ysym, "hdls": hdls, "ptrs": [[int(a) for a in h.buffer_ptrs] for h in hdls]}
            except Exception as exc:  # symmetric memory unavailable: NCCL all-gather instead
                ok.zero_()
                if rank == 0:
                    print(f"bench.py: symmetric memory unavailable ({exc!r}); using the NCCL all-gather", file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                p2p = None
            else:
                y_bufs = p2p["bufs"]

    def bcast(k, main):
        b = k % nbuf
        with torch.cuda.stream(cs):
            cs.wait_event(xy_free[b])          # the matvec that last read x_bufs[b] is done
            dist.broadcast(x_bufs[b], src=0)
            x_ready[b].record(cs)

    def run(nsteps):
        main = torch.cuda.current_stream()
        if not dist_on:
            for _ in range(nsteps):
                plan.matvec_device(x_bufs[0].data_ptr(), y_bufs[0].data_ptr(), accumulate=False,
                                   stream=main.cuda_stream)
            return
        for b in range(nbuf):
            xy_free[b].record(main)
            y_done[b].record(main)
        bcast(0, main)
        for k in range(nsteps):
            b = k % nbuf
            main.wait_event(x_ready[b])
            main.wait_event(y_done[b])         # the all-gather that last read y_bufs[b] is done
            if p2p is not None:
                plan.matvec_device_allgather(x_bufs[b].data_ptr(), p2p["ptrs"][b], rank, accumulate=False,
                                             stream=main.cuda_stream)
            else:
                plan.matvec_device(x_bufs[b].data_ptr(), y_bufs[b].data_ptr(), accumulate=False,
                                   stream=main.cuda_stream)
            xy_free[b].record(main)
            y_ready[b].record(main)
            if k + 1 < nsteps:
                bcast(k + 1, main)
            with torch.cuda.stream(cs):
                if gather:
                    cs.wait_event(y_ready[b])
                    if p2p is not None:
                        p2p["hdls"][b].barrier()   # every rank's rows have landed everywhere
                    else:
                        # row parts are balanced by bytes, not rows: gather padded slices
                        pad[: r1 - r0].copy_(y_bufs[b][r0:r1])
                        dist.all_gather_into_tensor(gathered, pad)
                        for q, (qa, qb) in enumerate(cuts):
                            if q != rank:
                                y_bufs[b][qa:qb].copy_(gathered[q * maxrows: q * maxrows + (qb - qa)])
                y_done[b].record(cs)
        main.wait_stream(cs)

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    run(max(args.warmup, 3))
    barrier()
    use_graph = dist_on and not args.no_graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if use_graph:
        # the whole K-step pipeline (our kernels + the NCCL collectives, two streams) as ONE
        # CUDA graph: at 0.3 ms of GPU work per step the Python/NCCL launch path is the bottleneck
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run(args.steps)
        barrier()
        graph.replay()                          # untimed replay (graph upload, NCCL warm-up)
        barrier()
        t_wall0 = time.time()
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        barrier()
        # per-stage kernel times for the roofline: the same steps launched eagerly
        plan.timing_begin(args.steps)
        run(args.steps)
        barrier()
    else:
        plan.timing_begin(args.steps)
        barrier()
        t_wall0 = time.time()
        e0.record(stream)
        run(args.steps)
        e1.record(stream)
        barrier()
    y_dev = y_bufs[(args.steps - 1) % nbuf]
    ms = e0.elapsed_time(e1)
    stage_ms, ncalls = plan.timing_end()
    clocks = sampler.stop(t_wall0) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps / (ms / 1e3)

    # per-stage roofline inputs of this rank
    b1 = 8 * st["part_v_words"] + 8 * st["ncols"]
    b2 = 8 * st["part_core_words"]
    b3 = 8 * (st["part_u_words"] + st["part_dense_words"]) + 8 * (r1 - r0)
    s1, s2, s3 = (m / max(ncalls, 1) for m in stage_ms)

    # ---- end to end through the host-pointer C ABI call (pinned host buffers) ----
    e2e = None
    if not args.no_e2e:
        xh = torch.from_numpy(v.copy()).pin_memory()
        yh = torch.zeros(n, dtype=torch.float64).pin_memory()
        xn, yn = xh.numpy(), yh.numpy()
        for _ in range(3):
            plan.matvec(xn, yn, accumulate=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            plan.matvec(xn, yn, accumulate=False)  # H2D x, 3 stages, D2H y rows, sync
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": args.steps / dt, "unit": "matvecs/s", "h2d_bytes_per_step": 8 * st["ncols"] * world,
               "d2h_bytes_per_step": 8 * st["nrows"], "ms_per_step": dt / args.steps * 1e3,
               "api": "hm_matvec (C ABI, host pointers)"}

    # ---- the same operator applied matrix-free (hm_assemble_kernel_free), measured beside the
    # headline stored path: nothing but the r x r cores is resident, the entries are evaluated
    # inside the matvec (FP64-bound).  Reported, not the headline: the roofline above is the
    # stored path's. ----
    mfree = None
    if world == 1 and not args.matrix_free and not args.no_matrix_free:
        t0 = time.perf_counter()
        Kf = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=local, matrix_free=True)
        pf = Kf.plan()
        t_setup = time.perf_counter() - t0
        yf = torch.zeros(n, dtype=torch.float64, device=dev)
        main = torch.cuda.current_stream()
        for _ in range(5):
            pf.matvec_device(x_bufs[0].data_ptr(), yf.data_ptr(), accumulate=False, stream=main.cuda_stream)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nmf = max(10, min(args.steps, 50))
        f0.record(main)
        for _ in range(nmf):
            pf.matvec_device(x_bufs[0].data_ptr(), yf.data_ptr(), accumulate=False, stream=main.cuda_stream)
        f1.record(main)
        torch.cuda.synchronize()
        ms_f = f0.elapsed_time(f1) / nmf
        ref_y = y_bufs[0]
        plan.matvec_device(x_bufs[0].data_ptr(), ref_y.data_ptr(), accumulate=False, stream=main.cuda_stream)
        torch.cuda.synchronize()
        dev_rel = float((yf - ref_y).abs().max() / ref_y.abs().max())
        mfree = {"value": 1e3 / ms_f, "unit": "matvecs/s", "ms_per_step": ms_f, "steps": nmf,
                 "resident_bytes": pf.stats()["stored_bytes"], "setup_s": round(t_setup, 3),
                 "relinf_vs_stored": dev_rel, "api": "hm_assemble_kernel_free + hm_matvec_device"}
        if not args.no_e2e:
            for _ in range(3):
                pf.matvec(xn, yn, accumulate=False)
            t0 = time.perf_counter()
            for _ in range(nmf):
                pf.matvec(xn, yn, accumulate=False)
            torch.cuda.synchronize()
            mfree["e2e"] = {"value": nmf / (time.perf_counter() - t0), "unit": "matvecs/s",
                            "api": "hm_matvec (C ABI, host pointers)"}
        del Kf, pf, yf

    y_host = y_dev.cpu().numpy() if (rank == 0 and (gather or not dist_on)) else None
    sampled = None
    if y_host is not None:
        # independent spot check (examples/Kernel.jl:78 on sampled rows): dense kernel rows in long double
        rows = np.unique(np.random.default_rng(1).integers(0, n, 48 if n <= (1 << 21) else 12))
        xl, yl, vl = px.astype(np.longdouble), py.astype(np.longdouble), v.astype(np.longdouble)
        dense = np.array([np.sum(vl / (xl[i] - yl)) for i in rows], dtype=np.float64)
        sampled = float(np.max(np.abs(y_host[rows] - dense)) / np.max(np.abs(dense)))

    if rank == 0:
        peak, peak_src = measured_peak()
        ach = (b1 + b3) / ((s1 + s3) / 1e3) / 1e9 if (s1 + s3) > 0 else None
        traffic = ncu_traffic() if (world == 1 and n == (1 << 20) and args.dist == "cheb") else None
        line = {
            "metric": "H-matvec matvecs/s", "value": value, "unit": "matvecs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(n, args.dist), "n": n, "dist": args.dist,
                       "partition": f"block-row x{world}" if world > 1 else "single GPU",
                       "collectives": ((("NCCL broadcast(x); all-gather(y) fused into stage 3 (stores into every rank's symmetric buffer "
                                         "over NVLink) + device barrier, pipelined on a second stream" if p2p is not None else
                                         "NCCL broadcast(x) + all-gather(y) per step, pipelined on a second stream")
                                        + (", whole loop replayed as one CUDA graph" if use_graph else "")) if gather else
                                       "NCCL broadcast(x) per step" if dist_on else "none"),
                       "l2": "inputs larger than L2 (%.1f GB streamed per step per GPU)" % (st["stored_bytes"] / 1e9),
                       "assembly_s": round(t_asm, 3), **({"matrix_free": True} if args.matrix_free else {})},
            "effective_gbs": st["algorithmic_bytes"] * value / 1e9,
            "algorithmic_bytes_per_matvec": st["algorithmic_bytes"],
            "roofline_frac_whole_step": st["algorithmic_bytes"] * value / 1e9 / (peak * world),
            "roofline": {"bound": "hbm", "kernel": ("hm_free1/hm_free3 (matrix-free: entries evaluated on the fly; 'achieved' is "
                                                   "the stored operator's algorithmic bytes per second, not HBM traffic)"
                                                   if args.matrix_free else
                                                   "hm_stream_kernel (stage 1 + stage 3 instantiations, rank 0)"),
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_launch": {"stage1": b1, "stage3": b3},
                         "ms_per_launch": {"stage1": s1, "stage2": s2, "stage3": s3}},
            "stages": {"stage1_gbs": b1 / (s1 / 1e3) / 1e9 if s1 > 0 else None,
                       "stage2_gbs": b2 / (s2 / 1e3) / 1e9 if s2 > 0 else None,
                       "stage3_gbs": b3 / (s3 / 1e3) / 1e9 if s3 > 0 else None},
            "e2e": e2e, "matrix_free": mfree, "gpu_launches": plan.launches_per_matvec * args.steps, "clocks": clocks,
            "leaves": {"dense": st["n_dense"], "bary2d": st["n_bary2d"]},
            "check_sampled_dense_rows_relerr": sampled,
        }
        if world == 1 and not args.no_cpu_baseline:
            del K, plan
            torch.cuda.empty_cache()
            cb, parity = cpu_baseline(n, args.dist, v, y_host)
            line["cpu_baseline"] = cb
            line["parity_relinf_vs_oracle"] = parity
        print(json.dumps(line), flush=True)
    if dist_on:
        # Tearing NCCL down while a captured graph still references its streams can hang:
        # release the graph, make sure every rank is done, then leave without the
        # communicator destructor.
        if use_graph:
            graph.reset()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
