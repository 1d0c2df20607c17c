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
    ap.add_argument("--no-graph", action="store_true",
                    help="N > 1: launch every step from Python instead of one CUDA graph of the whole loop")
    ap.add_argument("--dist", default="cheb", choices=["cheb", "unif"],
                    help="cheb = examples/Kernel.jl:61-62 point sets; unif = uniform interlaced")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary configurations (64-RHS product, 2^22 assembly + 16 RHS, 2^24 multi-GPU)")
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


# --------------------------------------------------------------------------- helpers
class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_view(torch, ptr, n, dev):
    """torch view (no copy) of n Float64 words at a device address handed out by the library."""
    return torch.as_tensor(_DevArr(ptr, n), device=dev)


_DGEMM = {}


def dgemm_peak(torch, dev):
    """cuBLAS DGEMM 8192^3 measured in this run (best of 5): the FP64-pipe denominator, which
    MEASURED_PEAKS.json does not carry."""
    if "tf" not in _DGEMM:
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
        _DGEMM["tf"] = 2 * 8192 ** 3 / (best / 1e3) / 1e12
        del A, B
        torch.cuda.empty_cache()
    return _DGEMM["tf"]


def sampled_rows_check(px, py, v, y_host, nrows_sampled, seed=1):
    """examples/Kernel.jl:78 on sampled rows: dense Cauchy rows in long double."""
    n = len(px)
    rows = np.unique(np.random.default_rng(seed).integers(0, n, nrows_sampled))
    xl, yl, vl = px.astype(np.longdouble), py.astype(np.longdouble), v.astype(np.longdouble)
    dense = np.array([np.sum(vl / (xl[i] - yl)) for i in rows], dtype=np.float64)
    return float(np.max(np.abs(y_host[rows] - dense)) / np.max(np.abs(dense)))


# --------------------------------------------------------------------------- multi-RHS product
def measure_matmat(torch, plan, st, px, py, dev, nrhs, steps, warmup):
    """Y = K X with X of `nrhs` columns through hm_matmat_device (FP64 tensor-core panel kernels,
    hm_panel.cu).  A step is one product; value = columns per second."""
    n = st["ncols"]
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.standard_normal((nrhs, n))).to(dev)   # column-major n x nrhs
    Y = torch.zeros((nrhs, st["nrows"]), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def run(k):
        for _ in range(k):
            plan.matmat_device(X.data_ptr(), n, Y.data_ptr(), st["nrows"], nrhs, accumulate=False,
                               stream=stream.cuda_stream)

    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    run(max(warmup, 3))
    torch.cuda.synchronize()
    npanels = (nrhs + 63) // 64
    plan.timing_begin(steps * npanels)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    run(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    stage_ms, ncalls = plan.timing_end()
    clocks = sampler.stop(t_wall0)
    dgemm_tflops = dgemm_peak(torch, dev)
    words = st["dense_words"] + st["lowrank_words"]
    form = plan.form
    flops_stored = 2.0 * words * nrhs
    if form == 3:
        # nested-basis form (hm_nest*): the work is that of the dense leaves, one 20-term moment / series
        # per point and side, and one 20 x 20 core per leaf; the box-to-box translations (~3 % more) are
        # not counted.  No operator bytes are read: the bound is the FP64 pipe.
        flops = 2.0 * nrhs * (st["dense_words"] + 20.0 * (st["ncols"] + st["nrows"]) + 400.0 * st["n_bary2d"])
        bytes_alg = 8 * (st["ncols"] + st["nrows"]) * nrhs
        kern = ("hm_nest_*_panel_kernel + hm_free3_panel_kernel (nested-basis form; FP64 DMMA m8n8k4, entries of the "
                "dense leaves evaluated into MMA fragments)")
    else:
        flops = flops_stored
        bytes_alg = (8 * words if form == 0 else 0) + 8 * (st["ncols"] + st["nrows"]) * nrhs
        kern = ("hm_panel kernels (FP64 DMMA m8n8k4; stage 1 + stage 2 + stage 3)" if form == 0 else
                "hm_free1/3_panel_kernel (U, V and dense entries evaluated into MMA fragments)")
    peak, peak_src = measured_peak()
    # spot check of two columns against dense kernel rows in long double
    Yh = Y.cpu().numpy()
    Xh = X.cpu().numpy()
    err = max(sampled_rows_check(px, py, Xh[c], Yh[c], 12, seed=2 + c) for c in (0, nrhs - 1))
    tf = flops / (ms / 1e3) / 1e12
    gbs = bytes_alg / (ms / 1e3) / 1e9
    bound = "tensor" if flops / (dgemm_tflops * 1e12) > bytes_alg / (peak * 1e9) else "hbm"
    out = {
        "value": nrhs / (ms / 1e3), "unit": "columns/s", "ms_per_step": ms, "steps": steps, "nrhs": nrhs,
        "panel_width": 16 if nrhs <= 16 else 32 if nrhs <= 32 else 64,
        "form": ["stored", "matrix-free (barycentric)", "matrix-free (Chebyshev, leaf by leaf)",
                 "matrix-free nested-basis"][form],
        "tflops": tf, "effective_gbs": gbs, "algorithmic_flops": flops, "algorithmic_bytes": bytes_alg,
        "stored_form_flops": flops_stored, "stored_form_equivalent_tflops": flops_stored / (ms / 1e3) / 1e12,
        "roofline": {"bound": bound,
                     "kernel": kern,
                     "achieved": tf if bound == "tensor" else gbs,
                     "peak": dgemm_tflops if bound == "tensor" else peak,
                     "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                     "frac": tf / dgemm_tflops if bound == "tensor" else gbs / peak,
                     "frac_fp64": tf / dgemm_tflops, "frac_hbm": gbs / peak,
                     "peak_tflops_source": "cuBLAS DGEMM 8192^3 measured in this run (best of 5)",
                     "peak_gbs_source": peak_src, "traffic": None,
                     "ms_per_launch": {"stage1+in": stage_ms[0] / max(ncalls, 1), "stage2": stage_ms[1] / max(ncalls, 1),
                                       "stage3+out": stage_ms[2] / max(ncalls, 1)}},
        "gpu_launches": steps * npanels * (plan.launches_per_matvec + 2), "clocks": clocks,
        "check_sampled_dense_entries_relerr": err, "api": "hm_matmat_device",
    }
    del X, Y
    return out


def run_matmat(args, torch, plan, st, px, py, dev, t_asm):
    """--nrhs mode: the product as the headline line (BASELINE configs[2])."""
    n, nrhs = args.n, args.nrhs
    o = measure_matmat(torch, plan, st, px, py, dev, nrhs, args.steps, args.warmup)
    line = {
        "metric": "H-matmat columns/s", "value": o["value"], "unit": "columns/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": o["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, args.dist).replace("single-vector mul!", f"{nrhs} right-hand sides"),
                   "n": n, "dist": args.dist, "nrhs": nrhs, "panel_width": o["panel_width"],
                   "assembly_s": round(t_asm, 3), "l2": "inputs larger than L2"},
        "tflops": o["tflops"], "effective_gbs": o["effective_gbs"], "roofline": o["roofline"],
        "gpu_launches": o["gpu_launches"], "clocks": o["clocks"],
        "check_sampled_dense_entries_relerr": o["check_sampled_dense_entries_relerr"],
    }
    print(json.dumps(line), flush=True)


def measure_cfg5(hm, torch, dev, local, dist_name, steps):
    """BASELINE configs[4]: KernelMatrix assembly of a 2^22-point Cauchy operator on the GPU
    (hm_assemble_kernel: host builds the range tree, device fills U, V, F and the dense tiles)
    followed by a 16-column product."""
    n = 1 << 22
    free, _total = torch.cuda.mem_get_info(dev)
    if free < 80e9:
        return {"skipped": f"needs ~70 GB of free device memory, {free / 1e9:.0f} GB available"}
    px, py = points(hm, n, dist_name)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=local)
    torch.cuda.synchronize()
    t_asm = time.perf_counter() - t0
    plan = K.plan()
    st = plan.stats()
    o = measure_matmat(torch, plan, st, px, py, dev, 16, steps, 3)
    o["workload"] = (f"on-GPU KernelMatrix assembly, N={n} ({dist_name}), then a 16-column product "
                     f"(BASELINE configs[4])")
    o["assembly_s"] = round(t_asm, 3)
    o["assembly_note"] = "wall clock of hm_assemble_kernel: host range tree + planner + device fill kernels"
    o["assembled_bytes"] = st["stored_bytes"]
    o["leaves"] = {"dense": st["n_dense"], "bary2d": st["n_bary2d"]}
    # the assembled operator also as a single-vector matvec (its HBM roofline)
    x1 = torch.from_numpy(np.random.default_rng(5).standard_normal(n)).to(dev)
    y1 = torch.zeros(n, dtype=torch.float64, device=dev)
    sm = torch.cuda.current_stream()
    for _ in range(3):
        plan.matvec_device(x1.data_ptr(), y1.data_ptr(), False, sm.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sm)
    for _ in range(steps):
        plan.matvec_device(x1.data_ptr(), y1.data_ptr(), False, sm.cuda_stream)
    e1.record(sm)
    torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1) / steps
    peak, _ = measured_peak()
    o["matvec"] = {"ms_per_step": ms1, "value": 1e3 / ms1, "unit": "matvecs/s",
                   "effective_gbs": st["algorithmic_bytes"] / ms1 / 1e6,
                   "frac_hbm": st["algorithmic_bytes"] / ms1 / 1e6 / peak,
                   "check_sampled_dense_rows_relerr": sampled_rows_check(px, py, x1.cpu().numpy(), y1.cpu().numpy(), 12)}
    del K, plan
    torch.cuda.empty_cache()
    # the same configuration on a matrix-free plan: nothing to assemble but the cores and the box tree
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Kf = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=local, matrix_free=True)
    pf = Kf.plan()
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    of = measure_matmat(torch, pf, pf.stats(), px, py, dev, 16, steps, 3)
    of["setup_s"] = round(t_setup, 3)
    of["resident_bytes"] = pf.stats()["stored_bytes"]
    for _ in range(3):
        pf.matvec_device(x1.data_ptr(), y1.data_ptr(), False, sm.cuda_stream)
    torch.cuda.synchronize()
    e0.record(sm)
    for _ in range(steps):
        pf.matvec_device(x1.data_ptr(), y1.data_ptr(), False, sm.cuda_stream)
    e1.record(sm)
    torch.cuda.synchronize()
    msf = e0.elapsed_time(e1) / steps
    of["matvec"] = {"ms_per_step": msf, "value": 1e3 / msf, "unit": "matvecs/s",
                    "check_sampled_dense_rows_relerr": sampled_rows_check(px, py, x1.cpu().numpy(), y1.cpu().numpy(), 12)}
    o["matrix_free"] = of
    del Kf, pf, x1, y1
    torch.cuda.empty_cache()
    return o


# --------------------------------------------------------------------------- multi-GPU pipeline (product API)
class DistLoop:
    """K steps of y = K x over all ranks through the product's hm_dist_* entry points.
    independent: every step gets a fresh x, replicated by the NCCL broadcast (hm_dist_bcast_x) on
        a second stream while the previous step computes.
    dependent: x_{k+1} = y_k (a Krylov-style iteration): y is already replicated by the fused
        all-gather, so no broadcast at all; chains of `chain` steps restart from the broadcast x."""

    def __init__(self, torch, plan, dev, rank, v):
        self.torch, self.plan, self.rank = torch, plan, rank
        (self.x0, self.x1), (self.y0, self.y1) = plan.dist_buffers()
        self.xs, self.ys = (self.x0, self.x1), (self.y0, self.y1)
        self.xd = torch.from_numpy(v).to(dev) if rank == 0 else None
        self.cs = torch.cuda.Stream(device=dev)
        self.x_ready = [torch.cuda.Event() for _ in range(2)]
        self.x_free = [torch.cuda.Event() for _ in range(2)]

    def _bcast(self, slot):
        torch = self.torch
        with torch.cuda.stream(self.cs):
            self.cs.wait_event(self.x_free[slot])  # the matvec that last read this x slot is done
            self.plan.dist_bcast_x(self.xd.data_ptr() if self.rank == 0 else 0, 0, slot, self.cs.cuda_stream)
            self.x_ready[slot].record(self.cs)

    def independent(self, nsteps):
        main = self.torch.cuda.current_stream()
        for b in range(2):
            self.x_free[b].record(main)
        self._bcast(0)
        for k in range(nsteps):
            b = k & 1
            main.wait_event(self.x_ready[b])
            self.plan.dist_matvec_device(self.xs[b], b, False, main.cuda_stream)
            self.x_free[b].record(main)
            if k + 1 < nsteps:
                self._bcast((k + 1) & 1)
        main.wait_stream(self.cs)
        return self.ys[(nsteps - 1) & 1]

    def independent_push(self, nsteps):
        """Same as `independent`, but x is replicated by the root's copy engines (hm_dist_push_x:
        peer cudaMemcpyAsync over NVLink, joined into the barrier of the step it overlaps) instead
        of an NCCL kernel: nothing but our own kernels occupies the SMs."""
        main = self.torch.cuda.current_stream()
        xp = self.xd.data_ptr() if self.rank == 0 else 0
        self.plan.dist_push_x(xp, 0, 0, main.cuda_stream)
        self.plan.dist_barrier(main.cuda_stream)
        for k in range(nsteps):
            if k + 1 < nsteps:
                self.plan.dist_push_x(xp, 0, (k + 1) & 1, main.cuda_stream)
            self.plan.dist_matvec_device(self.xs[k & 1], k & 1, False, main.cuda_stream)
        return self.ys[(nsteps - 1) & 1]

    def dependent(self, nsteps, chain=8):
        main = self.torch.cuda.current_stream()
        self.plan.dist_bcast_x(self.xd.data_ptr() if self.rank == 0 else 0, 0, 0, main.cuda_stream)
        for k in range(nsteps):
            j = k % chain
            src = self.xs[0] if j == 0 else self.ys[(j - 1) & 1]
            self.plan.dist_matvec_device(src, j & 1, False, main.cuda_stream)
        return self.ys[((nsteps - 1) % chain) & 1]


def timed_graph(torch, fn, barrier, stream):
    """Capture fn() once, replay it untimed (graph upload, NCCL warm-up), then time one replay."""
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ret = fn()
    barrier()
    g.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    g.replay()
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1), t_wall0, g, ret


def measure_cfg4(hm, torch, dist, dev, rank, world, local, steps):
    """BASELINE configs[3]: N = 2^24 uniform points, stored operator (289 GB) partitioned by block
    rows over the GPUs of the box, x replicated by the NCCL broadcast, y gathered by the fused
    stores.  Needs >= 2 GPUs to fit."""
    n = 1 << 24
    px, py = points(hm, n, "unif")
    whole = hm.KernelMatrix.layout_stats(px, py, 1.0, -1.0, 1.0, -1.0, rank, world)
    free, _total = torch.cuda.mem_get_info(dev)
    ok = torch.tensor([1.0 if whole["stored_bytes"] + 3e9 < free else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0:
        return {"skipped": f"part of {whole['stored_bytes'] / 1e9:.0f} GB does not fit {free / 1e9:.0f} GB free"}
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    K = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=local, part=rank, nparts=world)
    torch.cuda.synchronize()
    dist.barrier()
    t_asm = time.perf_counter() - t0
    plan = K.plan()
    st = plan.stats()
    plan.dist_init_torch()
    v = np.random.default_rng(0).standard_normal(n)
    loop = DistLoop(torch, plan, dev, rank, v)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    loop.independent(3)
    barrier()
    ms, t_wall0, g, ydev = timed_graph(torch, lambda: loop.independent(steps), barrier, torch.cuda.current_stream())
    plan.timing_begin(steps)
    loop.independent(steps)
    barrier()
    stage_ms, ncalls = plan.timing_end()
    clocks = sampler.stop(t_wall0) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    pb = torch.tensor([float(st["part_algorithmic_bytes"])], dtype=torch.float64, device=dev)
    dist.all_reduce(pb, op=dist.ReduceOp.MAX)
    plan.dist_check()
    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        yh = device_view(torch, ydev, n, dev).cpu().numpy()
        gbs = st["algorithmic_bytes"] / ms / 1e6
        s1, s2, s3 = (m / max(ncalls, 1) for m in stage_ms)
        out = {"workload": f"Cauchy KernelMatrix N={n} uniform interlaced points, stored operator "
                           f"({st['algorithmic_bytes'] / 1e9:.0f} GB) block-row x{world}, NCCL x-broadcast, fused y gather "
                           f"(BASELINE configs[3])",
               "value": 1e3 / ms, "unit": "matvecs/s", "ms_per_step": ms, "steps": steps, "n_gpus": world,
               "assembly_s": round(t_asm, 3), "effective_gbs": gbs,
               "roofline": {"bound": "hbm", "kernel": "hm_stream_kernel (stage 1 + stage 3), all ranks",
                            "achieved": gbs, "peak": peak * world, "unit": "GB/s", "frac": gbs / (peak * world),
                            "peak_source": f"{world} x {peak_src}", "traffic": None,
                            "ms_per_launch_rank0": {"stage1": s1, "stage2": s2, "stage3": s3}},
               "part_bytes_max_over_mean": float(pb.item()) * world / (st["algorithmic_bytes"] + 8 * n * (world - 1)),
               "check_sampled_dense_rows_relerr": sampled_rows_check(px, py, v, yh, 12),
               "clocks": clocks, "api": "hm_assemble_kernel(part) + hm_dist_init + hm_dist_bcast_x + hm_dist_matvec_device"}
    g.reset()
    del loop, K, plan
    torch.cuda.empty_cache()
    barrier()
    return out


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
    dist = None
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
        return run_matmat(args, torch, plan, st, px, py, dev, t_asm)
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
    stream = torch.cuda.current_stream()

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dependent = comm = x_push = nccl_loop = None
    x_exchange = None
    graph = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not dist_on:
        x_dev = torch.from_numpy(v).to(dev)
        y_dev = torch.zeros(n, dtype=torch.float64, device=dev)

        def run(k):
            for _ in range(k):
                plan.matvec_device(x_dev.data_ptr(), y_dev.data_ptr(), accumulate=False, stream=stream.cuda_stream)

        run(max(args.warmup, 3))
        barrier()
        plan.timing_begin(args.steps)
        t_wall0 = time.time()
        e0.record(stream)
        run(args.steps)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        stage_ms, ncalls = plan.timing_end()
    else:
        # the exchange is the product's: communicator, peer-mapped y / x buffers and barrier belong to the plan
        plan.dist_init_torch()
        loop = DistLoop(torch, plan, dev, rank, v)
        loop.independent(max(args.warmup, 3))
        barrier()
        if args.no_graph:
            t_wall0 = time.time()
            e0.record(stream)
            yptr = loop.independent(args.steps)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
        else:
            # the whole K-step pipeline (our kernels, barrier kernels, the NCCL broadcasts on the second
            # stream) as ONE CUDA graph: at 0.3 ms of GPU work per step the launch path would dominate
            ms, t_wall0, graph, yptr = timed_graph(torch, lambda: loop.independent(args.steps), barrier, stream)
        # per-stage kernel times for the roofline: the same steps launched eagerly
        plan.timing_begin(args.steps)
        loop.independent(args.steps)
        barrier()
        stage_ms, ncalls = plan.timing_end()
        y_dev = device_view(torch, yptr, n, dev)
        y_indep = y_dev.clone()
        # the same loop with x replicated by the copy engines instead of the NCCL kernel
        loop.independent_push(3)
        barrier()
        ms_push, _, gpush, ypush = timed_graph(torch, lambda: loop.independent_push(args.steps), barrier, stream)
        push_rel = float((device_view(torch, ypush, n, dev) - y_indep).abs().max() / y_indep.abs().max())
        gpush.reset()
        # dependent iteration x_{k+1} = y_k: no broadcast on the critical path at all
        loop.dependent(8)
        barrier()
        ms_dep, _, gdep, _ = timed_graph(torch, lambda: loop.dependent(args.steps), barrier, stream)
        gdep.reset()
        # the exchange pieces alone
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for _ in range(5):
            plan.dist_barrier(stream.cuda_stream)
            plan.dist_bcast_x(loop.xd.data_ptr() if rank == 0 else 0, 0, 0, stream.cuda_stream)
        barrier()
        ev[0].record(stream)
        for _ in range(50):
            plan.dist_barrier(stream.cuda_stream)
        ev[1].record(stream)
        for _ in range(50):
            plan.dist_bcast_x(loop.xd.data_ptr() if rank == 0 else 0, 0, 0, stream.cuda_stream)
        ev[2].record(stream)
        barrier()
        tt = torch.tensor([ms_dep, ev[0].elapsed_time(ev[1]) / 50, ev[1].elapsed_time(ev[2]) / 50, ms_push],
                          dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dep, t_bar, t_bc, ms_push = (float(a) for a in tt.tolist())
        x_push = {"value": args.steps / (ms_push / 1e3), "unit": "matvecs/s", "ms_per_step": ms_push / args.steps,
                  "relinf_vs_nccl_loop": push_rel,
                  "what": "independent right-hand sides with x replicated by hm_dist_push_x (the root's copy engines "
                          "write x(k+1) into every rank's buffer over NVLink during step k; joined into step k's barrier) "
                          "instead of the NCCL broadcast kernel"}
        dependent = {"value": args.steps / (ms_dep / 1e3), "unit": "matvecs/s", "ms_per_step": ms_dep / args.steps,
                     "what": "x_{k+1} = y_k (chains of 8 from the broadcast x): y is already on every rank after the "
                             "fused gather + barrier, so the step has no broadcast"}
        comm = {"barrier_kernel_ms": t_bar, "nccl_bcast_x_ms": t_bc,
                "note": "back-to-back on one stream, max over ranks; in the step the broadcast of x(k+1) runs on a "
                        "second stream under step k, the barrier is in line after stage 3"}
        plan.dist_check()
        y_dev = y_indep
        # headline = the faster of the two ways the product replicates a fresh x every step
        nccl_loop = {"value": None, "unit": "matvecs/s", "ms_per_step": None,
                     "what": "independent right-hand sides with x replicated by hm_dist_bcast_x (ncclBroadcast on a "
                             "second stream under the previous step)"}
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        nccl_loop["value"], nccl_loop["ms_per_step"] = args.steps / (ms / 1e3), ms / args.steps
        x_exchange = "nccl"
        if ms_push < ms:
            ms, x_exchange = ms_push, "copy_engines"
    clocks = sampler.stop(t_wall0) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    pbytes = torch.tensor([float(st["part_algorithmic_bytes"])], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(pbytes, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps / (ms / 1e3)

    # per-stage roofline inputs of this rank
    b1 = 8 * st["part_v_words"] + 8 * st["ncols"]
    b2 = 8 * st["part_core_words"]
    b3 = 8 * (st["part_u_words"] + st["part_dense_words"]) + 8 * (r1 - r0)
    s1, s2, s3 = (m / max(ncalls, 1) for m in stage_ms)

    # ---- end to end through the host-pointer C ABI call (pinned host buffers) ----
    e2e = None
    xn = yn = None
    if not args.no_e2e:
        if rank == 0:
            xh = torch.from_numpy(v.copy()).pin_memory()
            yh = torch.zeros(n, dtype=torch.float64).pin_memory()
            xn, yn = xh.numpy(), yh.numpy()

        def call():
            if dist_on:   # x is read on rank 0 only, the whole y lands in rank 0's host vector
                plan.dist_matvec(xn, yn, root=0, accumulate=False)
            else:
                plan.matvec(xn, yn, accumulate=False)  # H2D x, 3 stages, D2H y, sync

        for _ in range(3):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": args.steps / dt, "unit": "matvecs/s", "h2d_bytes_per_step": 8 * st["ncols"],
               "d2h_bytes_per_step": 8 * st["nrows"], "ms_per_step": dt / args.steps * 1e3,
               "api": ("hm_dist_matvec (C ABI, host pointers: x from rank 0's host memory, NCCL broadcast, whole y "
                       "back into rank 0's host memory)" if dist_on else "hm_matvec (C ABI, host pointers)")}
        if rank == 0 and dist_on:
            e2e["relinf_vs_device_loop"] = float(np.max(np.abs(yn - y_dev.cpu().numpy())) / np.max(np.abs(yn)))

    # ---- secondary objects, single GPU ----
    mfree = matmat64 = cfg5 = None
    if world == 1 and not args.matrix_free and not args.no_matrix_free:
        # the same operator applied matrix-free (hm_assemble_kernel_free): nothing but the r x r cores is
        # resident, the entries are evaluated inside the matvec (FP64-bound).  Reported, not the headline.
        t0 = time.perf_counter()
        Kf = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=local, matrix_free=True)
        pf = Kf.plan()
        t_setup = time.perf_counter() - t0
        yf = torch.zeros(n, dtype=torch.float64, device=dev)
        for _ in range(5):
            pf.matvec_device(x_dev.data_ptr(), yf.data_ptr(), accumulate=False, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nmf = max(10, min(args.steps, 50))
        pf.timing_begin(nmf)
        f0.record(stream)
        for _ in range(nmf):
            pf.matvec_device(x_dev.data_ptr(), yf.data_ptr(), accumulate=False, stream=stream.cuda_stream)
        f1.record(stream)
        torch.cuda.synchronize()
        ms_f = f0.elapsed_time(f1) / nmf
        fst, fn = pf.timing_end()
        dev_rel = float((yf - y_dev).abs().max() / y_dev.abs().max())
        # FP64 work of one matrix-free matvec, in lane-level FP64-pipe instructions.
        # Chebyshev form leaf by leaf: per (column, leaf) of stage 1 one mapping + 18 recurrence DFMA + 20
        # accumulations = 39, per (row, leaf) of stage 3 a 20-term Clenshaw sum = 41, per dense entry 5 (sub,
        # 3 refinement DFMA of the reciprocal, 1 accumulate; the rcp.approx itself runs on the XU pipe).
        # Nested-basis form: the same 39 / 41 once per column / row (not per leaf), 400 per leaf for its
        # core, 5 per dense entry; the box-to-box translations (~3 % more) are not counted.
        # peak = the DFMA issue rate measured by profiles/microbench/fp64_pipes.cu (34.2 TFLOP/s / 2)
        fform = pf.form
        if fform == 3:
            issues = 39.0 * st["ncols"] + 41.0 * st["nrows"] + 400.0 * st["n_bary2d"] + 5.0 * st["part_dense_words"]
            fkern = "hm_nest_base/up/core/down kernels + hm_nest_dense_kernel (dominant: the dense leaves)"
            fhow = ("algorithmic FP64-pipe instructions of the nested-basis form (39 per column, 41 per row, 400 per "
                    "leaf core, 5 per dense entry) / time")
            fname = ("nested basis: moments at the finest column boxes, translated up the dyadic box tree; shared "
                     "cores; coefficients translated down and evaluated once per row; dense leaves evaluated on the fly")
        else:
            issues = 39.0 * st["part_v_words"] / 20 + 41.0 * st["part_u_words"] / 20 + 5.0 * st["part_dense_words"]
            fkern = "hm_free1_kernel + hm_free3_kernel"
            fhow = ("algorithmic FP64-pipe instructions of the Chebyshev form (39 per column and leaf, 41 per row and "
                    "leaf, 5 per dense entry) / time")
            fname = "Chebyshev series (moments + Clenshaw), cores C F C' with the node correction"
        mfree = {"value": 1e3 / ms_f, "unit": "matvecs/s", "ms_per_step": ms_f, "steps": nmf,
                 "resident_bytes": pf.stats()["stored_bytes"], "setup_s": round(t_setup, 3),
                 "relinf_vs_stored": dev_rel, "api": "hm_assemble_kernel_free + hm_matvec_device",
                 "form": fname, "launches_per_matvec": pf.launches_per_matvec,
                 "ms_per_launch": {"stage1": fst[0] / max(fn, 1), "stage2": fst[1] / max(fn, 1), "stage3": fst[2] / max(fn, 1)},
                 "stored_form_equivalent_gbs": st["algorithmic_bytes"] / ms_f / 1e6,
                 "roofline": {"bound": "fp64", "kernel": fkern,
                              "achieved": issues / (ms_f / 1e3) / 1e12, "peak": 17.1,
                              "unit": "10^12 FP64-pipe instructions/s (lane-level)",
                              "frac": issues / (ms_f / 1e3) / 1e12 / 17.1,
                              "how": fhow + "; peak = DFMA rate of profiles/microbench/fp64_pipes.cu (34.2 TFLOP/s / 2)"}}
        if not args.no_e2e:
            for _ in range(3):
                pf.matvec(xn, yn, accumulate=False)
            t0 = time.perf_counter()
            for _ in range(nmf):
                pf.matvec(xn, yn, accumulate=False)
            torch.cuda.synchronize()
            mfree["e2e"] = {"value": nmf / (time.perf_counter() - t0), "unit": "matvecs/s",
                            "api": "hm_matvec (C ABI, host pointers)"}
        if not args.no_secondary:
            # BASELINE configs[2] on the matrix-free plan
            mm = measure_matmat(torch, pf, pf.stats(), px, py, dev, 64, 10, 3)
            mm["workload"] = workload_name(n, args.dist).replace("single-vector mul!", "64 right-hand sides") + \
                " (BASELINE configs[2]), matrix-free plan"
            mfree["matmat64"] = mm
        del Kf, pf, yf
    if world == 1 and not args.matrix_free and not args.no_secondary:
        # BASELINE configs[2]: the same operator applied to 64 right-hand sides
        matmat64 = measure_matmat(torch, plan, st, px, py, dev, 64, 10, 3)
        matmat64["workload"] = workload_name(n, args.dist).replace("single-vector mul!", "64 right-hand sides") + \
            " (BASELINE configs[2])"

    y_host = y_dev.cpu().numpy() if rank == 0 else None
    sampled = sampled_rows_check(px, py, v, y_host, 48 if n <= (1 << 21) else 12) if rank == 0 else None

    line = None
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
                       "collectives": (("product API hm_dist_*: a fresh x replicated from rank 0 every step "
                                        + ("by the root's copy engines over the NVLink peer mappings (hm_dist_push_x), under the "
                                           "previous step" if x_exchange == "copy_engines" else
                                           "by ncclBroadcast (hm_dist_bcast_x) on a second stream under the previous step")
                                        + "; all-gather(y) fused into stage 3 (stores into every rank's peer-mapped buffer over "
                                          "NVLink) + barrier kernel"
                                        + ("" if args.no_graph else "; whole loop replayed as one CUDA graph"))
                                       if dist_on else "none"),
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
            "e2e": e2e, "matrix_free": mfree, "matmat64": matmat64,
            "gpu_launches": (plan.launches_per_matvec + (1 if dist_on else 0)) * args.steps, "clocks": clocks,
            "leaves": {"dense": st["n_dense"], "bary2d": st["n_bary2d"]},
            "check_sampled_dense_rows_relerr": sampled,
        }
        if dist_on:
            line["dependent_iteration"] = dependent
            line["x_push_copy_engines"] = x_push
            line["x_bcast_nccl"] = nccl_loop
            line["comm"] = comm
            line["partition_balance"] = {"max_part_bytes": float(pbytes.item()),
                                         "ideal_part_bytes": st["algorithmic_bytes"] / world,
                                         "max_over_ideal": float(pbytes.item()) * world / st["algorithmic_bytes"]}
    if graph is not None:
        graph.reset()
    # release the headline operator before the large secondary configurations
    del K, plan
    if dist_on:
        del loop
    torch.cuda.empty_cache()
    if world == 1 and not args.matrix_free and not args.no_secondary:
        cfg5 = measure_cfg5(hm, torch, dev, local, args.dist, 10)
        line["cfg5_assemble_2p22_matmat16"] = cfg5
    if dist_on and not args.no_secondary:
        cfg4 = measure_cfg4(hm, torch, dist, dev, rank, world, local, 10)
        if rank == 0:
            line["cfg4_2p24_multi_gpu"] = cfg4
    if rank == 0:
        if not args.no_cpu_baseline and host_mem_ok(n):
            # the oracle on the host cores: baseline at N = 1, and the full-vector parity check at every N
            cb, parity = cpu_baseline(n, args.dist, v, y_host)
            if world == 1:
                line["cpu_baseline"] = cb
            line["parity_relinf_vs_oracle"] = parity
        print(json.dumps(line), flush=True)
    if dist_on:
        # leave without the communicator destructors (NCCL teardown order at exit can hang)
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
