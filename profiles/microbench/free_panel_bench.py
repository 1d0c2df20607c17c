"""Matrix-free many-RHS product (hm_free_panel.cu) at N = 2^20 (or argv[1]): ms per product and per
stage for 16 / 32 / 64 columns (argv[3] = comma list), beside the stored plan's panel kernels (argv[2] = 'stored' adds it)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import hmb200_loader
hm = hmb200_loader.load()

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
stored = len(sys.argv) > 2 and sys.argv[2] == "stored"
dev = torch.device("cuda:0")
px, py = hm.chebyshevpoints(n), hm.chebyshevpoints(n, 2)
plans = [("free", hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=0, matrix_free=True))]
if stored:
    plans.append(("stored", hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=0)))
stream = torch.cuda.current_stream()
for name, K in plans:
    plan = K.plan()
    st = plan.stats()
    words = st["dense_words"] + st["lowrank_words"]
    for nrhs in ([int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else (16, 32, 64)):
        X = torch.from_numpy(np.random.default_rng(0).standard_normal((nrhs, n))).to(dev)
        Y = torch.zeros((nrhs, n), dtype=torch.float64, device=dev)
        for _ in range(3):
            plan.matmat_device(X.data_ptr(), n, Y.data_ptr(), n, nrhs, accumulate=False, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        steps = 10
        plan.timing_begin(steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            plan.matmat_device(X.data_ptr(), n, Y.data_ptr(), n, nrhs, accumulate=False, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        sm, nc = plan.timing_end()
        # one column against the plan's own matvec
        y1 = torch.zeros(n, dtype=torch.float64, device=dev)
        plan.matvec_device(X[nrhs - 1].data_ptr(), y1.data_ptr(), False, stream.cuda_stream)
        torch.cuda.synchronize()
        err = float((Y[nrhs - 1] - y1).abs().max() / y1.abs().max())
        print(f"{name} n={n} nrhs={nrhs}: {ms:.3f} ms  ({2.0 * words * nrhs / ms / 1e9:.2f} TFLOP/s)  stages "
              f"{sm[0] / nc:.3f} / {sm[1] / nc:.3f} / {sm[2] / nc:.3f}  col-vs-matvec {err:.2e}", flush=True)
        del X, Y
