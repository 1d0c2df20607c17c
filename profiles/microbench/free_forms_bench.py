"""Matrix-free matvec by form at N = argv[1] (default 2^20), Chebyshev points: nested-basis (default),
per-leaf Chebyshev (HMB200_FREE_FORM=cheb), against the stored plan and sampled dense rows."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import hmb200_loader
hm = hmb200_loader.load()

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dist = sys.argv[2] if len(sys.argv) > 2 else "cheb"
dev = torch.device("cuda:0")
if dist == "cheb":
    px, py = hm.chebyshevpoints(n), hm.chebyshevpoints(n, 2)
else:
    i = np.arange(1, n + 1, dtype=np.float64)
    px, py = 1.0 - 2.0 * (i - 0.5) / n, 1.0 - 2.0 * (i - 0.25) / n
v = np.random.default_rng(0).standard_normal(n)
x = torch.from_numpy(v).to(dev)
stream = torch.cuda.current_stream()
rows = np.unique(np.random.default_rng(1).integers(0, n, 24))
xl, yl, vl = px.astype(np.longdouble), py.astype(np.longdouble), v.astype(np.longdouble)
dense = np.array([np.sum(vl / (xl[i] - yl)) for i in rows], dtype=np.float64)
ref = None
for name, env, free in (("stored", None, False), ("cheb", "cheb", True), ("nested", None, True)):
    if name == "stored" and n > (1 << 21):
        continue
    if os.environ.get("HMB200_ONLY") and os.environ["HMB200_ONLY"] != name:
        continue
    if env:
        os.environ["HMB200_FREE_FORM"] = env
    else:
        os.environ.pop("HMB200_FREE_FORM", None)
    t0 = time.perf_counter()
    K = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=0, matrix_free=free)
    plan = K.plan()
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    y = torch.zeros(n, dtype=torch.float64, device=dev)
    for _ in range(5):
        plan.matvec_device(x.data_ptr(), y.data_ptr(), False, stream.cuda_stream)
    torch.cuda.synchronize()
    steps = 50
    plan.timing_begin(steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        plan.matvec_device(x.data_ptr(), y.data_ptr(), False, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    sm, nc = plan.timing_end()
    yh = y.cpu().numpy()
    if ref is None:
        ref = yh
    err_d = float(np.max(np.abs(yh[rows] - dense)) / np.max(np.abs(dense)))
    err_s = float(np.max(np.abs(yh - ref)) / np.max(np.abs(ref)))
    print(f"{name:7s} form={plan.form} n={n}: {ms:.4f} ms ({1e3 / ms:.0f} matvecs/s)  stages "
          f"{sm[0] / nc:.4f} / {sm[1] / nc:.4f} / {sm[2] / nc:.4f}  setup {t_setup:.2f} s  launches {plan.launches_per_matvec}  "
          f"vs dense rows {err_d:.2e}  vs first {err_s:.2e}", flush=True)
    del K, plan, y
    torch.cuda.empty_cache()
