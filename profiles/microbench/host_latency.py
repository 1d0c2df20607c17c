import sys, time, os, numpy as np
sys.path.insert(0, os.getcwd())
import hmb200_loader; hm = hmb200_loader.load()
import torch
for N in (4096, 16384, 65536, 262144):
    x, y = hm.chebyshevpoints(N), hm.chebyshevpoints(N, kind=2)
    K = hm.KernelMatrix(hm.cauchykernel, x, y, 1.0, -1.0, 1.0, -1.0)
    P = K.plan()
    v = torch.randn(N, dtype=torch.float64).pin_memory().numpy()
    u = torch.zeros(N, dtype=torch.float64).pin_memory().numpy()
    for _ in range(20): P.matvec(v, u, accumulate=False)
    t0 = time.perf_counter()
    for _ in range(200): P.matvec(v, u, accumulate=False)
    print(N, os.environ.get("HMB200_NO_COPY_PIPELINE", "pipe"), (time.perf_counter() - t0) / 200 * 1e6, "us")
