"""One block-row part of the N = 2^20 operator on ONE GPU: what a rank of an 8-GPU run executes
per step, without the exchange.  Used to tune launch granularity (HMB200_RMAX / HMB200_CMAX0 ...)
and to compare against the ideal (whole-operator time / nparts).

  python profiles/microbench/part_bench.py --nparts 8 --part 0 --steps 200
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hmb200_loader  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nparts", type=int, default=8)
ap.add_argument("--part", type=int, default=0)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--free", action="store_true")
a = ap.parse_args()
hm = hmb200_loader.load()
n = a.n
px, py = hm.chebyshevpoints(n, 1), hm.chebyshevpoints(n, 2)
K = hm.KernelMatrix(hm.cauchykernel, px, py, 1.0, -1.0, 1.0, -1.0, device=0, part=a.part, nparts=a.nparts,
                    matrix_free=a.free)
plan = K.plan()
st = plan.stats()
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.zeros(n, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream()
for _ in range(5):
    plan.matvec_device(x.data_ptr(), y.data_ptr(), False, s.cuda_stream)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(a.steps):
        plan.matvec_device(x.data_ptr(), y.data_ptr(), False, torch.cuda.current_stream().cuda_stream)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
plan.timing_begin(a.steps)
for _ in range(a.steps):
    plan.matvec_device(x.data_ptr(), y.data_ptr(), False, s.cuda_stream)
torch.cuda.synchronize()
stage, nc = plan.timing_end()
b = st["part_algorithmic_bytes"]
print(json.dumps({"part": a.part, "nparts": a.nparts, "ms_graph": ms, "stage_ms": [v / nc for v in stage],
                  "part_GB": b / 1e9, "GBs": b / ms / 1e6, "items1": st["n_stage1_items"], "items3": st["n_stage3_items"],
                  "cores": st["n_stage2_blocks"], "launches": plan.launches_per_matvec,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("HMB200_")}}))
