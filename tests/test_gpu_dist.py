"""Multi-GPU mul! of the product (include/hmb200.h, hm_dist_*).  The two-rank test needs two GPUs
(`gpurun --gpus 2`); the single-rank test drives the same entry points -- communicator, barrier
kernel, slots -- on one GPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import TOL, relinf, device_view

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run_workers(world, env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    env.setdefault("NCCL_DEBUG", "WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("HM_DIST_RESULT ")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    return json.loads(line[-1][len("HM_DIST_RESULT "):])


@pytest.mark.gpu
@pytest.mark.parametrize("free", ["0", "1"])
def test_two_rank_mul_matches_oracle(free):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    world = min(_ngpu(), 4)
    res = _run_workers(world, {"HM_TEST_FREE": free})
    assert len(res) == world
    for r in res:
        for k in ("host", "host_acc", "host_strided", "dev", "dev_dependent", "push", "graph"):
            assert r[k] <= TOL, (k, r)
        assert r["identical"]


@pytest.mark.gpu
def test_single_rank_dist_entry_points(hm, O):
    """nranks = 1: hm_dist_init (NCCL communicator of one), broadcast, matvec + barrier kernel,
    host-pointer call, error paths."""
    import torch
    n = 6000
    x, y, (a, b, c, d) = O.example_points(n, "cheb")
    K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=0)
    plan = K.plan()
    with pytest.raises(hm.HmError) as ei:
        plan.dist_matvec_device(0, 0)
    assert ei.value.status == 5  # HM_ERR_STATE: not initialised
    uid = hm.dist_unique_id()
    assert len(uid) == 128
    with pytest.raises(hm.HmError):
        plan.dist_init(uid, 2, 0)  # the plan is part 0 of 1
    plan.dist_init(uid, 1, 0)
    with pytest.raises(hm.HmError):
        plan.dist_init(uid, 1, 0)  # twice
    ref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d).matvec
    v = np.random.default_rng(3).standard_normal(n)
    out = np.zeros(n)
    plan.dist_matvec(v, out, root=0, accumulate=False)
    assert relinf(out, ref(v)) <= TOL
    (x0, x1), (y0, y1) = plan.dist_buffers()
    vd = torch.from_numpy(v).cuda()
    s = torch.cuda.current_stream().cuda_stream
    plan.dist_bcast_x(vd.data_ptr(), 0, 1, s)
    plan.dist_matvec_device(x1, 1, False, s)
    plan.dist_barrier(s)
    torch.cuda.synchronize()
    assert relinf(device_view(y1, n, 0).cpu().numpy(), ref(v)) <= TOL
    plan.dist_check()
