"""Worker of tests/test_gpu_dist.py: one process per GPU (torchrun), rank r holds block-row part r.
Checks the product's multi-GPU mul! (hm_dist_*: NCCL broadcast of x, all-gather of y fused into
stage 3 through peer memory, barrier kernel) against the CPU oracle on every rank."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOL = 1e-12


def relinf(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    import hmb200_loader
    hm = hmb200_loader.load()
    from oracle import oracle as O
    from helpers import device_view
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("HM_TEST_N", "20000"))
    matrix_free = os.environ.get("HM_TEST_FREE", "0") == "1"
    x, y, (a, b, c, d) = O.example_points(n, "cheb")
    K = hm.KernelMatrix(hm.cauchykernel, x, y, a, b, c, d, device=local, part=rank, nparts=world,
                        matrix_free=matrix_free)
    plan = K.plan()
    plan.dist_init_torch()
    Kref = O.kernelmatrix(O.CAUCHY, x, y, a, b, c, d)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(n)
    ref = Kref.matvec(v)
    out = {}

    # host-pointer call: x only on the root, whole y on every rank
    yh = np.zeros(n)
    plan.dist_matvec(v if rank == 0 else None, yh, root=0, accumulate=False)
    out["host"] = relinf(yh, ref)
    ya = np.ones(n)
    plan.dist_matvec(v if rank == 0 else None, ya, root=0, accumulate=True)
    out["host_acc"] = relinf(ya, ref + 1.0)
    # a second root, strided arguments
    xs = np.zeros(2 * n)
    xs[::2] = v
    ys = np.zeros(3 * n)
    plan.dist_matvec(xs if rank == world - 1 else None, ys, root=world - 1, accumulate=False, incx=2, incy=3)
    out["host_strided"] = relinf(ys[::3], ref)

    # device path: broadcast, matvec, dependent matvec (x_{k+1} = y_k, no broadcast)
    (x0, x1), (y0, y1) = plan.dist_buffers()
    st = torch.cuda.current_stream()
    xd = torch.from_numpy(v).to(dev) if rank == 0 else None
    plan.dist_bcast_x(xd.data_ptr() if rank == 0 else 0, root=0, slot=0, stream=st.cuda_stream)
    plan.dist_matvec_device(x0, yslot=0, accumulate=False, stream=st.cuda_stream)
    plan.dist_matvec_device(y0, yslot=1, accumulate=False, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ref2 = Kref.matvec(ref)
    out["dev"] = relinf(device_view(y0, n, local).cpu().numpy(), ref)
    out["dev_dependent"] = relinf(device_view(y1, n, local).cpu().numpy(), ref2)

    # x replicated by the root's copy engines instead of the NCCL broadcast (hm_dist_push_x): complete
    # once the next barrier has passed
    xd2 = torch.from_numpy(2.0 * v).to(dev) if rank == 0 else None
    plan.dist_push_x(xd2.data_ptr() if rank == 0 else 0, root=0, slot=1, stream=st.cuda_stream)
    plan.dist_barrier(st.cuda_stream)
    plan.dist_push_x(xd.data_ptr() if rank == 0 else 0, root=0, slot=0, stream=st.cuda_stream)  # joins the next barrier
    plan.dist_matvec_device(x1, yslot=0, accumulate=False, stream=st.cuda_stream)
    plan.dist_matvec_device(x0, yslot=1, accumulate=False, stream=st.cuda_stream)
    torch.cuda.synchronize()
    out["push"] = max(relinf(device_view(y0, n, local).cpu().numpy(), 2.0 * ref),
                      relinf(device_view(y1, n, local).cpu().numpy(), ref))

    # the same sequence captured once and replayed (the barrier epoch lives on the device)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            s = torch.cuda.current_stream().cuda_stream
            plan.dist_bcast_x(xd.data_ptr() if rank == 0 else 0, root=0, slot=1, stream=s)
            plan.dist_matvec_device(x1, yslot=0, accumulate=False, stream=s)
            plan.dist_matvec_device(y0, yslot=1, accumulate=False, stream=s)
    for _ in range(3):
        device_view(y0, n, local).zero_()
        device_view(y1, n, local).zero_()
        dist.barrier()
        g.replay()
        torch.cuda.synchronize()
    out["graph"] = max(relinf(device_view(y0, n, local).cpu().numpy(), ref),
                       relinf(device_view(y1, n, local).cpu().numpy(), ref2))
    plan.dist_check()

    # all ranks hold bit-identical results
    t = device_view(y1, n, local).clone()
    lst = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    out["identical"] = all(bool(torch.equal(lst[0], u)) for u in lst)
    res = [None] * world
    dist.all_gather_object(res, out)
    if rank == 0:
        print("HM_DIST_RESULT " + json.dumps(res), flush=True)
    del g
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
