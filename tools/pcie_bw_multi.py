"""Concurrent pinned-memory PCIe bandwidth of the box with N GPUs copying at once (what caps the contract-complete e2e number
and its multi-GPU scaling).  Run under torchrun (N ranks) or plain python (1 rank); rank 0 prints one JSON line.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/pcie_bw_multi.py
torch is plumbing here (streams, pinned memory, the process group)."""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if dist is not None:
        dist.barrier()


MB = 1 << 20


def measure(h2d, d2h, size_mb, seconds=0.6):
    n = int(size_mb * MB)
    hu, du = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8, device="cuda")
    hd, dd = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8, device="cuda")
    su, sd = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    cnt = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(4):
            if h2d:
                with torch.cuda.stream(su):
                    du.copy_(hu, non_blocking=True)
            if d2h:
                with torch.cuda.stream(sd):
                    hd.copy_(dd, non_blocking=True)
            cnt += 1
        su.synchronize()
        sd.synchronize()
    dt = time.perf_counter() - t0
    gbs = cnt * n / dt / 1e9
    t = torch.tensor([gbs if h2d else 0.0, gbs if d2h else 0.0], dtype=torch.float64, device="cuda")
    tmin = t.clone()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    return {"h2d_total_gbs": round(float(t[0]), 1), "d2h_total_gbs": round(float(t[1]), 1), "h2d_min_rank_gbs": round(float(tmin[0]), 1),
            "d2h_min_rank_gbs": round(float(tmin[1]), 1)}


out = {"gpus": world, "host_cpus": len(os.sched_getaffinity(0)), "cases": {}}
measure(True, True, 8, 0.2)
for size in (64, 5.4):
    out["cases"]["h2d_only_%gMB" % size] = measure(True, False, size)
    out["cases"]["d2h_only_%gMB" % size] = measure(False, True, size)
    out["cases"]["full_duplex_%gMB" % size] = measure(True, True, size)
if rank == 0:
    print(json.dumps(out))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
