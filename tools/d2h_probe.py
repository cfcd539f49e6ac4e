#!/usr/bin/env python
"""Platform probe: concurrent device-to-host copy bandwidth with one process per GPU (torchrun), the transfer
pattern of bench.py's e2e leg without any rendering.  Prints per-rank and aggregate GB/s, the GPU<->CPU topology
and this process's CPU affinity.  Usage: python -m torch.distributed.run --nproc-per-node N tools/d2h_probe.py"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, device="cuda")
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
dst.zero_()


def run(reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


run(2)
together = run()
alone = None
for r in range(world):  # one rank at a time
    if world > 1:
        dist.barrier()
    if r == rank:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        alone = 4 * n / (time.perf_counter() - t0) / 1e9
if world > 1:
    dist.barrier()
t = torch.tensor([together, alone], dtype=torch.float64, device="cuda")
if world > 1:
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
else:
    g = [t]
if rank == 0:
    topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    out = {"n_gpus": world, "together_GBs_per_rank": [round(float(x[0]), 1) for x in g], "alone_GBs_per_rank": [round(float(x[1]), 1) for x in g],
           "aggregate_together_GBs": round(sum(float(x[0]) for x in g), 1), "cpu_count": os.cpu_count(), "affinity_rank0": sorted(os.sched_getaffinity(0))}
    print(json.dumps(out))
    print(topo)
    try:
        print(subprocess.run(["lscpu"], capture_output=True, text=True).stdout[:1500])
    except OSError:
        pass
if world > 1:
    dist.destroy_process_group()
