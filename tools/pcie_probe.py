#!/usr/bin/env python
"""Pinned-host <-> device copy bandwidth per GPU, alone and with all ranks copying at once (names the limiter of the e2e arm).

    python tools/pcie_probe.py                                  # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
GB = 2.0
n = int(GB * 1e9)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


res = {}
for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("h2d+d2h", both, 2 * n)):
    res[name + "_all_ranks_GBs"] = nbytes / timed(fn) / 1e9
if world > 1:       # one rank at a time
    for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n)):
        for r in range(world):
            dist.barrier()
            if r == rank:
                fn(); torch.cuda.synchronize(dev)
                t0 = time.perf_counter(); fn(); fn(); torch.cuda.synchronize(dev)
                res[name + "_alone_GBs"] = 2 * nbytes / (time.perf_counter() - t0) / 1e9
            dist.barrier()
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print(json.dumps({"world": world, "per_rank": out, "cpus": os.cpu_count()}, indent=1))
    dist.destroy_process_group()
else:
    print(json.dumps({"world": 1, "per_rank": [res], "cpus": os.cpu_count()}, indent=1))
