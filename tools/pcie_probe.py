#!/usr/bin/env python
"""Host <-> device copy bandwidth with ALL ranks copying at the same time (what bench.py's e2e leg does every step).
Answers: is the e2e figure limited by a link the GPUs of a box share, and does binding each rank (threads + pinned
buffers) to its GPU's NUMA node help?   torchrun --nproc-per-node N tools/pcie_probe.py [MiB per direction]
Prints one JSON line per rank and mode: {rank, numa_bound, h2d_GBs, d2h_GBs, both_GBs_per_direction}."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sylt2d_b200.numa import bind_to_gpu_numa, gpu_numa_info  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
ws = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(lr)
if ws > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = mib << 20


def barrier():
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(tag):
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    res = {"rank": rank, "mode": tag, "MiB": mib, "cpus": sorted(os.sched_getaffinity(0))[:4] + ["..."], "ncpus": len(os.sched_getaffinity(0))}
    for name, f in (("h2d", h2d), ("d2h", d2h), ("both", both)):
        f()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        barrier()
        res[name + "_GBs_per_direction"] = n / dt / 1e9
    print(json.dumps(res), flush=True)


if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
print(json.dumps({"rank": rank, "gpu_numa": gpu_numa_info(lr)}), flush=True)
measure("default affinity")
bind_to_gpu_numa(lr)
measure("bound to the GPU's NUMA node")
if ws > 1:
    dist.destroy_process_group()
