#!/usr/bin/env python
"""Host <-> device copy bandwidth with every rank copying at once (torchrun), with and without binding the rank to
the CPUs of its GPU's NUMA node before the pinned buffer is allocated.  Explains / tunes the e2e leg of bench.py at N > 1.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/probe_pcie.py
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from importlib import import_module  # noqa: E402

numa = import_module("parameter-variability_b200.numa")

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 150_000_000   # 1.2 GB of fp64
dev = torch.empty(n, dtype=torch.float64, device="cuda")


def measure(tag):
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    host.zero_()
    out = {}
    for name, fn in (("d2h", lambda: host.copy_(dev, non_blocking=True)), ("h2d", lambda: dev.copy_(host, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = 5 * n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(json.dumps({"rank": rank, "tag": tag, "cpus": len(os.sched_getaffinity(0)), **out}), flush=True)
    del host


info = numa.gpu_numa_info(local)
print(json.dumps({"rank": rank, "numa": info}), flush=True)
measure("default")
numa.bind_to_gpu_node(local)
measure("bound to the GPU's NUMA node")
if world > 1:
    dist.destroy_process_group()
