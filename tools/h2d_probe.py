"""Concurrent pinned host->device copy bandwidth of the box, per rank, with 1/2/4/8 ranks copying at once
(the platform ceiling of bench.py's e2e leg at N GPUs).  Launch with torchrun --nproc-per-node 8.
PROBE_BIND=1 pins every rank to its GPU's NUMA node first (what bench.py does)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
bound = bind_to_gpu_numa_node(lr) if os.environ.get("PROBE_BIND", "1") == "1" else None
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
GB = 4
host = torch.empty(GB << 30, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
devbuf = torch.empty(GB << 30, dtype=torch.uint8, device=dev)
res = {}
for n in (1, 2, 4, 8):
    if n > world:
        break
    for direction in ("h2d", "d2h"):
        dist.barrier(); torch.cuda.synchronize()
        t = torch.zeros(1, dtype=torch.float64, device=dev)
        if rank < n:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                if direction == "h2d":
                    devbuf.copy_(host, non_blocking=True)
                else:
                    host.copy_(devbuf, non_blocking=True)
            e1.record(); torch.cuda.synchronize()
            t[0] = 3 * GB * 1.073741824 / (e0.elapsed_time(e1) * 1e-3)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        res[f"{direction}_n{n}"] = [round(float(x.item()), 1) for x in allt[:n]]
if rank == 0:
    print(json.dumps({"bind": os.environ.get("PROBE_BIND", "1"), "cpus_bound_rank0": bound, "GBps_per_rank": res}))
dist.destroy_process_group()
