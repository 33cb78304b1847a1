"""Pinned host -> device bandwidth of this host with N ranks copying at the same time (the e2e arm
at N = 8 reached 22.7 GB/s per GPU against 43 GB/s at N = 1 -- is that the host's ceiling?).

    python scripts/micro/h2d_bw.py                                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/micro/h2d_bw.py                                                     # N ranks, one per GPU

Prints one JSON line per copy size on rank 0: GB/s per GPU (slowest rank) and aggregate.  The GPU boxes of this pool are
KVM guests with ONE visible NUMA node (lscpu, nvidia-smi topo -m), so there is no NUMA placement to choose from inside
the guest; what can be measured is the ceiling itself, which bench.py re-measures in every run (`e2e.h2d_frac`).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import measure_h2d_ceiling  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def sync_all():
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize(dev)


for mb in (10.6, 85, 340, 1024):
    gbs = measure_h2d_ceiling(dev, int(mb * 1e6), sync_all)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"copy_mb": mb, "ranks": world, "gbs_per_gpu_slowest": round(gbs, 2), "gbs_aggregate_lower_bound": round(gbs * world, 1)}), flush=True)
if world > 1:
    dist.destroy_process_group()
