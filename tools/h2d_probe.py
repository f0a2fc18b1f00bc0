#!/usr/bin/env python
"""Concurrent host-to-device bandwidth of one box (nvbandwidth-style): under torchrun every rank copies pinned host memory to
its own GPU at the same time; prints the per-GPU and the aggregate rate.  This is the ceiling of the end-to-end (host-buffer)
arm of bench.py at N GPUs: every rank feeds 4.09 GB of level-0 pixels per step from the same host memory system.
usage: python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py [MB per copy] [copies]"""
import json
import os
import sys

import torch
import torch.distributed as dist

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
host = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    dev.copy_(host, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = reps * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
if world > 1:
    lo, tot = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
else:
    lo, tot = t, t
if rank == 0:
    print(json.dumps({"gpus": world, "mb_per_copy": mb, "copies": reps, "h2d_GBps_per_gpu_min": float(lo.item()), "h2d_GBps_aggregate": float(tot.item()),
                      "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
