#!/usr/bin/env python
"""NCCL all-reduce of the flat decoder-gradient buffer (27.2 M fp32 = 108.7 MB), alone, CUDA-event timed.
   python -m torch.distributed.run --nproc-per-node N tools/allreduce_bench.py"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
for n, dt in [(27199761, torch.float32), (27199761, torch.bfloat16), (5120000, torch.float32), (1 << 20, torch.float32)]:
    x = torch.ones(n, dtype=dt, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        dist.all_reduce(x)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    nbytes = n * x.element_size()
    if rank == 0:
        print("all_reduce %9d x %s (%.1f MB), world %d: %.3f ms  algbw %.0f GB/s  busbw %.0f GB/s" % (
            n, str(dt).split(".")[-1], nbytes / 1e6, world, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 * 2 * (world - 1) / world), flush=True)
dist.destroy_process_group()
