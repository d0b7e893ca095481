#!/usr/bin/env python
"""GPU timeline of ONE data-parallel training step on rank 0 (torch.profiler / CUPTI): shows where the NCCL
all-reduces of the gradient buckets sit relative to the backward's kernels.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_timeline.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "indonesian-image-captioning_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import capdec  # noqa: E402
from capdec import parallel as cpar  # noqa: E402
from oracle import capdec_oracle as O  # noqa: E402
import bench  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
capdec.set_precision("bf16")
kind, dims, B, _ = bench.WORKLOADS["attention_scn_train"]
torch.manual_seed(0)
dec = bench.make_decoder(kind, dims).to(dev).train()
enc, tags, caps, caplens = [t.to(dev) for t in O.synthetic_batch(B, dims["V"], seed=1 + rank, lengths=[51] * B)]
red = cpar.GradReducer(dec, dist)


def step():
    res = dec(enc, tags, caps, caplens)
    loss, _ = dec.loss(res[0], res[1], res[2], res[3], alpha_c=1.0 / world, n_tokens=B * 50 * world)
    for p in dec.parameters():
        p.grad = None
    loss.backward()
    red.allreduce(res[0]._capdec_meta)
    return loss


for _ in range(6):
    step()
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    starts = [i for i, e in enumerate(evs) if "ce_bwd" in e.name]
    lo = starts[-1] if starts else 0
    t0 = evs[lo].time_range.start
    print("%9s %8s  %s" % ("start_us", "dur_us", "name"))
    for e in evs[lo:]:
        du = e.time_range.end - e.time_range.start
        if du < 3.0 and "nccl" not in e.name.lower():
            continue
        print("%9.1f %8.1f  %s" % (e.time_range.start - t0, du, e.name[:100]))
dist.barrier()
dist.destroy_process_group()
