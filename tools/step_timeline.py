#!/usr/bin/env python
"""GPU timeline of ONE training step (kernels, memcpys, memsets, idle gaps) from torch.profiler / CUPTI:
what a kernel list cannot show (copies, gaps between launches).  python tools/step_timeline.py [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "indonesian-image-captioning_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import capdec  # noqa: E402
from oracle import capdec_oracle as O  # noqa: E402
import bench  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
summary = "--summary" in sys.argv          # aggregate by kernel name instead of listing every activity
name = args[0] if args else "attention_scn_train"
capdec.set_precision("bf16")
capdec.set_graphs(True)
kind, dims, B, _ = bench.WORKLOADS[name]
torch.manual_seed(0)
dec = bench.make_decoder(kind, dims).cuda().train()
enc, tags, caps, caplens = [t.cuda() for t in O.synthetic_batch(B, dims["V"], seed=1, lengths=[51] * B)]


def step():
    res = dec(enc, caps, caplens) if kind == "pure_attention" else dec(enc, tags, caps, caplens)
    alphas = None if kind == "pure_scn" else res[3]
    loss, _ = dec.loss(res[0], res[1], res[2], alphas)
    for p in dec.parameters():
        p.grad = None
    loss.backward()
    return loss


for _ in range(6):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# the last step = events after the second-to-last recur_fwd kernel start
starts = [i for i, e in enumerate(evs) if "gather_features" in e.name]
lo = starts[-1] if starts else 0
sel = evs[lo:]
t0 = sel[0].time_range.start
end_prev = t0
busy = 0.0
gaps = []
if summary:
    import collections
    agg = collections.OrderedDict()
    for e in sel:
        k = e.name[:110]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    print("%6s %10s %8s  %s" % ("count", "total_us", "avg_us", "name"))
    for k, (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6d %10.1f %8.2f  %s" % (n, tot, tot / n, k))
print("%9s %8s %7s  %s" % ("start_us", "dur_us", "gap_us", "name"))
for e in sel:
    if summary:
        st, du = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = e.time_range.start - end_prev
        if gap > 0:
            gaps.append(gap)
        busy += du
        end_prev = max(end_prev, e.time_range.end)
        continue
    st, du = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - end_prev
    if gap > 0:
        gaps.append(gap)
    busy += du
    print("%9.1f %8.1f %7.1f  %s" % (st, du, max(gap, 0.0), e.name[:90]))
    end_prev = max(end_prev, e.time_range.end)
print("step span %.1f us, busy %.1f us, idle %.1f us in %d gaps, %d device activities" % (
    end_prev - t0, busy, sum(gaps), len(gaps), len(sel)))
