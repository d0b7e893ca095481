#!/usr/bin/env python
"""Host-side cost of one training step through the module API (config-3 shape): wall time of the three
host calls with the GPU idle at entry, next to the device time of the captured graphs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "indonesian-image-captioning_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import capdec  # noqa: E402
from oracle import capdec_oracle as O  # noqa: E402
import bench  # noqa: E402

capdec.set_precision("bf16")
capdec.set_graphs(True)
kind, dims, B, _ = bench.WORKLOADS["attention_scn_train"]
torch.manual_seed(0)
dec = bench.make_decoder(kind, dims).cuda().train()
enc, tags, caps, caplens = [t.cuda() for t in O.synthetic_batch(B, dims["V"], seed=1, lengths=[51] * B)]


def step(timed):
    ts = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = dec(enc, tags, caps, caplens)
    t1 = time.perf_counter()
    loss, _ = dec.loss(out[0], out[1], out[2], out[3])
    t2 = time.perf_counter()
    for p in dec.parameters():
        p.grad = None
    loss.backward()
    t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    return [1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t0, t4))]


for _ in range(5):
    step(False)
rows = [step(True) for _ in range(10)]
avg = [sum(r[i] for r in rows) / len(rows) for i in range(5)]
print("host ms: forward call %.3f | loss call %.3f | backward call %.3f | drain %.3f | total wall %.3f" % tuple(avg))
# back-to-back steps without syncs in between
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    out = dec(enc, tags, caps, caplens)
    loss, _ = dec.loss(out[0], out[1], out[2], out[3])
    for p in dec.parameters():
        p.grad = None
    loss.backward()
torch.cuda.synchronize()
print("back-to-back: %.3f ms per step" % (1e3 * (time.perf_counter() - t0) / 20))
