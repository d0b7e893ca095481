#!/usr/bin/env python
"""cProfile of the host side of the module forward (config-3 shape, graph mode)."""
import cProfile
import os
import pstats
import sys

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


def fwd():
    torch.cuda.synchronize()
    return dec(enc, tags, caps, caplens)


def full():
    out = fwd()
    loss, _ = dec.loss(out[0], out[1], out[2], out[3])
    for p in dec.parameters():
        p.grad = None
    loss.backward()


for _ in range(5):
    full()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    full()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
