#!/usr/bin/env python
"""Per-phase clock stamps of the persistent recurrence kernels (debug): runs one eager
attention_scn forward(+backward) at the config-3 shape with CAPDEC_RECUR_PROF=1; the library
prints, for CTA 0 (row group 0) and a few steps, the cycles between the phase boundaries (wait for the inputs + work)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "indonesian-image-captioning_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import capdec  # noqa: E402
from oracle import capdec_oracle as O  # noqa: E402
import bench  # noqa: E402

kind_name = sys.argv[1] if len(sys.argv) > 1 else "attention_scn"
capdec.set_precision("bf16")
capdec.set_graphs(False)          # the profiling path allocates and synchronises: never inside a capture
kind, dims, B, _ = bench.WORKLOADS[kind_name + "_train"]
B = int(os.environ.get("RECUR_PROF_BATCH", B))          # 16 = one row group alone
torch.manual_seed(0)
dec = bench.make_decoder(kind, dims).cuda().train()
enc, tags, caps, caplens = [t.cuda() for t in O.synthetic_batch(B, dims["V"], seed=1, lengths=[51] * B)]
for it in range(3):
    if it == 2:
        os.environ["CAPDEC_RECUR_PROF"] = "1"
    res = dec(enc, caps, caplens) if kind == "pure_attention" else dec(enc, tags, caps, caplens)
    alphas = None if kind == "pure_scn" else res[3]
    loss, _ = dec.loss(res[0], res[1], res[2], alphas)
    loss.backward()
    torch.cuda.synchronize()
print("done", loss.item())
