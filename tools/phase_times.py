#!/usr/bin/env python
"""Where does a training step go?  Times the captured forward / backward CUDA graphs of the
attention_scn decoder separately, for several caption lengths: the slope over T is the cost of one
recurrence step, the intercept is the batched (non-recurrent) part.

    python tools/phase_times.py [--kind attention_scn] [--batch 32]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "indonesian-image-captioning_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import capdec  # noqa: E402
from oracle import capdec_oracle as O  # noqa: E402  (synthetic inputs only)
import bench  # noqa: E402


def time_graph(g, n=10):
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="attention_scn")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    capdec.set_precision(args.precision)
    capdec.set_graphs(True)
    kind, dims, _, _ = bench.WORKLOADS[args.kind + "_train"]
    torch.manual_seed(0)
    dec = bench.make_decoder(kind, dims).cuda().train()
    out = []
    for cap_len in (51, 26, 11):
        enc, tags, caps, caplens = [t.cuda() for t in O.synthetic_batch(args.batch, dims["V"], seed=1,
                                                                        lengths=[cap_len] * args.batch)]
        meta = None
        for _ in range(4):
            res = dec(enc, caps, caplens) if kind == "pure_attention" else dec(enc, tags, caps, caplens)
            alphas = None if kind == "pure_scn" else res[3]
            loss, _ = dec.loss(res[0], res[1], res[2], alphas)
            for p in dec.parameters():
                p.grad = None
            loss.backward()
            meta = res[0]._capdec_meta
        torch.cuda.synchronize()
        plan = meta["plan"]
        row = {"T": cap_len - 1, "fwd_ms": time_graph(plan.graphs["fwd"]),
               "bwd_ms": time_graph(plan.graphs["bwd_fused"]),
               "fwd_kernels": plan.graph_nodes["fwd"], "bwd_kernels": plan.graph_nodes["bwd_fused"]}
        out.append(row)
        print(json.dumps(row))
    a, b = out[0], out[1]
    dT = a["T"] - b["T"]
    print(json.dumps({"fwd_us_per_step": 1e3 * (a["fwd_ms"] - b["fwd_ms"]) / dT,
                      "bwd_us_per_step": 1e3 * (a["bwd_ms"] - b["bwd_ms"]) / dT,
                      "fwd_fixed_ms": a["fwd_ms"] - a["T"] * (a["fwd_ms"] - b["fwd_ms"]) / dT,
                      "bwd_fixed_ms": a["bwd_ms"] - a["T"] * (a["bwd_ms"] - b["bwd_ms"]) / dT}))


if __name__ == "__main__":
    main()
