#!/usr/bin/env python
"""Kernel micro-benchmarks on one B200: every case is captured into a CUDA graph of N launches and
replayed, so the numbers are device time per launch without Python/ctypes launch overhead.

    python tools/microbench.py [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "indonesian-image-captioning_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
from capdec import functional as CF  # noqa: E402


def graph_time_us(fn, n=20, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, 1e3 * s.elapsed_time(e) / n)
    return best


def bench_attention(rows, precision, P=196, E=2048, A=512):
    dev = "cuda"
    ft = torch.bfloat16 if precision == "bf16" else torch.float32
    att1 = torch.randn(rows, P, A, device=dev).to(ft)
    enc = torch.randn(rows, P, E, device=dev).relu_().to(ft)
    g1 = torch.randn(rows, A + E, device=dev)
    w_f = torch.randn(A, device=dev) * 0.05
    b_f = torch.zeros(1, device=dev)
    us = graph_time_us(lambda: CF.attention_step(att1, enc, g1, A, w_f, b_f, precision=precision))
    nbytes = rows * P * (A + E) * (2 if precision == "bf16" else 4)
    return {"case": "attn_fwd rows=%d %s" % (rows, precision), "us": us, "GBps": nbytes / us / 1e3}


def bench_gemm(rows, N, K, precision, splitk=0):
    dev = "cuda"
    ft = torch.bfloat16 if precision == "bf16" else torch.float32
    X = torch.randn(rows, K, device=dev).to(ft)
    W = torch.randn(N, K, device=dev).to(ft)
    out = torch.zeros(rows, N, device=dev)
    us = graph_time_us(lambda: CF.gemm(X, W, precision=precision, splitk=splitk, out=out))
    return {"case": "gemm rows=%d N=%d K=%d %s splitk=%d" % (rows, N, K, precision, splitk), "us": us,
            "TFLOPs": 2.0 * rows * N * K / us / 1e6, "W_GBps": N * K * (2 if precision == "bf16" else 4) / us / 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    res = []
    for rows in (32, 128):
        res.append(bench_attention(rows, "bf16"))
    res.append(bench_attention(32, "fp32"))
    step_shapes = [(32, 4608, 512), (32, 2048, 2048), (32, 512, 1024), (32, 512, 2048), (32, 2048, 2048),
                   (32, 512, 2560)]
    for (r, n, k) in step_shapes:
        for sk in (0, -1):
            res.append(bench_gemm(r, n, k, "bf16", sk))
    for (r, n, k) in [(6272, 512, 2048), (1600, 10000, 512), (1600, 512, 10000), (10000, 512, 1600),
                      (1600, 2048, 512), (2048, 2048, 1600)]:
        res.append(bench_gemm(r, n, k, "bf16", 0))
    res.append(bench_gemm(32, 2048, 2048, "fp32", 0))
    for r in res:
        print(json.dumps(r))
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
