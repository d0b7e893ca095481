#!/usr/bin/env python
"""GEMM shapes of one beam-search step (625 images x beam 3 = 1875 rows), warm, graph-replayed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench as MB  # noqa: E402

for (r, n, k) in [(1875, 2048, 512), (1875, 2048, 2048), (1875, 4608, 512), (1875, 10000, 512), (1875, 512, 1024),
                  (1920, 2048, 2048), (1792, 2048, 2048)]:
    res = MB.bench_gemm(r, n, k, "bf16", 0)
    print("%-44s %8.1f us  %7.1f TFLOP/s" % (res["case"], res["us"], res["TFLOPs"]), flush=True)
