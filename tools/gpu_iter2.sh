#!/bin/bash
# iteration on the per-step chain kernels: all GPU tests (-x) + decode and scaled bench lines
TAG=${1:-it}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -15 > $OUT/${TAG}_tests.log
tail -4 $OUT/${TAG}_tests.log
for w in attention_scn_decode attention_scn_train_scaled; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$w.json"))
    print("$w", round(d["value"],1), "/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"])
except Exception as e:
    print("$w bench failed", e); print(open("$OUT/${TAG}_bench_$w.err").read()[-1500:])
PY
done
