#!/bin/bash
# quick B200 check: GPU tests (under a timeout so a hung kernel cannot eat the box), smoke, bench lines
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -25 > $OUT/${TAG}_tests.log
tail -8 $OUT/${TAG}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_train.json"))
    print("TRAIN", d["value"], "cap/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roof", d["roofline"]["frac"] if d["roofline"] else None)
except Exception as e:
    print("train bench failed", e); print(open("$OUT/${TAG}_bench_train.err").read()[-2000:])
PY
if [ "${2:-}" == "all" ]; then
for w in pure_scn_train pure_attention_train attention_scn_decode; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$w.json"))
    print("$w", d["value"], "cap/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
except Exception as e:
    print("$w bench failed", e); print(open("$OUT/${TAG}_bench_$w.err").read()[-1500:])
PY
done
fi
