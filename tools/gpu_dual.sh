#!/bin/bash
# check of the two-row-group persistent kernels: decoder parity tests first (short timeout: a hung
# cooperative kernel must not eat the box), then the full suite, per-phase stamps, bench lines
TAG=${1:-dg}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -25 > $OUT/${TAG}_tests_dec.log
tail -6 $OUT/${TAG}_tests_dec.log
grep -q "failed\|error\|Error" $OUT/${TAG}_tests_dec.log && exit 1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -25 > $OUT/${TAG}_tests.log
tail -4 $OUT/${TAG}_tests.log
timeout 200 python tools/recur_prof.py attention_scn > $OUT/${TAG}_prof.log 2>&1; tail -12 $OUT/${TAG}_prof.log
for w in attention_scn_train pure_scn_train pure_attention_train; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$w.json"))
    print("$w", round(d["value"]), "cap/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "roof", d["roofline"])
except Exception as e:
    print("$w bench failed", e); print(open("$OUT/${TAG}_bench_$w.err").read()[-1500:])
PY
done
