#!/bin/bash
# Round-end evidence: tests, smoke, every bench workload, reference arm, ncu launch list + full capture.
TAG=${1:-r1g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -8 > $OUT/${TAG}_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
for w in pure_scn_train pure_attention_train attention_scn_train_scaled; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
done
timeout 600 python bench.py --workload attention_scn_decode --steps 3 --warmup 3 > $OUT/${TAG}_bench_decode.json 2> $OUT/${TAG}_bench_decode.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'recur_fwd_kernel|recur_bwd_kernel|gemm_tc_persist' -s 6 -c 6 \
    -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu2.log 2>&1
tail -3 $OUT/${TAG}_tests.log; tail -2 $OUT/${TAG}_smoke.log
for f in $OUT/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("bench_")[-1], d.get("value"), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
