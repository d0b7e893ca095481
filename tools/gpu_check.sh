#!/bin/bash
# One B200 call: parity tests, smoke, the bench lines, and the ncu evidence (launch list + one
# full capture of the dominant kernels).  Usage: tools/gpu_check.sh <tag> [quick]
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -25 > $OUT/${TAG}_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
timeout 600 python bench.py --workload attention_scn_decode --steps 3 --warmup 3 > $OUT/${TAG}_bench_decode.json 2> $OUT/${TAG}_bench_decode.err
if [ "${2:-}" != "quick" ]; then
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
  timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'recur_fwd_kernel|recur_bwd_kernel' -s 4 -c 4 \
      -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu2.log 2>&1
fi
tail -5 $OUT/${TAG}_tests.log; cat $OUT/${TAG}_smoke.log | tail -3; cat $OUT/${TAG}_bench_train.json $OUT/${TAG}_bench_decode.json; tail -3 $OUT/${TAG}_bench_decode.err
