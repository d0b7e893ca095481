#!/bin/bash
# Round-2 evidence call: parity tests, smoke, the contract bench line (with its secondary block), ncu launch
# list + one full capture of the persistent kernels, per-phase stamps.  Usage: tools/gpu_r2.sh <tag> [notests]
set -u
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" != "notests" ]; then
  timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -25 > $OUT/${TAG}_tests.log
  timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
fi
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'recur_fwd_kernel|recur_bwd_kernel' -s 4 -c 4 \
    -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu2.log 2>&1
timeout 300 python tools/recur_prof.py > $OUT/${TAG}_recur_prof.txt 2>&1
tail -5 $OUT/${TAG}_tests.log 2>/dev/null; tail -3 $OUT/${TAG}_smoke.log 2>/dev/null; cat $OUT/${TAG}_bench_train.json; tail -3 $OUT/${TAG}_bench_train.err
