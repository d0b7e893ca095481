#!/bin/bash
# launch list of the scaled training workload (config 5 shape: 128 captions, D = F = 1024, V = 30k)
OUT=gpurun_out
python bench.py --workload attention_scn_train_scaled --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/sc_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv \
    --log-file $OUT/sc_launches.csv python bench.py --workload attention_scn_train_scaled --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/sc_ncu.log 2>&1
tail -1 $OUT/sc_plain.log | cut -c1-200
