#!/bin/bash
# launch list of the beam-search decode bench (one step = 625 images x beam 3, 51 decode steps)
OUT=gpurun_out
python bench.py --workload attention_scn_decode --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/dec_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1200 --csv \
    --log-file $OUT/dec_launches.csv python bench.py --workload attention_scn_decode --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/dec_ncu.log 2>&1
tail -2 $OUT/dec_plain.log | cut -c1-300
