#!/bin/bash
# full ncu capture of the top kernels of the beam-search decode bench (after a plain run exited 0)
OUT=gpurun_out
TAG=${1:-dec}
python bench.py --workload attention_scn_decode --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_wsum_stream|attn_scores_stream|gemm_tc_persist|beam_select' -s 40 -c 8 \
    -f -o $OUT/${TAG}_prof python bench.py --workload attention_scn_decode --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu2.log 2>&1
tail -1 $OUT/${TAG}_plain2.log | cut -c1-200
