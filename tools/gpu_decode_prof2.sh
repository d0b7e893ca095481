#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_beam.py tests/test_gpu_units.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload attention_scn_decode --steps 1 --warmup 3 --no-cpu-baseline > $OUT/dec2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|beam_select|attn_wsum|attn_scores' -s 300 -c 16 \
    -f -o $OUT/dec2_prof python bench.py --workload attention_scn_decode --steps 1 --warmup 3 --no-cpu-baseline > $OUT/dec2_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/dec2_plain.log").read().strip().splitlines()[-1])
print("DECODE", d["value"], d["ms_per_step"])
PY
