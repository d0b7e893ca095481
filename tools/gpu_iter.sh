#!/bin/bash
# kernel iteration: decoder parity tests (-x) + per-decode-step latency of the persistent kernels
TAG=${1:-it}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -15 > $OUT/${TAG}_tests.log
tail -4 $OUT/${TAG}_tests.log
LABEL="default B=32" 
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; o=r.get('other_direction',{})
print('value %8.1f ms/step %.3f  %s %.2f us/step  %s %.2f us/step' % (d['value'], d['ms_per_step'], r['kernel'][6:9], r.get('us_per_decode_step',0), o.get('kernel','      ')[6:9], o.get('us_per_decode_step',0)))" | tee $OUT/${TAG}_quick.txt
