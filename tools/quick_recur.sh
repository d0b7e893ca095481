#!/bin/bash
# quick sweep: per-decode-step latency of the two persistent kernels (bench.py roofline leg) under env / batch variations
run() {
  python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; o=r.get('other_direction',{})
print('%-40s value %8.1f ms/step %.3f  %s %.1f us/step  %s %.1f us/step' % (sys.argv[1], d['value'], d['ms_per_step'], r['kernel'][6:9], r.get('us_per_decode_step',0), o.get('kernel','      ')[6:9], o.get('us_per_decode_step',0)))" "$LABEL"
}
LABEL="default B=32" run
LABEL="B=16 (one row group)" run --batch 16
for s in; do
  LABEL="skew fwd=bwd=$s" CAPDEC_RECUR_SKEW=$s run
done
