#!/bin/bash
# sweep of the row-group start skew of the two persistent kernels (cycles)
run() {
  python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; o=r.get('other_direction',{})
print('%-28s value %8.1f ms/step %.3f  %s %.2f us/step  %s %.2f us/step' % (sys.argv[1], d['value'], d['ms_per_step'], r['kernel'][6:9], r.get('us_per_decode_step',0), o.get('kernel','      ')[6:9], o.get('us_per_decode_step',0)))" "$LABEL"
}
for s in 0 4000 8000 12000 16000 22000; do
  LABEL="fwd skew $s" CAPDEC_RECUR_SKEW_FWD=$s run
done
for s in 0 8000 14000 20000 26000 32000; do
  LABEL="bwd skew $s" CAPDEC_RECUR_SKEW_BWD=$s run
done
