run() {
  python bench.py --workload ${WL:-attention_scn_train_scaled} --steps 4 --warmup 2 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s value %8.1f ms/step %.3f' % (sys.argv[1], d['value'], d['ms_per_step']))" "$LABEL"
}
for rt in 128 64 32; do for tg in 148 296 444 592; do
LABEL="rowtile $rt target $tg" CAPDEC_GEMM_ROWTILE=$rt CAPDEC_GEMM_TARGET=$tg run
done; done
