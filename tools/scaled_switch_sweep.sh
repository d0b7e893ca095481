run() {
  python bench.py --workload attention_scn_train_scaled --steps 4 --warmup 2 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s value %8.1f ms/step %.3f' % (sys.argv[1], d['value'], d['ms_per_step']))" "$LABEL"
}
LABEL="default" run
LABEL="stream attention kernels forced" CAPDEC_WSUM_STREAM=1 run
LABEL="no PDL" CAPDEC_PDL=0 run
LABEL="fused epilogues" CAPDEC_FUSED_EPILOGUE=1 run
