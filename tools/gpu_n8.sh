#!/bin/bash
# eight-GPU check: the contract bench line at N=8 (one process per GPU over NCCL), with its secondary block
TAG=${1:-r2c}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_train_n$N.json 2> $OUT/${TAG}_bench_train_n$N.err
tail -c 400 $OUT/${TAG}_bench_train_n$N.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_train_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k,v in (d.get("secondary") or {}).items():
    print("  ", k, v.get("value"), v.get("ms_per_step"))
PY
