#!/bin/bash
OUT=gpurun_out
N=8
run() {
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 12 --warmup 4 --no-secondary --no-cpu-baseline 2>$OUT/r2c_proto_$LABEL.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-34s value %9.1f ms/step %.3f' % (sys.argv[1], d['value'], d['ms_per_step']))" "$LABEL"
}
PORT=29541 LABEL=proto_simple NCCL_PROTO=Simple run &
wait
PORT=29542 LABEL=proto_ll128 NCCL_PROTO=LL128 run
