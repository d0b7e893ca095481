#!/bin/bash
# eight-GPU look at the gradient exchange: which NCCL algorithm runs, where the all-reduces sit in the step (rank-0
# timeline), and the bench line under a few NCCL_ALGO / channel settings
OUT=gpurun_out
TAG=${1:-r2c}
N=${2:-8}
run() {
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 20 --warmup 5 --no-secondary --no-cpu-baseline 2>$OUT/${TAG}_nccl_$LABEL.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-34s value %9.1f ms/step %.3f' % (sys.argv[1], d['value'], d['ms_per_step']))" "$LABEL"
}
PORT=29531 LABEL=default run
PORT=29532 LABEL=info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING run
grep -i "nvls\|algo\|channels" $OUT/${TAG}_nccl_info.err | sed 's/^.*NCCL INFO//' | sort | uniq -c | sort -rn | head -12
PORT=29533 LABEL=algo_nvls NCCL_ALGO=NVLS run
PORT=29534 LABEL=algo_ring NCCL_ALGO=Ring run
PORT=29535 LABEL=algo_tree NCCL_ALGO=Tree run
PORT=29536 LABEL=no_early_fc CAPDEC_EARLY_FC=0 run
PORT=29537 LABEL=prio0 CAPDEC_AR_PRIORITY=0 run
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29538 \
  tools/dp_timeline.py > $OUT/${TAG}_dp_timeline_n$N.txt 2>&1
grep -n "nccl\|recur_bwd" $OUT/${TAG}_dp_timeline_n$N.txt | head -20
tail -3 $OUT/${TAG}_dp_timeline_n$N.txt
