#!/bin/bash
# two-GPU check of the overlapped gradient exchange: bench line at N=2 (one process per GPU) + rank-0 timeline
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_train_n2.json 2> $OUT/${TAG}_bench_train_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  tools/dp_timeline.py > $OUT/${TAG}_dp_timeline_n2.txt 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_train_n1_samebox.json 2>/dev/null
tail -c 600 $OUT/${TAG}_bench_train_n2.err; cut -c1-300 $OUT/${TAG}_bench_train_n2.json; cut -c1-200 $OUT/${TAG}_bench_train_n1_samebox.json; tail -40 $OUT/${TAG}_dp_timeline_n2.txt
