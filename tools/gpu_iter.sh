#!/bin/bash
# kernel iteration: decoder parity tests (-x) + per-decode-step latency of the persistent kernels
TAG=${1:-it}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | grep -v "^E  *+\|tensor(\[" | tail -15 > $OUT/${TAG}_tests.log
tail -4 $OUT/${TAG}_tests.log
bash tools/quick_recur.sh 2>&1 | tee $OUT/${TAG}_quick.txt
