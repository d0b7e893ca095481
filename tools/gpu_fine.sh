#!/bin/bash
# fine-grained phase stamps of the persistent kernels (debug build libcapdec_fine.so: tools/build_fine.sh first) + one-row-group comparison
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
CAPDEC_LIB=$PWD/indonesian-image-captioning_b200/libcapdec_fine.so timeout 200 python tools/recur_prof.py > $OUT/${TAG}_fine.txt 2>&1
bash tools/quick_recur.sh > $OUT/${TAG}_quick.txt 2>&1
cat $OUT/${TAG}_fine.txt | grep -v "all-CTA\|boundary" ; cat $OUT/${TAG}_quick.txt
