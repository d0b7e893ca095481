#!/bin/bash
# builds the debug variant of the library with tagged clock stamps inside the phases of the persistent kernels
# (-DCAPDEC_RECUR_FINE) next to the normal one: indonesian-image-captioning_b200/libcapdec_fine.so.  Use it with
#   CAPDEC_LIB=$PWD/indonesian-image-captioning_b200/libcapdec_fine.so python tools/recur_prof.py
set -e
cd "$(dirname "$0")/../indonesian-image-captioning_b200"
CAPDEC_NVCC_FLAGS=-DCAPDEC_RECUR_FINE python capdec/build.py --force > /dev/null
cp libcapdec.so libcapdec_fine.so
python capdec/build.py --force > /dev/null
ls -la libcapdec.so libcapdec_fine.so
