#!/bin/bash
# retry gpurun while the pod answers "busy" (exit code 3: nothing charged); usage: gpurun_retry.sh <timeout> [--gpus N] -- '<command>'
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
