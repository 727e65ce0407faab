#!/bin/bash
# tools/gpu_retry.sh <timeout_s> '<command>' -- gpurun with retries while the pod answers "busy" (exit code 3)
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
