#!/bin/bash
# tools/gpu.sh <timeout_s> '<command>'  -- gpurun with retries while the pod answers "no slot / draining" (exit 3)
t=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
