#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'   -- retries gpurun while the pod answers "transient" (no slot free)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.txt
exit $rc
