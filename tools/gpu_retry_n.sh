#!/bin/bash
# usage: tools/gpu_retry_n.sh <gpus> <timeout_s> '<command>'
N=$1; T=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@" > /tmp/gpurun_last_n.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last_n.txt || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -60 /tmp/gpurun_last_n.txt
exit $rc
