#!/bin/bash
# scratch: parity of the bf16 path under each recurrence mode, then per-kernel timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for mode in 2 3; do
  echo "=== RS_REC_MODE=$mode"
  RS_REC_MODE=$mode timeout 600 python -m pytest tests/test_bf16_gpu.py tests/test_varlen_gpu.py -x -q 2>&1 | tail -8
done
for mode in 2 3; do
  for B in 1024 8192; do
    echo "=== mode $mode B $B"
    RS_REC_MODE=$mode timeout 300 python tools/step_probe.py $B 2>&1 | tail -12
  done
done
