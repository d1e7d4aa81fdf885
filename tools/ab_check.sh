#!/bin/bash
# scratch: parity of the bf16 path with the current build, then A/B per-kernel timings against libroomslam_b200_base.so
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
for mode in 2 3; do
  echo "=== parity RS_REC_MODE=$mode"
  RS_REC_MODE=$mode timeout 900 python -m pytest tests/test_bf16_gpu.py tests/test_varlen_gpu.py tests/test_configs_gpu.py -x -q 2>&1 | tail -6
done
for B in ${AB_SIZES:-1024 8192}; do
  for lib in base new; do
    echo "=== $lib B $B"
    if [ $lib = base ]; then
      [ -f roomslam_b200/libroomslam_b200_base.so ] || { echo "(no roomslam_b200/libroomslam_b200_base.so: copy the library there before a change to A/B against it)"; continue; }
      export RS_LIB=$PWD/roomslam_b200/libroomslam_b200_base.so
    else unset RS_LIB; fi
    timeout 300 python tools/step_probe.py $B 2>&1 | tail -9
  done
done
} > gpurun_out/ab_check.txt 2>&1
tail -70 gpurun_out/ab_check.txt
