#!/bin/bash
# scratch: per-kernel timings of one training step under several environment settings.  usage: ab_env.sh "<B list>" "ENV1=.. ENV2=.." "..."
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
sizes=$1; shift
{
for B in $sizes; do
  for envs in "$@"; do
    echo "=== B $B  $envs"
    env $envs timeout 300 python tools/step_probe.py $B 2>&1 | grep "rec_\|ms/step" | cut -c1-60
  done
done
} > gpurun_out/ab_env.txt 2>&1
cat gpurun_out/ab_env.txt
