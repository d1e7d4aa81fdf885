#!/bin/bash
# the round's ncu captures (one GPU): step kernels at 8192, latency mode at 1024, H = 256, and the bench launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/step_once.py 8192 > gpurun_out/plain_step8192.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rec_|blk_' --launch-skip 7 --launch-count 7 -f \
    -o gpurun_out/r2e_step_b8192 python tools/step_once.py 8192 > gpurun_out/ncu_step8192.log 2>&1
python tools/rec_probe.py 1024 100 > gpurun_out/plain_rec.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rec_ --launch-skip 4 --launch-count 4 -f \
    -o gpurun_out/r2e_pair_b1024 python tools/rec_probe.py 1024 100 > gpurun_out/ncu_pair.log 2>&1
python tools/wide_once.py > gpurun_out/plain_wide.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rec_ --launch-skip 4 --launch-count 4 -f \
    -o gpurun_out/r2e_wide python tools/wide_once.py > gpurun_out/ncu_wide.log 2>&1
python bench.py --steps 2 --warmup 3 --skip-c4 --heatmap-traces 200000 > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2e_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-c4 --heatmap-traces 200000 > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
ls -la gpurun_out/r2e_* ; tail -2 gpurun_out/ncu_step8192.log gpurun_out/ncu_pair.log gpurun_out/ncu_wide.log
