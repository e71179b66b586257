#!/bin/bash
# Closing pass of the round (run under gpurun): launch list of one default-workload step and a
# --set full capture of k_int_hist (the one-pass raw-count quantile kernel).
TAG=${TAG:-r01t}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --streams 1 \
    --no-cpu-baseline --no-score-mode --config cfg2 > gpurun_out/launches_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_int_hist -s 1 -c 1 -f \
    -o gpurun_out/prof_${TAG}_k_int_hist python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline \
    --no-score-mode --config cfg2 > gpurun_out/prof_${TAG}_k_int_hist.log 2>&1
ls gpurun_out | grep ${TAG}
