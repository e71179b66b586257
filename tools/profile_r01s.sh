#!/bin/bash
# Profile pass after the fused Gaussian accumulation (run under gpurun): launch list of one
# default-workload step and --set full captures of k_gauss_vm / k_gauss_hm.
TAG=${TAG:-r01s}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --streams 1 \
    --no-cpu-baseline --no-score-mode --config cfg2 > gpurun_out/launches_${TAG}.log 2>&1
for k in k_gauss_vm k_gauss_hm; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 2 -c 1 -f \
      -o gpurun_out/prof_${TAG}_${k} python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline \
      --no-score-mode --config cfg1 > gpurun_out/prof_${TAG}_${k}.log 2>&1
done
ls gpurun_out | grep ${TAG}
