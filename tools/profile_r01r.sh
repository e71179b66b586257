#!/bin/bash
# Round-1 closing profile pass (run under gpurun): launch list of one default-workload step and
# --set full captures of the kernels that changed since r01p (k_gray6, k_score_warp x 3 classes).
TAG=${TAG:-r01r}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --streams 1 \
    --no-cpu-baseline --no-score-mode --config cfg2 > gpurun_out/launches_${TAG}.log 2>&1
for k in k_gray6 k_gauss_hm k_sobel_nms; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 2 -c 1 -f \
      -o gpurun_out/prof_${TAG}_${k} python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline \
      --no-score-mode --config cfg1 > gpurun_out/prof_${TAG}_${k}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_score_warp -s 3 -c 1 -f \
    -o gpurun_out/prof_${TAG}_k_score_warp python tools/score_bench.py --config cfg2 --chroms 1,2,3 --stripes 100000 --reps 1 \
    > gpurun_out/prof_${TAG}_k_score_warp.log 2>&1
python tools/score_bench.py --config cfg2 --stripes 100000 --reps 5 > gpurun_out/score_bench_${TAG}.json 2>&1
ls gpurun_out | grep ${TAG}
