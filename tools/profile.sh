#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of the bench command + one --set full
# capture of each hot kernel.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
TAG=${1:-r01}
mkdir -p gpurun_out
# launch list of one timed step of the default bench (warm-up launches skipped by count)
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-0} -c ${COUNT:-1200} --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline \
    --config ${CFG:-cfg1} > gpurun_out/launches_${TAG}.log 2>&1
for k in ${KERNELS:-k_gray k_gauss_v k_gauss_h k_sobel_nms k_search k_score k_frame_prepare k_radix_hist}; do
  ncu --set full --clock-control none --import-source on -k regex:^${k} -s 2 -c 1 -f \
      -o gpurun_out/prof_${TAG}_${k} python bench.py --steps 1 --warmup 0 --no-cpu-baseline --config cfg1 \
      > gpurun_out/prof_${TAG}_${k}.log 2>&1
done
ls -la gpurun_out
