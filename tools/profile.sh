#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of the bench command + one --set full
# capture of each hot kernel.  Outputs land in gpurun_out/ (summaries are copied to profiles/
# by tools/summarize_profiles.py).
#   TAG=r01c KERNELS="k_gray6 k_gauss_v8 ..." LAUNCHES=1 bash tools/profile.sh
TAG=${TAG:-${1:-r01}}
mkdir -p gpurun_out
if [ "${LAUNCHES:-1}" = "1" ]; then
  # launch list of one step of the bench (streams 1 so that launches are serial anyway)
  ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-0} -c ${COUNT:-4000} --csv \
      --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --streams 1 \
      --no-cpu-baseline --no-score-mode --config ${CFG:-cfg1} > gpurun_out/launches_${TAG}.log 2>&1
fi
for k in ${KERNELS:-k_gray6 k_gauss_v k_gauss_h k_sobel_nms k_search k_frame_prepare k_radix_hist}; do
  ncu --set full --clock-control none --import-source on -k regex:^${k} -s ${KSKIP:-2} -c 1 -f \
      -o gpurun_out/prof_${TAG}_${k} python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline \
      --no-score-mode --config cfg1 > gpurun_out/prof_${TAG}_${k}.log 2>&1
done
ls -la gpurun_out
