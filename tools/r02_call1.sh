#!/bin/bash
# round-2 call 1: pipe micro-benchmark + same-box baselines of the round-1 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/r02a_smi_before.txt
tools/bin/fp64_peak gpurun_out/r02a_fp64_peak.json
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/r02a_smi_after.txt
cat gpurun_out/r02a_fp64_peak.json
for cfg in cfg3 cfg2 cfg1; do
  python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_$cfg.json 2> gpurun_out/r02a_bench_$cfg.err
  tail -c 600 gpurun_out/r02a_bench_$cfg.json
done
